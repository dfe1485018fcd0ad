"""Probe (run under torchrun, N >= 2): can an NCCL all-reduce issued on a side stream be captured inside a CUDA graph
together with compute kernels, and what bandwidth does a 412 MB / 137 MB all-reduce reach?
usage: torchrun --nproc-per-node N tools/nccl_graph_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", rank=rank, world_size=world)
    dev = torch.device("cuda", local)
    big = torch.ones(34395 * 3000, device=dev)
    mid = torch.ones(1000 * 34405, device=dev)
    a = torch.randn(4096, 4096, device=dev)
    comm = torch.cuda.Stream()
    # bandwidth of plain all-reduces
    for name, t in (("412MB", big), ("137MB", mid)):
        for _ in range(3):
            td.all_reduce(t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            td.all_reduce(t)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if rank == 0:
            print(f"all_reduce {name}: {ms:.3f} ms -> algbw {t.numel() * 4 / ms / 1e6:.1f} GB/s", flush=True)
        t.fill_(1.0)

    def body():
        x = a @ a
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(comm):
            comm.wait_event(ev)
            td.all_reduce(big)
            done = torch.cuda.Event()
            done.record()
        y = x @ a  # overlaps with the all-reduce
        torch.cuda.current_stream().wait_event(done)
        return y.sum() + big[0]

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    ok = True
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = body()
        big.fill_(1.0)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 5
        val = big[0].item()
        expect = float(world) ** 5
        if rank == 0:
            print(f"graph capture of NCCL on a side stream: OK, replay {dt * 1e3:.3f} ms, big[0]={val} expect {expect}", flush=True)
    except Exception as ex:  # noqa: BLE001
        ok = False
        if rank == 0:
            print("graph capture of NCCL FAILED:", repr(ex)[:500], flush=True)
    td.barrier()
    td.destroy_process_group()
    return ok


if __name__ == "__main__":
    main()
