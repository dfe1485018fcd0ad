"""Fixed cost of a small contraction launch: 20 back-to-back (400 x 3000 x 512) GEMMs inside one event interval, with and
without a tiny elementwise kernel between them (which switches the SM's shared-memory carve-out back and forth).
usage: python tools/gemm_fixed_cost.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import kernels as K  # noqa: E402


def op(rows, cols, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.zeros(rows, K.round_up(cols, 64), dtype=torch.bfloat16, device="cuda")
    t[:, :cols] = (torch.randn(rows, cols, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    return t


def main():
    m, n, k = 400, 3000, 512
    a, b = op(m, k, 1), op(n, k, 2)
    out = torch.empty(m, n, device="cuda")
    tiny = torch.zeros(1024, device="cuda")
    for label, between in (("back-to-back", False), ("with a tiny elementwise kernel between", True)):
        for graph in (False, True):
            def body():
                for _ in range(20):
                    K.gemm([a], [b], m, n, [k], out_f32=out, splits=1)
                    if between:
                        tiny.add_(1.0)
            body()
            torch.cuda.synchronize()
            if graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    body()
                run = g.replay
            else:
                run = body
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if not graph:
                torch.cuda._sleep(int(1e7))
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            print(f"{label:42s} {'graph' if graph else 'eager'}: {e0.elapsed_time(e1) * 1e3 / 20:.2f} us per iteration")


if __name__ == "__main__":
    main()
