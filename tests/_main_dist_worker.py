"""Data-parallel run of the drop-in driver under torchrun (N >= 2, NCCL) on a tiny synthetic dataset with the big-matrix
machinery forced on (row-sharded AdamW, factor exchange, bf16 operand all-gather): 4 epochs uninterrupted == 2 epochs +
checkpoint + resume for 2 more, bit for bit, on every rank; ranks hold identical weights.
usage: torchrun --nproc-per-node N tests/_main_dist_worker.py <tmp dir>"""
import glob
import os
import sys

import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GDMCF_SHARD_MIN_BYTES"] = str(1 << 16)
from gdmcf_b200 import main as M  # noqa: E402
from gdmcf_b200.parse_args_util import parse_args  # noqa: E402


def run(tmp, name, extra):
    argv = ["--synthetic", "600,500,15000", "--dims", "64", "--batch_size", "100", "--steps", "5", "--noise_scale", "0.01",
            "--topN", "[10, 20]", "--lr", "0.001", "--eval_every", "2", "--debug", "true", "--log_name", tmp,
            "--dataset", "tiny", "--out_name", name] + extra
    out = sys.stdout
    try:
        res = M.main(parse_args(argv))
    finally:
        sys.stdout = out  # main() points non-zero ranks' stdout at /dev/null
    return res, {k: v.detach().clone() for k, v in M.main.last_model.state_dict().items()}


def main():
    tmp = sys.argv[1]
    rank = int(os.environ.get("RANK", "0"))
    # one process group for the three runs (main() only tears down a group it created itself)
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    td.init_process_group(backend="nccl", rank=rank, world_size=int(os.environ["WORLD_SIZE"]))
    res_full, sd_full = run(tmp, "full", ["--epochs", "4"])
    run(tmp, "part", ["--epochs", "2", "--checkpoint_every", "2"])
    ck = glob.glob(os.path.join(tmp, "tiny", "*", "part", "checkpoint.pt"))
    assert len(ck) == 1, ck
    res_res, sd_res = run(tmp, "resumed", ["--epochs", "4", "--resume", ck[0]])
    bad = [k for k in sd_full if not torch.equal(sd_full[k], sd_res[k])]
    ok = not bad and res_res == res_full
    if bad:
        print(f"rank {rank}: resumed run differs in {bad[:5]} "
              f"(max {max((sd_full[k].float() - sd_res[k].float()).abs().max().item() for k in bad):.3e})", flush=True)
    if res_res != res_full:
        print(f"rank {rank}: results differ: {res_full} vs {res_res}", flush=True)
    print(f"main_dist_check rank {rank}: {'OK' if ok else 'FAILED'}", file=sys.stderr, flush=True)
    td.barrier()
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
