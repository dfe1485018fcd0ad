"""The three weight-gradient contractions of the Yelp step (K = batch = 400, catalogue-sized fp32 output) next to the
device's write-only and copy bandwidth: they are bound by writing the gradient, not by the tensor pipe.
usage: python tools/wgrad_case.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import kernels as K  # noqa: E402


def op(rows, cols, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.zeros(rows, K.round_up(cols, 64), dtype=torch.bfloat16, device="cuda")
    t[:, :cols] = (torch.randn(rows, cols, generator=g, device="cuda") * 0.1).to(torch.bfloat16)
    return t


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, m, n in (("dE", 34395, 3000), ("dW1", 1000, 34405), ("dW1_U", 1000, 68800)):
        k = 400
        a, b = op(m, k, 1), op(n, k, 2)
        out = torch.empty(m, K.round_up(n, 4), device="cuda")
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.gemm([a], [b], m, n, [k], out_f32=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ref = a[:256, :k].float() @ b[:, :k].float().t()
        err = (out[:256, :n] - ref).abs().max().item() / ref.abs().max().item()
        t = sorted(ts[2:])[len(ts[2:]) // 2]
        print(f"{name}: {m} x {n} x {k}: {t:.1f} us  {2.0 * m * n * k / t / 1e6:.0f} TF/s  output {m * n * 4 / t / 1e3:.0f} GB/s  rel err {err:.1e}")


if __name__ == "__main__":
    main()


def hbm_probe():
    """Write-only and copy bandwidth of the device (what bounds a contraction whose output is catalogue-sized)."""
    n = 412 << 20
    x = torch.empty(n, dtype=torch.uint8, device="cuda")
    y = torch.empty(n, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, fn, bytes_ in (("fill (write only)", lambda: x.zero_(), n), ("copy (read + write)", lambda: y.copy_(x), 2 * n)):
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        print(f"{name}: {bytes_ / 1e6:.0f} MB in {t:.1f} us = {bytes_ / t / 1e3:.0f} GB/s")


if __name__ == "__main__" and os.environ.get("HBM_PROBE", "1") == "1":
    hbm_probe()
