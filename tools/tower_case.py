"""Time gdmcf_user_tower against the two contractions + mix it replaces (Yelp shape: 400 x 3000 -> 512 -> 3000).
usage: python tools/tower_case.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import kernels as K  # noqa: E402
from gdmcf_b200.kernels import Bf16Mat  # noqa: E402


def op(rows, cols, seed, scale=0.05):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.zeros(rows, K.round_up(cols, 64), dtype=torch.bfloat16, device="cuda")
    t[:, :cols] = (torch.randn(rows, cols, generator=g, device="cuda") * scale).to(torch.bfloat16)
    return Bf16Mat(t, None, rows, cols)


def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def main():
    B, d, H = 400, 1000, 512
    d3 = 3 * d
    hc_f32 = torch.randn(B, d3, device="cuda") * 0.3
    hc = K.cast_bf16(hc_f32)
    w1, w2 = op(H, d3, 1), op(d3, H, 2)
    b1, b2 = torch.zeros(H, device="cuda"), torch.zeros(d3, device="cuda")
    sumw = torch.tensor(0.7, device="cuda")
    out = Bf16Mat.empty(B, d3, "cuda")
    inv_u = torch.empty(B, device="cuda")
    g1 = Bf16Mat.empty(B, H, "cuda")
    g2 = torch.empty(B, d3, device="cuda")
    for budget in (0, 96, 64):
        t = timeit(lambda: K.user_tower(hc, hc_f32, w1, b1, w2, b2, sumw, B, out=out, inv_u=inv_u, max_ctas=budget))
        ws = [v for k, v in K._tower_ws.items()][0]
        print(f"user_tower (max_ctas {budget}): {t:.2f} us; CTA 0 ns since entry: phase1 done {ws[1][12].item()}, barrier A passed "
              f"{ws[1][13].item()}, reduce + barrier B {ws[1][14].item()}, phase 2 done {ws[1][15].item()}")

    def old():
        K.gemm([hc.hi], [w1.hi], B, H, [d3], act=K.ACT_RELU, bias=b1, out_bf16=g1.hi)
        K.gemm([g1.hi], [w2.hi], B, d3, [H], bias=b2, out_f32=g2)
        K.mix_rownorm(hc_f32, B, d3, g=g2, sumw=sumw, out=out, inv_norm=inv_u)
    print(f"conv1 + conv2 contractions + mix_rownorm: {timeit(old):.2f} us")
    print(f"conv1 only: {timeit(lambda: K.gemm([hc.hi], [w1.hi], B, H, [d3], act=K.ACT_RELU, bias=b1, out_bf16=g1.hi)):.2f} us")
    print(f"conv2 only: {timeit(lambda: K.gemm([g1.hi], [w2.hi], B, d3, [H], bias=b2, out_f32=g2)):.2f} us")
    print(f"mix only: {timeit(lambda: K.mix_rownorm(hc_f32, B, d3, g=g2, sumw=sumw, out=out, inv_norm=inv_u)):.2f} us")


if __name__ == "__main__":
    main()
