"""Stand-alone GEMM cases of the hot path for event timing and ncu captures.
usage: python tools/gemm_case.py <case> [iters]     cases: scorer_post | scorer_plain | dE | dE_plain | dW1 | enc"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import kernels as K  # noqa: E402


def op(rows, cols, seed, scale=0.05):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.zeros(rows, K.round_up(cols, 64), dtype=torch.bfloat16, device="cuda")
    t[:, :cols] = (torch.randn(rows, cols, generator=g, device="cuda") * scale).to(torch.bfloat16)
    return t


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "scorer_post"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    B, I, d = 400, 34395, 1000
    kw = {}
    if case.startswith("scorer"):
        m, n, k = B, I, 3 * d
        a, b = op(m, k, 1), op(n, k, 2)
        out = torch.empty(m, K.round_up(n, 4), device="cuda")
        kw = dict(out_f32=out, row_scale=torch.rand(m, device="cuda"), col_scale=torch.rand(n, device="cuda"))
        if case == "scorer_post":
            ob = K.Bf16Mat.empty(m, n, "cuda")
            kw.update(out_bf16=ob.hi, c1=torch.rand(5, device="cuda"), c2=torch.rand(5, device="cuda"),
                      xt=torch.randn(m, K.round_up(n, 4), device="cuda"), t_const=2)
    elif case == "dE":
        m, n, k = I, 3 * d, B
        a, b = op(m, k, 1), op(n, k, 2)
        out = torch.empty(m, n, device="cuda")
        kw = dict(out_f32=out, row_t=torch.arange(m, dtype=torch.int32, device="cuda"), c1=torch.ones(m, device="cuda"),
                  c2=torch.rand(m, device="cuda"), xt=torch.randn(m, n, device="cuda"))
    elif case == "dE_plain":  # the engine's form: plain fp32 store (the row-norm term is applied by the optimizer pass)
        m, n, k = I, 3 * d, B
        a, b = op(m, k, 1), op(n, k, 2)
        kw = dict(out_f32=torch.empty(m, n, device="cuda"))
    elif case == "dW1":
        m, n, k = d, I, B
        a, b = op(m, k, 1), op(n, k, 2)
        out = torch.empty(m, n + 10, device="cuda")
        kw = dict(out_f32=out)
    elif case == "enc":
        m, n, k = B, d, I
        a, b = op(m, k, 1), op(n, k, 2)
        ob = K.Bf16Mat.empty(m, n, "cuda")
        kw = dict(out_f32=torch.empty(m, n, device="cuda"), out_bf16=ob.hi, act=K.ACT_TANH, bias=torch.randn(5, n, device="cuda"),
                  ld_bias=n, row_t=torch.randint(0, 5, (m,), dtype=torch.int32, device="cuda"))
    else:
        raise SystemExit("unknown case")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.gemm([a], [b], m, n, [k], mode=mode, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    flops = 2.0 * m * n * k
    best = min(ts[1:]) if len(ts) > 1 else ts[0]
    print(f"{case} mode={mode}: m={m} n={n} k={k} ms={['%.3f' % t for t in ts]} best {flops / best / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
