"""The three heaviest contraction launches of the round-2 Yelp step as stand-alone launches for one ncu --set full capture:
  scorer   400 x 34395 x 3000   cosine epilogue (row / column scales), fp32 scores       (x2 per step)
  dE       34395 x 3000 x 400   plain fp32 store (item-table gradient)
  P        1000 x 3000 x 34395  split-K 3, bf16 output (projection operand of the reverse loop)
Two launches each (the second one is the one to read: operands warm in L2 like inside the step is NOT what ncu measures —
ncu flushes caches per replay — so both are cold).
usage: ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_bf16_tn_2cta -o out python tools/r2_ncu_cases.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gdmcf_b200 import kernels as K  # noqa: E402
from gemm_case import op  # noqa: E402


def main():
    B, I, d = 400, 34395, 1000
    cases = []
    a, b = op(B, 3 * d, 1), op(I, 3 * d, 2)
    cases.append(("scorer", a, b, B, I, 3 * d, dict(out_f32=torch.empty(B, K.round_up(I, 4), device="cuda"),
                                                    row_scale=torch.rand(B, device="cuda"), col_scale=torch.rand(I, device="cuda"))))
    a2, b2 = op(I, B, 3), op(3 * d, B, 4)
    cases.append(("dE", a2, b2, I, 3 * d, B, dict(out_f32=torch.empty(I, 3 * d, device="cuda"))))
    a3, b3 = op(d, I, 5), op(3 * d, I, 6)
    cases.append(("P", a3, b3, d, 3 * d, I, dict(out_bf16=K.Bf16Mat.empty(d, 3 * d, "cuda").hi)))
    for name, a_, b_, m, n, k, kw in cases:
        for it in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.gemm([a_], [b_], m, n, [k], **kw)
            e1.record()
            torch.cuda.synchronize()
        print(f"{name}: {m} x {n} x {k}: {e0.elapsed_time(e1) * 1e3:.1f} us  {2.0 * m * n * k / e0.elapsed_time(e1) / 1e9:.0f} TF/s")


if __name__ == "__main__":
    main()
