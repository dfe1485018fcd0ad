"""Stand-alone bf16-mode LightGCN propagation (spmm_bf16.cu) for event timing and ncu captures.
usage: python tools/spmm16_case.py [workload=yelp] [iters=6] [layers=3]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import data_utils, kernels as K  # noqa: E402
from gdmcf_b200.lightGCN import LightGCN  # noqa: E402

SHAPES = {"yelp": (54574, 34395, 1402736, 0), "amazon": (108822, 94949, 3146256, 1), "tiny": (2000, 1500, 40000, 3),
          "scaled": (1000000, 200000, 50000000, 2)}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "yelp"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    layers = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    U, I, P, seed = SHAPES[wl]
    tr, _, _ = data_utils.synthetic_interactions(U, I, P, seed)
    n_user, n_item = int(tr[:, 0].max()) + 1, int(tr[:, 1].max()) + 1
    lg = LightGCN({"user_id_idx": tr[:, 0], "item_id_idx": tr[:, 1]}, n_user, n_item, layers, 64, device="cuda", precision="bf16")
    E0 = lg.E0.weight.detach()
    out = torch.empty_like(E0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.lightgcn_propagate_bf16(lg.plan16, lg.dinv, E0, layers, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    dbg = lg.plan16.sync[33 + lg.plan16.n_long:].tolist()
    print("phase stamps (ns since entry; after phase 0, barrier, then per layer: hot staged, items done, barrier): CTA 0", dbg[:11],
          "last CTA", dbg[16:27])
    import numpy as _np
    per = _np.array(dbg[32:32 + 2 * 148]).reshape(-1, 2)
    for name, v in (("hub pieces done", per[:, 0]), ("rows done", per[:, 1])):
        print(f"layer 0 per CTA, {name}: min {v.min() / 1e3:.1f} median {_np.median(v) / 1e3:.1f} max {v.max() / 1e3:.1f} us")
    N, nnz = n_user + n_item, lg.plan16.col.numel()
    bytes_alg = layers * (nnz * 8 + (N + 1) * 4 + 2 * N * 64 * 4)
    t = sorted(ts[2:])[len(ts[2:]) // 2] if len(ts) > 2 else ts[-1]
    print(f"spmm bf16 {wl}: N={N} nnz={nnz} items={lg.plan16.n_items} long={lg.plan16.n_long} hot={lg.plan16.n_hot} "
          f"ms={['%.3f' % x for x in ts]} median {t:.3f} ms -> {bytes_alg / t / 1e6:.1f} GB/s on the fp32-interface bytes")


if __name__ == "__main__":
    main()
