"""Stand-alone LightGCN propagation (K layers, d=64, fp32) on a synthetic interaction graph, for event timing
and ncu captures. usage: python tools/spmm_case.py [workload=yelp] [iters=10] [layers=3]"""
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
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    layers = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    U, I, P, seed = SHAPES[wl]
    tr, _, _ = data_utils.synthetic_interactions(U, I, P, seed)
    n_user, n_item = int(tr[:, 0].max()) + 1, int(tr[:, 1].max()) + 1
    lg = LightGCN({"user_id_idx": tr[:, 0], "item_id_idx": tr[:, 1]}, n_user, n_item, layers, 64, device="cuda")
    _, col, val = lg.norm_adj_csr
    E0 = lg.E0.weight.detach()
    out = torch.empty_like(E0)
    sym = os.environ.get("SPMM_SYM", "1") != "0"
    work = K.lightgcn_sym_work(lg.plan, E0) if sym else (torch.empty_like(E0), torch.empty_like(E0),
                                                         torch.empty(max(lg.plan.n_slots, 1), 64, device="cuda"))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.lightgcn_propagate(lg.plan, col, val, E0, layers, out=out, work=work, dinv=lg.dinv if sym else None)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    N, nnz = n_user + n_item, col.numel()
    bytes_alg = layers * (nnz * 8 + (N + 1) * 4 + 2 * N * 64 * 4)
    t = sorted(ts[2:])[len(ts[2:]) // 2] if len(ts) > 2 else ts[-1]
    print(f"spmm {wl} ({'separable' if sym else 'value'} form): N={N} nnz={nnz} items={lg.plan.n_items} long={lg.plan.n_long} ms={['%.3f' % x for x in ts]} "
          f"median {t:.3f} ms -> {bytes_alg / t / 1e6:.1f} GB/s algorithmic ({bytes_alg / 1e6:.1f} MB)")


if __name__ == "__main__":
    main()
