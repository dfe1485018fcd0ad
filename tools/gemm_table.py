"""Per-contraction table of one train + rank step at the Yelp shape: shape, split-K factor, epilogue, time (CUDA events around
each call, synchronised, eager mode, L2 flushed before every call — so each number is that launch alone with cold inputs), tensor
throughput and the time its compulsory HBM bytes would take. Tells which contraction is far from which roofline.
usage: python tools/gemm_table.py [yelp|amazon]"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import data_utils, dist_utils  # noqa: E402
from gdmcf_b200 import kernels as K  # noqa: E402
from gdmcf_b200.engine import StepEngine  # noqa: E402
from gdmcf_b200.models import gaussian_diffusion as gd  # noqa: E402
from gdmcf_b200.models.DNN import DNNOneHotEmbeddingGCN  # noqa: E402
from gdmcf_b200.optim import FusedAdamW  # noqa: E402

SHAPES = {"yelp": (54574, 34395, 1402736, 0), "amazon": (108822, 94949, 3146256, 1)}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "yelp"
    U, I, P, seed = SHAPES[wl]
    dev = torch.device("cuda")
    tr, va, te = data_utils.synthetic_interactions(U, I, P, seed)
    n_user, n_item = int(tr[:, 0].max()) + 1, int(tr[:, 1].max()) + 1
    mk = lambda p: sp.csr_matrix((np.ones(len(p), dtype=np.float32), (p[:, 0], p[:, 1])), shape=(n_user, n_item))  # noqa: E731
    train_dev, test_dev = data_utils.DeviceInteractions(mk(tr), dev), data_utils.DeviceInteractions(mk(te), dev)
    torch.manual_seed(0)
    with torch.device(dev):
        model = DNNOneHotEmbeddingGCN([n_item, 1000], [1000, n_item], 10, item_num=n_item, user_num=n_user)
    diffusion = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", 0.01, 0.001, 0.01, 5, dev, discrete=0.9995,
                                             CatOneHot=True)
    diffusion.indexIn = True
    opt = FusedAdamW(model.parameters(), lr=1e-5, weight_decay=0.0, modules=[model], capturable=True)
    B = 400
    eng = StepEngine(model, diffusion, opt, dist_utils.Dist(), batch_size=B, n_item=n_item, topk=20, topN=[10, 20], cap_train_nnz=1,
                     cap_gt_nnz=1, graphs=False)
    eng.bind_resident(train_dev, gt_dev=test_dev)
    users = lambda b: torch.arange(b * B, (b + 1) * B, dtype=torch.int32, device=dev)  # noqa: E731
    for s in range(3):
        eng.load_users(users(s))
        eng.step()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    real = K.gemm

    def timed(a, b, m, n, k, **kw):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        real(a, b, m, n, k, **kw)
        e1.record()
        torch.cuda.synchronize()
        kk = int(sum(K.round_up(x, 64) for x in k))
        splits = kw.get("splits") or K.load().gdmcf_gemm_auto_splits(m, n, kk)
        out_b = sum(m * n * t.element_size() for t in (kw.get("out_f32"), kw.get("out_bf16"), kw.get("out_bf16_lo")) if t is not None)
        if kw.get("mode", K.EPI_STORE) != K.EPI_STORE and kw.get("xt") is not None:
            out_b += m * n * 4
        in_b = 2 * kk * (m + n)
        import traceback
        fr = [f for f in traceback.extract_stack()[:-1] if "gdmcf_b200" in f.filename][-1]
        rows.append((f"{os.path.basename(fr.filename)}:{fr.lineno}", m, n, kk, len(k), splits, kw.get("mode", 0), e0.elapsed_time(e1) * 1e3,
                     2.0 * m * n * kk, in_b + out_b))

    K.gemm = timed
    eng.load_users(users(5))
    eng.step()
    torch.cuda.synchronize()
    K.gemm = real
    tot = 0.0
    print(f"{'call site':28s} {'M':>6s} {'N':>6s} {'K':>6s} seg spl epi {'us':>7s} {'TF/s':>7s} {'hbm us':>7s} {'mma us':>7s}")
    for site, m, n, kk, seg, spl, epi, us, fl, by in rows:
        tot += us
        print(f"{site:28s} {m:6d} {n:6d} {kk:6d} {seg:3d} {spl:3d} {epi:3d} {us:7.1f} {fl / us / 1e6:7.0f} {by / 6515e3:7.1f} {fl / 1500e6:7.1f}")
    print(f"{len(rows)} contractions, {tot:.0f} us in total (each cold, serialised)")


if __name__ == "__main__":
    main()
