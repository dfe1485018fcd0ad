"""Factor exchange vs reduce-scatter of the item table's gradient (engine.StepEngine, world_size > 1): one step each from the
same weights; compares this rank's block of the summed gradient and the updated table.
usage: torchrun --nproc-per-node 2 tools/fx_check.py fg,fe,re   (f/r = factor exchange / reduce-scatter, g/e = graphs / eager;
every engine is compared with the last one)"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import data_utils, dist_utils  # noqa: E402
from gdmcf_b200.engine import StepEngine  # noqa: E402
from gdmcf_b200.models import gaussian_diffusion as gd  # noqa: E402
from gdmcf_b200.models.DNN import DNNOneHotEmbeddingGCN  # noqa: E402
from gdmcf_b200.optim import FusedAdamW  # noqa: E402


def main():
    kinds = (sys.argv[1] if len(sys.argv) > 1 else "fe,re").split(",")
    dist = dist_utils.init("nccl")
    G, rank = dist.world_size, dist.rank
    dev = torch.device("cuda", dist.local_rank)
    torch.cuda.set_device(dev)
    U, I, D, B, T, k = 900, 1203, 64, 64, 5, 20
    tr, va, te = data_utils.synthetic_interactions(U, I, 27000, 5)
    n_user, n_item = int(tr[:, 0].max()) + 1, int(tr[:, 1].max()) + 1
    mk = lambda p: sp.csr_matrix((np.ones(len(p), dtype=np.float32), (p[:, 0], p[:, 1])), shape=(n_user, n_item))  # noqa: E731
    train_dev, test_dev = data_utils.DeviceInteractions(mk(tr), dev), data_utils.DeviceInteractions(mk(te), dev)

    def make(fx, graphs):
        torch.manual_seed(0)
        model = DNNOneHotEmbeddingGCN([n_item, D], [D, n_item], 10, item_num=n_item, user_num=n_user).to(dev)
        diff = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", 0.01, 0.001, 0.01, T, dev, discrete=0.9995,
                                            CatOneHot=True)
        diff.indexIn = True
        diff.seed = model.seed = 77 + rank
        opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0, modules=[model], capturable=True)
        eng = StepEngine(model, diff, opt, dist, batch_size=B, n_item=n_item, topk=k, topN=[10, k], cap_train_nnz=27000,
                         cap_gt_nnz=27000, graphs=graphs, nccl_sms=32, shard_min_bytes=1 << 16, factor_exchange=fx)
        return model, eng

    engines = [make(kd[0] == "f", kd[1] == "g") for kd in kinds]
    name = "embedding_item.weight"
    for _, e in engines:
        e.load_resident(train_dev, test_dev, rank * B, (rank + 1) * B)
        e.capture(warmup=2, preserve_state=os.environ.get("FX_PRESERVE", "1") == "1")
    for s in range(3):
        lo = ((s * G + rank) * B) % (n_user - B)
        for _, e in engines:
            e.load_resident(train_dev, test_dev, lo, lo + B)
            e.step()
        torch.cuda.synchronize()
        m0, e0 = engines[-1]
        for kd, (m1, e1) in zip(kinds[:-1], engines[:-1]):
            sh1, sh0 = e1._shards[name], e0._shards[name]
            R = sh1["R"]
            g1 = sh1["gbuf"][rank * R:(rank + 1) * R, :3 * D]
            g0 = sh0["gbuf"][rank * R:(rank + 1) * R, :3 * D]
            d = (g1 - g0).abs()
            print(f"rank {rank} step {s} {kd} vs {kinds[-1]}: gradient block max |diff| {d.max().item():.3e} (max |g| {g0.abs().max().item():.3e}), "
                  f"rows differing {(d.max(1).values > 1e-6 * g0.abs().max()).sum().item()} of {R}", flush=True)
            for (n, pa), (_, pb) in zip(m1.named_parameters(), m0.named_parameters()):
                dd = (pa - pb).abs().max().item()
                if dd > 1e-6:
                    print(f"   rank {rank} step {s} {kd}: {n} differs by {dd:.3e}", flush=True)
    dist.barrier()
    dist.shutdown()


if __name__ == "__main__":
    main()
