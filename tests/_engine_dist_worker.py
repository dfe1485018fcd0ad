"""Data-parallel check of engine.StepEngine under torchrun (N >= 2, NCCL): the graph-segment engine with overlapped
all-reduces must equal the same program run eagerly, ranks must stay bit-identical, and the sparse user-row exchange
+ row-sparse AdamW with exact catch-up must equal a dense all-reduce of the user table's gradient + dense AdamW, and the reduce-scatter + sharded AdamW + all-gather path
must equal the all-reduce + replicated AdamW path.
usage: torchrun --nproc-per-node N tests/_engine_dist_worker.py  (driven by tests/test_engine_gpu.py::test_engine_data_parallel_torchrun)"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import data_utils, dist_utils  # noqa: E402
from gdmcf_b200.engine import StepEngine  # noqa: E402
from gdmcf_b200.models import gaussian_diffusion as gd  # noqa: E402
from gdmcf_b200.models.DNN import DNNOneHotEmbeddingGCN  # noqa: E402
from gdmcf_b200.optim import FusedAdamW  # noqa: E402


def main():
    dist = dist_utils.init("nccl")
    G, rank = dist.world_size, dist.rank
    dev = torch.device("cuda", dist.local_rank)
    torch.cuda.set_device(dev)
    U, I, D, B, T, k = 900, 1203, 64, 64, 5, 20
    tr, va, te = data_utils.synthetic_interactions(U, I, 27000, 5)
    n_user, n_item = int(tr[:, 0].max()) + 1, int(tr[:, 1].max()) + 1
    mk = lambda p: sp.csr_matrix((np.ones(len(p), dtype=np.float32), (p[:, 0], p[:, 1])), shape=(n_user, n_item))  # noqa: E731
    train_sp, test_sp = mk(tr), mk(te)
    train_dev, test_dev = data_utils.DeviceInteractions(train_sp, dev), data_utils.DeviceInteractions(test_sp, dev)

    def make(graphs, sparse=True, shard=True):
        torch.manual_seed(0)
        model = DNNOneHotEmbeddingGCN([n_item, D], [D, n_item], 10, item_num=n_item, user_num=n_user).to(dev)
        diff = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", 0.01, 0.001, 0.01, T, dev,
                                            discrete=0.9995, CatOneHot=True)
        diff.indexIn = True
        diff.seed = model.seed = 77 + rank
        opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0, modules=[model], capturable=True)
        eng = StepEngine(model, diff, opt, dist, batch_size=B, n_item=n_item, topk=k, topN=[10, k],
                         cap_train_nnz=int(train_sp.nnz), cap_gt_nnz=int(test_sp.nnz), graphs=graphs, nccl_sms=32,
                         shard_optimizer=shard, shard_min_bytes=1 << 16, lazy_user_rows=sparse)
        eng.sparse_user_rows = eng.sparse_user_rows and sparse
        return model, diff, eng

    # graph segments + sharded optimizer | same program eagerly | eager, dense user-table all-reduce, replicated optimizer
    engines = [make(True), make(False), make(False, sparse=False, shard=False)]
    for _, _, e in engines:
        e.load_resident(train_dev, test_dev, rank * B, (rank + 1) * B)
        e.capture(warmup=2)
    n_seg = len(engines[0][2]._segments)
    for s in range(1, 5):
        lo = ((s * G + rank) * B) % (n_user - B)
        outs = []
        for _, _, e in engines:
            e.load_resident(train_dev, test_dev, lo, lo + B)
            loss, idx, sums = e.step()
            outs.append((loss.clone(), idx.clone(), sums.clone()))
        for j, o in enumerate(outs[1:], 1):
            same = torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) and torch.equal(outs[0][2], o[2])
            if not same:
                print(f"rank {rank} step {s}: engine 0 vs {j}: loss {outs[0][0].item():.9f} vs {o[0].item():.9f}, "
                      f"top-k index agreement {(outs[0][1] == o[1]).float().mean().item():.4f}", flush=True)
    for _, _, e in engines:
        e.flush()  # row-sparse user-table updates: replay the pending zero-gradient steps before comparing weights
    torch.cuda.synchronize()
    ok = True
    for (n, pg), (_, pe), (_, pd) in zip(*[m.named_parameters() for m, _, _ in engines]):
        if not torch.equal(pg, pe):
            ok = False
            print(f"rank {rank}: graph != eager for {n}: {(pg - pe).abs().max().item():.3e}", flush=True)
        if not torch.allclose(pg, pd, rtol=0, atol=1e-6):
            ok = False
            print(f"rank {rank}: sharded/sparse != replicated/dense exchange for {n}: {(pg - pd).abs().max().item():.3e}", flush=True)
        # ranks identical
        ref = pg.detach().clone()
        td.broadcast(ref, src=0)
        if not torch.equal(ref, pg):
            ok = False
            print(f"rank {rank}: parameter {n} differs from rank 0", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    if rank == 0:
        print(f"engine_dist_check: world {G}, {n_seg} graph segments per step, "
              f"{'OK' if flag.item() == 1.0 else 'FAILED'}", flush=True)
    dist.barrier()
    dist.shutdown()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
