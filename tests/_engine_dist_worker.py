"""Data-parallel check of engine.StepEngine under torchrun (N >= 2, NCCL): the graph-segment engine with overlapped
collectives must equal the same program run eagerly (bit for bit), ranks must stay bit-identical; the sparse user-row
exchange + row-sparse AdamW with exact catch-up must equal a dense all-reduce of the user table's gradient + dense AdamW;
reduce-scatter + sharded AdamW + all-gather must equal all-reduce + replicated AdamW; and the factor exchange of the item
table's gradient (all-to-all of Gs^T row blocks + all-gather of hc'^T, one contraction with K over the ranks) must give
the reduce-scattered gradient block to fp32 rounding.
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

    def make(graphs, sparse=True, shard=True, fx=True):
        torch.manual_seed(0)
        model = DNNOneHotEmbeddingGCN([n_item, D], [D, n_item], 10, item_num=n_item, user_num=n_user).to(dev)
        diff = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", 0.01, 0.001, 0.01, T, dev,
                                            discrete=0.9995, CatOneHot=True)
        diff.indexIn = True
        diff.seed = model.seed = 77 + rank
        opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0, modules=[model], capturable=True)
        eng = StepEngine(model, diff, opt, dist, batch_size=B, n_item=n_item, topk=k, topN=[10, k],
                         cap_train_nnz=int(train_sp.nnz), cap_gt_nnz=int(test_sp.nnz), graphs=graphs, nccl_sms=32,
                         shard_optimizer=shard, shard_min_bytes=1 << 16, lazy_user_rows=sparse, factor_exchange=fx,
                         bf16_gather=fx)
        eng.sparse_user_rows = eng.sparse_user_rows and sparse
        return model, diff, eng

    # 0: graph segments + sharded optimizer + factor exchange of the item table's gradient + its rows all-gathered as the
    #    bf16 operand | 1: the same program eagerly | 2: eager, item table reduce-scattered / fp32 all-gather like the other
    #    matrices | 3: eager, dense user-table all-reduce, replicated optimizer
    engines = [make(True), make(False), make(False, fx=False), make(False, sparse=False, shard=False)]
    for _, _, e in engines:
        e.load_resident(train_dev, test_dev, rank * B, (rank + 1) * B)
        e.capture(warmup=2, preserve_state=True)
    assert "fx" in engines[0][2]._shards["embedding_item.weight"] and "fx" not in engines[2][2]._shards["embedding_item.weight"]
    assert engines[0][2]._bf16_gather_active and not engines[2][2]._bf16_gather_active
    n_seg = len(engines[0][2]._segments)
    ok = True
    for s in range(1, 5):
        lo = ((s * G + rank) * B) % (n_user - B)
        outs = []
        for _, _, e in engines:
            e.load_resident(train_dev, test_dev, lo, lo + B)
            loss, idx, sums = e.step()
            outs.append((loss.clone(), idx.clone(), sums.clone()))
        if not (torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])):
            ok = False
            print(f"rank {rank} step {s}: graph vs eager: loss {outs[0][0].item():.9f} vs {outs[1][0].item():.9f}, "
                  f"top-k index agreement {(outs[0][1] == outs[1][1]).float().mean().item():.4f}", flush=True)
        if s == 1:
            # all engines started this step from the same weights: this rank's block of the summed item-table gradient, formed
            # from the exchanged factors (one contraction, K over the ranks) vs reduce-scattered (per-rank products summed by
            # NCCL) — the same sum in another fp32 association
            torch.cuda.synchronize()
            sh0, sh2 = engines[0][2]._shards["embedding_item.weight"], engines[2][2]._shards["embedding_item.weight"]
            R, c3 = sh0["R"], sh0["p"].shape[1]
            g0, g2 = sh0["gbuf"][rank * R:(rank + 1) * R, :c3], sh2["gbuf"][rank * R:(rank + 1) * R, :c3]
            rel = ((g0 - g2).abs().max() / g2.abs().max()).item()
            print(f"rank {rank}: item-table gradient block, factor exchange vs reduce-scatter: max rel diff {rel:.2e}", flush=True)
            if not rel < 2e-6:
                ok = False
    for _, _, e in engines:
        e.flush()  # row-sparse user-table updates: replay the pending zero-gradient steps before comparing weights
    torch.cuda.synchronize()
    lr = 1e-3
    for (n, pg), (_, pe), (_, pr), (_, pd) in zip(*[m.named_parameters() for m, _, _ in engines]):
        if not torch.equal(pg, pe):
            ok = False
            print(f"rank {rank}: graph != eager for {n}: {(pg - pe).abs().max().item():.3e}", flush=True)
        if not torch.allclose(pr, pd, rtol=0, atol=1e-6):
            ok = False
            print(f"rank {rank}: sharded/sparse != replicated/dense exchange for {n}: {(pr - pd).abs().max().item():.3e}", flush=True)
        # factor exchange vs reduce-scatter: gradients equal to fp32 rounding (checked above); AdamW's first steps (second
        # moment ~ g^2 of very few steps) amplify the rounding, so after 4 steps the weights agree to a fraction of one step
        if not torch.allclose(pg, pr, rtol=0, atol=0.2 * lr):
            ok = False
            print(f"rank {rank}: factor exchange != reduce-scatter for {n}: {(pg - pr).abs().max().item():.3e}", flush=True)
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
