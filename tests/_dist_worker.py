"""Worker for tests/test_dist_cpu.py (world_size 2, gloo, CPU): exercises gdmcf_b200.dist_utils with the CPU oracle
standing in for the denoiser, writes its results to <out>/rank<r>.pt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import dist_utils  # noqa: E402
from oracle import gdmcf_oracle as O  # noqa: E402


def main(out_dir):
    dist = dist_utils.init("gloo")
    assert dist.world_size == 2
    torch.manual_seed(100 + dist.rank)  # different init per rank on purpose: broadcast must fix it
    I, D, U, B, T = 120, 16, 50, 8, 5
    model = O.OracleGDMCF([I, D], [D, I], 10, item_num=I, user_num=U)
    big = torch.nn.Parameter(torch.randn(2_200_000))  # > 8 MB: exercises the in-place large-tensor path
    model.register_parameter("big_extra", big)
    dist.broadcast_parameters(model)
    # batches dealt round-robin: rank r takes logical batches r, r+G, ...
    n_batches = 7
    mine = list(range(dist.rank, n_batches - n_batches % dist.world_size, dist.world_size))
    g = torch.Generator().manual_seed(7)  # same data on both ranks
    dense = (torch.rand(U, I, generator=g) < 0.1).float()
    users = torch.randperm(U, generator=g)[: 2 * B].reshape(2, B)
    draws = []
    for b in range(2):
        draws.append(dict(ts1=torch.randint(0, T, (B,), generator=g), ts=torch.randint(0, T, (B,), generator=g),
                          noise=torch.randn(B, I, generator=g), u=torch.rand(B, I, generator=g),
                          kx=torch.rand(B, I, generator=g) >= 0.5, kxu=torch.rand(B, 2 * I, generator=g) >= 0.5))
    diff = O.OracleDiffusion(steps=T)
    d = draws[dist.rank]
    idx = users[dist.rank]
    model.train()
    terms = diff.training_losses(model, dense[idx], idx, d["ts1"], d["ts"], d["noise"], d["u"], d["kx"], d["kxu"])
    loss = terms["loss"].mean() + 1e-3 * (model.big_extra ** 2).sum() * (dist.rank + 1)
    loss.backward()
    dist.all_reduce_gradients(model)
    sums = torch.full((2, 4), float(dist.rank + 1), dtype=torch.float64)
    dist.all_reduce(sums)
    # asynchronous form used by engine.StepEngine: large, flattened small and padded-view tensors
    def make(rank):
        gen = torch.Generator().manual_seed(11 + rank)
        pad = torch.randn(6, 12, generator=gen)
        return [torch.randn(2_100_000, generator=gen), torch.randn(7, generator=gen), torch.randn(3, 5, generator=gen), pad[:, :10]]
    own, other = make(dist.rank), make(1 - dist.rank)
    expect = [a + b for a, b in zip(own, other)]
    works, after = dist.all_reduce_async(own)
    for w in works:
        w.wait()
    for fn in after:
        fn()
    async_ok = all(torch.equal(a, b) for a, b in zip(own, expect))
    # factor exchange of a rank-B gradient (engine.StepEngine, item table): block q of the summed product Gs^T hc' from the
    # exchanged factors == rows of the all-reduced full product
    G, R, Bf, C = dist.world_size, 37, 8, 24
    gen = torch.Generator().manual_seed(500 + dist.rank)
    gsT, hcT = torch.randn(G * R, Bf, generator=gen), torch.randn(C, Bf, generator=gen)
    recv_rows, recv_small = torch.zeros(G, R, Bf), torch.zeros(G, C, Bf)
    for w in dist.exchange_factors(gsT, hcT, recv_rows, recv_small):
        w.wait()
    block = sum(recv_rows[r] @ recv_small[r].t() for r in range(G))
    full = gsT @ hcT.t()
    dist.all_reduce(full)
    fx_err = (block - full[dist.rank * R:(dist.rank + 1) * R]).abs().max().item()
    torch.save({"rank": dist.rank, "mine": mine, "sums": sums, "async_ok": async_ok, "fx_err": fx_err,
                "params": {k: v.detach().clone() for k, v in model.named_parameters()},
                "grads": {k: (v.grad.clone() if v.grad is not None else None) for k, v in model.named_parameters()},
                "draws": draws, "users": users, "dense": dense},
               os.path.join(out_dir, f"rank{dist.rank}.pt"))
    dist.barrier()
    dist.shutdown()


if __name__ == "__main__":
    main(sys.argv[1])
