"""World-size-2 gloo test of the data-parallel plumbing (gdmcf_b200/dist_utils.py): parameter broadcast, round-robin
batch dealing, gradient all-reduce (large in-place + flattened small tensors), metric all-reduce, and the equivalence
`G-rank step == gradients averaged over G consecutive batches` against a single-process evaluation of the oracle."""
import os
import socket
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_equivalence(tmp_path):
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), str(tmp_path)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-3000:]
    r0 = torch.load(tmp_path / "rank0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "rank1.pt", weights_only=False)
    # broadcast: identical parameters; dealing: disjoint cover of the even number of batches; metric sum
    for k in r0["params"]:
        assert torch.equal(r0["params"][k], r1["params"][k]), k
    assert r0["mine"] == [0, 2, 4] and r1["mine"] == [1, 3, 5]
    assert r0["async_ok"] and r1["async_ok"]  # all_reduce_async: large / flattened small / padded-view tensors
    assert r0["fx_err"] < 1e-4 and r1["fx_err"] < 1e-4  # factor exchange: block mapping of all-to-all + all-gather
    assert torch.equal(r0["sums"], torch.full((2, 4), 3.0, dtype=torch.float64)) and torch.equal(r0["sums"], r1["sums"])
    # all-reduced gradients identical on both ranks and equal to the single-process sum over the two batches
    from oracle import gdmcf_oracle as O
    I, D, U, T = 120, 16, 50, 5
    model = O.OracleGDMCF([I, D], [D, I], 10, item_num=I, user_num=U)
    model.register_parameter("big_extra", torch.nn.Parameter(torch.zeros(2_200_000)))
    model.load_state_dict({k: v for k, v in r0["params"].items()})
    model.train()
    total = {k: torch.zeros_like(v) for k, v in model.named_parameters()}
    for b in range(2):
        model.zero_grad()
        d = r0["draws"][b]
        idx = r0["users"][b]
        diff = O.OracleDiffusion(steps=T)
        terms = diff.training_losses(model, r0["dense"][idx], idx, d["ts1"], d["ts"], d["noise"], d["u"], d["kx"], d["kxu"])
        (terms["loss"].mean() + 1e-3 * (model.big_extra ** 2).sum() * (b + 1)).backward()
        for k, v in model.named_parameters():
            if v.grad is not None:
                total[k] += v.grad
    for k, v in model.named_parameters():
        g0, g1 = r0["grads"][k], r1["grads"][k]
        if g0 is None:
            assert g1 is None and total[k].abs().max() == 0, k  # out_layers: dead in the reference too
            continue
        assert torch.equal(g0, g1), k
        assert torch.allclose(g0, total[k], rtol=1e-5, atol=1e-7 * total[k].abs().max().item() + 1e-12), k
