"""CPU-side tests: the C-ABI library builds, loads and exports every symbol include/gdmcf_sm100.h declares (no compute
calls without a GPU), host logic of the mirror (CLI, data path, SpMM plan, Lt_history update, metrics finalisation),
and the reference arm of bench.py."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "gdmcf_sm100.h")).read()
    declared = set(re.findall(r"\b(gdmcf_[a-z0-9_]+)\s*\(", header))
    declared -= {"gdmcf_stream_t"}
    assert len(declared) >= 35
    from gdmcf_b200 import _lib
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == declared
    assert lib.gdmcf_abi_version() == 1


def test_library_is_sm100a_only():
    so = os.path.join(ROOT, "gdmcf_b200", "libgdmcf_sm100.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gdmcf_b200 import evaluate_utils, kernels
    from gdmcf_b200._lib import GdmcfError
    with pytest.raises(GdmcfError):
        kernels.row_inv_norm(torch.zeros(4, 4))
    with pytest.raises(RuntimeError):
        evaluate_utils.computeTopNAccuracy([[1]], [[1, 2]], [1])


def test_spmm_plan_host(lib):
    from gdmcf_b200 import kernels as K
    rowptr = np.array([0, 3, 3, 103, 110, 400], dtype=np.int32)
    plan = K.spmm_plan(rowptr, chunk=32, device="cpu")
    items = plan.items.numpy()[:plan.n_items]
    longs = plan.long_rows.numpy()[:plan.n_long]
    # every non-zero covered exactly once, chunks <= 32, long rows first
    cover = np.zeros(400, dtype=int)
    for r, b, e, slot in items:
        assert e - b <= 32 and rowptr[r] <= b <= e <= rowptr[r + 1]
        cover[b:e] += 1
        assert (slot >= 0) == (rowptr[r + 1] - rowptr[r] > 32)
    assert (cover == 1).all()
    assert plan.n_long == 2 and plan.n_slots == 4 + 10 and sorted(longs[:, 0]) == [2, 4]
    assert sum(1 for it in items if it[3] < 0) == 3  # rows 0, 1 (empty), 3
    with pytest.raises(AssertionError):
        K.spmm_plan(np.array([0, 5, 3], dtype=np.int32), chunk=32, device="cpu")


def test_cli_matches_reference_flags():
    from gdmcf_b200.parse_args_util import parse_args
    a = parse_args([])
    assert (a.lr, a.batch_size, a.topN, a.emb_size, a.steps, a.noise_scale, a.discrete, a.reweight) == \
        (0.0001, 400, '[10, 20, 50, 100]', 10, 100, 0.1, 0.9995, True)
    a = parse_args(["--dims=[1000]", "--steps=5", "--noise_scale=0.01", "--reweight", "False", "--lr=0.00001"])
    assert a.dims == [1000] and a.steps == 5 and a.reweight is False
    assert parse_args(["--dims", "200", "--dims", "600"]).dims == [200, 600]
    cfg = os.path.join(ROOT, "tests", "golden", "yelpOneEmbGcn.yaml")
    a = parse_args(["-c", cfg, "--batch_size", "400"])
    assert (a.backbone, a.OneHotMatrix, a.steps, a.dims, a.batch_size, a.sampling_steps) == \
        ("DNNOneHotEmbeddingGCN", 2, 5, [1000], 400, 0)


def test_synthetic_data_and_loader(tmp_path):
    from gdmcf_b200 import data_utils
    tr, va, te = data_utils.synthetic_interactions(500, 300, 9000, 0)
    for name, arr in (("train", tr), ("valid", va), ("test", te)):
        np.save(tmp_path / f"{name}_list.npy", arr)
    train, valid, test, n_user, n_item = data_utils.data_load(str(tmp_path / "train_list.npy"), str(tmp_path / "valid_list.npy"),
                                                              str(tmp_path / "test_list.npy"))
    assert (n_user, n_item) == (500, 300) and train.shape == (500, 300) and train.dtype == np.float64
    assert train.nnz == len(tr) and (train.multiply(test)).nnz == 0 and (np.diff(train.indptr) >= 1).all()
    assert 0.6 < len(tr) / (len(tr) + len(va) + len(te)) < 0.8
    ds = data_utils.DataDiffusion(train)
    row, idx = ds[7]
    assert idx == 7 and row.shape == (300,) and row.sum().item() == train[7].sum()
    dev = data_utils.DeviceInteractions(train, "cpu")
    b = dev.batch([3, 1, 4])
    assert b.shape == (3, 300) and b.users.dtype == torch.int32


def test_history_update_matches_reference_loop():
    from gdmcf_b200.models import gaussian_diffusion as gd
    from oracle import gdmcf_oracle as O
    torch.manual_seed(0)
    d = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", 0.01, 0.001, 0.01, 5, "cpu", discrete=0.9995, CatOneHot=True)
    o = O.OracleDiffusion(steps=5)
    for it in range(25):
        Bn = 7 if it % 3 else 64
        ts = torch.randint(0, 5, (Bn,))
        loss = torch.rand(Bn, dtype=torch.float64)
        d._update_history(ts, loss)
        o.update_history(ts, loss)
        assert torch.equal(d.Lt_history, o.Lt_history) and torch.equal(d.Lt_count, o.Lt_count)
        assert torch.allclose(d._pt_for(ts).double(), o.pt_for(ts).double())
    for k in ("betas", "alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped"):
        assert torch.equal(getattr(d, k), getattr(o.sch, k)), k


def test_engine_modules_match_oracle_state_dict():
    from gdmcf_b200.models.DNN import DNN, DNNOneHotEmbeddingGCN
    from oracle import gdmcf_oracle as O
    e = DNNOneHotEmbeddingGCN([150, 32], [32, 150], 10, item_num=150, user_num=40)
    o = O.OracleGDMCF([150, 32], [32, 150], 10, item_num=150, user_num=40)
    assert {k: tuple(v.shape) for k, v in e.state_dict().items()} == {k: tuple(v.shape) for k, v in o.state_dict().items()}
    assert sum(p.numel() for p in e.parameters()) == sum(p.numel() for p in o.parameters())
    e2, o2 = DNN([150, 32], [32, 150], 10), O.OracleDNN([150, 32], [32, 150], 10)
    assert {k: tuple(v.shape) for k, v in e2.state_dict().items()} == {k: tuple(v.shape) for k, v in o2.state_dict().items()}
    # dims with more than one entry (main.py:198-206): same modules, keys and shapes as the reference's construction
    for dims in ([150, 48, 32], [150, 48, 40, 32]):
        e3 = DNNOneHotEmbeddingGCN(list(dims), list(dims[::-1]), 10, item_num=150, user_num=40)
        o3 = O.OracleGDMCF(list(dims), list(dims[::-1]), 10, item_num=150, user_num=40)
        assert {k: tuple(v.shape) for k, v in e3.state_dict().items()} == {k: tuple(v.shape) for k, v in o3.state_dict().items()}
        assert e3.deep and e3.d1 == dims[1] and e3.hidden == dims[-1]
        e4, o4 = DNN(list(dims), list(dims[::-1]), 10), O.OracleDNN(list(dims), list(dims[::-1]), 10)
        assert {k: tuple(v.shape) for k, v in e4.state_dict().items()} == {k: tuple(v.shape) for k, v in o4.state_dict().items()}
        assert [n for n, _ in e4._middle()] == [f"in_layers.{i}" for i in range(1, len(dims) - 1)] + \
            [f"out_layers.{j}" for j in range(len(dims) - 2)]
    # parameter counts at the Yelp shape with n_user=3000
    n = lambda i, d, u: (i + 10) * d + d + (2 * i + 10) * d + d + (2 * d) * i + i + i * 3 * d + u * d + 110 + 3 * d * 512 + 512 + 512 * 3 * d + 3 * d + 1  # noqa: E731
    assert n(34395, 1000, 3000) == 281292018


def test_bench_reference_arm_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--dims", "64",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "users/s"


def test_operand_cache_refreshes_in_place_and_adopts():
    """_OperandCache (models/DNN.py): stale values are handed back to the builder (in-place refresh keeps addresses stable
    for captured CUDA graphs); mark_fresh adopts a value refreshed by someone else (FusedAdamW's fused refresh)."""
    import torch
    from gdmcf_b200.models.DNN import _OperandCache
    cache = _OperandCache()
    p = torch.nn.Parameter(torch.zeros(3))
    built = []

    def build(prev):
        built.append(prev)
        out = prev if prev is not None else torch.empty(3)
        out.copy_(p.detach() * 2)
        return out

    a = cache.get("x", [p], build)
    assert built == [None] and cache.get("x", [p], build) is a and len(built) == 1  # hit
    with torch.no_grad():
        p.add_(1.0)                      # torch bumps the version counter
    b = cache.get("x", [p], build)
    assert b is a and built[-1] is a and torch.equal(a, torch.full((3,), 2.0))       # rebuilt in place
    cache.epoch += 1                     # raw-pointer update (FusedAdamW): stale ...
    assert cache.peek("x") is a
    cache.mark_fresh("x", [p])           # ... unless the optimizer refreshed it itself
    assert cache.get("x", [p], build) is a and len(built) == 2
    cache.epoch += 1
    cache.get("x", [p], build)
    assert len(built) == 3


def test_cli_engine_flags():
    from gdmcf_b200.parse_args_util import parse_args
    a = parse_args(["--dims=[64]", "--checkpoint_every", "2", "--resume", "ck.pt", "--synthetic", "600,500,15000"])
    assert a.dims == [64] and a.checkpoint_every == 2 and a.resume == "ck.pt" and a.synthetic == "600,500,15000"
    b = parse_args([])
    assert b.checkpoint_every == 0 and b.resume == "" and b.dims == [1000]
