"""StepEngine (CUDA-graph replay of train + denoise + rank) must be the eager public API, bit for bit: two identical
models are stepped over the same batches, one through graph replays, one kernel by kernel. Also the device-side
timestep sampler (gaussian_diffusion.py:959-986) against its closed form."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu


def _setup(seed_model=0, U=700, I=1203, D=64, B=64, T=5, k=20, dims=None):
    from gdmcf_b200 import data_utils, dist_utils
    from gdmcf_b200.engine import StepEngine
    from gdmcf_b200.models import gaussian_diffusion as gd
    from gdmcf_b200.models.DNN import DNNOneHotEmbeddingGCN
    from gdmcf_b200.optim import FusedAdamW
    dev = torch.device("cuda")
    tr, va, te = data_utils.synthetic_interactions(U, I, 21000, 5)
    n_user, n_item = int(tr[:, 0].max()) + 1, int(tr[:, 1].max()) + 1
    mk = lambda p: sp.csr_matrix((np.ones(len(p), dtype=np.float32), (p[:, 0], p[:, 1])), shape=(n_user, n_item))  # noqa: E731
    train_sp, test_sp = mk(tr), mk(te)
    train_dev, test_dev = data_utils.DeviceInteractions(train_sp, dev), data_utils.DeviceInteractions(test_sp, dev)

    def make(graphs, **engine_kw):
        torch.manual_seed(seed_model)
        in_dims = [n_item] + list(dims or [D])
        model = DNNOneHotEmbeddingGCN(in_dims, in_dims[::-1], 10, item_num=n_item, user_num=n_user).to(dev)
        diff = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", 0.01, 0.001, 0.01, T, dev,
                                            discrete=0.9995, CatOneHot=True)
        diff.indexIn = True
        diff.seed = model.seed = 77
        opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0, modules=[model], capturable=True)
        eng = StepEngine(model, diff, opt, dist_utils.Dist(), batch_size=B, n_item=n_item, topk=k, topN=[10, k],
                         cap_train_nnz=int(train_sp.nnz), cap_gt_nnz=int(test_sp.nnz), graphs=graphs, **engine_kw)
        return model, diff, eng

    return make, train_dev, test_dev, train_sp, test_sp, B, n_user


def test_graph_step_equals_eager_step_deep_encoders():
    """dims = [96, 64] (two-layer tanh encoders): the captured step (deep forward + backward, operands of the small layers
    re-derived inside the graph) == the eager step, bit for bit, and every deep parameter trains."""
    make, train_dev, test_dev, _, _, B, n_user = _setup(dims=[96, 64])
    m_g, d_g, e_g = make(True)
    m_e, d_e, e_e = make(False)
    assert m_g.deep
    before = {n: p.detach().clone() for n, p in m_g.named_parameters() if n.startswith(("in_layers.1", "in_layers2.1"))}
    assert len(before) == 4
    for e in (e_g, e_e):
        e.load_resident(train_dev, test_dev, 0, B)
        e.capture(warmup=2)
    for s in range(1, 5):
        lo = (s * B) % (n_user - B)
        outs = []
        for e in (e_g, e_e):
            e.load_resident(train_dev, test_dev, lo, lo + B)
            loss, idx, sums = e.step()
            outs.append((loss.clone(), idx.clone(), sums.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2]), s
    e_g.flush(); e_e.flush()
    for (n, pg), (_, pe) in zip(m_g.named_parameters(), m_e.named_parameters()):
        assert torch.equal(pg, pe), n
    for n, p0 in before.items():
        assert not torch.equal(p0, dict(m_g.named_parameters())[n].detach()), n


def test_graph_step_equals_eager_step():
    make, train_dev, test_dev, _, _, B, n_user = _setup()
    m_g, d_g, e_g = make(True)
    m_e, d_e, e_e = make(False)
    for e in (e_g, e_e):
        e.load_resident(train_dev, test_dev, 0, B)
        e.capture(warmup=3)
    assert e_g.launches_per_step > 50 and e_e.launches_per_step == 0
    for s in range(1, 6):
        lo = (s * B) % (n_user - B)
        outs = []
        for e in (e_g, e_e):
            e.load_resident(train_dev, test_dev, lo, lo + B)
            loss, idx, sums = e.step()
            outs.append((loss.clone(), idx.clone(), sums.clone()))
        (lg, ig, sg), (le, ie, se) = outs
        assert torch.equal(lg, le), (s, lg.item(), le.item())
        assert torch.equal(ig, ie)
        assert torch.equal(sg, se)
    for (n, pg), (_, pe) in zip(m_g.named_parameters(), m_e.named_parameters()):
        assert torch.equal(pg, pe), n
    assert torch.equal(d_g.Lt_history, d_e.Lt_history) and torch.equal(d_g.Lt_count, d_e.Lt_count)
    # the eager API stays coherent after replays: an eager rank on the graph-stepped model equals the eager model's
    from gdmcf_b200.models.gaussian_diffusion import CsrBatch
    users = torch.arange(5, 5 + B, dtype=torch.int32, device="cuda")
    batch = train_dev.batch(users)
    m_g.eval(); m_e.eval()
    assert isinstance(batch, CsrBatch)
    assert torch.equal(d_g.rank(m_g, batch, 20, hist=train_dev.csr), d_e.rank(m_e, batch, 20, hist=train_dev.csr))


def test_graph_step_host_inputs_match_resident():
    make, train_dev, test_dev, train_sp, test_sp, B, n_user = _setup(seed_model=1)
    _, _, e_a = make(True)
    _, _, e_b = make(True)
    for e in (e_a, e_b):
        e.load_resident(train_dev, test_dev, 0, B)
        e.capture(warmup=2)
    lo, hi = 130, 130 + B
    e_a.load_resident(train_dev, test_dev, lo, hi)
    hb = []
    for m in (train_sp, test_sp):
        rp = (m.indptr[lo:hi + 1] - m.indptr[lo]).astype(np.int32)
        cl = m.indices[m.indptr[lo]:m.indptr[hi]].astype(np.int32)
        hb.append((torch.from_numpy(rp).pin_memory(), torch.from_numpy(cl).pin_memory()))
    e_b.load_host(torch.arange(lo, hi, dtype=torch.int32).pin_memory(), hb[0], hb[1])
    la, ia, sa = e_a.step()
    lb, ib, sb = e_b.step()
    assert torch.equal(la, lb) and torch.equal(ia, ib) and torch.equal(sa, sb)


def test_sample_timesteps_kernel():
    from gdmcf_b200 import kernels as K
    T, H, B = 7, 10, 200000
    g = torch.Generator().manual_seed(3)
    hist = (torch.rand(T, H, generator=g, dtype=torch.float64) * 3).cuda()
    full = torch.full((T,), H, dtype=torch.int64, device="cuda")
    epoch = torch.tensor([5], dtype=torch.int64, device="cuda")
    ts, pt = K.sample_timesteps(hist, full, B, uniform_prob=0.001, seed=11, offset=1 << 40, epoch=epoch)
    p = torch.sqrt((hist ** 2).mean(-1))
    p = p / p.sum() * (1 - 0.001) + 0.001 / T
    assert ts.min() >= 0 and ts.max() < T
    torch.testing.assert_close(pt, p[ts] * T, rtol=1e-12, atol=0)
    freq = torch.bincount(ts, minlength=T).double() / B
    # binomial 5-sigma band per timestep
    sigma = torch.sqrt(p * (1 - p) / B)
    assert ((freq - p).abs() < 5 * sigma).all(), (freq, p)
    # fresh draws per epoch, same draws for the same epoch
    ts2, _ = K.sample_timesteps(hist, full, B, seed=11, offset=1 << 40, epoch=epoch)
    assert torch.equal(ts, ts2)
    epoch += 1
    ts3, _ = K.sample_timesteps(hist, full, B, seed=11, offset=1 << 40, epoch=epoch)
    assert not torch.equal(ts, ts3)
    # history not full yet -> uniform draws, pt = 1 (gaussian_diffusion.py:961-962)
    part = full.clone()
    part[3] = H - 1
    tu, pu = K.sample_timesteps(hist, part, B, seed=11, offset=1 << 40, epoch=epoch)
    assert (pu == 1).all()
    fu = torch.bincount(tu, minlength=T).double() / B
    assert ((fu - 1.0 / T).abs() < 5 * (1.0 / T * (1 - 1.0 / T) / B) ** 0.5).all()
    # injected timesteps: only pt
    tin = torch.randint(0, T, (B,), device="cuda")
    t_same, p_in = K.sample_timesteps(hist, full, B, ts_in=tin)
    assert t_same is tin
    torch.testing.assert_close(p_in, p[tin] * T, rtol=1e-12, atol=0)


def test_explicit_stages_match_autograd_path():
    """fused_train_stages (engine path, no autograd) vs training_losses(...).mean().backward() on identical draws."""
    from gdmcf_b200.train_step import fused_train_stages
    make, train_dev, test_dev, _, _, B, n_user = _setup(seed_model=2)
    m_a, d_a, _ = make(False)
    m_b, d_b, _ = make(False)
    users = torch.arange(40, 40 + B, dtype=torch.int32, device="cuda")
    batch = train_dev.batch(users)
    m_a.train(); m_b.train()
    la = d_a.training_losses(m_a, batch, True, index=users)["loss"].mean()
    la.backward()
    stages = fused_train_stages(d_b, m_b, batch, True, index=users)
    _, lb = next(stages)
    grads = {}
    for _, g in stages:
        grads.update(g)
    assert torch.equal(la.detach(), lb)
    pa = dict(m_a.named_parameters())
    for n, g in grads.items():
        ref = pa[n].grad
        assert ref is not None, n
        err = (g - ref).norm() / ref.norm().clamp_min(1e-30)
        assert err < 1e-6, (n, err.item())
    assert {n for n, p in pa.items() if p.grad is not None} == set(grads)


def test_deferred_item_norm_term_equals_epilogue_form():
    """The engine applies the cosine scorer's row-wise norm gradient (-E_i * ri^2 * c_i) inside the AdamW pass; the
    autograd path applies it in the wgrad contraction's epilogue. The effective gradients must agree."""
    make, train_dev, test_dev, _, _, B, n_user = _setup(seed_model=3)
    m_a, _, e_a = make(False)
    m_b, _, e_b = make(False)
    e_b.defer_item_norm = False
    assert e_a.defer_item_norm
    grads = []
    for m, e in ((m_a, e_a), (m_b, e_b)):
        e.load_resident(train_dev, test_dev, 0, B)
        e.capture(warmup=1)
        e.load_resident(train_dev, test_dev, 200, 200 + B)
        before = m.embedding_item.weight.detach().clone()
        e.step()
        g = m.embedding_item.weight.grad.clone()
        if e.defer_item_norm:
            g = g + m._item_grad_rowcoef[:, None] * before
        grads.append(g)
    err = (grads[0] - grads[1]).norm() / grads[1].norm()
    assert err < 1e-6, err.item()


def test_adamw_row_coef_kernel():
    from gdmcf_b200 import kernels as K
    rows, cols = 130, 192
    g0 = torch.Generator(device="cuda").manual_seed(5)
    p = torch.randn(rows, cols, device="cuda", generator=g0) * 0.1
    g = torch.randn(rows, cols, device="cuda", generator=g0)
    coef = torch.randn(rows, device="cuda", generator=g0)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    p2, m2, v2 = p.clone(), m.clone(), v.clone()
    op = K.cast_bf16(p)
    K.adamw_refresh(p, g, m, v, lr=1e-2, weight_decay=0.0, step=1, op=op, row_coef=coef)
    K.adamw_fused(p2.view(-1), torch.addcmul(g, coef[:, None], p2).view(-1), m2.view(-1), v2.view(-1), lr=1e-2, weight_decay=0.0, step=1)
    torch.testing.assert_close(m, m2, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(p, p2, rtol=0, atol=1e-5)
    assert torch.equal(op.hi, K.cast_bf16(p).hi)


def test_engine_dnn_backbone_graph_equals_eager():
    """The DiffRec-style DNN backbone (models/DNN.py:11-88) through the same engine: graph replays == eager program."""
    from gdmcf_b200 import data_utils, dist_utils
    from gdmcf_b200.engine import StepEngine
    from gdmcf_b200.models import gaussian_diffusion as gd
    from gdmcf_b200.models.DNN import DNN
    from gdmcf_b200.optim import FusedAdamW
    dev = torch.device("cuda")
    tr, va, te = data_utils.synthetic_interactions(500, 1203, 15000, 7)
    n_user, n_item = int(tr[:, 0].max()) + 1, int(tr[:, 1].max()) + 1
    mk = lambda p: sp.csr_matrix((np.ones(len(p), dtype=np.float32), (p[:, 0], p[:, 1])), shape=(n_user, n_item))  # noqa: E731
    train_sp, test_sp = mk(tr), mk(te)
    train_dev, test_dev = data_utils.DeviceInteractions(train_sp, dev), data_utils.DeviceInteractions(test_sp, dev)
    B, k = 64, 20
    engines = []
    for graphs in (True, False):
        torch.manual_seed(4)
        model = DNN([n_item, 64], [64, n_item], 10).to(dev)
        diff = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", 0.01, 0.001, 0.01, 5, dev, discrete=0.9995)
        diff.seed = model.seed = 5
        opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0, modules=[model], capturable=True)
        eng = StepEngine(model, diff, opt, dist_utils.Dist(), batch_size=B, n_item=n_item, topk=k, topN=[10, k],
                         cap_train_nnz=int(train_sp.nnz), cap_gt_nnz=int(test_sp.nnz), graphs=graphs)
        eng.load_resident(train_dev, test_dev, 0, B)
        eng.capture(warmup=2)
        engines.append((model, eng))
    for s in range(1, 4):
        outs = []
        for _, e in engines:
            e.load_resident(train_dev, test_dev, s * B, (s + 1) * B)
            outs.append(tuple(t.clone() for t in e.step()))
        assert all(torch.equal(a, b) for a, b in zip(*outs))
        assert torch.isfinite(outs[0][0])
    for (n, pa), (_, pb) in zip(engines[0][0].named_parameters(), engines[1][0].named_parameters()):
        assert torch.equal(pa, pb), n


@pytest.mark.parametrize("world", [2])
def test_engine_data_parallel_torchrun(world):
    """The data-parallel engine under torchrun + NCCL (tests/_engine_dist_worker.py): graph segments with overlapped
    collectives == the same program run eagerly (bit for bit); ranks stay bit-identical; sparse user-row exchange == dense
    all-reduce; reduce-scatter + row-sharded AdamW + all-gather == all-reduce + replicated AdamW."""
    import os
    import socket
    import subprocess
    import sys
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_engine_dist_worker.py")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(port), worker],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert f"engine_dist_check: world {world}" in res.stdout and "OK" in res.stdout, res.stdout[-2000:]


def test_overlapped_optimizer_equals_serial():
    """world_size 1 with overlap_sms > 0 (AdamW of the big matrices on a side stream, confined to a set of SMs by
    gdmcf_adamw_partitioned, + one refresh-only pass) must produce the same weights, operands and results as the serial
    fused update, bit for bit."""
    make, train_dev, test_dev, _, _, B, n_user = _setup()
    outs = []
    for ov in (0, 8):
        model, diff, eng = make(True, overlap_sms=ov)
        eng.load_resident(train_dev, test_dev, 0, B)
        eng.capture(warmup=2)
        res = []
        for s in range(1, 4):
            eng.load_resident(train_dev, test_dev, s * B, (s + 1) * B)
            loss, idx, sums = eng.step()
            res.append((loss.clone(), idx.clone(), sums.clone()))
        torch.cuda.synchronize()
        outs.append((res, [p.detach().clone() for p in model.parameters()]))
    for (l0, i0, s0), (l1, i1, s1) in zip(outs[0][0], outs[1][0]):
        assert torch.equal(l0, l1) and torch.equal(i0, i1) and torch.equal(s0, s1)
    for p0, p1 in zip(outs[0][1], outs[1][1]):
        assert torch.equal(p0, p1)


def test_lazy_user_rows_equal_dense_adamw():
    """The user table's gradient has B non-zero rows per step. The engine's row-sparse AdamW (replay of the skipped
    zero-gradient steps when a row is next used, gdmcf_adamw_rows_lazy) must equal the dense torch.optim.AdamW-style pass
    over the whole table bit for bit: per-step results while training, and every parameter after flush()."""
    make, train_dev, test_dev, _, _, B, n_user = _setup()
    runs = []
    for lazy in (True, False):
        model, diff, eng = make(True, lazy_user_rows=lazy)
        eng.load_resident(train_dev, test_dev, 0, B)
        eng.capture(warmup=2)
        res = []
        for s in (1, 3, 1, 5, 2, 3, 7):  # users come back after different numbers of skipped steps
            lo = (s * B) % (n_user - B)
            eng.load_resident(train_dev, test_dev, lo, lo + B)
            loss, idx, sums = eng.step()
            res.append((loss.clone(), idx.clone(), sums.clone()))
        if lazy:
            stale = model.embedding_user.weight.detach().clone()
        eng.flush()
        torch.cuda.synchronize()
        runs.append((res, {n: p.detach().clone() for n, p in model.named_parameters()}))
    for (l0, i0, s0), (l1, i1, s1) in zip(runs[0][0], runs[1][0]):
        assert torch.equal(l0, l1) and torch.equal(i0, i1) and torch.equal(s0, s1)
    for n in runs[0][1]:
        assert torch.equal(runs[0][1][n], runs[1][1][n]), n
    assert not torch.equal(stale, runs[0][1]["embedding_user.weight"])  # the flush really had pending steps to replay
