"""The CLI's ablation switches on the CUDA path against tests/golden/ablations.npz (outputs of the UNMODIFIED reference,
oracle/make_golden_ablations.py): sampling_noise (models/gaussian_diffusion.py:745-750), mean_type=eps (:895-898,924-928,
1085-1090,1106-1111), gcnLayerNum in {0,1} (models/DNN.py:1078-1103,1278), noise_type in {1,2} (models/DNN.py:1236-1259).
Tolerances as in test_models_gpu.py (normwise relative): fp32 mode 5e-5 scores / 3e-4 gradients; bf16 mode 5e-3 / 3e-2."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
B, I, U, D, E, T = 12, 150, 40, 32, 10, 5
TOL = {"fp32": dict(score=5e-5, grad=3e-4, loss=2e-4), "bf16": dict(score=5e-3, grad=3e-2, loss=2e-2)}
# sums that cancel (see test_models_gpu.py) and relu-mask flips of the 12-row batch in bf16 mode
GRAD_SCALE = {"sumW": 60.0, "gcn_model.conv1.bias": 4.0, "gcn_model.conv1.lin.weight": 4.0}


@pytest.fixture(scope="module")
def g(lib):
    assert lib.gdmcf_device_check() == 0
    return dict(np.load(os.path.join(GOLD, "ablations.npz")))


def rel(got, ref):
    got, ref = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def sd(g, tag):
    pre = tag + "sd."
    return {k[len(pre):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(pre)}


def diffusion(mean="x0", cat=True, index_in=True):
    from gdmcf_b200.models import gaussian_diffusion as gd
    d = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X if mean == "x0" else gd.ModelMeanType.EPSILON, "linear-var",
                                     0.01, 0.001, 0.01, T, "cuda", discrete=0.9995, CatOneHot=cat)
    d.indexIn = index_in
    return d


def gdmcf(g, tag, precision, **args):
    from gdmcf_b200.models.DNN import DNNOneHotEmbeddingGCN
    a = types.SimpleNamespace(user_guided=1, gcnLayerNum=2, noise_type=0)
    a.__dict__.update(args)
    m = DNNOneHotEmbeddingGCN([I, D], [D, I], E, item_num=I, user_num=U, args=a, precision=precision)
    m.load_state_dict(sd(g, tag), strict=True)
    return m.cuda()


def replay(g, tag, model, diff, is_gdmcf, precision):
    tol = TOL[precision]
    x0, index = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["index"]).cuda()
    model.train()
    n = g[tag + "loss"].shape[0]
    for it in range(n):
        model.zero_grad()
        inj = dict(ts=torch.from_numpy(g[tag + "ts"][it]).long().reshape(-1).cuda(), noise=torch.from_numpy(g[tag + "noise"][it]).cuda(),
                   keep_x=torch.from_numpy(g[tag + "keep_x"][it]).cuda())
        if is_gdmcf:
            inj.update(ts_discrete=torch.from_numpy(g[tag + "ts_discrete"][it]).long().reshape(-1).cuda(),
                       u_keep=torch.from_numpy(g[tag + "u_keep"][it]).cuda(), keep_xU=torch.from_numpy(g[tag + "keep_xU"][it]).cuda())
        terms = diff.training_losses(model, x0, True, index=index, inject=inj)
        assert rel(terms["loss"], g[tag + "loss"][it]) < tol["loss"], (tag, it, rel(terms["loss"], g[tag + "loss"][it]))
        terms["loss"].mean().backward()
    bad = {}
    for k, p in model.named_parameters():
        ref = g[f"{tag}grad.{k}"]
        if ref.size == 0 or np.abs(ref).max() == 0:
            assert p.grad is None or p.grad.abs().max().item() <= 1e-6 * max(1.0, float(p.abs().max())), k
            continue
        assert p.grad is not None, k
        e = rel(p.grad, ref)
        if e >= tol["grad"] * GRAD_SCALE.get(k, 1.0):
            bad[k] = e
    assert not bad, (tag, bad)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sampling_noise(g, precision):
    m = gdmcf(g, "sn.", precision).eval()
    x0, index = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["index"]).cuda()
    d = diffusion()
    out = d.p_sample(m, x0, 0, True, index=index, inject=dict(sampling_noise=[torch.from_numpy(z).cuda() for z in g["sn.noise"]]))
    assert rel(out, g["sn.out"]) < TOL[precision]["score"]
    # in-kernel Philox draws: a different sample every call, same mean, spread set by the posterior variance
    a, b = d.p_sample(m, x0, 0, True, index=index), d.p_sample(m, x0, 0, True, index=index)
    quiet = d.p_sample(m, x0, 0, False, index=index)
    assert not torch.equal(a, b) and rel(a, quiet) < 0.5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mean_type_eps(g, precision):
    from gdmcf_b200.models.DNN import DNN
    x0, index = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["index"]).cuda()
    m = DNN([I, D], [D, I], E, precision=precision)
    m.load_state_dict(sd(g, "eps_dnn."))
    m.cuda()
    p = diffusion("eps", cat=False, index_in=False).p_sample(m.eval(), x0, 0, index=index)
    assert rel(p, g["eps_dnn.p_sample_s0"]) < TOL[precision]["score"] * 4  # the eps form subtracts two O(1) terms
    replay(g, "eps_dnn.", m, diffusion("eps", cat=False, index_in=False), False, precision)
    m = gdmcf(g, "eps_gdmcf.", precision)
    p = diffusion("eps").p_sample(m.eval(), x0, 0, index=index)
    assert rel(p, g["eps_gdmcf.p_sample_s0"]) < TOL[precision]["score"] * 4
    replay(g, "eps_gdmcf.", m, diffusion("eps"), True, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,kw", [("gcn0.", dict(gcnLayerNum=0)), ("gcn1.", dict(gcnLayerNum=1)),
                                    ("nt1.", dict(noise_type=1)), ("nt2.", dict(noise_type=2))])
def test_gcn_layers_and_noise_types(g, tag, kw, precision):
    m = gdmcf(g, tag, precision, **kw).eval()
    x0, index = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["index"]).cuda()
    x_U = torch.nn.functional.one_hot(x0.long(), 2).float()
    tol = TOL[precision]["score"]
    out = m(torch.from_numpy(g["fwd_x"]).cuda(), torch.from_numpy(g["fwd_ts"]).cuda(), x_U, index=index, graph=x_U.long())
    assert rel(out, g[tag + "fwd_eval"]) < tol
    d = diffusion()
    assert rel(d.p_sample(m, x0, 0, index=index), g[tag + "p_sample_s0"]) < tol
    # CSR input takes the sparse one-hot encoder where the configuration allows it
    import scipy.sparse as sp
    from gdmcf_b200 import data_utils
    full = np.zeros((U, I), dtype=np.float32)
    full[g["index"]] = g["x0"]
    batch = data_utils.DeviceInteractions(sp.csr_matrix(full), "cuda").batch(g["index"].astype(np.int32))
    assert rel(d.p_sample(m, batch, 0), g[tag + "p_sample_s0"]) < tol
    replay(g, tag, m, diffusion(), True, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_faithful_graph_gcn_all_nodes(g, precision):
    """faithful_graph: LayerGCN over all B + I nodes with the user -> item edges (models/DNN.py:1217-1219,1277-1280) through the
    tcgen05 contractions + the CSR SpMM kernel. Item rows against the reference's own GCN output (golden, PyG GCNConv
    restated in oracle/ref_harness.py); user rows and the model output identical to the default closed-form path."""
    x0, index = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["index"]).cuda()
    x_U = torch.nn.functional.one_hot(x0.long(), 2).float()
    xt, ts = torch.from_numpy(g["fwd_x"]).cuda(), torch.from_numpy(g["fwd_ts"]).cuda()
    m = gdmcf(g, "graph.", precision).eval()
    base = m(xt, ts, x_U, index=index, graph=x_U.long()).clone()
    m.faithful_graph = True
    out = m(xt, ts, x_U, index=index, graph=x_U.long())
    tol = TOL[precision]["score"]
    assert rel(out, g["graph.fwd_eval"]) < tol and rel(out, base) < (1e-6 if precision == "fp32" else 2e-3)
    all_rows = m.last_gcn_all
    assert all_rows.shape == (B + I, 3 * D)
    assert rel(all_rows[B:], g["graph.gcn_out"][B:]) < tol * 2      # item rows: neighbour sums of the user rows
    assert rel(all_rows[:B], g["graph.gcn_out"][:B]) < tol * 2      # user rows: self loop only
    # an empty edge set leaves the user rows unchanged and reduces item rows to their self loop
    out0 = m(xt, ts, x_U, index=index, graph=torch.zeros_like(x_U).long())
    assert rel(out0, base) < (1e-6 if precision == "fp32" else 2e-3)


def test_faithful_graph_bookkeeping(g):
    """gdmcf_graph_noise_step against the reference's recorded draws (exact), its flip probability (statistical), and
    p_sample in faithful mode: same scores as the default path whatever edges are drawn."""
    from gdmcf_b200 import kernels as K
    state = torch.zeros(B, I, dtype=torch.uint8, device="cuda")
    deg = torch.from_numpy(g["x0"]).sum(1)
    deg_frac = (deg / deg.max()).float().cuda()
    for step in range(T):
        t = T - 1 - step
        # recorded multinomial outcomes -> uniforms that reproduce them: entry flips iff u >= P(0 -> 0); guide iff u < deg_frac
        u_e = torch.from_numpy(g["graph.entry_draws"][step]).float().cuda().contiguous()
        u_u = (1.0 - torch.from_numpy(g["graph.user_draws"][step]).float()).cuda().contiguous()
        K.graph_noise_step(state, t, B, deg_frac=deg_frac, discrete=0.9995, user_guided=True, u_entry=u_e, u_user=u_u)
        assert torch.equal(state.cpu().long(), torch.from_numpy(g["graph.states"][step]).long()), step
    # Philox draws: flip rate of a class-0 entry = (1 - t/batch)(1 - p) for guided users (all users with deg_frac = 1)
    rows, cols, t, batch, p = 512, 4096, 3, 512, 0.99
    st = torch.zeros(rows, cols, dtype=torch.uint8, device="cuda")
    K.graph_noise_step(st, t, batch, deg_frac=torch.ones(rows, device="cuda"), discrete=p, user_guided=True, seed=3, offset=1 << 40)
    rate, want = st.float().mean().item(), (1 - t / batch) * (1 - p)
    assert abs(rate - want) < 5 * (want / (rows * cols)) ** 0.5
    st.zero_()
    half = torch.full((rows,), 0.5, device="cuda")
    K.graph_noise_step(st, t, batch, deg_frac=half, discrete=p, user_guided=True, seed=4, offset=2 << 40)
    guided = (st.sum(1) > 0).float().mean().item()
    assert abs(guided - 0.5) < 0.12 and abs(st.float().mean().item() - 0.5 * want) < 0.25 * want
    # p_sample: the random edges cannot change the scores (they only reach item rows)
    x0, index = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["index"]).cuda()
    m = gdmcf(g, "graph.", "fp32").eval()
    d = diffusion()
    d.args = types.SimpleNamespace(user_guided=1)
    base = d.p_sample(m, x0, 0, index=index)
    m.faithful_graph = True
    faithful = d.p_sample(m, x0, 0, index=index)
    # (the default path carries the recurrence in the encoder's pre-activation space: same values up to fp32 rounding order)
    assert rel(faithful, base) < 1e-5 and m.last_gcn_all.shape == (B + I, 3 * D)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["deep2.", "deep3."])
def test_deep_encoders(lib, tag, precision):
    """dims with more than one entry (`--dims '[d_a, d_b]'`, main.py:198-206): the encoders are tanh MLPs (models/DNN.py:1240-1252),
    the user tower works on the LAST width, the projected reverse loop carries the FIRST layer's pre-activation. Against
    tests/golden/deep.npz (unmodified reference): eval forward, p_sample (dense and CSR input), training steps with every
    gradient incl. the deep layers'; and the captured engine step == the eager one."""
    from gdmcf_b200.models.DNN import DNNOneHotEmbeddingGCN
    g = dict(np.load(os.path.join(GOLD, "deep.npz")))
    dims = [int(x) for x in g[tag + "in_dims"]]
    a = types.SimpleNamespace(user_guided=1, gcnLayerNum=2, noise_type=0)
    m = DNNOneHotEmbeddingGCN(list(dims), list(dims[::-1]), E, item_num=I, user_num=U, args=a, precision=precision)
    m.load_state_dict(sd(g, tag), strict=True)
    m = m.cuda().eval()
    assert m.deep and m.d1 == dims[1] and m.hidden == dims[-1]
    x0, index = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["index"]).cuda()
    x_U = torch.nn.functional.one_hot(x0.long(), 2).float()
    tol = TOL[precision]["score"]
    out = m(torch.from_numpy(g["fwd_x"]).cuda(), torch.from_numpy(g["fwd_ts"]).cuda(), x_U, index=index, graph=x_U.long())
    assert rel(out, g[tag + "fwd_eval"]) < tol
    d = diffusion()
    assert m.can_project()
    assert rel(d.p_sample(m, x0, 0, index=index), g[tag + "p_sample_s0"]) < tol          # projected loop
    os.environ["GDMCF_PROJECTED_LOOP"] = "0"
    try:
        assert rel(d.p_sample(m, x0, 0, index=index), g[tag + "p_sample_s0"]) < tol      # step-by-step loop
    finally:
        del os.environ["GDMCF_PROJECTED_LOOP"]
    import scipy.sparse as sp
    from gdmcf_b200 import data_utils
    full = np.zeros((U, I), dtype=np.float32)
    full[g["index"]] = g["x0"]
    batch = data_utils.DeviceInteractions(sp.csr_matrix(full), "cuda").batch(g["index"].astype(np.int32))
    assert rel(d.p_sample(m, batch, 0), g[tag + "p_sample_s0"]) < tol                     # sparse one-hot encoder
    replay(g, tag, m, diffusion(), True, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_deep_dnn_backbone(lib, precision):
    """Plain DNN backbone with dims=[a, b] (two in_layers + two out_layers, tanh between, models/DNN.py:72-88) against
    tests/golden/deep.npz: eval forward, p_sample, 2 training steps with every gradient."""
    from gdmcf_b200.models.DNN import DNN
    g, tag = dict(np.load(os.path.join(GOLD, "deep.npz"))), "dnn_deep."
    dims = [int(x) for x in g[tag + "in_dims"]]
    m = DNN(list(dims), list(dims[::-1]), E, precision=precision)
    m.load_state_dict(sd(g, tag), strict=True)
    m = m.cuda().eval()
    assert m.deep and m.hidden == dims[1] and m.d_dec == dims[1]
    x0, index = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["index"]).cuda()
    tol = TOL[precision]["score"]
    assert rel(m(torch.from_numpy(g["fwd_x"]).cuda(), torch.from_numpy(g["fwd_ts"]).cuda()), g[tag + "fwd_eval"]) < tol
    d = diffusion(cat=False, index_in=False)
    # 5 reverse steps through 4 layers each: bf16 mode re-rounds x_t and three activations per step (measured 5.9e-3)
    assert rel(d.p_sample(m, x0, 0, index=index), g[tag + "p_sample_s0"]) < tol * (2 if precision == "bf16" else 1)
    replay(g, tag, m, diffusion(cat=False, index_in=False), False, precision)
