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
