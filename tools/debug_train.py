"""Layer-by-layer diff of the engine's fused train step against an autograd evaluation of the same math on the
CPU (golden inputs). Prints normwise relative errors of every forward intermediate and every backward quantity.
Debug tool (GPU box): python tools/debug_train.py [fp32|bf16]"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import train_step  # noqa: E402
from gdmcf_b200.models import DNN as M  # noqa: E402
from gdmcf_b200.models import gaussian_diffusion as gd  # noqa: E402
from oracle import gdmcf_oracle as O  # noqa: E402

B, I, U, D, E, T = 12, 150, 40, 32, 10, 5
precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "gdmcf_backbone.npz")))
sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


it = 0
x0 = torch.from_numpy(g["x0"])
index = torch.from_numpy(g["index"])
ts = torch.from_numpy(g["train.ts"][it]).long().reshape(-1)
ts1 = torch.from_numpy(g["train.ts_discrete"][it]).long().reshape(-1)
noise = torch.from_numpy(g["train.noise"][it])
u_keep = torch.from_numpy(g["train.u_keep"][it])
kx = torch.from_numpy(g["train.keep_x"][it])
kxu = torch.from_numpy(g["train.keep_xU"][it])

# ---------------- CPU autograd reference with retained intermediates
om = O.OracleGDMCF([I, D], [D, I], E, item_num=I, user_num=U)
om.load_state_dict(sd)
om.train()
sch = O.Schedule(steps=T)
x_t = O.q_sample(sch, x0, ts, noise)
x_tU = O.apply_noise_and_mask(x0, ts1, 0.9995, u_keep)
xd = x_t * kx.float() * 2.0
xud = x_tU.reshape(B, -1) * kxu.float() * 2.0
emb = om.emb_layer(O.timestep_embedding(ts, E))
h = torch.tanh(om.in_layers[0](torch.cat([xd, emb], -1))); h.retain_grad()
hU = torch.tanh(om.in_layers2[0](torch.cat([xud, emb], -1))); hU.retain_grad()
closs = O.nt_xent_loss(h, hU)
eu = om.embedding_user(index)
hc = torch.cat([h, hU, eu], 1); hc.retain_grad()
g1 = torch.relu(F.linear(hc, om.gcn_model.conv1.lin.weight) + om.gcn_model.conv1.bias); g1.retain_grad()
g2 = F.linear(g1, om.gcn_model.conv2.lin.weight) + om.gcn_model.conv2.bias; g2.retain_grad()
hcp = hc * om.sumW + g2 * (1 - om.sumW); hcp.retain_grad()
E_item = om.embedding_item.weight
out = torch.mm(hcp, E_item.t()) / (torch.norm(hcp, dim=1, keepdim=True) * torch.norm(E_item, dim=1)); out.retain_grad()
mse = ((x0 - out) ** 2).mean(1)
w = sch.reweight(ts)
loss = w * mse + 0.1 * closs
loss.mean().backward()
print("reference loss vs golden:", rel(loss.detach(), g["train.loss"][it]))

# ---------------- engine
model = M.DNNOneHotEmbeddingGCN([I, D], [D, I], E, item_num=I, user_num=U, precision=precision)
model.load_state_dict(sd)
model.cuda().train()
diff = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", 0.01, 0.001, 0.01, T, "cuda", discrete=0.9995,
                                    CatOneHot=True)
diff.indexIn = True
inj = dict(noise=noise.cuda(), u_keep=u_keep.cuda(), keep_x=kx.cuda(), keep_xU=kxu.cuda())
x0c = torch.zeros(B, 152, device="cuda"); x0c[:, :I] = x0.cuda()
c = train_step._gdmcf_forward(model, diff, x0c[:, :I], B, I, index.cuda().int(), ts1.cuda().int(), ts.cuda().int(), inj)
d = D
print("A1   ", rel(c.A1.float(), xd))
print("A2   ", rel(c.A2[:, :2 * I].float(), xud))
print("h    ", rel(c.hc_f32[:, :d], h))
print("hU   ", rel(c.hc_f32[:, d:2 * d], hU))
print("eu   ", rel(c.hc_f32[:, 2 * d:], eu))
print("S    ", rel(c.S[:, :B], (h @ hU.t())))
print("closs", rel(c.closs_rows.mean(), closs))
print("g1   ", rel(c.g1_f32, g1))
print("g2   ", rel(c.g2, g2))
print("hcp  ", rel(c.hcp_f32, hcp))
print("inv_u", rel(c.inv_u, 1 / hcp.norm(dim=1)))
print("out  ", rel(c.out[:, :I], out))
print("mse  ", rel(c.mse, mse))

# backward with the same upstream gradients
g_mse = (w / B).float().cuda()
g_closs = torch.tensor(0.1).cuda()
orig = train_step.K.mix_backward
cap = {}


def spy_mix(d_hcp, hc_, g2_, sw, d_hc, d_g2, dw, rows, cols):
    cap["d_hcp"] = d_hcp
    orig(d_hcp, hc_, g2_, sw, d_hc, d_g2, dw, rows, cols)
    cap["d_g2"], cap["d_hc"], cap["dw"] = d_g2, d_hc, dw


train_step.K.mix_backward = spy_mix
orig_ew = train_step.K.ew_binary
ews = []


def spy_ew(op, a, b, rows, cols, **kw):
    orig_ew(op, a, b, rows, cols, **kw)
    ews.append((op, a, kw.get("out_f32")))


train_step.K.ew_binary = spy_ew
grads = train_step._gdmcf_backward(model, diff, c, g_mse, g_closs)
print("d_hcp", rel(cap["d_hcp"], hcp.grad))
print("d_g2 ", rel(cap["d_g2"], g2.grad))
print("dw   ", cap["dw"].sum().item(), om.sumW.grad.item())
print("d_g1 ", rel(ews[0][1], g1.grad))
print("dh_tot", rel(ews[1][1], h.grad), " dhU_tot", rel(ews[2][1], hU.grad))
for k, p in om.named_parameters():
    if p.grad is None:
        continue
    print(f"grad {k:32s} rel {rel(grads[k], p.grad):.3e}   golden {rel(grads[k], g[f'train.grad{it}.{k}']):.3e}")
