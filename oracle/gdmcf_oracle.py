"""TEST INFRASTRUCTURE — CPU restatement (PyTorch fp32/fp64 on the host) of the reference algorithm for GDMCF's
train-and-rank hot path. It is the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg;
the product (gdmcf_b200/) never imports it.

Why torch-on-CPU rather than C/numpy: the path is floating-point tensor algebra whose reference *is* PyTorch
CPU code; restating it with the same ATen ops keeps the oracle's own rounding behaviour identical to the
reference's (golden vectors agree to ~1e-7). Every function cites the reference lines it restates.

Pinning (tests/test_oracle_golden.py): each function here is checked against outputs of the unmodified reference
modules executed in the authoring container through oracle/ref_harness.py, committed as tests/golden/*.npz by
oracle/make_golden.py, plus the known-answer vectors of SURVEY.md §8c (tests/golden/kat.json).
PARITY UNPINNED for one boundary: torch_geometric's GCNConv is not in the container; `gcn_user_rows` restates
its published algorithm on the only rows the model consumes (user rows: self loop, degree 1).

Randomness is always injected (timesteps, gaussian noise, discrete keep decisions, dropout masks) so that the
oracle, the reference and the engine can be driven with identical draws.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------
# Schedules — models/gaussian_diffusion.py:616-666, :1138-1144
# --------------------------------------------------------------------------------------------------------
def betas_from_linear_variance(steps: int, variance: np.ndarray, max_beta: float = 0.999) -> np.ndarray:
    """gaussian_diffusion.py:1138-1144."""
    alpha_bar = 1 - variance
    betas = [1 - alpha_bar[0]]
    for i in range(1, steps):
        betas.append(min(1 - alpha_bar[i] / alpha_bar[i - 1], max_beta))
    return np.array(betas)


def get_betas(noise_schedule: str, noise_scale: float, noise_min: float, noise_max: float, steps: int) -> np.ndarray:
    """gaussian_diffusion.py:616-637."""
    if noise_schedule in ("linear", "linear-var"):
        start, end = noise_scale * noise_min, noise_scale * noise_max
        lin = np.linspace(start, end, steps, dtype=np.float64)
        return lin if noise_schedule == "linear" else betas_from_linear_variance(steps, lin)
    if noise_schedule == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        return np.array([min(1 - f((i + 1) / steps) / f(i / steps), 0.999) for i in range(steps)])
    if noise_schedule == "binomial":
        return np.array([1 / (steps - t + 1) for t in range(steps)])
    raise NotImplementedError(f"unknown beta schedule: {noise_schedule}!")


class Schedule:
    """Constants of GaussianDiffusionDiscrete.__init__/calculate_for_diffusion (gaussian_diffusion.py:553-595,
    :639-666); all float64 1-D tensors of length `steps`."""

    def __init__(self, noise_schedule="linear-var", noise_scale=0.01, noise_min=0.001, noise_max=0.01, steps=5,
                 beta_fixed=True):
        self.steps = steps
        self.noise_scale = noise_scale
        if noise_scale == 0.0:
            return
        betas = torch.tensor(get_betas(noise_schedule, noise_scale, noise_min, noise_max, steps), dtype=torch.float64)
        if beta_fixed:
            betas[0] = 0.00001
        assert betas.dim() == 1 and len(betas) == steps
        assert (betas > 0).all() and (betas <= 1).all(), "betas out of range"
        self.betas = betas
        alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(alphas, 0)
        self.alphas_cumprod_prev = torch.cat([torch.tensor([1.0]), self.alphas_cumprod[:-1]])
        self.alphas_cumprod_next = torch.cat([self.alphas_cumprod[1:], torch.tensor([0.0])])
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = torch.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = torch.log(
            torch.cat([self.posterior_variance[1].unsqueeze(0), self.posterior_variance[1:]]))
        self.posterior_mean_coef1 = betas * torch.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * torch.sqrt(alphas) / (1.0 - self.alphas_cumprod)

    def SNR(self, t: torch.Tensor) -> torch.Tensor:
        """gaussian_diffusion.py:1113-1118 (negative t wraps like the reference's tensor indexing)."""
        return self.alphas_cumprod[t] / (1 - self.alphas_cumprod[t])

    def reweight(self, ts: torch.Tensor) -> torch.Tensor:
        """START_X branch of gaussian_diffusion.py:919-922 (float64)."""
        w = self.SNR(ts - 1) - self.SNR(ts)
        return torch.where(ts == 0, 1.0, w)


def extract(arr: torch.Tensor, t: torch.Tensor, shape) -> torch.Tensor:
    """_extract_into_tensor, gaussian_diffusion.py:1120-1135: gather float64 coefficient, cast to float32."""
    res = arr[t].float()
    while res.dim() < len(shape):
        res = res[..., None]
    return res.expand(shape)


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """models/DNN.py:1806-1825."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


# --------------------------------------------------------------------------------------------------------
# Noising — gaussian_diffusion.py:988-996 (q_sample), :770-831 + :597-614 + :999-1039 (apply_noise)
# --------------------------------------------------------------------------------------------------------
def q_sample(sch: Schedule, x_start: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    return (extract(sch.sqrt_alphas_cumprod, t, x_start.shape) * x_start
            + extract(sch.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)


def discrete_keep_prob(ts: torch.Tensor, batch_size: int, discrete: float):
    """Probability that apply_noise keeps the true class, per row, for x0 == 1 and x0 == 0.
    Q_bar = a*I + (1-a)*u_x with a = ts.float()/batch_size (gaussian_diffusion.py:775-776, :597-603) and
    u_x = [[p, 1-p], [p, 1-p]] (:591-592); probX = onehot(x0) @ Q_bar (:790) -> P(sample == x0) = Q_bar[x0, x0]."""
    a = ts.float() / batch_size
    u_x = torch.tensor([[discrete, 1 - discrete], [discrete, 1 - discrete]])  # fp32, like the reference
    q = a[:, None, None] * torch.eye(2)[None] + (1 - a[:, None, None]) * u_x[None]
    return q[:, 1, 1], q[:, 0, 0]


def apply_noise_and_mask(x_start: torch.Tensor, ts: torch.Tensor, discrete: float, u_keep: torch.Tensor) -> torch.Tensor:
    """x_tU = apply_noise(ts, one_hot(x0)) & one_hot(x0)  (gaussian_diffusion.py:849-852): [B, I, 2] float.
    u_keep[b, i] in [0,1): the class survives iff u < Q_bar[x0, x0] (injected stand-in for multinomial(1), :1032)."""
    q_one, q_zero = discrete_keep_prob(ts, x_start.shape[0], discrete)
    cls = x_start.long()
    q = torch.where(cls == 1, q_one[:, None], q_zero[:, None])
    kept = u_keep < q
    return F.one_hot(cls, 2).float() * kept[..., None].float()


# --------------------------------------------------------------------------------------------------------
# Denoisers — models/DNN.py:11-88 (DNN), :1105-1327 (DNNOneHotEmbeddingGCN), :1077-1103 (LayerGCN), :479-508
# --------------------------------------------------------------------------------------------------------
def _init_linear(layer: nn.Linear):
    fan_out, fan_in = layer.weight.shape
    std = np.sqrt(2.0 / (fan_in + fan_out))
    layer.weight.data.normal_(0.0, std)
    layer.bias.data.normal_(0.0, 0.001)


def _dropout(x: torch.Tensor, keep: Optional[torch.Tensor], p: float) -> torch.Tensor:
    if keep is None:
        return x
    return x * keep.to(x.dtype) / (1.0 - p)


class OracleDNN(nn.Module):
    """models/DNN.py:11-88 (time_type='cat'); state_dict keys identical to the reference."""

    def __init__(self, in_dims, out_dims, emb_size, norm=False, dropout=0.5):
        super().__init__()
        assert out_dims[0] == in_dims[-1]
        self.in_dims, self.out_dims, self.time_emb_dim, self.norm, self.p = in_dims, out_dims, emb_size, norm, dropout
        self.emb_layer = nn.Linear(emb_size, emb_size)
        in_t = [in_dims[0] + emb_size] + list(in_dims[1:])
        self.in_layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(in_t[:-1], in_t[1:])])
        self.out_layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(out_dims[:-1], out_dims[1:])])
        for l in list(self.in_layers) + list(self.out_layers) + [self.emb_layer]:
            _init_linear(l)

    def forward(self, x, timesteps, keep_x=None):
        emb = self.emb_layer(timestep_embedding(timesteps, self.time_emb_dim))
        if self.norm:
            x = F.normalize(x)
        x = _dropout(x, keep_x, self.p)
        h = torch.cat([x, emb], dim=-1)
        for layer in self.in_layers:
            h = torch.tanh(layer(h))
        for i, layer in enumerate(self.out_layers):
            h = layer(h)
            if i != len(self.out_layers) - 1:
                h = torch.tanh(h)
        return h


def nt_xent_loss(z1, z2, temperature=0.1, eps=1e-5):
    """models/DNN.py:479-508 (returns loss2)."""
    n = z1.size(0)
    sim = torch.mm(z1, z2.t()) / temperature
    mask = torch.eye(n).bool()
    dist = F.softmax(sim, dim=-1)
    negatives = dist.masked_select(~mask).view(n, -1)
    return -torch.log((torch.diag(dist) + eps) / negatives.sum(dim=1)).mean()


class _GCNLin(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))
        a = math.sqrt(6.0 / (i + o))
        nn.init.uniform_(self.weight, -a, a)


class _GCNConvParams(nn.Module):
    """Parameter container with PyG GCNConv's state_dict keys (`bias`, `lin.weight`)."""

    def __init__(self, i, o):
        super().__init__()
        self.lin = _GCNLin(i, o)
        self.bias = nn.Parameter(torch.zeros(o))


class _LayerGCNParams(nn.Module):
    """LayerGCN.__init__ (models/DNN.py:1078-1086): gcnLayerNum == 1 builds a single in -> out convolution, every other
    value builds conv1 (in -> hidden) and conv2 (hidden -> out) (gcnLayerNum == 0 builds them too but never calls them)."""

    def __init__(self, i, h, o, n_layers: int = 2):
        super().__init__()
        self.n_layers = n_layers
        if n_layers == 1:
            self.conv1 = _GCNConvParams(i, o)
        else:
            self.conv1 = _GCNConvParams(i, h)
            self.conv2 = _GCNConvParams(h, o)


def gcn_user_rows(gcn: _LayerGCNParams, hc: torch.Tensor) -> torch.Tensor:
    """LayerGCN.forward (models/DNN.py:1093-1103) restricted to the user rows.
    Edges are user -> item only (DNN.py:1217-1219) and GCNConv aggregates at the target, so a user node receives
    only its own self loop with normalisation deg^-1/2 * 1 * deg^-1/2 = 1: conv(x)[user] = x W^T + b.
    relu then LeakyReLU(0.1) (:1097-1098) is relu. PARITY UNPINNED against real torch_geometric (absent)."""
    h = F.linear(hc, gcn.conv1.lin.weight) + gcn.conv1.bias
    if getattr(gcn, "n_layers", 2) == 1:
        return h  # gcnLayerNum == 1: conv1 only, no activation (:1095-1096)
    h = F.leaky_relu(torch.relu(h), 0.1)
    return F.linear(h, gcn.conv2.lin.weight) + gcn.conv2.bias


def gcn_all_rows(gcn: _LayerGCNParams, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """LayerGCN.forward on ALL B + I nodes with the edge set (models/DNN.py:1093-1103, 1277-1280): restated GCNConv
    (torch_geometric 2.5.3 defaults: self loops added, symmetric in-degree normalisation, aggregation at the target).
    out[i] = sum_{j -> i} deg_j^-1/2 deg_i^-1/2 (x_j W^T) + b, deg = in-degree including the self loop."""
    def conv(c, x):
        n = x.shape[0]
        row, col = edge_index[0], edge_index[1]
        ar = torch.arange(n)
        row, col = torch.cat([row, ar]), torch.cat([col, ar])
        deg = torch.zeros(n, dtype=x.dtype).scatter_add_(0, col, torch.ones(col.numel(), dtype=x.dtype))
        dis = deg.pow(-0.5)
        w = dis[row] * dis[col]
        h = F.linear(x, c.lin.weight)
        return torch.zeros_like(h).index_add_(0, col, h[row] * w[:, None]) + c.bias
    h = conv(gcn.conv1, x)
    if getattr(gcn, "n_layers", 2) == 1:
        return h
    return conv(gcn.conv2, F.leaky_relu(torch.relu(h), 0.1))


class OracleGDMCF(nn.Module):
    """DNNOneHotEmbeddingGCN (models/DNN.py:1105-1327), noise_type=0, gcnLayerNum=2, graph-free closed form."""

    def __init__(self, in_dims, out_dims, emb_size, item_num, user_num, norm=False, dropout=0.5, noise_type=0, gcnLayerNum=2):
        super().__init__()
        in_dims, out_dims = list(in_dims), list(out_dims)
        assert out_dims[0] == in_dims[-1]
        self.time_emb_dim, self.norm, self.p = emb_size, norm, dropout
        self.noise_type, self.gcnLayerNum = noise_type, gcnLayerNum  # ablation switches (DNN.py:1236-1259, :1278)
        in_dims2 = list(in_dims)
        in_dims2[0] *= 2
        self.emb_layer = nn.Linear(emb_size, emb_size)
        in_t = [in_dims[0] + emb_size] + in_dims[1:]
        in_t2 = [in_dims2[0] + emb_size] + in_dims2[1:]
        out_t = list(out_dims)
        out_t[0] += in_dims2[-1]  # DNN.py:1128 (out_layers is built but never used in forward)
        self.in_layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(in_t[:-1], in_t[1:])])
        self.in_layers2 = nn.ModuleList([nn.Linear(a, b) for a, b in zip(in_t2[:-1], in_t2[1:])])
        self.out_layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(out_t[:-1], out_t[1:])])
        e_user = in_t[-1]
        e_item = in_t[-1] + e_user + in_t2[-1]
        self.embedding_item = nn.Embedding(item_num, e_item)
        self.embedding_user = nn.Embedding(user_num, e_user)
        self.gcn_model = _LayerGCNParams(e_item, 512, e_item, gcnLayerNum)
        for l in list(self.in_layers) + list(self.in_layers2) + list(self.out_layers) + [self.emb_layer]:
            _init_linear(l)
        nn.init.xavier_uniform_(self.embedding_item.weight)
        nn.init.xavier_uniform_(self.embedding_user.weight)
        self.sumW = nn.Parameter(torch.tensor(1.0))

    def forward(self, x, timesteps, x_U, index, RCloss=False, keep_x=None, keep_xU=None):
        x_U = x_U.reshape(x_U.shape[0], -1)  # interleaved [i0c0, i0c1, i1c0, ...]  DNN.py:1224
        emb = self.emb_layer(timestep_embedding(timesteps, self.time_emb_dim))
        if self.norm:
            x, x_U = F.normalize(x), F.normalize(x_U)
        x = _dropout(x, keep_x, self.p)
        x_U = _dropout(x_U, keep_xU, self.p)
        # noise_type 1: the continuous branch is fed the first n_item columns of the INTERLEAVED one-hot matrix (:1236-1237);
        # noise_type 2: the one-hot branch is fed [x, x] (:1246-1247); both zero the contrastive loss (:1258-1259)
        h = torch.cat([x_U[:, :x.shape[1]] if self.noise_type == 1 else x, emb], dim=-1)
        for layer in self.in_layers:
            h = torch.tanh(layer(h))
        h_U = torch.cat([x, x, emb] if self.noise_type == 2 else [x_U, emb], dim=-1)
        for layer in self.in_layers2:
            h_U = torch.tanh(layer(h_U))
        closs = nt_xent_loss(h, h_U) if RCloss else None
        if closs is not None and self.noise_type != 0:
            closs = closs * 0
        e_item = self.embedding_item.weight
        e_user = self.embedding_user(index)
        hc = torch.cat([h, h_U, e_user], dim=1)
        g = gcn_user_rows(self.gcn_model, hc) if self.gcnLayerNum > 0 else hc  # :1278: no GCN at all with 0 layers
        hc = hc * self.sumW + g * (1 - self.sumW)  # DNN.py:1288
        user_norms = torch.norm(hc, dim=1, keepdim=True)  # DNN.py:1320-1325
        item_norms = torch.norm(e_item, dim=1)
        out = torch.mm(hc, e_item.t()) / (user_norms * item_norms.t())
        return (out, closs) if RCloss else out


# --------------------------------------------------------------------------------------------------------
# GaussianDiffusionDiscrete — models/gaussian_diffusion.py:552-1135
# --------------------------------------------------------------------------------------------------------
class OracleDiffusion:
    def __init__(self, noise_schedule="linear-var", noise_scale=0.01, noise_min=0.001, noise_max=0.01, steps=5,
                 history_num_per_term=10, discrete=0.9995, CatOneHot=True, indexIn=True, mean_type="x0"):
        self.mean_type = mean_type  # "x0" (ModelMeanType.START_X) or "eps" (EPSILON), gaussian_diffusion.py:10-12
        self.sch = Schedule(noise_schedule, noise_scale, noise_min, noise_max, steps)
        self.steps, self.noise_scale, self.discrete = steps, noise_scale, discrete
        self.CatOneHot, self.indexIn = CatOneHot, indexIn
        self.history_num_per_term = history_num_per_term
        self.Lt_history = torch.zeros(steps, history_num_per_term, dtype=torch.float64)
        self.Lt_count = torch.zeros(steps, dtype=torch.int64)

    # -- gaussian_diffusion.py:959-986
    def importance_ready(self) -> bool:
        return bool((self.Lt_count == self.history_num_per_term).all())

    def importance_probs(self, uniform_prob=0.001) -> torch.Tensor:
        Lt_sqrt = torch.sqrt(torch.mean(self.Lt_history ** 2, dim=-1))
        pt_all = Lt_sqrt / torch.sum(Lt_sqrt)
        pt_all = pt_all * (1 - uniform_prob)
        pt_all = pt_all + uniform_prob / len(pt_all)
        return pt_all

    def pt_for(self, ts: torch.Tensor) -> torch.Tensor:
        """pt returned next to an (injected) draw `ts`: ones (float32) while warming up, else pt_all[t]*T (float64)."""
        if not self.importance_ready():
            return torch.ones_like(ts).float()
        pt_all = self.importance_probs()
        return pt_all.gather(0, ts) * len(pt_all)

    # -- gaussian_diffusion.py:935-949
    def update_history(self, ts: torch.Tensor, loss: torch.Tensor):
        for t, l in zip(ts.tolist(), loss.detach().tolist()):
            if self.Lt_count[t] == self.history_num_per_term:
                old = self.Lt_history.clone()
                self.Lt_history[t, :-1] = old[t, 1:]
                self.Lt_history[t, -1] = l
            else:
                self.Lt_history[t, self.Lt_count[t]] = l
                self.Lt_count[t] += 1

    # -- gaussian_diffusion.py:834-957
    def training_losses(self, model, x_start, index, ts_discrete, ts, noise, u_keep, keep_x, keep_xU, reweight=True):
        """Returns dict(loss [B] f64, model_output, mse, closs, x_t, x_tU). Draws are injected:
        ts_discrete (first sample_timesteps, :845), ts (second, :865), noise (:868), u_keep (multinomial, :1032),
        keep_x / keep_xU (the two nn.Dropout masks, DNN.py:1232-1233)."""
        pt = self.pt_for(ts)  # both sample_timesteps calls see the same Lt state; only the second pt is used
        x_t = q_sample(self.sch, x_start, ts, noise) if self.noise_scale != 0.0 else x_start
        closs = None
        if self.CatOneHot:
            x_tU = apply_noise_and_mask(x_start, ts_discrete, self.discrete, u_keep)
            if self.indexIn:
                out, closs = model(x_t, ts, x_tU, index, RCloss=True, keep_x=keep_x, keep_xU=keep_xU)
            else:
                raise NotImplementedError("only the GDMCF (indexIn) and DNN backbones are on the hot path")
        else:
            x_tU = None
            out = model(x_t, ts, keep_x=keep_x)
        assert out.shape == x_start.shape
        target = x_start if self.mean_type == "x0" else noise           # :895-898
        mse = ((target - out) ** 2).mean(dim=1)
        lossv = mse
        if not reweight:
            weight = torch.ones(len(ts))
        elif self.mean_type == "x0":
            weight = self.sch.reweight(ts)
        else:  # :924-928
            sch = self.sch
            weight = (1 - sch.alphas_cumprod[ts]) / ((1 - sch.alphas_cumprod_prev[ts]) ** 2 * (1 - sch.betas[ts]))
            weight = torch.where(ts == 0, 1.0, weight)
            likelihood = ((x_start - self.predict_xstart_from_eps(x_t, ts, out)) ** 2 / 2.0).mean(dim=1)
            lossv = torch.where(ts == 0, likelihood, mse)
        loss = weight * lossv
        self.update_history(ts, loss)
        loss = loss / pt
        if closs is not None:
            loss = loss + closs * 0.1
        return {"loss": loss, "model_output": out, "mse": mse, "closs": closs, "x_t": x_t, "x_tU": x_tU}

    # -- gaussian_diffusion.py:668-768, :1041-1103 (graph bookkeeping :710-729 cannot affect user rows: omitted)
    def predict_xstart_from_eps(self, x_t, t, eps):
        """gaussian_diffusion.py:1106-1111."""
        return (extract(self.sch.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - extract(self.sch.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def p_sample(self, model, x_start, sampling_steps, index=None, noise=None, ts_u_keep=None, sampling_noise=None):
        """sampling_noise: None (= sampling_noise False) or the list of N(0,1) draws of the reverse steps, one [B, I]
        tensor per step in the order t = T-1 .. 0 (:745-750; the t = 0 draw is made but masked out)."""
        assert sampling_steps <= self.steps, "Too much steps in inference."
        B = x_start.shape[0]
        if self.CatOneHot:
            if sampling_steps == 0:
                x_tU = F.one_hot(x_start.long(), 2).float()
            else:
                t = torch.tensor([sampling_steps - 1] * B)
                x_tU = apply_noise_and_mask(x_start, t, self.discrete, ts_u_keep)
        if sampling_steps == 0:
            x_t = x_start
        else:
            t = torch.tensor([sampling_steps - 1] * B)
            x_t = q_sample(self.sch, x_start, t, noise)
        for i in reversed(range(self.steps)):
            t = torch.tensor([i] * B)
            if self.CatOneHot:
                out = model(x_t, t, x_tU, index)
            else:
                out = model(x_t, t)
            if self.noise_scale == 0.0:
                x_t = out
                continue
            # p_mean_variance + q_posterior_mean_variance (:1041-1103); sampling_noise=False -> x_t = mean
            pred = out if self.mean_type == "x0" else self.predict_xstart_from_eps(x_t, t, out)
            mean = (extract(self.sch.posterior_mean_coef1, t, x_t.shape) * pred
                    + extract(self.sch.posterior_mean_coef2, t, x_t.shape) * x_t)
            if sampling_noise is not None:
                z = sampling_noise[self.steps - 1 - i]
                nonzero = (t != 0).float().view(-1, 1)
                logvar = extract(self.sch.posterior_log_variance_clipped, t, x_t.shape)
                x_t = mean + nonzero * torch.exp(0.5 * logvar) * z
            else:
                x_t = mean
        return x_t


def graph_bookkeeping_step(state: torch.Tensor, entry_sample: torch.Tensor, user_sample: torch.Tensor, user_guided: bool = True):
    """One reverse step of p_sample's random graph bookkeeping (gaussian_diffusion.py:710-729) given the outcomes of its
    two multinomial draws: entry_sample [B, I] = class sampled by apply_noise(t, one_hot(state)) for every entry,
    user_sample [B] = the degree-guided draw (class 1 w.p. deg_b / max deg). x_start_io = x_start_i & one_hot(user draw)
    keeps class 1 only where both are 1 (:721-722); the result is OR-ed into the accumulated state (:726).
    The probabilities: an entry of class 0 is sampled as 1 w.p. (1 - a)(1 - p), a = t / B (get_Qt_bar :597-614, :775)."""
    x_io = (entry_sample.bool() & user_sample.bool()[:, None]) if user_guided else entry_sample.bool()
    return (state.bool() | x_io).long()


# --------------------------------------------------------------------------------------------------------
# Ranking — main.py:267-310, evaluate_utils.py:6-52
# --------------------------------------------------------------------------------------------------------
def mask_topk(prediction: torch.Tensor, history_rows: Sequence[Sequence[int]], k: int):
    """main.py:299-301: -inf at the history coordinates, then torch.topk."""
    pred = prediction.clone()
    for r, items in enumerate(history_rows):
        if len(items):
            pred[r, torch.as_tensor(list(items), dtype=torch.long)] = -np.inf
    vals, idx = torch.topk(pred, k)
    return vals, idx


def computeTopNAccuracy(GroundTruth: List[List[int]], predictedIndices: List[List[int]], topN: List[int]):
    """evaluate_utils.py:6-52, loop for loop (pure Python: use on small inputs)."""
    precision, recall, NDCG, MRR = [], [], [], []
    for index in range(len(topN)):
        sumP = sumR = sumN = sumM = 0
        for i in range(len(predictedIndices)):
            if len(GroundTruth[i]) != 0:
                mrrFlag, userHit, userMRR, dcg, idcg = True, 0, 0, 0, 0
                idcgCount = len(GroundTruth[i])
                ndcg = 0
                for j in range(topN[index]):
                    if predictedIndices[i][j] in GroundTruth[i]:
                        dcg += 1.0 / math.log2(j + 2)
                        if mrrFlag:
                            userMRR = 1.0 / (j + 1.0)
                            mrrFlag = False
                        userHit += 1
                    if idcgCount > 0:
                        idcg += 1.0 / math.log2(j + 2)
                        idcgCount -= 1
                if idcg != 0:
                    ndcg += dcg / idcg
                sumP += userHit / topN[index]
                sumR += userHit / len(GroundTruth[i])
                sumN += ndcg
                sumM += userMRR
        n = len(predictedIndices)
        precision.append(round(sumP / n, 4))
        recall.append(round(sumR / n, 4))
        NDCG.append(round(sumN / n, 4))
        MRR.append(round(sumM / n, 4))
    return precision, recall, NDCG, MRR


# --------------------------------------------------------------------------------------------------------
# LightGCN propagation — lightGCN.py:145-194
# --------------------------------------------------------------------------------------------------------
def lightgcn_norm_adj(users: np.ndarray, items: np.ndarray, n_users: int, n_items: int):
    """get_A_tilda (lightGCN.py:145-178) as a scipy CSR float32 matrix (vectorised; same arithmetic:
    float32 rowsum, d_inv = (rowsum + 1e-9)^-0.5, D A D)."""
    import scipy.sparse as sp
    n = n_users + n_items
    R = sp.coo_matrix((np.ones(len(users), dtype=np.float32), (users, items)), shape=(n_users, n_items)).tocsr()
    R.data[:] = 1.0  # duplicates collapse to 1 like the dok assignment at :147
    adj = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    rowsum = np.array(adj.sum(1), dtype=np.float32)
    d_inv = np.power(rowsum + 1e-9, -0.5).flatten().astype(np.float32)
    d_inv[np.isinf(d_inv)] = 0.0
    D = sp.diags(d_inv)
    return D.dot(adj).dot(D).tocsr().astype(np.float32)


def lightgcn_propagate(norm_adj, E0: torch.Tensor, n_layers: int) -> torch.Tensor:
    """propagate_through_layers (lightGCN.py:180-194): mean over [E0, A E0, ..., A^K E0] via torch.sparse.mm."""
    coo = norm_adj.tocoo()
    A = torch.sparse_coo_tensor(np.vstack((coo.row, coo.col)), coo.data, coo.shape)
    layers, E = [E0], E0
    for _ in range(n_layers):
        E = torch.sparse.mm(A, E)
        layers.append(E)
    return torch.mean(torch.stack(layers), dim=0)
