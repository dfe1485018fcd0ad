"""CLI — mirror of the reference's parse_args_util.py:3-49 (same flag names and defaults) on argparse + yaml
(configargparse is not a dependency). Differences, all fixes of defects listed in SURVEY.md §0:
`-c/--config` is optional and merged as defaults (D10); `--dims` accepts `--dims 1000`, repeated flags, and the
README's `--dims=[1000]` (D11); boolean flags parse real booleans (D13). Added flags: --n_user (D3: the
reference hard-codes 3000; 0 = all users), --precision, --eval_batch_size, --synthetic, --eager, --nccl_sms,
--checkpoint_every / --resume (the reference only pickles the best model, main.py:375)."""
from __future__ import annotations

import argparse
import ast

import yaml


def _bool(s):
    if isinstance(s, bool):
        return s
    if s.lower() in ("1", "true", "t", "yes", "y"):
        return True
    if s.lower() in ("0", "false", "f", "no", "n"):
        return False
    raise argparse.ArgumentTypeError(f"not a boolean: {s}")


def _dims(s):
    v = ast.literal_eval(s) if s.strip().startswith("[") else int(s)
    return list(v) if isinstance(v, (list, tuple)) else [int(v)]


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('-c', '--config', default=None, help='Config file path (yaml)')
    parser.add_argument('--dataset', type=str, default='ml-1m_clean', help='choose the dataset')
    parser.add_argument('--data_path', type=str, default='../Datasets/yelp_clean/', help='load data path')
    parser.add_argument('--lr', type=float, default=0.0001, help='learning rate')
    parser.add_argument('--weight_decay', type=float, default=0.0)
    parser.add_argument('--batch_size', type=int, default=400)
    parser.add_argument('--random_seed', type=int, default=1)
    parser.add_argument('--epochs', type=int, default=1000, help='upper epoch limit')
    parser.add_argument('--topN', type=str, default='[10, 20, 50, 100]')
    parser.add_argument('--tst_w_val', action='store_true', help='test with validation')
    parser.add_argument('--cuda', action='store_true', help='use CUDA')
    parser.add_argument('--gpu', type=str, default='0', help='gpu card ID')
    parser.add_argument('--save_path', type=str, default='./saved_models/', help='save model path')
    parser.add_argument('--log_name', type=str, default='log', help='the log name')
    parser.add_argument('--round', type=int, default=1, help='record the experiment')
    parser.add_argument('--out_name', type=str, default='GDMCF', help='output name')
    parser.add_argument('--debug', type=_bool, default=False, help='debug')
    parser.add_argument('--noise_type', type=int, default=0, help='continous noise type')
    parser.add_argument('--gcnLayerNum', type=int, default=2, help='the number of GCN layer')
    parser.add_argument('--user_guided', type=int, default=1, help='user-guided or not')
    # params for the model
    parser.add_argument('--time_type', type=str, default='cat', help='cat or add')
    parser.add_argument('--dims', type=_dims, action='append', help='the dims for the projection')
    parser.add_argument('--norm', type=_bool, default=False, help='Normalize the input or not')
    parser.add_argument('--emb_size', type=int, default=10, help='timestep embedding size')
    parser.add_argument('--backbone', type=str, default='DNNOneHotEmbeddingGCN', help='projection network type')
    parser.add_argument('--OneHotMatrix', type=int, default=2, help='use descrete noise or not')
    # params for diffusion
    parser.add_argument('--mean_type', type=str, default='x0', help='MeanType for diffusion: x0, eps')
    parser.add_argument('--steps', type=int, default=100, help='diffusion steps')
    parser.add_argument('--noise_schedule', type=str, default='linear-var', help='the schedule for noise generating')
    parser.add_argument('--noise_scale', type=float, default=0.1, help='noise scale of for continous noise generating')
    parser.add_argument('--noise_min', type=float, default=0.001, help='noise lower bound')
    parser.add_argument('--noise_max', type=float, default=0.01, help='noise upper bound')
    parser.add_argument('--sampling_noise', type=_bool, default=False, help='sampling with noise or not')
    parser.add_argument('--sampling_steps', type=int, default=0, help='steps of the forward process during inference')
    parser.add_argument('--reweight', type=_bool, default=True, help='assign different weight to different timestep or not')
    parser.add_argument('--discrete', type=float, default=0.9995, help='discrete value of diffusion')
    # engine additions
    parser.add_argument('--n_user', type=int, default=0, help='train/evaluate on the first n users (reference: 3000); 0 = all')
    parser.add_argument('--precision', type=str, default='bf16', choices=['bf16', 'fp32'])
    parser.add_argument('--eval_batch_size', type=int, default=0, help='physical inference batch (0 = batch_size)')
    parser.add_argument('--eval_every', type=int, default=5)
    parser.add_argument('--checkpoint_every', type=int, default=0, help='write <out_path>/checkpoint.pt every N epochs (0 = never)')
    parser.add_argument('--resume', type=str, default='', help='checkpoint.pt to continue from (weights, AdamW state, Lt_history, RNG)')
    parser.add_argument('--faithful_graph', action='store_true',
                        help="run the reference's per-step random edge bookkeeping and LayerGCN over all B + I nodes (item rows "
                             "included) instead of the user-row closed form; same results, for parity work (implies --eager)")
    parser.add_argument('--eager', action='store_true',
                        help='run the loop call by call through the reference-shaped API instead of the captured StepEngine programs')
    parser.add_argument('--nccl_sms', type=int, default=32, help='torchrun: SMs the contractions leave to NCCL while collectives are in flight')
    parser.add_argument('--synthetic', type=str, default='', help="'yelp' | 'amazon' | 'U,I,pairs': generate data instead of loading")
    return parser


def parse_args(argv=None):
    parser = build_parser()
    pre, _ = parser.parse_known_args(argv)
    if pre.config:
        with open(pre.config) as f:
            cfg = yaml.safe_load(f) or {}
        if "dims" in cfg:
            cfg["dims"] = [list(cfg["dims"])] if isinstance(cfg["dims"], (list, tuple)) else [[int(cfg["dims"])]]
        if "gpu" in cfg:
            cfg["gpu"] = str(cfg["gpu"])
        unknown = set(cfg) - {a.dest for a in parser._actions}
        if unknown:
            parser.error(f"unknown keys in {pre.config}: {sorted(unknown)}")
        parser.set_defaults(**cfg)
    args = parser.parse_args(argv)
    flat = [d for group in (args.dims or [[1000]]) for d in group]
    args.dims = flat
    return args
