"""gdmcf_b200 — B200-native (sm_100a) engine for GDMCF's train-and-rank hot path.

Layout mirrors the reference so it drops in for that path:
  gdmcf_b200.models.gaussian_diffusion  <- models/gaussian_diffusion.py (GaussianDiffusionDiscrete)
  gdmcf_b200.models.DNN                 <- models/DNN.py (DNN, DNNOneHotEmbeddingGCN)
  gdmcf_b200.evaluate_utils / data_utils / parse_args_util / main / lightGCN
All device arithmetic lives in libgdmcf_sm100.so (csrc/, C ABI in include/gdmcf_sm100.h).
"""
__version__ = "0.1.0"
