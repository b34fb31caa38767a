"""B200-native LightGCN training + full-rank evaluation hot path.

Host layer (Python, mirroring the reference's duck-typed API) over the C ABI in
include/lgcn_b200.h (hand-written sm_100a CUDA).  Importing the package does not
load the library; the first op does and fails loudly if it is missing.
"""
from .dataloader import BasicDataset, Loader  # noqa: F401
from .model import FusedAdam, LightGCN  # noqa: F401
from .model_ssm import LightGCNSSM  # noqa: F401
from .model_variants import DDPLightGCN, RGCN, rAdjGCN  # noqa: F401
from .negative_sample import UniformSample, UniformSampleCapped, UniformSampling, set_seed  # noqa: F401
from .trainer import Trainer, minibatch, shuffle  # noqa: F401

__all__ = ["BasicDataset", "Loader", "LightGCN", "LightGCNSSM", "rAdjGCN", "RGCN", "DDPLightGCN", "FusedAdam", "UniformSample",
           "UniformSampleCapped", "UniformSampling", "set_seed", "Trainer",
           "minibatch", "shuffle"]
