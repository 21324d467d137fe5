"""B200-native MC-parallel variational layers (hot path of pytorch_ProbabilisticLayers).

Python here is only the host-side mirror of the reference's nn.Module API; all arithmetic runs in
the hand-written sm_100a kernels of ``lib/libpl_b200.so`` (C ABI: include/pl_b200.h).
"""
from . import _cabi, functional  # noqa: F401
from .layers import (BayesAdaptiveInit_FlattenAndLinear, BayesConv2d, BayesLinear,  # noqa: F401
                     BayesianNeuralNetwork, GaussianPrior, LaplacePrior, MC_BatchNorm1D,
                     MC_BatchNorm2D, MC_ExpansionLayer, MC_MaxPool2d, SIVILayer, VariationalNormal,
                     manual_seed, set_eps_source, set_mc_shard, set_precision)
from .loss import MC_Accuracy, MC_CrossEntropyLoss, MC_NLL  # noqa: F401

__version__ = "0.1.0"
