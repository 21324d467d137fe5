from pytorch_probabilisticlayers_b200.layers import *  # noqa: F401,F403
from pytorch_probabilisticlayers_b200.layers import (BayesAdaptiveInit_FlattenAndLinear, BayesConv2d,  # noqa: F401
                                                    BayesLinear, BayesianNeuralNetwork, GaussianPrior,
                                                    LaplacePrior, MC_BatchNorm1D, MC_BatchNorm2D,
                                                    MC_ExpansionLayer, MC_MaxPool2d, SIVILayer,
                                                    VariationalNormal)
from pytorch_probabilisticlayers_b200.loss import MC_Accuracy, MC_CrossEntropyLoss, MC_NLL  # noqa: F401
