"""Drop-in import shim: the reference's experiments import
``pytorch_ProbabilisticLayers.{src,data}.*`` (experiments/BNN.py:42-47, RegressionBNN.py:35-41,
SIVILastLayer.py:35-41).  Every name resolves to the B200-native implementation in
``pytorch_probabilisticlayers_b200``."""
