from pytorch_probabilisticlayers_b200.loss import MC_Accuracy, MC_CrossEntropyLoss, MC_NLL  # noqa: F401
