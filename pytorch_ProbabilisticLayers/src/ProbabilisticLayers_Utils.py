from pytorch_probabilisticlayers_b200.utils import *  # noqa: F401,F403
from pytorch_probabilisticlayers_b200.utils import (AverageMeter, MC_GradientCorrection,  # noqa: F401
                                                   RunningAverageMeter, Timer, str2bool)
