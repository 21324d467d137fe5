"""Small host-side helpers the reference's experiments import next to the layers
(reference: src/ProbabilisticLayers_Utils.py:27-230).  Not on the hot path."""
import time

import torch
from torch.optim.optimizer import Optimizer, required


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise ValueError('Boolean value expected.')


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class RunningAverageMeter(object):
    """Exponential moving average with a non-improvement counter."""

    def __init__(self, momentum=0.99, _max_non_improving=20, init_avg=1.0):
        self.momentum = momentum
        self.last_value = 0
        self.non_improving_counter = 0
        self.step = 0
        self.max_non_improving_steps = _max_non_improving
        self.best_val = 10000
        self.avg = init_avg

    def reset(self):
        self.val = self.last_val = self.avg = self.last_avg = None
        self.step = 0

    def update(self, val):
        self.last_val = val
        self.step += 1
        if self.avg is None:
            self.avg = self.last_avg = val
        else:
            self.last_avg = self.avg
            self.avg = self.avg * self.momentum + val * (1 - self.momentum)

    def stopping_criterion(self):
        if self.avg <= self.best_val:
            self.best_val = self.avg
            self.non_improving_counter = 0
        else:
            self.non_improving_counter += 1
        return self.non_improving_counter >= self.max_non_improving_steps


class Timer(object):
    def __init__(self):
        self.t0 = time.time()

    def elapsed(self):
        return time.time() - self.t0


class MC_GradientCorrection(Optimizer):
    """Multiplies gradients by num_MC (the reference documents it as unnecessary; kept importable)."""

    def __init__(self, params, num_MC=required):
        super().__init__(params, dict(num_MC=num_MC))

    def step(self):
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is not None:
                    p.grad.data.mul_(group['num_MC'])
