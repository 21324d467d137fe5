"""Likelihood halves of the ELBO and the accuracy metric, MC-first
(reference: src/ProbabilisticLayers_Loss.py:27-120, duplicated at src/ProbabilisticLayers.py:22-117).

On CUDA MC_CrossEntropyLoss runs as ONE pass over the logits (csrc/ce.cu: online log-sum-exp, target
pick, argmax and the gradient in the same sweep); the number of argmax hits
of the last call is kept on ``.last_correct`` so ELBOTrainStep gets the accuracy without re-reading the
logits.  MC_Accuracy returns a device tensor instead of forcing a ``.item()`` host sync.  MC_NLL and the
CPU path of the cross entropy stay torch ops.
"""
import math

import torch
import torch.nn.functional as F

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


class MC_CrossEntropyLoss(torch.nn.Module):
    def __init__(self, num_samples):
        super().__init__()
        assert num_samples > 0
        assert type(num_samples) is int
        self.num_samples = num_samples

    def forward(self, pred, target):
        assert pred.dim() == 3, f'Prediction should be of shape [MC, BS, C] but is only of shape {pred.shape}'
        assert target.dim() == 1, f'Targets should be of shape [BS] with each correct label but is only of shape {target.shape}'
        MC, BS, C = pred.shape
        assert target.shape == torch.Size([BS]), f"{target.shape=}"
        if pred.is_cuda and pred.dtype == torch.float32:
            from .functional import mc_cross_entropy
            loss, out = mc_cross_entropy(pred, target, self.num_samples)
            self.last_correct = out[1]
            return loss
        # mean over MC*BS of the per-sample CE == CE against the MC-tiled targets, without
        # materialising target.repeat(MC)
        logp = F.log_softmax(pred, dim=-1)
        picked = logp.gather(-1, target.long().view(1, BS, 1).expand(MC, BS, 1))
        return -picked.mean() * self.num_samples


def MC_Accuracy(pred, target):
    if pred.dim() == 3:
        assert target.dim() == 1, f"{target.dim()=}"
        assert target.shape[0] == pred.shape[1], f"{target.shape=}!={pred.shape=}[1]={pred.shape[1]}"
        correct = pred.argmax(dim=-1).eq(target.view(1, -1))
        return correct.float().mean().reshape(1)
    elif pred.dim() == 2:
        assert target.dim() == 1, f"{target.dim()=}"
        assert target.shape[0] == pred.shape[0], f"{target.shape=}!={pred.shape=}[1]={pred.shape[1]}"
        return pred.argmax(dim=-1).eq(target).float().mean().reshape(1)
    raise ValueError(f"{pred.dim()=}")


class MC_NLL(torch.nn.Module):
    def __init__(self, num_samples):
        super().__init__()
        self.num_samples = num_samples  # rescales the log-likelihood to the full dataset

    def forward(self, pred, target):
        assert pred.dim() == target.dim() + 1
        mu, std = pred.mean(dim=0), pred.std(dim=0) + 1e-3
        assert mu.shape == std.shape == target.shape, f'{mu.shape=} {std.shape=} {target.shape=}'
        nll = _HALF_LOG_2PI + torch.log(std) + (target - mu) ** 2 / (2 * std ** 2)
        return nll.mean() * self.num_samples
