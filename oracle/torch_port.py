"""Op-for-op torch-CPU port of the reference's hot path (layers + one ELBO train step).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used as the checker in tests and as the thing
timed by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs (kind "port": the reference is
Python and cannot travel to the GPU box, so its exact sequence of torch calls is restated here).
Pinned against tests/golden/*.npz in tests/test_oracle_golden.py::test_torch_port_*.

Every forward below issues the same torch calls, in the same order and with the same RNG
consumption, as the cited reference lines -- including the parts that make the reference slow
(materialised eps and W of shape [MC,I,O], MC-fold ``repeat`` of the input, ``Normal`` objects
for the KL).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch.distributions import Normal, kl_divergence


class VariationalNormalPort(torch.nn.Module):
    """src/ProbabilisticLayers.py:811-841"""

    def __init__(self, loc, scale):
        super().__init__()
        self.loc = torch.nn.Parameter(loc)
        self.logscale = torch.nn.Parameter(torch.log(torch.exp(scale) - 1))

    def dist(self):
        return Normal(self.loc, F.softplus(self.logscale))

    def rsample(self, sample_shape):
        shape = torch.Size(sample_shape) + self.loc.shape
        self.eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)   # :835
        return self.loc + self.eps * F.softplus(self.logscale)                        # :836


class BayesLinearPort(torch.nn.Module):
    """src/ProbabilisticLayers.py:844-958"""

    def __init__(self, in_features, out_features, prior=1., bias=True):
        super().__init__()
        self.mu_init_std = torch.sqrt(torch.scalar_tensor(2 / (in_features + out_features)))
        self.weight = VariationalNormalPort(
            torch.empty(in_features, out_features).normal_(0., float(self.mu_init_std)),
            torch.full((in_features, out_features), 0.001))
        self.bias = VariationalNormalPort(
            torch.empty(out_features).normal_(0., float(self.mu_init_std)),
            torch.full((out_features,), 0.001)) if bias else None
        self.prior_scale = float(prior)
        torch.nn.init.kaiming_uniform_(self.weight.loc.data, a=math.sqrt(5))            # :880
        self.weight.logscale.data.fill_(
            float(torch.log(torch.exp(self.mu_init_std / out_features) - 1)))            # :881
        if bias:
            bound = 1 / math.sqrt(out_features)                                          # :883-885
            torch.nn.init.uniform_(self.bias.loc.data, -bound, bound)
            self.bias.logscale.data.fill_(float(torch.log(torch.exp(self.mu_init_std) - 1)))

    def forward(self, x):
        assert x.dim() == 3
        mc = x.shape[0]
        self.sampled_w = self.weight.rsample((mc,))                                      # :925
        if self.bias is not None:
            self.sampled_b = self.bias.rsample((mc,))                                    # :926
            out = torch.baddbmm(self.sampled_b.unsqueeze(1), x, self.sampled_w)          # :934
        else:
            out = torch.bmm(x, self.sampled_w)                                           # :935
        self.kl_div = kl_divergence(self.weight.dist(), Normal(0, self.prior_scale)).sum()  # :955
        self.entropy = self.weight.dist().entropy().sum()                                # :956
        return out


class BayesConv2dPort(torch.nn.Module):
    """src/ProbabilisticLayers.py:963-1061"""

    def __init__(self, cin, cout, k, stride=1, padding=0, dilation=1, num_MC=1, prior_scale=1):
        super().__init__()
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.num_MC, self.prior_scale, self.cout = num_MC, prior_scale, cout
        std = torch.scalar_tensor(1. / math.sqrt(cin * k ** 2))
        self.weight = VariationalNormalPort(torch.empty(cout, cin, k, k).normal_(0, float(std)),
                                            torch.full((cout, cin, k, k), 0.001))
        self.bias = VariationalNormalPort(torch.empty(cout).normal_(0., float(std)),
                                          torch.full((cout,), 0.001))
        torch.nn.init.kaiming_uniform_(self.weight.loc.data, a=math.sqrt(5))
        self.weight.logscale.data.fill_(float(torch.log(torch.exp(std / cin) - 1)))
        bound = 1 / math.sqrt(cin * k * k)
        torch.nn.init.uniform_(self.bias.loc.data, -bound, bound)
        self.bias.logscale.data.fill_(float(torch.log(torch.exp(std) - 1)))

    def forward(self, x):
        mc, bs, c, h, w = x.shape
        out = x.permute(1, 0, 2, 3, 4).flatten(1, 2).contiguous()                        # :1032-1033
        wt = self.weight.rsample((mc,)).flatten(0, 1).contiguous()                       # :1038,1044
        b = self.bias.rsample((mc,)).flatten().contiguous()                              # :1039,1045
        out = F.conv2d(out, wt, groups=mc, bias=b, stride=self.stride, padding=self.padding,
                       dilation=self.dilation)                                           # :1051
        out = out.reshape(bs, mc, self.cout, out.shape[-2], out.shape[-1]).contiguous()
        out = out.permute(1, 0, 2, 3, 4)                                                 # :1056-1057
        self.kl_div = kl_divergence(self.weight.dist(), Normal(0, self.prior_scale)).sum()  # :1059
        return out


class MCExpansionPort(torch.nn.Module):
    """src/ProbabilisticLayers.py:765-785 (materialising repeat)"""

    def __init__(self, num_MC, input_dim=2):
        super().__init__()
        self.num_MC, self.input_dim = num_MC, input_dim

    def forward(self, x):
        if x.dim() == self.input_dim:
            return x.unsqueeze(0).repeat(self.num_MC, *(x.dim() * (1,)))
        return x


class MCBatchNorm2DPort(torch.nn.modules.batchnorm._BatchNorm):
    """src/ProbabilisticLayers.py:736-762 (statistics pooled over MC, B, H, W)"""

    def _check_input_dim(self, input):
        return

    def forward(self, x):
        factor = 0.0 if self.momentum is None else self.momentum
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked += 1
            factor = 1.0 / float(self.num_batches_tracked) if self.momentum is None else self.momentum
        out = x.permute(1, 2, 3, 4, 0)                                                   # :757
        out = F.batch_norm(out, self.running_mean, self.running_var, self.weight, self.bias,
                           self.training or not self.track_running_stats, factor, self.eps)
        return out.permute(4, 0, 1, 2, 3)                                                # :761


class MCMaxPool2dPort(torch.nn.Module):
    """src/ProbabilisticLayers.py:207-236 (fold MC into channels, pool, unfold: 3 copies)"""

    def __init__(self, kernel_size=2, stride=2):
        super().__init__()
        self.kernel_size, self.stride = kernel_size, stride

    def forward(self, x):
        mc, bs, c = x.shape[:3]
        out = x.permute(1, 0, 2, 3, 4).flatten(1, 2).contiguous()
        out = F.max_pool2d(out, kernel_size=self.kernel_size, stride=self.stride)
        out = out.reshape(bs, mc, c, out.shape[-2], out.shape[-1]).contiguous()
        return out.permute(1, 0, 2, 3, 4).contiguous()


class FlattenAndLinearPort(torch.nn.Module):
    """src/ProbabilisticLayers.py:787-808"""

    def __init__(self, in_features, num_hidden):
        super().__init__()
        self.linear = BayesLinearPort(in_features, num_hidden)

    def forward(self, x):
        return self.linear(x.flatten(2, -1))


def build_convnet(num_mc, in_dims=(3, 32, 32), num_hidden=200, num_classes=10, chans=None):
    """experiments/BNN.py:228-253 ('cbnn'): 4 x [BayesConv2d(k5,p2) - MC_BatchNorm2D - LeakyReLU -
    MC_MaxPool2d(2,2)] with channels 3-96-128-256-128, then flatten - BayesLinear - LeakyReLU -
    BayesLinear.  ``chans`` overrides the channel list (toy-sized golden trajectory)."""
    chans = [in_dims[0], 96, 128, 256, 128] if chans is None else list(chans)
    mods = [MCExpansionPort(num_mc, 4)]
    hw = in_dims[1]
    for i in range(len(chans) - 1):
        mods += [BayesConv2dPort(chans[i], chans[i + 1], 5, stride=1, padding=2, num_MC=num_mc),
                 MCBatchNorm2DPort(chans[i + 1]), torch.nn.LeakyReLU(), MCMaxPool2dPort(2, 2)]
        hw //= 2
    mods += [FlattenAndLinearPort(chans[-1] * hw * hw, num_hidden), torch.nn.LeakyReLU(),
             BayesLinearPort(num_hidden, num_classes)]
    return torch.nn.Sequential(*mods)


class MCCrossEntropyPort(torch.nn.Module):
    """src/ProbabilisticLayers.py:22-43"""

    def __init__(self, num_samples):
        super().__init__()
        self.num_samples = num_samples

    def forward(self, pred, target):
        mc = pred.shape[0]
        return F.cross_entropy(pred.flatten(0, 1), target.repeat(mc).long(),
                               reduction='mean') * self.num_samples


class MCNLLPort(torch.nn.Module):
    """src/ProbabilisticLayers_Loss.py:95-120"""

    def __init__(self, num_samples):
        super().__init__()
        self.num_samples = num_samples

    def forward(self, pred, target):
        mu, std = pred.mean(dim=0), pred.std(dim=0) + 1e-3
        return -Normal(mu, std).log_prob(target).mean() * self.num_samples


def mc_accuracy_port(pred, target):
    """src/ProbabilisticLayers.py:46-67 (includes the .item() host sync)"""
    mc = pred.shape[0]
    t = target.repeat(mc)
    correct = pred.argmax(dim=-1).flatten().eq(t).sum().item()
    return correct / t.numel()


def build_mlp(dims, num_mc, act=torch.nn.LeakyReLU):
    """experiments/BNN.py:212-226 ('bnn') / experiments/RegressionBNN.py:73-81 assembly."""
    mods = [MCExpansionPort(num_mc, 2)]
    for i in range(len(dims) - 1):
        mods.append(BayesLinearPort(dims[i], dims[i + 1]))
        if i + 2 < len(dims):
            mods.append(act())
    return torch.nn.Sequential(*mods)


def collect_kl_div(net):
    """experiments/BNN.py:269-276: BayesLinear only (conv KL never reaches the loss)."""
    kl = torch.zeros(1)
    for m in net.modules():
        if isinstance(m, BayesLinearPort) and hasattr(m, "kl_div"):
            kl = kl + m.kl_div
    return kl


def train_step(net, criterion, opt, x, y, num_samples, with_accuracy=True):
    """experiments/BNN.py:287-313 + Lightning's zero_grad/backward/step around it."""
    opt.zero_grad()
    pred = net(x)
    acc = mc_accuracy_port(pred, y) if with_accuracy else None
    ll = criterion(pred, y) / num_samples
    kl = collect_kl_div(net) / num_samples
    loss = ll + kl
    loss.backward()
    opt.step()
    return {"loss": float(loss.detach()), "LL": float(ll.detach()), "KL": float(kl.detach()), "ACC": acc}
