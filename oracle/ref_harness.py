"""Train-step harness around the UNMODIFIED reference modules staged in oracle/_ref (oracle/make_ref.py).

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference's experiment scripts need an old
pytorch_lightning and the author's private ``Utils`` package, so the ~40 lines of Lightning glue are
restated here; every layer, loss and metric that runs is the reference's own class:

  build_mlp / build_convnet   experiments/BNN.py:212-226 ('bnn') and :228-253 ('cbnn') assembly
  train_step                  experiments/BNN.py:287-313 (training_step: forward, MC_Accuracy, LL, collect_kl_div,
                              collect_entropy, loss) + the zero_grad / backward / Adam.step Lightning wraps around it
                              (configure_optimizers: Adam, BNN.py:349)

Device-agnostic: ``net.to('cuda')`` times the reference eagerly on the GPU (BASELINE.md section 4's "honest bar").
"""
import torch

from . import make_ref


def modules():
    return make_ref.load()


def build_mlp(dims, num_mc, act=torch.nn.LeakyReLU):
    PL, _ = modules()
    mods = [PL.MC_ExpansionLayer(num_MC=num_mc, input_dim=2)]
    for i in range(len(dims) - 1):
        mods.append(PL.BayesLinear(dims[i], dims[i + 1], prior=1., bias=True))
        if i + 2 < len(dims):
            mods.append(act())
    return torch.nn.Sequential(*mods)


def build_convnet(num_mc, in_dims=(3, 32, 32), num_hidden=200, num_classes=10):
    PL, _ = modules()
    chans = [in_dims[0], 96, 128, 256, 128]
    args = {'kernel_size': 5, 'padding': 2, 'stride': 1, 'num_MC': num_mc}
    mods = [PL.MC_ExpansionLayer(num_MC=num_mc, input_dim=4)]
    for i in range(4):
        mods += [PL.BayesConv2d(in_channels=chans[i], out_channels=chans[i + 1], **args), PL.MC_BatchNorm2D(chans[i + 1]),
                 torch.nn.LeakyReLU(), PL.MC_MaxPool2d(kernel_size=2, stride=2)]
    mods += [PL.BayesAdaptiveInit_FlattenAndLinear(num_hidden), torch.nn.LeakyReLU(), PL.BayesLinear(num_hidden, num_classes)]
    net = torch.nn.Sequential(*mods)
    net(torch.randn(1, *in_dims, dtype=torch.float32))       # the reference's dry run (BNN.py:253)
    return net


def criterion(num_samples):
    PL, _ = modules()
    return PL.MC_CrossEntropyLoss(num_samples=int(num_samples))


def train_step(net, crit, opt, x, y, num_samples):
    PL, _ = modules()
    opt.zero_grad()
    pred = net(x)
    acc = PL.MC_Accuracy(pred, y)                           # includes the reference's .item() host sync
    ll = crit(pred, y) / num_samples
    kl = torch.zeros(1, device=pred.device)
    ent = torch.zeros(1, device=pred.device)
    for m in net.modules():                                 # collect_kl_div / collect_entropy: BayesLinear only
        if isinstance(m, PL.BayesLinear):
            if hasattr(m, 'kl_div'):
                kl = kl + m.kl_div
            if hasattr(m, 'entropy'):
                ent = ent + m.entropy
    kl = kl / num_samples
    loss = ll + kl
    loss.backward()
    opt.step()
    return loss.detach(), ll.detach(), kl.detach(), acc


# ---- regression configs (C1: experiments/RegressionBNN.py, C5: experiments/SIVILastLayer.py) ----------------
def build_regression_mlp(num_mc, num_hidden=50):
    """experiments/RegressionBNN.py:73-81: 1-49-51-52-1 BayesLinear MLP with ReLU."""
    PL, _ = modules()
    h = num_hidden
    return torch.nn.Sequential(PL.MC_ExpansionLayer(num_MC=num_mc, input_dim=2), PL.BayesLinear(1, h - 1), torch.nn.ReLU(),
                               PL.BayesLinear(h - 1, h + 1), torch.nn.ReLU(), PL.BayesLinear(h + 1, h + 2), torch.nn.ReLU(),
                               PL.BayesLinear(h + 2, 1))


class SIVILastLayerNet(torch.nn.Module):
    """experiments/SIVILastLayer.py (LastLayerSIVI): Linear(1,50) - LeakyReLU - Linear(50,50) - LeakyReLU - repeat over MC -
    SIVILayer(50, 1, 5); forward as :255-272."""

    def __init__(self, num_mc, num_hidden=50, dim_noise=5):
        super().__init__()
        PL, _ = modules()
        self.num_mc = num_mc
        self.fc1 = torch.nn.Linear(1, num_hidden)
        self.fc2 = torch.nn.Linear(num_hidden, num_hidden)
        self.fc3 = PL.SIVILayer(num_hidden, 1, dim_noise)

    def forward(self, x):
        out = torch.nn.functional.leaky_relu(self.fc1(x))
        out = torch.nn.functional.leaky_relu(self.fc2(out))
        out = out.unsqueeze(0).repeat(self.num_mc, 1, 1)
        return self.fc3(out)


def nll_criterion(num_samples):
    _, LOSS = modules()
    return LOSS.MC_NLL(num_samples=num_samples)


def train_step_regression(net, crit, opt, x, y, num_samples):
    """experiments/RegressionBNN.py:109-120 / SIVILastLayer.py:287-296: MSE (logged), NLL, KL over BayesLinear / SIVILayer."""
    PL, _ = modules()
    opt.zero_grad()
    pred = net(x)
    mse = torch.nn.functional.mse_loss(pred.mean(dim=0), y)
    ll = crit(pred, y) / num_samples
    kl = 0.
    for m in net.modules():
        if isinstance(m, (PL.BayesLinear, PL.SIVILayer)) and hasattr(m, 'kl_div'):
            kl = kl + m.kl_div
    kl = kl / num_samples
    loss = ll + kl
    loss.backward()
    opt.step()
    return loss.detach(), ll.detach(), kl.detach() if torch.is_tensor(kl) else kl, mse.detach()
