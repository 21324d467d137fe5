#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference at /root/reference.

Run in the builder container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests or fixtures, so these files are the parity
pins: every array is produced by the reference's own classes (loaded by path, read-only, after
stubbing the import-time-only dependencies that are absent here -- SURVEY.md appendix A) under
fixed torch seeds on CPU fp32.  eps is captured from ``layer.weight.eps`` / ``layer.bias.eps``
(src/ProbabilisticLayers.py:835) so that the kernels can be run in injected-eps mode.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "future"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].rcParams = {}
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.lines"].Line2D = object

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    pl = load("ref_PL", f"{REF}/src/ProbabilisticLayers.py")
    loss = load("ref_Loss", f"{REF}/src/ProbabilisticLayers_Loss.py")
    return pl, loss


def npy(t):
    return t.detach().cpu().numpy().copy()


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path)} bytes, {len(arrays)} arrays")


def randomise_logscale(layer, g):
    """Fresh init gives sigma ~1e-5 which would hide eps-path bugs; spread sigma."""
    with torch.no_grad():
        for vn in (layer.weight, layer.bias):
            if vn is not None:
                vn.logscale.copy_(torch.empty_like(vn.logscale).uniform_(-4.0, 0.5, generator=g))


def linear_case(PL, name, mc, b, i, o, bias, seed, x_2d=False):
    g = torch.Generator().manual_seed(seed)
    layer = PL.BayesLinear(i, o, bias=bias, prior=0.7)
    randomise_logscale(layer, g)
    x = torch.randn(mc, b, i, generator=g)
    if x_2d:                               # MC_ExpansionLayer broadcast, then the layer
        x2 = torch.randn(b, i, generator=g)
        x = PL.MC_ExpansionLayer(num_MC=mc, input_dim=2)(x2)
    x.requires_grad_(True)
    gy = torch.randn(mc, b, o, generator=g)
    torch.manual_seed(seed + 1)
    out = layer(x)
    (out * gy).sum().backward()
    arrs = dict(x=npy(x), gy=npy(gy), out=npy(out), dx=npy(x.grad),
                w_mu=npy(layer.weight.loc), w_rho=npy(layer.weight.logscale),
                eps_w=npy(layer.weight.eps), dw_mu=npy(layer.weight.loc.grad),
                dw_rho=npy(layer.weight.logscale.grad), sampled_w=npy(layer.sampled_w),
                prior=np.float64(0.7))
    if bias:
        arrs.update(b_mu=npy(layer.bias.loc), b_rho=npy(layer.bias.logscale),
                    eps_b=npy(layer.bias.eps), db_mu=npy(layer.bias.loc.grad),
                    db_rho=npy(layer.bias.logscale.grad))
    # KL / entropy and their own gradients (weights only; bias gets none -- appendix B.1)
    layer.zero_grad()
    torch.manual_seed(seed + 1)
    layer(x.detach())
    kl, ent = layer.kl_div, layer.entropy
    gk = torch.autograd.grad(kl, [layer.weight.loc, layer.weight.logscale], retain_graph=True)
    ge = torch.autograd.grad(ent, [layer.weight.logscale])
    arrs.update(kl=npy(kl), entropy=npy(ent), dkl_mu=npy(gk[0]), dkl_rho=npy(gk[1]),
                dent_rho=npy(ge[0]))
    save(name, **arrs)


def conv_case(PL, name, mc, b, cin, cout, h, w, k, stride, padding, dilation, seed):
    g = torch.Generator().manual_seed(seed)
    layer = PL.BayesConv2d(cin, cout, k, stride=stride, padding=padding, dilation=dilation,
                           num_MC=mc, prior_scale=1.3)
    randomise_logscale(layer, g)
    x = torch.randn(mc, b, cin, h, w, generator=g, requires_grad=True)
    torch.manual_seed(seed + 1)
    out = layer(x)
    gy = torch.randn(*out.shape, generator=g)
    (out * gy).sum().backward()
    arrs = dict(x=npy(x), gy=npy(gy), out=npy(out), dx=npy(x.grad),
                out_strides=np.array(out.stride()),
                w_mu=npy(layer.weight.loc), w_rho=npy(layer.weight.logscale),
                b_mu=npy(layer.bias.loc), b_rho=npy(layer.bias.logscale),
                eps_w=npy(layer.weight.eps), eps_b=npy(layer.bias.eps),
                dw_mu=npy(layer.weight.loc.grad), dw_rho=npy(layer.weight.logscale.grad),
                db_mu=npy(layer.bias.loc.grad), db_rho=npy(layer.bias.logscale.grad),
                hyper=np.array([k, stride, padding, dilation]), prior=np.float64(1.3))
    layer.zero_grad()
    torch.manual_seed(seed + 1)
    layer(x.detach())
    gk = torch.autograd.grad(layer.kl_div, [layer.weight.loc, layer.weight.logscale])
    arrs.update(kl=npy(layer.kl_div), dkl_mu=npy(gk[0]), dkl_rho=npy(gk[1]))
    save(name, **arrs)


def mlp_train_case(PL, LOSS, name, dims, mc, b, act, loss_kind, num_samples, lr, steps, seed,
                   regression_data=False):
    """N optimiser steps of the ELBO exactly as experiments/BNN.py:287-313 (CE) or
    experiments/RegressionBNN.py:109-120 (NLL) assemble it, Adam (BNN.py:349)."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    mods = [PL.MC_ExpansionLayer(num_MC=mc, input_dim=2)]
    for li in range(len(dims) - 1):
        mods.append(PL.BayesLinear(dims[li], dims[li + 1]))
        if li + 2 < len(dims):
            mods.append(act())
    net = torch.nn.Sequential(*mods)
    layers = [m for m in net if isinstance(m, PL.BayesLinear)]
    for l in layers:
        randomise_logscale(l, g)
    if regression_data:
        rs = np.random.RandomState(2)
        xs = np.linspace(-0.35, 0.55, 1000)
        xn = rs.normal(0.0, 0.01, size=xs.shape)
        yn = rs.normal(loc=0, scale=np.linspace(0, 0.2, 1000))
        ys = xs + 0.3 * np.sin(2 * np.pi * (xs + xn)) + 0.3 * np.sin(4 * np.pi * (xs + xn)) + yn
        xt = torch.from_numpy(xs.reshape(-1, 1)).float()
        yt = torch.from_numpy(ys.reshape(-1, 1)).float()
        xt = (xt - xt.mean(0)) / (xt.std(0) + 0.1)      # experiments/RegressionBNN.py:200-201
        yt = (yt - yt.mean(0)) / (yt.std(0) + 0.1)
        idx = torch.randperm(1000, generator=g)[:b]
        x, y = xt[idx], yt[idx]
    else:
        x = torch.randn(b, dims[0], generator=g)
        y = torch.randint(0, dims[-1], (b,), generator=g)
    crit = (LOSS.MC_NLL(num_samples) if loss_kind == "nll"
            else PL.MC_CrossEntropyLoss(num_samples))
    opt = torch.optim.Adam(net.parameters(), lr)
    arrs = dict(x=npy(x), y=npy(y), dims=np.array(dims), mc=np.int64(mc),
                num_samples=np.int64(num_samples), lr=np.float64(lr), steps=np.int64(steps))
    for li, l in enumerate(layers):
        arrs[f"init_w_mu{li}"] = npy(l.weight.loc)
        arrs[f"init_w_rho{li}"] = npy(l.weight.logscale)
        arrs[f"init_b_mu{li}"] = npy(l.bias.loc)
        arrs[f"init_b_rho{li}"] = npy(l.bias.logscale)
    for s in range(steps):
        opt.zero_grad()
        torch.manual_seed(seed + 100 + s)
        pred = net(x)
        ll = crit(pred, y) / num_samples
        kl = sum(l.kl_div for l in layers) / num_samples
        loss = ll + kl
        loss.backward()
        arrs[f"s{s}_pred"] = npy(pred)
        arrs[f"s{s}_LL"] = npy(ll)
        arrs[f"s{s}_KL"] = npy(kl)
        arrs[f"s{s}_loss"] = npy(loss)
        for li, l in enumerate(layers):
            arrs[f"s{s}_eps_w{li}"] = npy(l.weight.eps)
            arrs[f"s{s}_eps_b{li}"] = npy(l.bias.eps)
            if s == 0:
                arrs[f"s0_g_w_mu{li}"] = npy(l.weight.loc.grad)
                arrs[f"s0_g_w_rho{li}"] = npy(l.weight.logscale.grad)
                arrs[f"s0_g_b_mu{li}"] = npy(l.bias.loc.grad)
                arrs[f"s0_g_b_rho{li}"] = npy(l.bias.logscale.grad)
        opt.step()
    for li, l in enumerate(layers):
        arrs[f"final_w_mu{li}"] = npy(l.weight.loc)
        arrs[f"final_w_rho{li}"] = npy(l.weight.logscale)
        arrs[f"final_b_mu{li}"] = npy(l.bias.loc)
        arrs[f"final_b_rho{li}"] = npy(l.bias.logscale)
    save(name, **arrs)


def sivi_case(PL, name, mc, b, i, seed):
    torch.manual_seed(seed)
    layer = PL.SIVILayer(i, 1, 5)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(mc, b, i, generator=g, requires_grad=True)
    gy = torch.randn(mc, b, 1, generator=g)
    torch.manual_seed(seed + 1)
    out = layer(x)
    for t in (layer.w_mu, layer.w_std, layer.b_mu, layer.b_std):
        t.retain_grad()
    (out * gy).sum().backward(retain_graph=True)
    # draw order noise -> eps_w -> eps_b: replay to capture eps
    torch.manual_seed(seed + 1)
    noise = torch.randn(mc, 5)
    eps_w = torch.randn(mc, i, 1)
    eps_b = torch.randn(mc, 1)
    arrs = dict(x=npy(x), gy=npy(gy), out=npy(out), dx=npy(x.grad), noise=npy(noise),
                w_mu=npy(layer.w_mu), w_std=npy(layer.w_std), b_mu=npy(layer.b_mu),
                b_std=npy(layer.b_std), eps_w=npy(eps_w), eps_b=npy(eps_b),
                sampled_w=npy(layer.sampled_w), sampled_b=npy(layer.sampled_b),
                dw_mu=npy(layer.w_mu.grad), dw_std=npy(layer.w_std.grad),
                db_mu=npy(layer.b_mu.grad), db_std=npy(layer.b_std.grad),
                kl=npy(layer.kl_div))
    sd = layer.state_dict()
    for k_, v in sd.items():
        arrs["sd__" + k_.replace(".", "__")] = npy(v)
    grads = torch.autograd.grad(layer.kl_div, list(layer.sivi_net.parameters()))
    for (n_, _), g_ in zip(layer.sivi_net.named_parameters(), grads):
        arrs["dkl__" + n_.replace(".", "__")] = npy(g_)
    save(name, **arrs)


def loss_cases(PL, LOSS):
    g = torch.Generator().manual_seed(77)
    pred = torch.randn(6, 9, 5, generator=g, requires_grad=True)
    tgt = torch.randint(0, 5, (9,), generator=g)
    ce = PL.MC_CrossEntropyLoss(1000)(pred, tgt)
    ce.backward()
    acc = PL.MC_Accuracy(pred, tgt)
    pr = torch.randn(7, 11, 1, generator=g, requires_grad=True)
    yr = torch.randn(11, 1, generator=g)
    nll = LOSS.MC_NLL(500)(pr, yr)
    nll.backward()
    save("losses", ce_pred=npy(pred), ce_target=npy(tgt), ce_loss=npy(ce), ce_dpred=npy(pred.grad),
         ce_acc=npy(acc), nll_pred=npy(pr), nll_target=npy(yr), nll_loss=npy(nll),
         nll_dpred=npy(pr.grad))


def aux_layer_cases(PL):
    g = torch.Generator().manual_seed(5)
    x5 = torch.randn(3, 4, 6, 8, 8, generator=g)
    bn2 = PL.MC_BatchNorm2D(6)
    y_bn2 = bn2(x5)
    mp = PL.MC_MaxPool2d(2, 2)(x5)
    x3 = torch.randn(3, 5, 7, generator=g)
    bn1 = PL.MC_BatchNorm1D(7)
    y_bn1 = bn1(x3)
    ex = PL.MC_ExpansionLayer(num_MC=4, input_dim=2)(x3[0])
    save("aux_layers", x5=npy(x5), bn2=npy(y_bn2), bn2_rm=npy(bn2.running_mean),
         bn2_rv=npy(bn2.running_var), maxpool=npy(mp), x3=npy(x3), bn1=npy(y_bn1),
         bn1_rm=npy(bn1.running_mean), bn1_rv=npy(bn1.running_var), expand=npy(ex))


def init_case(PL):
    """Fresh-init statistics of BayesLinear(784,1200) (SURVEY 8c pin 6) + conv init."""
    torch.manual_seed(0)
    l = PL.BayesLinear(784, 1200)
    x = torch.zeros(1, 1, 784)
    l(x)
    c = PL.BayesConv2d(3, 96, 5, padding=2, num_MC=1)
    save("init_stats", lin_kl=npy(l.kl_div), lin_entropy=npy(l.entropy),
         lin_w_rho0=npy(l.weight.logscale.flatten()[0]), lin_b_rho0=npy(l.bias.logscale.flatten()[0]),
         lin_w_mu_absmax=npy(l.weight.loc.abs().max()), lin_b_mu_absmax=npy(l.bias.loc.abs().max()),
         conv_w_rho0=npy(c.weight.logscale.flatten()[0]), conv_b_rho0=npy(c.bias.logscale.flatten()[0]),
         conv_w_mu_absmax=npy(c.weight.loc.abs().max()), conv_b_mu_absmax=npy(c.bias.loc.abs().max()))


def sivi_train_case(PL, LOSS, name, mc, b, steps, seed):
    """Config 5 (experiments/SIVILastLayer.py:244-296): Linear(1,50)-LeakyReLU-Linear(50,50)-LeakyReLU-
    repeat(MC)-SIVILayer(50,1,5), MC_NLL likelihood + SIVI kl_div, Adam lr 1e-2."""
    torch.manual_seed(seed)
    fc1, fc2 = torch.nn.Linear(1, 50), torch.nn.Linear(50, 50)
    sivi = PL.SIVILayer(50, 1, 5)
    params = list(fc1.parameters()) + list(fc2.parameters()) + list(sivi.parameters())
    rs = np.random.RandomState(2)
    xs = np.linspace(-0.35, 0.55, 1000)
    ys = xs + 0.3 * np.sin(2 * np.pi * xs) + 0.3 * np.sin(4 * np.pi * xs) + rs.normal(0, 0.05, 1000)
    idx = rs.permutation(1000)[:b]
    x = torch.from_numpy(xs[idx].reshape(-1, 1)).float()
    y = torch.from_numpy(ys[idx].reshape(-1, 1)).float()
    num_samples = 1000
    crit = LOSS.MC_NLL(num_samples)
    opt = torch.optim.Adam(params, 1e-2)
    arrs = dict(x=npy(x), y=npy(y), mc=np.int64(mc), steps=np.int64(steps), num_samples=np.int64(num_samples))
    for k_, v in fc1.state_dict().items():
        arrs["fc1__" + k_] = npy(v)
    for k_, v in fc2.state_dict().items():
        arrs["fc2__" + k_] = npy(v)
    for k_, v in sivi.state_dict().items():
        arrs["sivi__" + k_.replace(".", "__")] = npy(v)
    for s_ in range(steps):
        opt.zero_grad()
        h = torch.nn.functional.leaky_relu(fc2(torch.nn.functional.leaky_relu(fc1(x))))
        h = h.unsqueeze(0).repeat(mc, 1, 1)                      # SIVILastLayer.py:268
        torch.manual_seed(seed + 100 + s_)
        pred = sivi(h)
        torch.manual_seed(seed + 100 + s_)                       # replay the layer's draws
        arrs[f"s{s_}_noise"] = npy(torch.randn(mc, 5))
        arrs[f"s{s_}_eps_w"] = npy(torch.randn(mc, 50, 1))
        arrs[f"s{s_}_eps_b"] = npy(torch.randn(mc, 1))
        ll = crit(pred, y) / num_samples
        kl = sivi.kl_div / num_samples
        loss = ll + kl
        loss.backward()
        arrs[f"s{s_}_pred"] = npy(pred)
        arrs[f"s{s_}_LL"] = npy(ll)
        arrs[f"s{s_}_KL"] = npy(kl)
        arrs[f"s{s_}_loss"] = npy(loss)
        opt.step()
    arrs["final_fc1_weight"] = npy(fc1.weight)
    arrs["final_sivi_last_bias"] = npy(sivi.sivi_net[4].bias)
    save(name, **arrs)


def seeded_init_case(PL):
    """state_dicts of freshly constructed layers under a fixed global seed (drop-in init parity)."""
    arrs = {}
    torch.manual_seed(3)
    for k, v in PL.BayesLinear(7, 5).state_dict().items():
        arrs["lin__" + k.replace(".", "__")] = npy(v)
    torch.manual_seed(4)
    for k, v in PL.BayesConv2d(2, 3, 3).state_dict().items():
        arrs["conv__" + k.replace(".", "__")] = npy(v)
    torch.manual_seed(5)
    for k, v in PL.SIVILayer(6, 1, 5).state_dict().items():
        arrs["sivi__" + k.replace(".", "__")] = npy(v)
    save("seeded_init", **arrs)


def convnet_train_case(PL, name, mc, b, hw, chans, hidden, classes, num_samples, lr, steps, seed):
    """The reference's 'cbnn' block structure (experiments/BNN.py:228-253) at toy size, trained exactly as
    BNN.training_step assembles the ELBO (:287-313): BayesConv2d(k5,p2) -> MC_BatchNorm2D -> LeakyReLU ->
    MC_MaxPool2d per block, then BayesAdaptiveInit_FlattenAndLinear -> LeakyReLU -> BayesLinear; KL is collected
    from BayesLinear layers only (collect_kl_div, :269-276: conv KL never reaches the loss)."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    mods = [PL.MC_ExpansionLayer(num_MC=mc, input_dim=4)]
    for ci in range(len(chans) - 1):
        mods += [PL.BayesConv2d(chans[ci], chans[ci + 1], kernel_size=5, padding=2, stride=1, num_MC=mc),
                 PL.MC_BatchNorm2D(chans[ci + 1]), torch.nn.LeakyReLU(), PL.MC_MaxPool2d(kernel_size=2, stride=2)]
    mods += [PL.BayesAdaptiveInit_FlattenAndLinear(hidden), torch.nn.LeakyReLU(), PL.BayesLinear(hidden, classes)]
    net = torch.nn.Sequential(*mods)
    net(torch.randn(1, chans[0], hw, hw))                      # the reference's dry run creates the flatten layer
    convs = [m for m in net if isinstance(m, PL.BayesConv2d)]
    lins = [net[-3].linear, net[-1]]
    for l in convs + lins:
        randomise_logscale(l, g)
    x = torch.randn(b, chans[0], hw, hw, generator=g)
    y = torch.randint(0, classes, (b,), generator=g)
    crit = PL.MC_CrossEntropyLoss(num_samples)
    opt = torch.optim.Adam(net.parameters(), lr)
    arrs = dict(x=npy(x), y=npy(y), chans=np.array(chans), hw=np.int64(hw), hidden=np.int64(hidden),
                classes=np.int64(classes), mc=np.int64(mc), num_samples=np.int64(num_samples),
                lr=np.float64(lr), steps=np.int64(steps))
    for li, l in enumerate(convs + lins):
        arrs[f"init_w_mu{li}"] = npy(l.weight.loc)
        arrs[f"init_w_rho{li}"] = npy(l.weight.logscale)
        arrs[f"init_b_mu{li}"] = npy(l.bias.loc)
        arrs[f"init_b_rho{li}"] = npy(l.bias.logscale)
    net.train()
    for s in range(steps):
        opt.zero_grad()
        torch.manual_seed(seed + 100 + s)
        pred = net(x)
        ll = crit(pred, y) / num_samples
        kl = sum(l.kl_div for l in lins) / num_samples
        loss = ll + kl
        loss.backward()
        arrs[f"s{s}_pred"] = npy(pred)
        arrs[f"s{s}_loss"] = npy(loss)
        for li, l in enumerate(convs + lins):
            arrs[f"s{s}_eps_w{li}"] = npy(l.weight.eps)
            arrs[f"s{s}_eps_b{li}"] = npy(l.bias.eps)
            if s == 0:
                arrs[f"s0_g_w_mu{li}"] = npy(l.weight.loc.grad)
                arrs[f"s0_g_w_rho{li}"] = npy(l.weight.logscale.grad)
        opt.step()
    for li, l in enumerate(convs + lins):
        arrs[f"final_w_mu{li}"] = npy(l.weight.loc)
        arrs[f"final_w_rho{li}"] = npy(l.weight.logscale)
    bns = [m for m in net if isinstance(m, PL.MC_BatchNorm2D)]
    for bi, bn in enumerate(bns):
        arrs[f"final_bn_mean{bi}"] = npy(bn.running_mean)
        arrs[f"final_bn_var{bi}"] = npy(bn.running_var)
        arrs[f"final_bn_w{bi}"] = npy(bn.weight)
        arrs[f"final_bn_b{bi}"] = npy(bn.bias)
    save(name, **arrs)


def laplace_case(PL, name, mc, b, i, o, steps, seed, prior="laplace"):
    """BayesLinear(prior='laplace' | 'laplace_clamp') (src/ProbabilisticLayers.py:131-159, :867-870): the prior
    scale is an EMA of 1/grad^2 of weight.loc, updated by a tensor hook during backward; KL against the
    per-element prior.  `steps` SGD-free steps (gradients only, parameters fixed) so the EMA is what is pinned."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    layer = PL.BayesLinear(i, o, prior=prior)
    randomise_logscale(layer, g)
    x = torch.randn(mc, b, i, generator=g)
    gy = torch.randn(mc, b, o, generator=g)
    arrs = dict(x=npy(x), gy=npy(gy), w_mu=npy(layer.weight.loc), w_rho=npy(layer.weight.logscale),
                b_mu=npy(layer.bias.loc), b_rho=npy(layer.bias.logscale), steps=np.int64(steps))
    for s in range(steps):
        layer.zero_grad()
        torch.manual_seed(seed + 10 + s)
        out = layer(x)
        loss = (out * gy).sum() + 0.01 * layer.kl_div
        arrs[f"s{s}_prior_scale_in"] = npy(layer.prior.dist().scale * torch.ones_like(layer.weight.loc))
        arrs[f"s{s}_kl"] = npy(layer.kl_div)
        loss.backward()
        arrs[f"s{s}_eps_w"] = npy(layer.weight.eps)
        arrs[f"s{s}_eps_b"] = npy(layer.bias.eps)
        arrs[f"s{s}_out"] = npy(out)
        arrs[f"s{s}_g_w_mu"] = npy(layer.weight.loc.grad)
        arrs[f"s{s}_g_w_rho"] = npy(layer.weight.logscale.grad)
        arrs[f"s{s}_ema"] = npy(layer.prior.scale)
    save(name, **arrs)


def main():
    torch.set_num_threads(4)
    torch.manual_seed(0)          # the first case constructs its layer before any other seeding: make it reproducible
    PL, LOSS = load_reference()
    linear_case(PL, "linear_small", mc=3, b=7, i=5, o=4, bias=True, seed=11)
    linear_case(PL, "linear_nobias", mc=2, b=3, i=6, o=9, bias=False, seed=12)
    linear_case(PL, "linear_odd", mc=5, b=33, i=70, o=67, bias=True, seed=13)
    linear_case(PL, "linear_expand", mc=4, b=9, i=8, o=16, bias=True, seed=14, x_2d=True)
    linear_case(PL, "linear_tile", mc=2, b=128, i=128, o=128, bias=True, seed=15)
    conv_case(PL, "conv_small", mc=2, b=3, cin=3, cout=4, h=6, w=5, k=3, stride=1, padding=1,
              dilation=1, seed=21)
    conv_case(PL, "conv_strided", mc=3, b=2, cin=2, cout=5, h=9, w=8, k=3, stride=2, padding=2,
              dilation=2, seed=22)
    conv_case(PL, "conv_k5", mc=2, b=2, cin=3, cout=6, h=8, w=8, k=5, stride=1, padding=2,
              dilation=1, seed=23)
    mlp_train_case(PL, LOSS, "train_c1_regression", [1, 49, 51, 52, 1], mc=10, b=50,
                   act=torch.nn.ReLU, loss_kind="nll", num_samples=1000, lr=1e-2, steps=3,
                   seed=31, regression_data=True)
    mlp_train_case(PL, LOSS, "train_mlp_ce", [12, 16, 16, 5], mc=4, b=6, act=torch.nn.LeakyReLU,
                   loss_kind="ce", num_samples=100, lr=1e-3, steps=3, seed=32)
    sivi_case(PL, "sivi_small", mc=4, b=3, i=6, seed=41)
    sivi_train_case(PL, LOSS, "train_c5_sivi", mc=128, b=100, steps=3, seed=51)
    convnet_train_case(PL, "train_c3_convnet", mc=3, b=4, hw=8, chans=[3, 6, 8], hidden=10, classes=5,
                       num_samples=200, lr=1e-3, steps=3, seed=71)
    laplace_case(PL, "laplace_prior", mc=3, b=5, i=6, o=4, steps=3, seed=61)
    laplace_case(PL, "laplace_prior_clamp", mc=2, b=4, i=5, o=8, steps=2, seed=62, prior="laplace_clamp")
    loss_cases(PL, LOSS)
    aux_layer_cases(PL)
    init_case(PL)
    seeded_init_case(PL)


if __name__ == "__main__":
    main()
