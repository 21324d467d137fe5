"""numpy restatement of the reference's MC-parallel variational-layer arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity pinned against
tests/golden/*.npz, which were produced by the unmodified reference
(tests/golden/make_golden.py).

All functions take numpy arrays, compute in ``dtype`` (float64 unless stated) and follow
the reference line by line; citations are relative to /root/reference.

Conventions (README.md:12): dim 0 = MC sample, dim 1 = batch.  Linear weights are laid
out [in_features, out_features] (src/ProbabilisticLayers.py:858); conv weights
[Cout, Cin, k, k] (:993).
"""
from __future__ import annotations

import math

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------
# elementwise pieces
# ----------------------------------------------------------------------------------------
def softplus(x, dtype=np.float64):
    """F.softplus(beta=1, threshold=20) as used at src/ProbabilisticLayers.py:826,836."""
    x = np.asarray(x, dtype=dtype)
    safe = np.minimum(x, 20.0)
    return np.where(x > 20.0, x, np.log1p(np.exp(safe)))


def sigmoid(x, dtype=np.float64):
    """d softplus / dx (1 above the threshold, like torch's SoftplusBackward)."""
    x = np.asarray(x, dtype=dtype)
    return np.where(x > 20.0, 1.0, 1.0 / (1.0 + np.exp(-x)))


def rsample(loc, logscale, eps, dtype=np.float64):
    """VariationalNormal.rsample, src/ProbabilisticLayers.py:828-841: loc + eps*softplus(logscale).

    eps has a leading MC axis; loc/logscale broadcast against it.
    """
    return np.asarray(loc, dtype)[None] + np.asarray(eps, dtype) * softplus(logscale, dtype)[None]


# ----------------------------------------------------------------------------------------
# BayesLinear  (src/ProbabilisticLayers.py:888-958)
# ----------------------------------------------------------------------------------------
def linear_fwd(x, w_mu, w_rho, eps_w, b_mu=None, b_rho=None, eps_b=None, dtype=np.float64):
    """out = baddbmm(sampled_b.unsqueeze(1), x, sampled_w)  (:934) or bmm (:935).

    x [MC,B,I] (or [B,I], broadcast over MC like MC_ExpansionLayer :778-785),
    w_* [I,O], eps_w [MC,I,O], b_* [O], eps_b [MC,O]  ->  [MC,B,O]
    """
    x = np.asarray(x, dtype)
    w = rsample(w_mu, w_rho, eps_w, dtype)                      # [MC,I,O]
    if x.ndim == 2:
        x = np.broadcast_to(x[None], (w.shape[0],) + x.shape)
    out = np.matmul(x, w)                                       # per-MC GEMM (BLAS; einsum's C loop is ~60x slower)
    if b_mu is not None:
        out = out + rsample(b_mu, b_rho, eps_b, dtype)[:, None, :]
    return out


def linear_bwd(dy, x, w_mu, w_rho, eps_w, b_rho=None, eps_b=None, dtype=np.float64):
    """Closed-form autograd of rsample + baddbmm (verified against the
    reference's autograd in tests/test_oracle_golden.py).

    Returns dict(dx, dw_mu, dw_rho, db_mu, db_rho); bias entries None without a bias.
    """
    dy = np.asarray(dy, dtype)
    x = np.asarray(x, dtype)
    eps_w = np.asarray(eps_w, dtype)
    w = rsample(w_mu, w_rho, eps_w, dtype)
    if x.ndim == 2:
        x = np.broadcast_to(x[None], (w.shape[0],) + x.shape)
    dx = np.matmul(dy, w.transpose(0, 2, 1))
    dw = np.matmul(x.transpose(0, 2, 1), dy)                    # per-MC weight gradient [MC,I,O]
    out = {
        "dx": dx,
        "dw_mu": dw.sum(0),
        "dw_rho": sigmoid(w_rho, dtype) * (dw * eps_w).sum(0),
        "db_mu": None,
        "db_rho": None,
    }
    if b_rho is not None:
        dyb = dy.sum(1)                                         # [MC,O]
        out["db_mu"] = dyb.sum(0)
        out["db_rho"] = sigmoid(b_rho, dtype) * (dyb * np.asarray(eps_b, dtype)).sum(0)
    return out


# ----------------------------------------------------------------------------------------
# SIVILayer contraction (src/ProbabilisticLayers.py:307-333): per-MC mu / sigma
# ----------------------------------------------------------------------------------------
def sivi_linear_fwd(x, w_mu, w_std, eps_w, b_mu, b_std, eps_b, dtype=np.float64):
    """baddbmm(sampled_b.unsqueeze(-1), x, sampled_w) with per-MC Normal(w_mu, w_std) (:321-326).

    Only out_features == 1 is well defined in the reference:
    sampled_b [MC,1] -> unsqueeze(-1) -> [MC,1,1] broadcast over [MC,B,1].
    """
    x = np.asarray(x, dtype)
    w = np.asarray(w_mu, dtype) + np.asarray(w_std, dtype) * np.asarray(eps_w, dtype)
    b = np.asarray(b_mu, dtype) + np.asarray(b_std, dtype) * np.asarray(eps_b, dtype)
    return np.einsum("mbi,mio->mbo", x, w) + b[:, None, :]


def sivi_linear_bwd(dy, x, w_mu, w_std, eps_w, eps_b, dtype=np.float64):
    dy = np.asarray(dy, dtype)
    x = np.asarray(x, dtype)
    eps_w = np.asarray(eps_w, dtype)
    w = np.asarray(w_mu, dtype) + np.asarray(w_std, dtype) * eps_w
    dw = np.einsum("mbi,mbo->mio", x, dy)
    dyb = dy.sum(1)
    return {
        "dx": np.einsum("mbo,mio->mbi", dy, w),
        "dw_mu": dw,
        "dw_std": dw * eps_w,
        "db_mu": dyb,
        "db_std": dyb * np.asarray(eps_b, dtype),
    }


def sivi_kl(w_mu, w_std, dtype=np.float64):
    """kl(N(mu_mc, std_mc) || N(0,1)).mean(0).sum(), weights only (:330-333)."""
    w_mu = np.asarray(w_mu, dtype)
    w_std = np.asarray(w_std, dtype)
    kl = 0.5 * (w_std ** 2 + w_mu ** 2 - 1.0 - np.log(w_std ** 2))
    return kl.mean(0).sum()


# ----------------------------------------------------------------------------------------
# KL(q || p) and entropy (src/ProbabilisticLayers.py:955-956, :1059;
# torch/distributions/kl.py:468-471, torch/distributions/normal.py:114-115)
# ----------------------------------------------------------------------------------------
def kl_fwd(mu, rho, prior_scale=1.0, dtype=np.float64):
    """Returns (kl_div, entropy) summed over all weight elements.  Weights only: the
    reference never regularises the bias posterior."""
    mu = np.asarray(mu, dtype)
    sigma = softplus(rho, dtype)
    sp = np.asarray(prior_scale, dtype)
    var_ratio = (sigma / sp) ** 2
    t1 = (mu / sp) ** 2
    kl = 0.5 * (var_ratio + t1 - 1.0 - np.log(var_ratio))
    ent = 0.5 + 0.5 * LOG_2PI + np.log(sigma)
    return kl.sum(), ent.sum()


def kl_bwd(mu, rho, prior_scale=1.0, dtype=np.float64):
    """d kl_div / d(mu, rho): mu/sp^2 and (sigma/sp^2 - 1/sigma) * sigmoid(rho)."""
    mu = np.asarray(mu, dtype)
    sigma = softplus(rho, dtype)
    sp = np.asarray(prior_scale, dtype)
    return mu / sp ** 2, (sigma / sp ** 2 - 1.0 / sigma) * sigmoid(rho, dtype)


def entropy_bwd(rho, dtype=np.float64):
    """d entropy / d rho = sigmoid(rho) / softplus(rho)."""
    return sigmoid(rho, dtype) / softplus(rho, dtype)


# ----------------------------------------------------------------------------------------
# BayesConv2d (src/ProbabilisticLayers.py:1014-1061) == per-MC conv2d (SURVEY.md pin 4)
# ----------------------------------------------------------------------------------------
def _im2col(x, k, stride, padding, dilation):
    """x [N,C,H,W] -> patches [N, Ho, Wo, C, k, k]."""
    n, c, h, w = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding), (padding, padding)))
    ho = (h + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    wo = (w + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    out = np.empty((n, ho, wo, c, k, k), dtype=x.dtype)
    for kh in range(k):
        for kw in range(k):
            hs, ws = kh * dilation, kw * dilation
            out[:, :, :, :, kh, kw] = xp[:, :, hs:hs + stride * ho:stride,
                                         ws:ws + stride * wo:stride].transpose(0, 2, 3, 1)
    return out, ho, wo


def conv2d_fwd(x, w_mu, w_rho, eps_w, b_mu, b_rho, eps_b, stride=1, padding=0, dilation=1,
               dtype=np.float64):
    """x [MC,B,Cin,H,W], w_* [Cout,Cin,k,k], eps_w [MC,Cout,Cin,k,k], b_* [Cout],
    eps_b [MC,Cout] -> [MC,B,Cout,H',W'] (logical shape of the view returned at :1057)."""
    x = np.asarray(x, dtype)
    w = rsample(w_mu, w_rho, eps_w, dtype)
    b = rsample(b_mu, b_rho, eps_b, dtype)
    mc = w.shape[0]
    k = w.shape[-1]
    outs = []
    for m in range(mc):
        p, ho, wo = _im2col(x[m], k, stride, padding, dilation)     # [B,Ho,Wo,C,k,k]
        bsz = p.shape[0]
        y = np.matmul(p.reshape(bsz * ho * wo, -1), w[m].reshape(w.shape[1], -1).T)     # [B*Ho*Wo, Cout] (BLAS)
        y = y.reshape(bsz, ho, wo, -1).transpose(0, 3, 1, 2) + b[m][None, :, None, None]
        outs.append(y)
    return np.stack(outs, 0)


def conv2d_bwd(dy, x, w_mu, w_rho, eps_w, b_rho, eps_b, stride=1, padding=0, dilation=1,
               dtype=np.float64):
    dy = np.asarray(dy, dtype)
    x = np.asarray(x, dtype)
    eps_w = np.asarray(eps_w, dtype)
    w = rsample(w_mu, w_rho, eps_w, dtype)
    mc, bsz, cin, h, wd = x.shape
    k = w.shape[-1]
    dx = np.zeros_like(x)
    dw = np.zeros_like(w)
    for m in range(mc):
        p, ho, wo = _im2col(x[m], k, stride, padding, dilation)
        cout = w.shape[1]
        pm = p.reshape(bsz * ho * wo, -1)                            # [B*Ho*Wo, Cin*k*k]
        dym = dy[m].transpose(0, 2, 3, 1).reshape(bsz * ho * wo, cout)
        dw[m] = np.matmul(dym.T, pm).reshape(w[m].shape)
        dp = np.matmul(dym, w[m].reshape(cout, -1)).reshape(p.shape)  # grad wrt patches
        dxp = np.zeros((bsz, cin, h + 2 * padding, wd + 2 * padding), dtype)
        for kh in range(k):
            for kw in range(k):
                hs, ws = kh * dilation, kw * dilation
                dxp[:, :, hs:hs + stride * ho:stride, ws:ws + stride * wo:stride] += \
                    dp[:, :, :, :, kh, kw].transpose(0, 3, 1, 2)
        dx[m] = dxp[:, :, padding:padding + h, padding:padding + wd]
    dyb = dy.sum((1, 3, 4))                                          # [MC,Cout]
    return {
        "dx": dx,
        "dw_mu": dw.sum(0),
        "dw_rho": sigmoid(w_rho, dtype) * (dw * eps_w).sum(0),
        "db_mu": dyb.sum(0),
        "db_rho": sigmoid(b_rho, dtype) * (dyb * np.asarray(eps_b, dtype)).sum(0),
    }


# ----------------------------------------------------------------------------------------
# Likelihood halves of the ELBO (src/ProbabilisticLayers_Loss.py:27-120,
# duplicated at src/ProbabilisticLayers.py:22-117)
# ----------------------------------------------------------------------------------------
def mc_cross_entropy(pred, target, num_samples, dtype=np.float64):
    """F.cross_entropy(pred.flatten(0,1), target.repeat(MC), 'mean') * num_samples (:41).
    Returns (loss, dloss/dpred)."""
    pred = np.asarray(pred, dtype)
    mc, bsz, c = pred.shape
    z = pred - pred.max(-1, keepdims=True)
    lse = np.log(np.exp(z).sum(-1, keepdims=True))
    logp = z - lse
    tgt = np.asarray(target).astype(np.int64)
    picked = np.take_along_axis(logp, np.broadcast_to(tgt[None, :, None], (mc, bsz, 1)), -1)
    loss = -picked.mean() * num_samples
    onehot = np.zeros_like(pred)
    np.put_along_axis(onehot, np.broadcast_to(tgt[None, :, None], (mc, bsz, 1)), 1.0, -1)
    dpred = (np.exp(logp) - onehot) * (num_samples / (mc * bsz))
    return loss, dpred


def mc_accuracy(pred, target):
    """MC_Accuracy for 3-D predictions (src/ProbabilisticLayers.py:46-67)."""
    pred = np.asarray(pred)
    return float((pred.argmax(-1) == np.asarray(target)[None, :]).mean())


def mc_nll(pred, target, num_samples, dtype=np.float64):
    """-Normal(pred.mean(0), pred.std(0)+1e-3).log_prob(target).mean() * num_samples
    (src/ProbabilisticLayers_Loss.py:103-120; unbiased std).  Returns (loss, dloss/dpred)."""
    pred = np.asarray(pred, dtype)
    y = np.asarray(target, dtype)
    mc = pred.shape[0]
    m = pred.mean(0)
    sd = pred.std(0, ddof=1)
    s = sd + 1e-3
    nll = 0.5 * LOG_2PI + np.log(s) + (y - m) ** 2 / (2.0 * s ** 2)
    loss = nll.mean() * num_samples
    scale = num_samples / nll.size
    dl_ds = 1.0 / s - (y - m) ** 2 / s ** 3
    dl_dm = -(y - m) / s ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        ds_dp = (pred - m[None]) / ((mc - 1) * sd[None])
    dpred = scale * (dl_ds[None] * ds_dp + dl_dm[None] / mc)
    return loss, dpred


# ----------------------------------------------------------------------------------------
# Small helpers the reference's network assembly uses around the hot path
# ----------------------------------------------------------------------------------------
def mc_expand(x, num_mc):
    """MC_ExpansionLayer.forward (src/ProbabilisticLayers.py:778-785) for x.dim()==input_dim."""
    x = np.asarray(x)
    return np.broadcast_to(x[None], (num_mc,) + x.shape).copy()


def leaky_relu(x, slope=0.01):
    return np.where(x > 0, x, slope * x)


def mlp_elbo_step(x, target, params, eps, num_samples, prior_scale=1.0, act_slope=0.01,
                  loss="ce", dtype=np.float64):
    """One forward+backward of a BayesLinear MLP ELBO exactly as experiments/BNN.py:287-313 /
    experiments/RegressionBNN.py:109-120 assemble it:

        loss = LL(pred, y)/num_samples + sum_layers kl_div / num_samples

    params: list of dicts(w_mu, w_rho, b_mu, b_rho); eps: list of dicts(w [MC,I,O], b [MC,O]).
    act_slope: LeakyReLU slope between layers (0.0 -> ReLU).  Returns dict with pred, LL, KL,
    loss and per-layer gradients (likelihood + KL parts already summed, as autograd would).
    """
    mc = eps[0]["w"].shape[0]
    h = mc_expand(np.asarray(x, dtype), mc)
    acts, pre = [h], []
    for li, (p, e) in enumerate(zip(params, eps)):
        z = linear_fwd(h, p["w_mu"], p["w_rho"], e["w"], p["b_mu"], p["b_rho"], e["b"], dtype)
        pre.append(z)
        h = leaky_relu(z, act_slope) if li + 1 < len(params) else z
        acts.append(h)
    pred = h
    if loss == "ce":
        ll, dpred = mc_cross_entropy(pred, target, num_samples, dtype)
    else:
        ll, dpred = mc_nll(pred, target, num_samples, dtype)
    ll_scaled = ll / num_samples
    g = dpred / num_samples
    grads = [None] * len(params)
    kl_total = 0.0
    for li in reversed(range(len(params))):
        p, e = params[li], eps[li]
        if li + 1 < len(params):
            g = g * np.where(pre[li] > 0, 1.0, act_slope)
        b = linear_bwd(g, acts[li], p["w_mu"], p["w_rho"], e["w"], p["b_rho"], e["b"], dtype)
        kl, _ = kl_fwd(p["w_mu"], p["w_rho"], prior_scale, dtype)
        dkm, dkr = kl_bwd(p["w_mu"], p["w_rho"], prior_scale, dtype)
        kl_total += kl
        grads[li] = {
            "w_mu": b["dw_mu"] + dkm / num_samples,
            "w_rho": b["dw_rho"] + dkr / num_samples,
            "b_mu": b["db_mu"],
            "b_rho": b["db_rho"],
        }
        g = b["dx"]
    return {"pred": pred, "LL": ll_scaled, "KL": kl_total / num_samples,
            "loss": ll_scaled + kl_total / num_samples, "grads": grads}
