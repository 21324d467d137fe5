"""Statistical gate for the eps stream (the 8-normals-per-Philox-call contract, 20-bit radius / 12-bit
angle, is an explicit quality trade and must be gated by tests).  Shared by the CPU test on oracle/philox.py and the GPU
test on pl_eps_dump.  All thresholds are > 4.5 sigma of chance, so a correct stream fails with probability < 1e-4 overall."""
import numpy as np
from scipy import stats


def check_normal_stream(eps, label=""):
    """eps: float array [MC, rows, cols], cols % 8 == 0.  Raises AssertionError with the failing statistic."""
    eps = np.asarray(eps, np.float64)
    mc, rows, cols = eps.shape
    flat = eps.reshape(-1)
    n = flat.size
    z = 4.5                                            # sigmas of chance allowed per statistic

    # moments
    assert abs(flat.mean()) < z / np.sqrt(n), (label, "mean", flat.mean())
    assert abs(flat.var() - 1.0) < z * np.sqrt(2.0 / n), (label, "var", flat.var())
    assert abs((flat ** 3).mean()) < z * np.sqrt(15.0 / n), (label, "skew", (flat ** 3).mean())
    assert abs((flat ** 4).mean() - 3.0) < z * np.sqrt(96.0 / n), (label, "kurtosis", (flat ** 4).mean())

    # Kolmogorov-Smirnov against N(0,1) on the whole sample
    ks = stats.kstest(flat, "norm")
    assert ks.pvalue > 1e-5, (label, "KS", ks.statistic, ks.pvalue)

    # chi-square on |z| bins out to the tail (last bin [5, inf): the 20-bit radius tops out at 5.27)
    edges = np.array([0, 0.5, 1, 1.5, 2, 2.5, 3, 3.5, 4, 4.5, 5, np.inf])
    obs, _ = np.histogram(np.abs(flat), bins=edges)
    p = np.diff(2 * stats.norm.cdf(edges) - 1)
    exp = p * n
    chi2 = float(((obs - exp) ** 2 / exp).sum())
    dof = len(exp) - 1
    assert chi2 < dof + z * np.sqrt(2 * dof) + 10, (label, "chi2 tails", chi2, obs.tolist(), exp.tolist())
    assert np.abs(flat).max() < 5.3

    def corr(a, b):
        a, b = a.reshape(-1), b.reshape(-1)
        return float(((a - a.mean()) * (b - b.mean())).mean() / (a.std() * b.std()))

    # within a Box-Muller pair (columns 2j, 2j+1 share radius and angle): values and squares uncorrelated
    a, b = eps[:, :, 0::2], eps[:, :, 1::2]
    m = a.size
    assert abs(corr(a, b)) < z / np.sqrt(m), (label, "pair corr", corr(a, b))
    assert abs(corr(a ** 2, b ** 2)) < z / np.sqrt(m), (label, "pair corr of squares", corr(a ** 2, b ** 2))
    # across the 8 normals of one Philox call (4 words): adjacent words, and first vs last
    g = eps.reshape(mc, rows, cols // 8, 8)
    for i, j in ((0, 2), (1, 3), (2, 4), (0, 7), (3, 6)):
        c = corr(g[..., i], g[..., j])
        assert abs(c) < z / np.sqrt(g[..., i].size), (label, f"group corr {i},{j}", c)
        c2 = corr(g[..., i] ** 2, g[..., j] ** 2)
        assert abs(c2) < z / np.sqrt(g[..., i].size), (label, f"group corr of squares {i},{j}", c2)
    # neighbouring calls: next column group, next row, next MC sample (counter words 0, 1, 2)
    if cols >= 16:
        c = corr(eps[:, :, :-8], eps[:, :, 8:])
        assert abs(c) < z / np.sqrt(eps[:, :, 8:].size), (label, "column-group corr", c)
    if rows >= 2:
        c = corr(eps[:, :-1], eps[:, 1:])
        assert abs(c) < z / np.sqrt(eps[:, 1:].size), (label, "row corr", c)
    if mc >= 2:
        c = corr(eps[:-1], eps[1:])
        assert abs(c) < z / np.sqrt(eps[1:].size), (label, "MC corr", c)
        # per-element mean over MC behaves like N(0, 1/MC): variance of the MC-mean
        vm = eps.mean(0).var() * mc
        assert abs(vm - 1.0) < z * np.sqrt(2.0 / (rows * cols)), (label, "variance of the MC mean", vm)
