"""numpy restatement of the counter-based eps stream the CUDA kernels synthesise on-chip.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws eps with ``_standard_normal`` (src/ProbabilisticLayers.py:835), i.e. the
global torch generator (mt19937 on CPU, Philox on CUDA, both order dependent).  The B200
kernels never store eps, so they need a stateless stream: Philox4x32-10 (Salmon et al.,
"Parallel random numbers: as easy as 1, 2, 3", SC'11 -- the same generator ATen's CUDA
``normal_`` uses) keyed by the layer seed and indexed by *what* is being sampled:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (col // 8, row, global_mc_index, stream)         stream = 2*layer_id + is_bias
    4 x u32 -> 8 normals for columns 8*(col//8) .. +7 of ``row``: word j is one Box-Muller pair
    (columns 2j, 2j+1) with a 20-bit radius and a 12-bit angle

``row``/``col`` index the parameter viewed as a 2-D matrix: Linear weight [I,O] -> (i, o);
conv weight [Cout,Cin,k,k] -> (cout, cin*k*k + kh*k + kw); bias [O] -> (0, o).  The value of
eps(mc,row,col) therefore does not depend on tiling, on the number of GPUs or on which kernel
(forward, dX, dW) asks for it.

Known-answer pins: Random123's kat_vectors for philox4x32_10 (tests/test_oracle_golden.py).
The integer stream is compared bit-exactly with ``pl_eps_dump(raw=1)``; the normals within a
small absolute tolerance (the kernels use MUFU lg2/sin/cos approximations).
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-R.  Inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(rounds):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _bits_float(bits):
    """float32 in [1,2) from a uint32 bit pattern 0x3F8xxxxx (what the kernels build with SHF/LOP3)."""
    return np.asarray(bits, np.uint32).view(np.float32).astype(np.float64)


def box_muller(w):
    """(n_cos, n_sin) from ONE uint32 word, the exact recipe of csrc/common.cuh::box_muller:
         f = 2 - t20 in (0,1] (top 20 bits);  r = sqrt(-2 ln f)
         theta = 2*pi*t12 - 3*pi (low 12 bits, mid-bin) in (-pi, pi)
    """
    w = np.asarray(w, np.uint32)
    f = 2.0 - _bits_float(((w >> np.uint32(12)) << np.uint32(3)) | np.uint32(0x3F800000))
    t = _bits_float(((w & np.uint32(0xFFF)) << np.uint32(11)) | np.uint32(0x3F800400))
    r = np.sqrt(-2.0 * np.log(f))
    theta = 2.0 * np.pi * t - 3.0 * np.pi
    return r * np.cos(theta), r * np.sin(theta)


def raw_words(seed, stream, mc0, num_mc, rows, cols):
    """uint32 [num_mc, rows, 4*ceil(cols/8)]: the raw Philox words in column order (one call per 8 cols)."""
    groups = (cols + 7) // 8
    g = np.arange(groups, dtype=np.uint64)[None, None, :]
    r = np.arange(rows, dtype=np.uint64)[None, :, None]
    m = (np.arange(num_mc, dtype=np.uint64) + np.uint64(mc0))[:, None, None]
    g, r, m = np.broadcast_arrays(g, r, m)
    w = philox4x32_10(g, r, m, np.full(g.shape, stream, np.uint64),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, -1).reshape(num_mc, rows, groups * 4)


def eps_normal(seed, stream, mc0, num_mc, rows, cols):
    """float64 [num_mc, rows, cols] standard normals of the on-chip stream."""
    w = raw_words(seed, stream, mc0, num_mc, rows, cols)
    n_cos, n_sin = box_muller(w)                       # word j -> columns 2j, 2j+1
    out = np.stack([n_cos, n_sin], -1).reshape(num_mc, rows, -1)
    return out[:, :, :cols]


def weight_stream(layer_id):
    return 2 * int(layer_id)


def bias_stream(layer_id):
    return 2 * int(layer_id) + 1
