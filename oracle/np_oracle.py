"""Independent numpy restatement of JWave's FWT / WPT path - TEST INFRASTRUCTURE ONLY.

Second opinion for oracle/jw_oracle.c: written separately, with its
own tap derivation (the analytic families are re-derived here; the literal families are parsed
from the text of oracle/jw_taps_literal.inc), and with the Java operation order kept - separate
multiply and add, `j` ascending for the forward sum, `i`-outer scatter for the reverse - so the
two oracles are expected to agree BIT FOR BIT.

Citations are relative to /root/reference/src/main/java/jwave/.
"""
import math
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _build(s_de):
    """transforms/wavelets/Wavelet.java:104-122"""
    s_de = np.asarray(s_de, dtype=np.float64)
    L = len(s_de)
    w_de = np.array([s_de[L - 1 - i] if i % 2 == 0 else -s_de[L - 1 - i] for i in range(L)])
    return s_de, w_de, s_de.copy(), w_de.copy()


def _analytic():
    out = {}
    r2 = math.sqrt(2.0)
    # haar/Haar1.java:52-68
    s = np.array([1.0 / r2, 1.0 / r2])
    w = np.array([s[1], -s[0]])
    out["Haar1"] = (s, w, s.copy(), w.copy())
    # daubechies/Daubechies2.java:53-63
    r3 = math.sqrt(3.0)
    out["Daubechies2"] = _build([((1.0 + r3) / 4.0) / r2, ((3.0 + r3) / 4.0) / r2,
                                 ((3.0 - r3) / 4.0) / r2, ((1.0 - r3) / 4.0) / r2])
    # daubechies/Daubechies3.java:54-66
    r10 = math.sqrt(10.0)
    cA = math.sqrt(5.0 + 2.0 * r10)
    d3 = [(1.0 + 1.0 * r10 + 1.0 * cA) / 16.0, (5.0 + 1.0 * r10 + 3.0 * cA) / 16.0,
          (10.0 - 2.0 * r10 + 2.0 * cA) / 16.0, (10.0 - 2.0 * r10 - 2.0 * cA) / 16.0,
          (5.0 + 1.0 * r10 - 3.0 * cA) / 16.0, (1.0 + 1.0 * r10 - 1.0 * cA) / 16.0]
    out["Daubechies3"] = _build([v / r2 for v in d3])
    # coiflet/Coiflet1.java:52-62
    q2 = 1.4142135623730951
    r15 = math.sqrt(15.0)
    out["Coiflet1"] = _build([q2 * (r15 - 3.0) / 32.0, q2 * (1.0 - r15) / 32.0, q2 * (6.0 - 2 * r15) / 32.0,
                              q2 * (2.0 * r15 + 6.0) / 32.0, q2 * (r15 + 13.0) / 32.0, q2 * (9.0 - r15) / 32.0])
    # legendre/Legendre{1,2,3}.java
    out["Legendre1"] = _build([-1.0 / r2, -1.0 / r2])
    out["Legendre2"] = _build([(-5.0 / 8.0) / r2, (-3.0 / 8.0) / r2, (-3.0 / 8.0) / r2, (-5.0 / 8.0) / r2])
    out["Legendre3"] = _build([(v / 128.0) / r2 for v in (-63.0, -35.0, -30.0, -30.0, -35.0, -63.0)])
    return out


def _literal():
    txt = open(os.path.join(_HERE, "jw_taps_literal.inc")).read()
    out = {}
    for m in re.finditer(r'JW_LITERAL_WAVELET\("(\w+)",\s*"([^"]+)",\s*(\d+),([^)]*)\)', txt):
        taps = [float(t) for t in m.group(4).replace("\n", " ").split(",")]
        assert len(taps) == int(m.group(3))
        out[m.group(1)] = _build(taps)
    return out


def _bior():
    """transforms/wavelets/biorthogonal/BiOrthogonal{11..68}.java (+ BiOrthogonal.java:43-66 where the
    constructor builds the reconstruction filters) and haar/Haar1Orthogonal.java:137-161."""
    txt = open(os.path.join(_HERE, "jw_taps_bior.inc")).read()
    out = {}
    for m in re.finditer(r'JW_BIOR_WAVELET\("(\w+)",\s*"([^"]+)",\s*(\d+),\s*(\d),([^)]*)\)', txt):
        L, built = int(m.group(3)), int(m.group(4))
        v = np.array([float(t) for t in m.group(5).replace("\n", " ").split(",")])
        assert len(v) == (2 if built else 4) * L
        s_de, w_de = v[:L], v[L:2 * L]
        if built:
            sign = np.where(np.arange(L) % 2 == 0, -1.0, 1.0)
            s_re, w_re = sign * w_de, sign * s_de
        else:
            s_re, w_re = v[2 * L:3 * L], v[3 * L:]
        out[m.group(1)] = (s_de, w_de, s_re, w_re)
    one = np.array([1.0, 1.0])
    out["Haar1Orthogonal"] = (one, np.array([1.0, -1.0]), one.copy(), np.array([1.0, -1.0]))
    return out


WAVELETS = {}
WAVELETS.update(_analytic())
WAVELETS.update(_literal())
WAVELETS.update(_bior())
# haar/Haar1Orthogonal.java:39, :197-199: every reconstruction term times _energyCorrectionFactor
RECON_FACTOR = {"Haar1Orthogonal": 0.5}


def wavelet_forward(name, x, n):
    """transforms/wavelets/Wavelet.java:236-260, vectorised over i, j ascending."""
    s_de, w_de, _, _ = WAVELETS[name]
    h = n >> 1
    i2 = 2 * np.arange(h)
    a = np.zeros(h)
    d = np.zeros(h)
    for j in range(len(s_de)):
        v = x[(i2 + j) % n]
        a = a + v * s_de[j]
        d = d + v * w_de[j]
    return np.concatenate([a, d])


def wavelet_reverse(name, c, n):
    """transforms/wavelets/Wavelet.java:277-303: scatter-add in (i outer, j inner) order."""
    _, _, s_re, w_re = WAVELETS[name]
    L = len(s_re)
    h = n >> 1
    i = np.repeat(np.arange(h), L)
    j = np.tile(np.arange(L), h)
    k = (2 * i + j) % n
    vals = (c[i] * s_re[j]) + (c[i + h] * w_re[j])
    if name in RECON_FACTOR:
        vals = RECON_FACTOR[name] * vals
    t = np.zeros(n)
    np.add.at(t, k, vals)  # unbuffered, applied in index order == the Java loop order
    return t


def _levels(n, level):
    if n <= 0 or n & (n - 1):
        raise ValueError("not 2^p")
    p = n.bit_length() - 1
    if level is None:
        level = p
    if level < 0 or level > p:
        raise ValueError("level out of range")
    return p, level


def fwt_forward(name, x, level=None):
    """transforms/FastWaveletTransform.java:71-101"""
    x = np.array(x, dtype=np.float64)
    n = len(x)
    _, level = _levels(n, level)
    h, l = n, 0
    while h >= 2 and l < level:
        x[:h] = wavelet_forward(name, x, h)
        h >>= 1
        l += 1
    return x


def fwt_reverse(name, c, level=None):
    """transforms/FastWaveletTransform.java:119-153"""
    c = np.array(c, dtype=np.float64)
    n = len(c)
    p, level = _levels(n, level)
    h = 2 << (p - level)
    while 2 <= h <= n:
        c[:h] = wavelet_reverse(name, c, h)
        h <<= 1
    return c


def wpt_forward(name, x, level=None):
    """transforms/WaveletPacketTransform.java:73-124"""
    x = np.array(x, dtype=np.float64)
    n = len(x)
    _, level = _levels(n, level)
    h, l = n, 0
    while h >= 2 and l < level:
        for p in range(n // h):
            x[p * h:(p + 1) * h] = wavelet_forward(name, x[p * h:(p + 1) * h].copy(), h)
        h >>= 1
        l += 1
    return x


def wpt_reverse(name, c, level=None):
    """transforms/WaveletPacketTransform.java:141-191"""
    c = np.array(c, dtype=np.float64)
    n = len(c)
    p, level = _levels(n, level)
    h = 2 << (p - level)
    while 2 <= h <= n:
        for q in range(n // h):
            c[q * h:(q + 1) * h] = wavelet_reverse(name, c[q * h:(q + 1) * h].copy(), h)
        h <<= 1
    return c


_ONE_D = {("fwt", "forward"): fwt_forward, ("fwt", "reverse"): fwt_reverse,
          ("wpt", "forward"): wpt_forward, ("wpt", "reverse"): wpt_reverse}


def transform_2d(kind, direction, name, m, lvlM, lvlN):
    """transforms/BasicTransform.java:361-399 / :436-474"""
    f = _ONE_D[(kind, direction)]
    m = np.array(m, dtype=np.float64)
    rows, cols = m.shape

    def do_rows():
        for i in range(rows):
            m[i, :] = f(name, m[i, :], lvlN)

    def do_cols():
        for j in range(cols):
            m[:, j] = f(name, m[:, j], lvlM)

    if direction == "forward":
        do_rows()
        do_cols()
    else:
        do_cols()
        do_rows()
    return m


def transform_3d(kind, direction, name, s, lvlP, lvlQ, lvlR):
    """transforms/BasicTransform.java:509-566 / :602-659 (2-D slices get (lvlP, lvlQ): F5)"""
    f = _ONE_D[(kind, direction)]
    s = np.array(s, dtype=np.float64)
    P, Q, R = s.shape
    for i in range(P):
        s[i] = transform_2d(kind, direction, name, s[i], lvlP, lvlQ)
    for j in range(Q):
        for k in range(R):
            s[:, j, k] = f(name, s[:, j, k], lvlR)
    return s
