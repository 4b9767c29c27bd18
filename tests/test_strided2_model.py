"""CPU model of the second-generation strided-axis kernels (jwave_b200/csrc/jwc_fwt_strided2.cu).

The kernels cannot run here (no GPU), but everything that can go wrong in them short of a typo is index
arithmetic: which rows a tile stages and where the periodic wrap falls, the halo counts per level, the
in-place rounds (a round's results overwrite rows that later rounds must not read any more), the periodic
extension rows of resident mode, the reverse kernel's left extensions F_k, the staged slot ranges O_k and the
placement of d_{k-1} behind the rows a_{k-1} will be written to.  This file restates those formulas exactly as
the launchers / kernels compute them, runs them on one column with numpy, and compares with the oracle's
FastWaveletTransform (np_oracle) - forward, reverse, tile and resident mode, every even filter length class
that matters (2, 4, 8, 16, 30, 40) and several (h, T, m).  It also asserts the in-place safety conditions
directly, so a change of the round structure that breaks them fails here rather than on the GPU box."""
import numpy as np
import pytest

from oracle import np_oracle as npo

R = 4            # kR2
BOXF, BOXR = 16, 8
MAXTHR = 320     # kMaxThr2


def round_up(v, q):
    return (v + q - 1) // q * q


def taps_of(L):
    name = {2: "Haar1", 4: "Daubechies2", 8: "Daubechies4", 16: "Symlet8", 30: "Coiflet5", 40: "Daubechies20"}[L]
    s_de, w_de, s_re, w_re = npo.WAVELETS[name]
    return name, np.asarray(s_de), np.asarray(w_de), np.asarray(s_re), np.asarray(w_re)


def fwd_task(X, g, lo_t, hi_t, L, with_hi=True):
    """fwd_run2: outputs R g .. R g + R - 1 from rows 2 R g + s"""
    lo, hi = np.zeros(R), np.zeros(R)
    for s in range(2 * R + L - 2):
        v = X[2 * R * g + s]
        for r in range(R):
            j = s - 2 * r
            if 0 <= j < L:
                lo[r] += v * lo_t[j]
                if with_hi:
                    hi[r] += v * hi_t[j]
    return lo, hi


def model_forward_pass(src, h, T, m, L, lo_t, hi_t, resident, ngrp):
    """One launch of k_fwt_fwd_str2 on one column: returns (d rows dict level -> array, a_m)."""
    tiles = 1 if resident else h // T
    outD = {k: np.full(h >> k, np.nan) for k in range(1, m + 1)}
    outA = np.full(h >> m, np.nan)
    for tile in range(tiles):
        Tt = h if resident else T
        n0 = h + (L - 2) if resident else T + ((1 << m) - 1) * (L - 2)
        groups1 = max(1, (h // 2) // R) if resident else ((T >> 1) + ((1 << (m - 1)) - 1) * (L - 2) + R - 1) // R
        rows0 = round_up(h + L - 2, BOXF) if resident else round_up(max(n0, 2 * R * groups1 + L - 2), BOXF)
        X = np.full(rows0, np.nan)
        boxes = (n0 + BOXF - 1) // BOXF
        for bx in range(boxes):  # TMA boxes wrap at box granularity
            s = (tile * Tt + bx * BOXF) & (h - 1)
            X[bx * BOXF:(bx + 1) * BOXF] = src[s:s + BOXF]
        for k in range(1, m + 1):
            last = k == m
            h_in = h >> (k - 1)
            n_det = (h_in >> 1) if resident else (Tt >> k)
            rowD = 0 if resident else tile * n_det
            if (not resident) or n_det >= R:
                gkeep = n_det // R
                groups = gkeep if resident else (n_det + ((1 << (m - k)) - 1) * (L - 2) + R - 1) // R
                for g0 in range(0, groups, ngrp):
                    res = {}
                    for g in range(g0, min(g0 + ngrp, groups)):
                        assert 2 * R * g + 2 * R + L - 3 < rows0
                        lo, hi = fwd_task(X, g, lo_t, hi_t, L, with_hi=g < gkeep)
                        if g < gkeep:
                            outD[k][rowD + R * g:rowD + R * g + R] = hi
                            if last:
                                outA[rowD + R * g:rowD + R * g + R] = lo
                        res[g] = lo
                    if last:
                        continue
                    # barrier; in-place write.  Safety: later rounds read rows >= 2 R (g0 + ngrp) only
                    hi_written = R * min(g0 + ngrp, groups)
                    if g0 + ngrp < groups:
                        assert hi_written <= 2 * R * (g0 + ngrp)
                    for g, lo in res.items():
                        X[R * g:R * g + R] = lo
            else:
                mask = h_in - 1
                res = {}
                for i in range(n_det):
                    lo = sum(X[(2 * i + j) & mask] * lo_t[j] for j in range(L))
                    hi = sum(X[(2 * i + j) & mask] * hi_t[j] for j in range(L))
                    outD[k][i] = hi
                    if last:
                        outA[i] = lo
                    res[i] = lo
                if not last:
                    for i, lo in res.items():
                        X[i] = lo
            if last:
                break
            if resident and n_det >= 2 * R:
                for row in range(L - 2):
                    assert n_det + row < rows0
                    X[n_det + row] = X[row & (n_det - 1)]
    return outD, outA


def rev_task(A, D, top, lo_t, hi_t, L):
    """rev_run2: slots top - R + 1 .. top -> 2 R outputs"""
    t = np.zeros(2 * R)
    for s in range(R + L // 2 - 1):
        av, dv = A(top - s), D(top - s)
        for pp in range(R):
            q = s - (R - 1 - pp)
            if 0 <= q < L // 2:
                t[2 * pp] += av * lo_t[2 * q] + dv * hi_t[2 * q]
                t[2 * pp + 1] += av * lo_t[2 * q + 1] + dv * hi_t[2 * q + 1]
    return t


def rev_geometry(L, T, m):
    """launch_rev2_L, tile mode"""
    ru = round_up(L // 2 - 1, R)
    F = [0] * (m + 2)
    N = 0
    for k in range(1, m + 1):
        F[k] = round_up((N + 1) // 2, R)
        N = F[k] + L // 2 - 1
    if (F[m] + ru) % BOXR:
        ru += R
    length, s0 = [0] * (m + 1), [0] * (m + 1)
    gmax = 0
    for k in range(1, m + 1):
        length[k] = (T >> k) + F[k] + ru if k == m else (T >> k) + 2 * F[k + 1]
        s0[k] = ru if k == m else 2 * F[k + 1] - F[k]
        assert length[k] % BOXR == 0
        if k >= 2:
            gmax = max(gmax, ((T >> k) + F[k]) // R)
    offD = [0] * (m + 1)
    offD[m] = length[m]
    end = 2 * length[m]
    for k in range(m - 1, 0, -1):
        offD[k] = max(end, length[k])
        end = offD[k] + length[k]
    nmain = max(128, min(256, T // 2 // 128 * 128))
    nthr = max(nmain, round_up(gmax * 8, 32))
    return ru, F, length, s0, offD, end, nthr


def model_reverse_pass(coef, a_m, h0, T, m, L, lo_t, hi_t, resident):
    """One launch of k_fwt_rev_str2 on one column.  coef: line with d_k at rows [h0 >> k, 2 (h0 >> k));
    a_m: the coarsest approximation (width h0 >> m).  Returns a_0 (width h0)."""
    out = np.full(h0, np.nan)
    if resident:
        X = np.array(coef[:h0], dtype=np.float64)
        X[:h0 >> m] = a_m
        ngrp = max(64, round_up(max(1, (h0 // 4) // R) * 8, 32)) // 8
        for k in range(m, 0, -1):
            half = h0 >> k
            mask = half - 1
            last = k == 1
            if half >= R:
                groups = half // R
                if not last:
                    assert groups <= ngrp
                res = {}
                for g in range(groups):
                    top = R * g + R - 1
                    res[g] = rev_task(lambda i: X[i & mask], lambda i: X[half + (i & mask)], top, lo_t, hi_t, L)
                for g, t in res.items():
                    if last:
                        out[2 * R * g:2 * R * g + 2 * R] = t
                    else:
                        X[2 * R * g:2 * R * g + 2 * R] = t
            else:
                res = {}
                for p in range(half):
                    t0 = t1 = 0.0
                    for q in range(L // 2):
                        i = (p - q) & mask
                        t0 += X[i] * lo_t[2 * q] + X[half + i] * hi_t[2 * q]
                        t1 += X[i] * lo_t[2 * q + 1] + X[half + i] * hi_t[2 * q + 1]
                    res[p] = (t0, t1)
                for p, (t0, t1) in res.items():
                    if last:
                        out[2 * p], out[2 * p + 1] = t0, t1
                    else:
                        X[2 * p], X[2 * p + 1] = t0, t1
        return out
    ru, F, length, s0, offD, rows, nthr = rev_geometry(L, T, m)
    assert nthr <= MAXTHR
    ngrp = nthr // 8
    for tile in range(h0 // T):
        t0 = tile * T
        X = np.full(rows, np.nan)
        for k in range(m, 0, -1):
            wk = h0 >> k
            O = ((t0 >> k) - F[k] - ru) if k == m else 2 * ((t0 >> (k + 1)) - F[k + 1])
            assert O % BOXR == 0 and wk % BOXR == 0
            for j in range(length[k] // BOXR):
                slot = (O + j * BOXR) & (wk - 1)
                if k == m:
                    X[j * BOXR:(j + 1) * BOXR] = a_m[slot:slot + BOXR]
                X[offD[k] + j * BOXR:offD[k] + (j + 1) * BOXR] = coef[wk + slot:wk + slot + BOXR]
        for k in range(m, 0, -1):
            groups = ((T >> k) + F[k]) // R
            if k > 1:
                assert groups <= ngrp
                assert 2 * R * groups == length[k - 1] and offD[k - 1] >= length[k - 1]
            res = {}
            for g in range(groups):
                top = s0[k] + R * g + R - 1
                assert top - (R + L // 2 - 2) >= 0 and top < length[k]
                res[g] = rev_task(lambda i: X[i], lambda i: X[offD[k] + i], top, lo_t, hi_t, L)
            for g, t in res.items():
                if k > 1:
                    X[2 * R * g:2 * R * g + 2 * R] = t
                else:
                    out[t0 + 2 * R * g:t0 + 2 * R * g + 2 * R] = t
    return out


def fwd_tile_threads(T):
    """launch_fwd2_L: whole multiples of 4 warps"""
    return max(128, min(256, T // 2 // 128 * 128))


def plan_levels(L, T, want=0):
    m = 1
    while ((m < want) if want > 0 else (((1 << (m + 1)) - 1) * (L - 2) <= T // 4)) and (T >> (m + 1)) >= BOXF:
        m += 1
    return m


CASES = [  # (L, h, T, cap)
    (2, 256, 64, 32), (4, 512, 128, 64), (8, 1024, 128, 64), (8, 512, 512, 512), (16, 1024, 256, 64),
    (30, 1024, 512, 256), (40, 2048, 512, 512), (40, 512, 512, 512), (30, 64, 512, 512), (16, 16, 512, 512),
]


@pytest.mark.parametrize("L,h,T,cap", CASES)
def test_forward_model_matches_oracle(L, h, T, cap):
    name, s_de, w_de, _, _ = taps_of(L)
    x = np.random.default_rng(L * 1000 + h).standard_normal(h)
    level = int(np.log2(h))
    want = npo.fwt_forward(name, x, level)
    got = np.full(h, np.nan)
    src, width, left = x, h, level
    while left > 0:
        resident = width <= cap or width < T
        m = left if resident else min(left, plan_levels(L, T))
        groups1 = max(1, (width // 2) // R)
        nthr = min(256, max(64, round_up((groups1 + 1) // 2 * 8, 32))) if resident else fwd_tile_threads(T)
        outD, outA = model_forward_pass(src, width, T, m, L, s_de, w_de, resident, nthr // 8)
        for k, d in outD.items():
            got[width >> k:2 * (width >> k)] = d
        if m == left:
            got[:width >> m] = outA
        src, width, left = outA, width >> m, left - m
    assert not np.isnan(got).any()
    assert np.abs(got - want).max() <= 1e-12 * np.abs(x).max()


@pytest.mark.parametrize("L,h,T,cap", CASES)
def test_reverse_model_matches_oracle(L, h, T, cap):
    name, _, _, s_re, w_re = taps_of(L)
    c = np.random.default_rng(L * 77 + h).standard_normal(h)
    level = int(np.log2(h))
    want = npo.fwt_reverse(name, c, level)
    rev_m = 2 if L >= 20 else (3 if L >= 12 else 5)
    while (T >> rev_m) < 8:
        rev_m -= 1
    widths, wv = [], h
    while wv > 1:
        widths.append(wv)
        if wv <= cap or wv < T or (wv >> rev_m) <= 1:
            break
        wv >>= rev_m
    cur, a = 1, c[:1].copy()
    for h0 in reversed(widths):
        resident = h0 <= cap or h0 < T
        m = int(np.log2(h0 // cur))
        a = model_reverse_pass(c, a, h0, T, m, L, s_re, w_re, resident)
        assert not np.isnan(a).any()
        cur = h0
    assert np.abs(a - want).max() <= 1e-12 * np.abs(c).max()


def test_launch_shapes_fit_the_cta_bound():
    """Default plan (T = 512): CTA sizes stay within kMaxThr2 for every filter length."""
    for L in range(2, 42, 2):
        assert fwd_tile_threads(512) % 128 == 0 and fwd_tile_threads(512) <= MAXTHR
        rev_m = 2 if L >= 20 else (3 if L >= 12 else 5)
        assert rev_geometry(L, 512, rev_m)[-1] <= MAXTHR
