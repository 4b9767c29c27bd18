"""CPU tests: pin the oracle against every known-answer vector and analytic property the
reference's own tests hold for this path, and against the independent
numpy restatement (bit for bit)."""
import numpy as np
import pytest

import reference_suite as rs
from adapters import OracleFWT, OracleWPT, rng_signal
from oracle import c_oracle as co
from oracle import np_oracle as no


def make(kind, cls):
    return OracleFWT(cls) if kind == "fwt" else OracleWPT(cls)


def test_registry_is_the_in_scope_set():
    names = co.wavelet_names()
    assert len(names) == 63 and len(set(names)) == 63  # 47 orthonormal + Haar1Orthogonal + 15 BiOrthogonal
    assert names[:47] == rs.ORTHONORMAL and set(names[47:]) == set(rs.FOUR_FILTER)
    assert {"Haar1", "Daubechies4", "Symlet8", "Daubechies20", "Coiflet5", "Legendre3"} <= set(names)
    for cls, L in (("Haar1", 2), ("Daubechies4", 8), ("Symlet8", 16), ("Coiflet5", 30), ("Daubechies20", 40)):
        assert co.wavelet(cls).contents.motherWavelength == L


def test_haar_known_answer():
    rs.check_haar_kat(make)


def test_filter_fixtures():
    rs.check_haar_filters()


def test_fixture_files_no_reference_test_reads():
    rs.check_unused_reference_fixtures(make)


@pytest.mark.parametrize("cls", rs.ORTHONORMAL)
def test_tap_identities(cls):
    """sum h = sqrt 2 (Legendre: -sqrt 2), g[i] = (-1)^i h[L-1-i], recon == decomp; and
    sum h^2 = 1 plus double-shift orthogonality wherever SURVEY.md F8 says the table has them."""
    s_de, w_de, s_re, w_re = co.wavelet(cls).contents.taps()
    L = len(s_de)
    sign = -1.0 if cls.startswith("Legendre") else 1.0
    assert abs(s_de.sum() - sign * np.sqrt(2.0)) < 1e-9
    assert np.array_equal(s_re, s_de) and np.array_equal(w_re, w_de)
    assert np.array_equal(w_de, np.array([s_de[L - 1 - i] * (1 if i % 2 == 0 else -1) for i in range(L)]))
    if cls in ("Legendre2", "Legendre3"):
        assert abs((s_de ** 2).sum() - 1.0) > 0.3  # not orthonormal (F8): forward parity only
        return
    tol = 1e-6 if cls.startswith("Coiflet") else 1e-9
    assert abs((s_de ** 2).sum() - 1.0) < tol
    for k in range(1, L // 2):
        assert abs(np.dot(s_de[2 * k:], s_de[: L - 2 * k])) < tol


@pytest.mark.parametrize("cls", rs.FOUR_FILTER)
def test_four_filter_identities(cls):
    """biorthogonal/BiOrthogonal.java:43-66 where the constructor builds the reconstruction pair; perfect
    reconstruction of one level (sum over shifts of analysis x synthesis = identity) exactly for the
    members the reference keeps in create2arr (WaveletBuilder.java:427-502) and for Haar1Orthogonal."""
    from jwave_b200._taps_bior import BIOR_TAPS
    w = co.wavelet(cls).contents
    s_de, w_de, s_re, w_re = w.taps()
    L = len(s_de)
    if cls in BIOR_TAPS and BIOR_TAPS[cls][1]:
        sign = np.where(np.arange(L) % 2 == 0, -1.0, 1.0)
        assert np.array_equal(s_re, sign * w_de) and np.array_equal(w_re, sign * s_de)
    n = 64
    eye = np.eye(n)
    back = np.stack([co.transform_1d(co.FWT, co.REVERSE, cls, co.transform_1d(co.FWT, co.FORWARD, cls, e, 1), 1)
                     for e in eye])
    err = np.abs(back - eye).max()
    if cls in rs.CREATE2ARR or cls == "Haar1Orthogonal":
        assert err < 1e-12
    else:
        assert err > 0.1  # as in the reference: these members do not reconstruct


@pytest.mark.parametrize("cls", rs.CREATE2ARR)
def test_stepping_ladder(cls):
    rs.check_stepping(make, cls)


@pytest.mark.parametrize("cls", rs.CREATE2ARR)
def test_decompose_ladder(cls):
    rs.check_decompose(make, cls)


@pytest.mark.parametrize("cls", rs.CREATE2ARR + rs.LEGENDRE)
def test_rounding(cls):
    rs.check_rounding(make, cls, fwt_iters=20, wpt_iters=6)


@pytest.mark.parametrize("cls", rs.CREATE2ARR)
def test_general_example(cls):
    rs.check_general_example(make, cls, n_random=1 << 12)


def test_sampling():
    rs.check_sampling(make)


def test_properties():
    rs.check_properties(make)


def test_error_paths():
    rs.check_error_paths(make)


def test_parallel_equals_sequential_wpt():
    """transforms/ParallelWPTTest.java:153-180 and ParallelWPTPerformanceTest.java:56-100: the
    packet-parallel driver is the same arithmetic as WaveletPacketTransform (here: bit-equal)."""
    for n, level in ((256, 4), (512, 5), (4096, 6)):
        x = rng_signal(42, 5, n)
        seq = co.transform_1d(co.WPT, co.FORWARD, "Daubechies4", x, level)
        par = co.parallel_wpt(co.FORWARD, "Daubechies4", x, level, threads=4)
        assert np.array_equal(seq, par)
        assert np.array_equal(co.transform_1d(co.WPT, co.REVERSE, "Daubechies4", seq, level),
                              co.parallel_wpt(co.REVERSE, "Daubechies4", par, level, threads=4))
    with pytest.raises(co.OracleError):  # PooledWaveletPacketTransform.java:29
        co.parallel_wpt(co.FORWARD, "Daubechies4", rng_signal(1, 2, 64), 0)


@pytest.mark.parametrize("cls", co.wavelet_names())
def test_c_oracle_equals_numpy_restatement(cls):
    """Two independent restatements of Wavelet.java:236-303 + the level loops agree bit for bit."""
    s = no.WAVELETS[cls]
    for a, b in zip(co.wavelet(cls).contents.taps(), s):
        assert np.array_equal(a, b)
    for n in (2, 4, 16, 128):
        x = rng_signal(n, n)
        for level in (None, 1, max(0, n.bit_length() - 3)):
            f = co.transform_1d(co.FWT, co.FORWARD, cls, x, level)
            assert np.array_equal(f, no.fwt_forward(cls, x, level))
            assert np.array_equal(co.transform_1d(co.FWT, co.REVERSE, cls, f, level), no.fwt_reverse(cls, f, level))
            p = co.transform_1d(co.WPT, co.FORWARD, cls, x, level)
            assert np.array_equal(p, no.wpt_forward(cls, x, level))
            assert np.array_equal(co.transform_1d(co.WPT, co.REVERSE, cls, p, level), no.wpt_reverse(cls, p, level))


@pytest.mark.parametrize("kind", ["fwt", "wpt"])
def test_2d_3d_drivers_three_ways(kind):
    """2-D / 3-D have no reference vectors.  Pin them by composition: the C
    driver, the numpy driver and the product's BasicTransform host driver (with the oracle's
    1-D transform plugged in) must agree bit for bit, including the 3-D level shift (F5)."""
    k = co.FWT if kind == "fwt" else co.WPT
    t = make(kind, "Coiflet2")
    m = rng_signal(3, 8, 16)
    for lv in ((3, 4), (1, 2), (0, 3)):
        f = co.transform_2d(k, co.FORWARD, "Coiflet2", m, *lv)
        assert np.array_equal(f, no.transform_2d(kind, "forward", "Coiflet2", m, *lv))
        assert np.array_equal(f, t.forward(m, *lv))
        r = co.transform_2d(k, co.REVERSE, "Coiflet2", f, *lv)
        assert np.array_equal(r, no.transform_2d(kind, "reverse", "Coiflet2", f, *lv))
        assert np.array_equal(r, t.reverse(f, *lv))
    assert np.array_equal(co.transform_2d(k, co.FORWARD, "Coiflet2", m), t.forward(m))
    s = rng_signal(4, 4, 8, 16)
    for lv in ((3, 4, 2), (1, 2, 1)):  # (lvlP, lvlQ, lvlR): Q=8 gets lvlP, R=16 gets lvlQ, P=4 gets lvlR
        f = co.transform_3d(k, co.FORWARD, "Coiflet2", s, *lv)
        assert np.array_equal(f, no.transform_3d(kind, "forward", "Coiflet2", s, *lv))
        assert np.array_equal(f, t.forward(s, *lv))
        r = co.transform_3d(k, co.REVERSE, "Coiflet2", f, *lv)
        assert np.array_equal(r, t.reverse(f, *lv))
    with pytest.raises(co.OracleError):  # lvlP = 4 > log2(Q = 8): the shift makes this fail
        co.transform_3d(k, co.FORWARD, "Coiflet2", s, 4, 3, 2)
    c = rng_signal(5, 8, 8, 8)
    assert np.array_equal(co.transform_3d(k, co.FORWARD, "Coiflet2", c), t.forward(c))


def test_ancient_egyptian_decomposition():
    """AncientEgyptianDecomposition.java:97-183 + MathToolKit.decompose (:57-84): no vectors in the
    reference, so the restatement is pinned by composition - the result is the wrapped transform
    applied at full depth to every 2^p block of the binary expansion, largest block first."""
    import jwave_b200 as jw
    assert co.decompose(13) == [3, 2, 0] and co.decompose(1) == [0] and co.decompose(64) == [6]
    for n in (1, 2, 3, 13, 100, 1000, 4097):
        assert co.decompose(n) == jw.AncientEgyptianDecomposition.decompose(n)
        assert sum(1 << p for p in co.decompose(n)) == n
        x = rng_signal(n, n)
        for kind in (co.FWT, co.WPT):
            f = co.aed(kind, co.FORWARD, "Daubechies4", x)
            off, parts = 0, []
            for p in co.decompose(n):
                parts.append(co.transform_1d(kind, co.FORWARD, "Daubechies4", x[off:off + (1 << p)]))
                off += 1 << p
            assert np.array_equal(f, np.concatenate(parts))
            assert np.abs(co.aed(kind, co.REVERSE, "Daubechies4", f) - x).max() < 1e-9


def test_compressor_magnitude_restatement():
    """compressions/CompressorMagnitude.java:52-68 + Compressor.java:97-110 (no vectors in the
    reference): magnitude is the mean |c|; coefficients below magnitude * threshold become zero."""
    x = np.array([1.0, -2.0, 3.0, -4.0, 0.5, -0.25, 6.0, 0.0])
    out, mag = co.compress_magnitude(x, 1.0)
    assert mag == np.abs(x).sum() / 8
    assert np.array_equal(out, np.where(np.abs(x) >= mag, x, 0.0))
    out2, _ = co.compress_magnitude(x, 0.1)
    assert np.array_equal(out2, np.where(np.abs(x) >= 0.1 * mag, x, 0.0))


def test_parallel_transform_ports_match_the_sequential_drivers():
    """ParallelTransform (ParallelTransform.java:70-134, :137-213), the CPU baseline of configs 4 and 5: the
    2-D form and the 3-D forward run the same 1-D transforms in the same axis order as BasicTransform, so they
    are bit-equal to jwo_2d / jwo_3d; the 3-D reverse runs the outer axis FIRST (:193), which only commutes
    up to rounding."""
    rng = np.random.default_rng(11)
    imgs = rng.standard_normal((3, 32, 64))
    for direction in (co.FORWARD, co.REVERSE):
        want = np.stack([co.transform_2d(co.FWT, direction, "Daubechies4", m, 5, 6) for m in imgs])
        for threads in (1, 4):
            assert np.array_equal(co.parallel_2d(co.FWT, direction, "Daubechies4", imgs, 5, 6, threads), want)
    vol = rng.standard_normal((16, 8, 32))
    for kind in (co.FWT, co.WPT):
        f = co.transform_3d(kind, co.FORWARD, "Coiflet1", vol, 3, 5, 4)   # (lvlP, lvlQ, lvlR) with the F5 shift: Q = 8 gets 3, R = 32 gets 5, P = 16 gets 4
        for threads in (1, 3):
            assert np.array_equal(co.parallel_3d(kind, co.FORWARD, "Coiflet1", vol, 3, 5, 4, threads), f)
        r = co.transform_3d(kind, co.REVERSE, "Coiflet1", f, 3, 5, 4)
        pr = co.parallel_3d(kind, co.REVERSE, "Coiflet1", f, 3, 5, 4, 4)
        assert np.abs(pr - r).max() <= 1e-12 * np.abs(f).max()
        assert np.abs(pr - vol).max() <= 1e-9
    with pytest.raises(co.OracleError):
        co.parallel_3d(co.FWT, co.FORWARD, "Haar1", np.zeros((4, 6, 8)), 2, 2, 3)
