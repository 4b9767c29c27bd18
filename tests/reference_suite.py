"""The checks JWave's own test-suite makes on the FWT / WPT path, restated once and run
against whichever implementation a `make(kind, wavelet_cls)` factory returns ("fwt" | "wpt").

Each function cites the reference test it follows (paths relative to
/root/reference/src/test/java/jwave/) and keeps that test's tolerance."""
import json
import math
import os

import numpy as np

from jwave_b200 import Transform, WaveletBuilder
from jwave_b200.wavelets import WAVELET_CLASSES

KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kat.json")))

# WaveletBuilder.create2arr() (WaveletBuilder.java:427-502): the orthonormal families without Legendre,
# and BiOrthogonal 1/1, 1/3, 1/5, 3/1 .. 3/9 (the members the reference keeps in its own test loops)
_NAMES = {WaveletBuilder.create(c).getName(): c for c in WAVELET_CLASSES}
CREATE2ARR = [_NAMES[w.getName()] for w in WaveletBuilder.create2arr()]
LEGENDRE = [c for c in WAVELET_CLASSES if c.startswith("Legendre")]
# four independent filters; the ones outside CREATE2ARR do not
# reconstruct in the reference either (WaveletBuilder.java:481-492 comments them out)
FOUR_FILTER = [c for c in WAVELET_CLASSES if c.startswith("BiOrthogonal") or c == "Haar1Orthogonal"]
ORTHONORMAL = [c for c in WAVELET_CLASSES if c not in FOUR_FILTER]


def assert_array(expected, actual, delta):
    expected = np.asarray(expected)
    actual = np.asarray(actual)
    assert expected.shape == actual.shape
    err = np.abs(expected - actual).max() if expected.size else 0.0
    assert err <= delta, f"max abs error {err} > {delta}"


def check_haar_kat(make):
    """transforms/CrossValidationTest.java:186-208 - the only non-constant known answer."""
    t = Transform(make("fwt", "Haar1"))
    out = t.forward(KAT["haar_simple_input"], 1)
    assert_array(KAT["haar_level1_approx_manual"], out[:4], 1e-10)
    assert_array(KAT["haar_level1_detail_manual"], out[4:], 1e-10)


def check_haar_filters():
    """transforms/CrossValidationTest.java:158-181 (+ the rec_* and db fixtures beside them)."""
    w = WaveletBuilder.create("Haar")
    assert_array(KAT["filter_haar_dec_lo"], w.getScalingDeComposition(), 1e-10)
    assert_array(KAT["filter_haar_dec_hi"], w.getWaveletDeComposition(), 1e-10)
    assert_array(KAT["filter_haar_rec_lo"], w.getScalingReConstruction(), 1e-10)
    # filter_haar_rec_hi.txt holds PyWavelets' sign convention (-g); JWave reuses the
    # decomposition filter (Haar1.java:66-72), so compare up to the global sign.
    assert_array(np.abs(KAT["filter_haar_rec_hi"]), np.abs(w.getWaveletReConstruction()), 1e-10)
    d2 = WaveletBuilder.create("Daubechies 2")  # PyWavelets 'db2' fixtures are named db4 (4 taps)
    assert_array(KAT["filter_db4_dec_lo"], d2.getScalingDeComposition(), 1e-10)
    assert_array(KAT["filter_db4_dec_hi"], d2.getWaveletDeComposition(), 1e-10)


def check_unused_reference_fixtures(make):
    """Fixture files that sit beside the ones above in T/../resources/testdata but that no reference test
    reads: filter_db2_dec_lo.txt (a 2-tap filter, i.e. Haar), haar_constant_input.txt and
    haar_linear_input.txt (inputs without expected outputs).  They still pin something: the 2-tap table, and
    the closed forms of a Haar level on a constant (details 0, approximations c * sqrt 2) and on a ramp
    (details -1 / sqrt 2, approximations (4 i + 1) / sqrt 2)."""
    assert_array(KAT["filter_db2_dec_lo"], WaveletBuilder.create("Haar").getScalingDeComposition(), 1e-10)
    t = Transform(make("fwt", "Haar1"))
    c = np.array(KAT["haar_constant_input"])
    out = t.forward(c, 1)
    assert_array(c[:4] * math.sqrt(2.0), out[:4], 1e-10)
    assert_array(np.zeros(4), out[4:], 1e-10)
    assert_array(c, t.reverse(out, 1), 1e-10)
    r = np.array(KAT["haar_linear_input"])
    out = t.forward(r, 1)
    assert_array((4.0 * np.arange(4) + 1.0) / math.sqrt(2.0), out[:4], 1e-10)
    assert_array(np.full(4, -1.0 / math.sqrt(2.0)), out[4:], 1e-10)
    assert_array(r, t.reverse(out, 1), 1e-10)


def ladder(n, level):
    """All-ones input: level l gives 2^(l/2) on the first n / 2^l entries, 0 elsewhere."""
    out = np.zeros(n)
    out[: n >> level] = math.sqrt(2.0) ** level
    return out


def check_stepping(make, wavelet_cls):
    """SteppingTest.java:37-315: ones(4) and ones(64), every level, FWT and WPT, delta 1e-8."""
    for kind in ("fwt", "wpt"):
        t = Transform(make(kind, wavelet_cls))
        for n in (4, 64):
            ones = np.ones(n)
            for level in range(int(math.log2(n)) + 1):
                hilb = t.forward(ones, level)
                assert_array(ladder(n, level), hilb, 1e-8)
                assert_array(ones, t.reverse(hilb, level), 1e-8)


def check_decompose(make, wavelet_cls):
    """DecomposeTest.java:30-171: decompose / recompose ladder, delta 1e-8."""
    for kind in ("fwt", "wpt"):
        t = Transform(make(kind, wavelet_cls))
        ones = np.ones(64)
        mat = t.decompose(ones)
        assert mat.shape == (7, 64)
        for level in range(7):
            assert_array(ladder(64, level), mat[level], 1e-8)
            assert_array(ones, t.recompose(mat, level), 1e-8)


def check_rounding(make, wavelet_cls, fwt_iters=50, wpt_iters=16):
    """RoundingTest.java:37-204: repeated forward/reverse on ones(1024) stays within 1e-8
    (the reference runs 1000 / 256 rounds; fewer here, the drift is linear)."""
    for kind, iters in (("fwt", fwt_iters), ("wpt", wpt_iters)):
        t = Transform(make(kind, wavelet_cls))
        arr = np.ones(1024)
        for _ in range(iters):
            arr = t.reverse(t.forward(arr))
        assert_array(np.ones(1024), arr, 1e-8)


def check_general_example(make, wavelet_cls, n_random=1 << 14):
    """GeneralTest.java:36-74: the 8-sample vector and a long random array round-trip, delta
    1e-6 ("due to a lot of wavelets with different precisions")."""
    t = Transform(make("fwt", wavelet_cls))
    x = np.array(KAT["general_test_example"])
    assert_array(x, t.reverse(t.forward(x)), 1e-6)
    r = np.random.default_rng(7).random(n_random)
    assert_array(r, t.reverse(t.forward(r)), 1e-6)


def check_sampling(make, n=1 << 16, oscillations=64):
    """SamplingTest.java:29-75: Haar FWT and WPT round-trip of a sampled sine / cosine, 1e-10."""
    phase = 2.0 * math.pi * oscillations * np.arange(n) / n
    for kind in ("fwt", "wpt"):
        t = Transform(make(kind, "Haar1"))
        for sig in (np.sin(phase), np.cos(phase)):
            assert_array(sig, t.reverse(t.forward(sig)), 1e-10)


def check_properties(make):
    """transforms/PropertyBasedTest.java:137-230, :279-310, :359-384 (seed 42, tolerance 1e-8)."""
    rng = np.random.default_rng(42)
    tol = 1e-8
    for cls in ("Haar1", "Daubechies4", "Symlet4"):
        t = Transform(make("fwt", cls))
        for _ in range(8):
            n = 1 << int(rng.integers(3, 8))
            x = rng.uniform(-10, 10, n)
            c = t.forward(x)
            e = float(np.sum(x * x))
            assert abs(e - float(np.sum(c * c))) <= tol * e          # energy conservation
            assert_array(x, t.reverse(c), tol)                          # perfect reconstruction
    haar = Transform(make("fwt", "Haar1"))
    for _ in range(8):
        n = 1 << int(rng.integers(3, 8))
        const = float(rng.uniform(-100, 100))
        c = haar.forward(np.full(n, const))
        assert_array(np.zeros(n - 1), c[1:], tol * max(1.0, abs(const)))  # details vanish
        assert abs(c[0] - const * math.sqrt(n)) <= tol * abs(const * math.sqrt(n))
        x, y = rng.uniform(-10, 10, n), rng.uniform(-10, 10, n)
        a, b = rng.uniform(-5, 5, 2)
        assert_array(a * haar.forward(x) + b * haar.forward(y), haar.forward(a * x + b * y), tol)  # linearity
        dc = float(np.sum(x)) / math.sqrt(n)
        assert abs(haar.forward(x)[0] - dc) <= tol * max(1.0, abs(dc))  # sum preservation


def check_error_paths(make):
    """transforms/ParallelWPTTest.java:127-151 and FastWaveletTransform.java:74-83: a non-2^p
    length and a level outside [0, log2 N] are JWaveFailures; the Transform facade prints them
    and returns null (Transform.java:81-90)."""
    from jwave_b200 import JWaveFailure
    import pytest
    for kind in ("fwt", "wpt"):
        bt = make(kind, "Daubechies4")
        for bad in (np.ones(100), np.ones(3)):
            with pytest.raises(JWaveFailure):
                bt.forward(bad)
            with pytest.raises(JWaveFailure):
                bt.reverse(bad)
        for level in (-1, 9):
            with pytest.raises(JWaveFailure):
                bt.forward(np.ones(256), level)
            with pytest.raises(JWaveFailure):
                bt.reverse(np.ones(256), level)
        assert Transform(bt).forward(np.ones(100)) is None
        assert Transform(bt).reverse(np.ones(256), 9) is None
        # length 1 is legal: isBinary(1) is true and there are zero levels
        assert_array([3.5], bt.forward(np.array([3.5])), 0.0)
        assert_array([3.5], bt.reverse(np.array([3.5])), 0.0)
