"""GPU parity tests: the CUDA path, called through the C ABI (host buffers) and through the
device-resident entry points, against the CPU oracle on the same seeded inputs.

Bar: max |gpu - cpu| <= 1e-12 * ||x||_inf for forward AND reverse, fp64,
FMA / reordering permitted.  Round trips are held to 1e-10 only where the reference itself
achieves it; elsewhere the GPU round trip must equal the CPU round trip."""
import numpy as np
import pytest

import jwave_b200 as jw
import reference_suite as rs
from adapters import rng_signal
from jwave_b200 import _lib
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

REL = 1e-12
ALL = list(jw.WAVELET_CLASSES)
CONFIG_WAVELETS = ["Haar1", "Daubechies4", "Symlet8", "Daubechies20", "Coiflet5"]


def make(kind, cls):
    w = jw.WaveletBuilder.create(cls)
    return jw.CudaFastWaveletTransform(w) if kind == "fwt" else jw.CudaWaveletPacketTransform(w)


def okind(kind):
    return co.FWT if kind == "fwt" else co.WPT


def close(gpu, cpu, scale):
    err = float(np.abs(gpu - cpu).max()) if gpu.size else 0.0
    assert err <= REL * max(scale, 1e-300), f"max abs err {err:.3e} > {REL * scale:.3e}"


# ---- the reference's own tests, run on the GPU ---------------------------------------------------

def test_ref_haar_known_answer():
    rs.check_haar_kat(make)
    rs.check_unused_reference_fixtures(make)


@pytest.mark.parametrize("cls", rs.CREATE2ARR)
def test_ref_stepping(cls):
    rs.check_stepping(make, cls)


@pytest.mark.parametrize("cls", CONFIG_WAVELETS + ["Coiflet1", "Symlet20"])
def test_ref_decompose(cls):
    rs.check_decompose(make, cls)


@pytest.mark.parametrize("cls", CONFIG_WAVELETS + rs.LEGENDRE)
def test_ref_rounding(cls):
    rs.check_rounding(make, cls, fwt_iters=20, wpt_iters=6)


@pytest.mark.parametrize("cls", rs.CREATE2ARR)
def test_ref_general_example(cls):
    rs.check_general_example(make, cls, n_random=1 << 16)


def test_ref_sampling():
    rs.check_sampling(make, n=1 << 20, oscillations=1024)


def test_ref_properties():
    rs.check_properties(make)


def test_ref_error_paths():
    rs.check_error_paths(make)


# ---- parity against the oracle ---------------------------------------------------------------------

@pytest.mark.parametrize("cls", ALL)
@pytest.mark.parametrize("kind", ["fwt", "wpt"])
def test_1d_parity_all_wavelets(kind, cls):
    """Every in-scope wavelet, lengths 2 .. 4096 (so every h < L wrap case occurs), full depth
    and partial levels, forward and reverse."""
    t = make(kind, cls)
    for n in (2, 4, 8, 32, 256, 4096):
        x = rng_signal(n + 11, n)
        p = n.bit_length() - 1
        for level in sorted({p, 1, p // 2}):
            cf = co.transform_1d(okind(kind), co.FORWARD, cls, x, level)
            gf = t.forward(x, level)
            close(gf, cf, np.abs(x).max())
            close(t.reverse(cf, level), co.transform_1d(okind(kind), co.REVERSE, cls, cf, level), np.abs(cf).max())


@pytest.mark.parametrize("cls", CONFIG_WAVELETS + ["BiOrthogonal68", "BiOrthogonal13"])
@pytest.mark.parametrize("kind", ["fwt", "wpt"])
def test_batch_parity(kind, cls):
    """forwardBatch / reverseBatch on ragged batch sizes and the config lengths (2^14, 2^16)."""
    t = make(kind, cls)
    for batch, n, level in ((1, 1 << 16, None), (3, 1 << 14, None), (37, 512, 5), (130, 64, 6), (5, 1 << 16, 6)):
        x = rng_signal(batch * 7 + n, batch, n)
        lv = n.bit_length() - 1 if level is None else level
        cf = co.batch_1d(okind(kind), co.FORWARD, cls, x, lv)
        close(t.forwardBatch(x, lv), cf, np.abs(x).max())
        close(t.reverseBatch(cf, lv), co.batch_1d(okind(kind), co.REVERSE, cls, cf, lv), np.abs(cf).max())
    assert t.forwardBatch(np.empty((0, 64))).shape == (0, 64)  # empty batch


def test_input_is_not_mutated_and_output_is_fresh():
    t = make("fwt", "Daubechies4")
    x = rng_signal(1, 1024)
    keep = x.copy()
    y = t.forward(x)
    assert np.array_equal(x, keep) and y is not x  # FastWaveletTransform.java:85


@pytest.mark.parametrize("kind", ["fwt", "wpt"])
@pytest.mark.parametrize("cls", ["Haar1", "Daubechies4", "Daubechies20", "Coiflet5", "BiOrthogonal37", "Haar1Orthogonal"])
def test_2d_parity(kind, cls):
    t = make(kind, cls)
    for rows, cols, lv in ((64, 64, None), (16, 128, (2, 5)), (256, 8, (8, 0)), (1, 32, (0, 5)), (128, 256, None),
                           (2048, 32, None), (1024, 16, (3, 2))):
        m = rng_signal(rows * 3 + cols, rows, cols)
        args = () if lv is None else lv
        cf = co.transform_2d(okind(kind), co.FORWARD, cls, m, *args)
        close(t.forward(m, *args), cf, np.abs(m).max())
        close(t.reverse(cf, *args), co.transform_2d(okind(kind), co.REVERSE, cls, cf, *args), np.abs(cf).max())


@pytest.mark.parametrize("cls", ["Daubechies4", "Daubechies20", "Coiflet5", "Haar1"])
def test_strided_axis_tile_and_resident(cls):
    """Columns long enough for the tiled strided kernels (rows > 512 -> tile passes + resident
    tail) and short ones (resident only), full and partial depth, via the axis primitive."""
    import torch
    from jwave_b200.device import DeviceTransforms
    dev = DeviceTransforms(jw.WaveletBuilder.create(cls))
    for outer, n, inner, level in ((2, 4096, 16, 12), (1, 2048, 40, 5), (3, 512, 8, 9), (1, 8192, 8, 13), (2, 64, 24, 6)):
        x = rng_signal(n + inner, outer, n, inner)
        def columns(direction, arr):  # 1-D transform of every (outer, inner) line along axis 1
            lines = np.ascontiguousarray(arr.transpose(0, 2, 1).reshape(-1, n))
            res = co.batch_1d(co.FWT, direction, cls, lines, level)
            return np.ascontiguousarray(res.reshape(outer, inner, n).transpose(0, 2, 1))
        ref = columns(co.FORWARD, x)
        xd = torch.from_numpy(x).cuda()
        fd = dev.axis(_lib.FWT, _lib.FORWARD, xd, outer, n, inner, level)
        close(fd.cpu().numpy(), ref, np.abs(x).max())
        back = columns(co.REVERSE, ref)
        rd = dev.axis(_lib.FWT, _lib.REVERSE, torch.from_numpy(ref).cuda(), outer, n, inner, level)
        close(rd.cpu().numpy(), back, np.abs(ref).max())
    dev.close()


@pytest.mark.parametrize("cls", ["Haar1", "Daubechies4", "Symlet8", "Daubechies20"])
def test_strided_axis_packet_transform_through_transposes(cls, monkeypatch):
    """WPT along a strided axis = batched transpose -> fused contiguous plan -> transpose (jwc_transpose.cu,
    jwc_plan.cu::wpt_transposed): any inner stride incl. ones that are not multiples of the 32-wide tile, partial and
    full depth, against the oracle AND bit for bit against the one-level kernels (JWC_TUNE wpt_transpose=0)."""
    import torch
    from jwave_b200.device import DeviceTransforms
    w = jw.WaveletBuilder.create(cls)
    monkeypatch.delenv("JWC_TUNE", raising=False)
    dev = DeviceTransforms(w)  # jwc_create reads JWC_TUNE once
    monkeypatch.setenv("JWC_TUNE", "wpt_transpose=0")
    plain = DeviceTransforms(w)
    monkeypatch.delenv("JWC_TUNE")
    for outer, n, inner, level in ((2, 256, 24, 5), (1, 64, 3, 6), (3, 32, 40, 2), (1, 1024, 17, 10), (2, 4096, 8, 6)):
        x = rng_signal(n + inner, outer, n, inner)
        def columns(direction, arr):
            lines = np.ascontiguousarray(arr.transpose(0, 2, 1).reshape(-1, n))
            res = co.batch_1d(co.WPT, direction, cls, lines, level)
            return np.ascontiguousarray(res.reshape(outer, inner, n).transpose(0, 2, 1))
        ref = columns(co.FORWARD, x)
        xd = torch.from_numpy(x).cuda()
        fd = dev.axis(_lib.WPT, _lib.FORWARD, xd, outer, n, inner, level)
        close(fd.cpu().numpy(), ref, np.abs(x).max())
        assert torch.equal(fd, plain.axis(_lib.WPT, _lib.FORWARD, xd, outer, n, inner, level))
        back = columns(co.REVERSE, ref)
        rd = dev.axis(_lib.WPT, _lib.REVERSE, torch.from_numpy(ref).cuda(), outer, n, inner, level)
        close(rd.cpu().numpy(), back, np.abs(ref).max())
    dev.close()
    plain.close()


@pytest.mark.parametrize("cls", ["Haar1", "Daubechies2", "Daubechies4", "Symlet8", "Coiflet5", "Daubechies20"])
def test_strided_axis_second_generation_shapes(cls):
    """inner % 16 == 0 takes the 16-column kernels (jwc_fwt_strided2.cu): tile passes + resident tail, lines
    that are resident from the start, lines shorter than one TMA box (declined -> first generation), partial
    depth (the last tile pass is not followed by a resident pass), several column blocks and outer slices."""
    import torch
    from jwave_b200.device import DeviceTransforms
    dev = DeviceTransforms(jw.WaveletBuilder.create(cls))
    shapes = ((2, 4096, 32, 12), (1, 1024, 48, 10), (3, 512, 16, 9), (1, 8192, 16, 13), (2, 64, 32, 6),
              (2, 32, 16, 5), (1, 16, 16, 4), (1, 2048, 64, 3), (2, 1024, 16, 1), (2, 8, 16, 3), (1, 2048, 16, 6))
    for outer, n, inner, level in shapes:
        x = rng_signal(n + inner + level, outer, n, inner)
        def columns(direction, arr):
            lines = np.ascontiguousarray(arr.transpose(0, 2, 1).reshape(-1, n))
            res = co.batch_1d(co.FWT, direction, cls, lines, level)
            return np.ascontiguousarray(res.reshape(outer, inner, n).transpose(0, 2, 1))
        ref = columns(co.FORWARD, x)
        fd = dev.axis(_lib.FWT, _lib.FORWARD, torch.from_numpy(x).cuda(), outer, n, inner, level)
        close(fd.cpu().numpy(), ref, np.abs(x).max())
        back = columns(co.REVERSE, ref)
        rd = dev.axis(_lib.FWT, _lib.REVERSE, torch.from_numpy(ref).cuda(), outer, n, inner, level)
        close(rd.cpu().numpy(), back, np.abs(ref).max())
    dev.close()


def test_2d_equals_row_by_row_composition():
    """The whole-array 2-D pass equals BasicTransform's row-by-row driver over the GPU 1-D
    transform (BasicTransform.java:361-399) - same kernels, so bit-equal up to tile effects."""
    t = make("fwt", "Symlet8")
    m = rng_signal(9, 32, 64)
    whole = t.forward(m)
    composed = jw.BasicTransform._forward2(t, m, 5, 6)
    close(whole, composed, np.abs(m).max())


def test_batched_2d_parity():
    t = make("fwt", "Daubechies4")
    mats = rng_signal(21, 5, 64, 128)
    ref = np.stack([co.transform_2d(co.FWT, co.FORWARD, "Daubechies4", m) for m in mats])
    close(t.forwardBatch2D(mats), ref, np.abs(mats).max())
    close(t.reverseBatch2D(ref), mats, 1e3 * np.abs(mats).max())


@pytest.mark.parametrize("kind", ["fwt", "wpt"])
@pytest.mark.parametrize("cls", ["Haar1", "Coiflet5", "Daubechies4"])
def test_3d_parity(kind, cls):
    t = make(kind, cls)
    for shape, lv in (((16, 16, 16), None), ((4, 8, 32), (3, 5, 2)), ((32, 4, 8), (1, 1, 1)), ((8, 8, 8), (0, 3, 0)),
                      ((64, 32, 64), (5, 6, 6)), ((1024, 8, 8), (3, 3, 10)), ((32, 32, 32), None)):
        s = rng_signal(sum(shape), *shape)
        args = () if lv is None else lv
        cf = co.transform_3d(okind(kind), co.FORWARD, cls, s, *args)
        close(t.forward(s, *args), cf, np.abs(s).max())
        close(t.reverse(cf, *args), co.transform_3d(okind(kind), co.REVERSE, cls, cf, *args), np.abs(cf).max())


def test_3d_level_shift_is_reproduced():
    """(lvlP, lvlQ, lvlR) = (4, 3, 2) on a 4 x 8 x 16 volume asks for 4 levels on
    the length-8 axis -> JWaveFailure, exactly as the reference's driver would throw."""
    t = make("fwt", "Haar1")
    with pytest.raises(jw.JWaveFailure):
        t.forward(rng_signal(2, 4, 8, 16), 4, 3, 2)


def test_independent_filters_take_the_one_level_kernels():
    """The ABI accepts four independent filters (BiOrthogonal-style).  A set whose high pass is not
    the mirrored low pass cannot use the fused kernels (they derive hi from lo) and must still be
    right - checked against the numpy restatement with the same custom taps."""
    from oracle import np_oracle as no
    rng = np.random.default_rng(3)
    L = 6
    taps = tuple(rng.standard_normal(L) for _ in range(4))
    no.WAVELETS["custom6"] = taps

    class Custom(jw.Wavelet):
        def __init__(self):
            super().__init__("custom", taps[0], taps[1])
            self._scalingReCon, self._waveletReCon = taps[2].copy(), taps[3].copy()

    for T, f, r in ((jw.CudaFastWaveletTransform, no.fwt_forward, no.fwt_reverse),
                    (jw.CudaWaveletPacketTransform, no.wpt_forward, no.wpt_reverse)):
        t = T(Custom())
        x = rng_signal(12, 512)
        for level in (9, 3):
            ref_f, ref_r = f("custom6", x, level), r("custom6", x, level)  # unnormalised taps: scale by the result
            close(t.forward(x, level), ref_f, 10 * np.abs(ref_f).max())
            close(t.reverse(x, level), ref_r, 10 * np.abs(ref_r).max())


@pytest.mark.parametrize("kind", ["fwt", "wpt"])
def test_ancient_egyptian_decomposition(kind):
    """AncientEgyptianDecomposition over the CUDA transforms: arbitrary lengths, single signals and
    batches, against the oracle's restatement (AncientEgyptianDecomposition.java:97-183)."""
    for cls in ("Haar1", "Daubechies4", "Coiflet5"):
        aed = jw.AncientEgyptianDecomposition(make(kind, cls))
        for n in (1, 3, 13, 100, 1000, 4097, 65537):
            x = rng_signal(n + 1, n)
            cf = co.aed(okind(kind), co.FORWARD, cls, x)
            close(aed.forward(x), cf, np.abs(x).max())
            close(aed.reverse(cf), co.aed(okind(kind), co.REVERSE, cls, cf), np.abs(cf).max())
        # 1234: gathered blocks; 1000 and 4100 (multiples of 4): FWT blocks run in place of the signals (line pitch n)
        for width in (1234, 1000, 4100):
            xb = rng_signal(5, 37, width)
            cb = co.aed(okind(kind), co.FORWARD, cls, xb)
            close(aed.forwardBatch(xb), cb, np.abs(xb).max())
            close(aed.reverseBatch(cb), co.aed(okind(kind), co.REVERSE, cls, cb), np.abs(cb).max())
        assert jw.Transform(aed).forward(np.ones(12)) is not None
        with pytest.raises(jw.JWaveError):
            aed.forward(np.ones(12), 2)


@pytest.mark.parametrize("kind", ["fwt", "wpt"])
@pytest.mark.parametrize("cls", ["Haar1", "Daubechies4", "Symlet8", "Daubechies20", "BiOrthogonal35"])
def test_decompose_matches_forward_at_every_level(kind, cls):
    """jwc_decompose1d (WaveletTransform.java:136-146): row p == forward(x, p) of the oracle for every p,
    single signals and batches, lengths down to 2 (h < L wraps); recompose (:166-176) from the oracle's
    rows; TransformBuilder routes its names to the GPU classes."""
    t = make(kind, cls)
    for n in (2, 8, 64, 1024):
        x = rng_signal(n * 5 + 1, n)
        p_max = n.bit_length() - 1
        want = np.stack([co.transform_1d(okind(kind), co.FORWARD, cls, x, p) for p in range(p_max + 1)])
        got = t.decompose(x)
        assert got.shape == (p_max + 1, n)
        close(got, want, np.abs(x).max())
        for p in (0, p_max // 2, p_max):
            close(t.recompose(want, p), co.transform_1d(okind(kind), co.REVERSE, cls, want[p], p), np.abs(want[p]).max())
    xb = rng_signal(77, 5, 256)
    gb = t.decomposeBatch(xb)
    assert gb.shape == (5, 9, 256)
    for b in range(5):
        for p in (0, 1, 4, 8):
            close(gb[b, p], co.transform_1d(okind(kind), co.FORWARD, cls, xb[b], p), np.abs(xb).max())
    name = "Fast Wavelet Transform" if kind == "fwt" else "Wavelet Packet Transform"
    tb = jw.TransformBuilder.create("Cuda " + name, jw.WaveletBuilder.create(cls).getName())
    assert jw.TransformBuilder.identify(tb) == name
    close(tb.forward(xb[0], 3), co.transform_1d(okind(kind), co.FORWARD, cls, xb[0], 3), np.abs(xb).max())


def test_compressor_magnitude():
    """CompressorMagnitude (compressions/CompressorMagnitude.java:52-118) on 1-D / 2-D / 3-D coefficient
    arrays against the oracle: same magnitude to rounding, same zero pattern (coefficients sitting
    exactly on the cut are excluded from the comparison - the GPU sums |c| in a different order)."""
    t = make("fwt", "Daubechies4")
    for shape, thr in (((4096,), 1.0), ((256, 512), 0.5), ((16, 32, 64), 2.0), ((1 << 20,), 1.3)):
        x = rng_signal(sum(shape), *shape)
        c = t.forward(x) if len(shape) == 1 else x  # compress real coefficients in the 1-D cases
        ref, mag = co.compress_magnitude(c, thr)
        comp = jw.CompressorMagnitude(thr)
        got = comp.compress(c)
        assert abs(comp.getMagnitude() - mag) <= 1e-13 * mag
        safe = np.abs(np.abs(c) - mag * thr) > 1e-12 * mag  # not within rounding of the cut
        assert np.array_equal(got[safe], ref[safe])
        assert got.shape == c.shape
        assert abs(comp.calcCompressionRate(got) - comp.calcCompressionRate(ref)) < 1e-3
    assert jw.CompressorMagnitude(-3.0).getThreshold() == 1.0  # Compressor.java:52-66


@pytest.mark.parametrize("kind", ["fwt", "wpt"])
def test_forward_and_compress_in_one_call(kind):
    """jwc_forward1d_compress_dev == CompressorMagnitude.compress(forward(x)): the |c| sum runs chunk by chunk behind
    the transform (several chunks here: 48 MB each), the threshold pass in place."""
    import torch
    from jwave_b200.device import DeviceTransforms
    cls, n, level, batch = "Daubechies4", 4096, (12 if kind == "fwt" else 5), 4000   # 125 MB of coefficients
    dev = DeviceTransforms(jw.WaveletBuilder.create(cls))
    x = rng_signal(123, batch, n)
    K = _lib.FWT if kind == "fwt" else _lib.WPT
    xd = torch.from_numpy(x).cuda()
    got, mag = dev.forward_compress1d(K, xd, level, 0.8)
    c = co.batch_1d(okind(kind), co.FORWARD, cls, x, level)
    ref, rmag = co.compress_magnitude(c, 0.8)
    assert abs(float(mag) - rmag) <= 1e-12 * rmag
    g = got.cpu().numpy()
    safe = np.abs(np.abs(c) - rmag * 0.8) > 1e-11 * rmag
    kept = ref != 0.0
    assert np.array_equal((g != 0.0)[safe], kept[safe])
    assert np.abs(g[safe & kept] - ref[safe & kept]).max() <= 1e-12 * np.abs(x).max()
    again, mag2 = dev.forward_compress1d(K, xd, level, 0.8)   # the CTA counter is back at zero: same result
    assert torch.equal(again, got) and float(mag2) == float(mag)
    dev.close()


def test_abi_status_codes():
    """Raw C-ABI status codes (include/jwave_cuda.h) without the Python pre-checks."""
    import ctypes as C
    ctx = jw.CudaContext.default()
    wid = ctx.register(jw.WaveletBuilder.create("Haar"))
    L = ctx._lib
    a = np.ones(100)
    b = np.empty(100)
    assert L.jwc_fwt1d(ctx.handle, wid, 0, a.ctypes.data, b.ctypes.data, 1, 100, 1) == _lib.ERR_NOT_BINARY
    assert L.jwc_fwt1d(ctx.handle, wid, 0, a.ctypes.data, b.ctypes.data, 1, 64, 7) == _lib.ERR_LEVEL
    assert L.jwc_fwt1d(ctx.handle, wid, 0, a.ctypes.data, b.ctypes.data, 1, 64, -1) == _lib.ERR_LEVEL
    assert L.jwc_fwt1d(ctx.handle, 999, 0, a.ctypes.data, b.ctypes.data, 1, 64, 1) == _lib.ERR_ARG
    assert L.jwc_fwt1d(ctx.handle, wid, 0, None, b.ctypes.data, 1, 64, 1) == _lib.ERR_ARG
    assert L.jwc_fwt2d(ctx.handle, wid, 0, a.ctypes.data, b.ctypes.data, 1, 10, 10, 1, 1) == _lib.ERR_NOT_BINARY
    assert L.jwc_fwt3d(ctx.handle, wid, 0, a.ctypes.data, b.ctypes.data, 4, 4, 4, 3, 1, 1) == _lib.ERR_LEVEL
    odd = np.ones(3)
    w = C.c_int()
    dp = C.POINTER(C.c_double)
    p = odd.ctypes.data_as(dp)
    assert L.jwc_set_wavelet(ctx.handle, 3, p, p, p, p, C.byref(w)) == _lib.ERR_ARG
    assert b"even" in L.jwc_last_error(ctx.handle)


def test_device_resident_entry_points():
    """torch tensors through the *_dev functions == host-buffer path (same kernels)."""
    import torch
    from jwave_b200.device import DeviceTransforms
    w = jw.WaveletBuilder.create("Symlet8")
    dev = DeviceTransforms(w)
    host = jw.CudaWaveletPacketTransform(w)
    x = rng_signal(5, 12, 4096)
    xd = torch.from_numpy(x).cuda()
    before = dev.launch_count()
    yd = dev.transform1d(_lib.WPT, _lib.FORWARD, xd, 6)
    assert dev.launch_count() > before
    assert np.array_equal(yd.cpu().numpy(), host.forwardBatch(x, 6))
    assert torch.equal(xd.cpu(), torch.from_numpy(x))
    back = dev.transform1d(_lib.WPT, _lib.REVERSE, yd, 6)
    close(back.cpu().numpy(), x, 1e2 * np.abs(x).max())
    with pytest.raises(jw.JWaveFailure):  # overlap is refused
        dev.transform1d(_lib.WPT, _lib.FORWARD, xd, 6, out=xd)
    v = rng_signal(6, 8, 16, 32)
    vd = torch.from_numpy(v).cuda()
    fd = dev.transform3d(_lib.FWT, _lib.FORWARD, vd, 4, 5, 3)
    close(fd.cpu().numpy(), co.transform_3d(co.FWT, co.FORWARD, "Symlet8", v, 4, 5, 3), np.abs(v).max())
    # the axis primitive: middle axis of [outer][n][inner]
    ax = dev.axis(_lib.FWT, _lib.FORWARD, vd, 8, 16, 32, 4)
    ref = np.stack([co.transform_2d(co.FWT, co.FORWARD, "Symlet8", s, 4, 0) for s in v])
    close(ax.cpu().numpy(), ref, np.abs(v).max())
    dev.close()


def test_staging_pipeline_chunks():
    """Host path with a tiny staging chunk (many chunks, both slots reused) == one chunk."""
    ctx = jw.CudaContext(0)
    t = jw.CudaFastWaveletTransform(jw.WaveletBuilder.create("Daubechies4"), context=ctx)
    x = rng_signal(8, 301, 256)
    one = t.forwardBatch(x)
    ctx.set_staging_bytes(256 * 8 * 7)  # 7 signals per chunk -> 43 chunks
    many = t.forwardBatch(x)
    assert np.array_equal(one, many)
    close(one, co.batch_1d(co.FWT, co.FORWARD, "Daubechies4", x, 8), np.abs(x).max())
    ctx.close()


def _tuned_context(monkeypatch, tune):
    """A context created under JWC_TUNE (jwc_create reads the launch-shape switches once)."""
    if tune:
        monkeypatch.setenv("JWC_TUNE", tune)
    else:
        monkeypatch.delenv("JWC_TUNE", raising=False)
    return jw.CudaContext(0)


@pytest.mark.parametrize("cls", ["Haar1", "Legendre1", "Haar1Orthogonal"])
def test_two_tap_shuffle_kernel(monkeypatch, cls):
    """k_fwt_fwd_shfl2 (jwc_shfl.cu): forward FWT of the 2-tap filters in registers and warp shuffles, every level
    count from 1 to full depth on widths that are and are not multiples of 256, against the oracle AND bit for bit
    against the shared-memory tile kernels (JWC_TUNE shfl=0) - both evaluate fma(x1, h1, x0 * h0)."""
    w = jw.WaveletBuilder.create(cls)
    on = jw.CudaFastWaveletTransform(w, context=_tuned_context(monkeypatch, ""))
    off = jw.CudaFastWaveletTransform(w, context=_tuned_context(monkeypatch, "shfl=0"))
    for n, batch in ((256, 7), (512, 5), (1024, 3), (1 << 16, 2), (128, 4)):
        x = rng_signal(n, batch, n)
        for level in sorted({1, 2, 3, 4, 7, 8, 9, n.bit_length() - 1}):
            if level > n.bit_length() - 1:
                continue
            got = on.forwardBatch(x, level)
            want = np.stack([co.transform_1d(co.FWT, co.FORWARD, cls, x[b], level) for b in range(batch)])
            close(got, want, np.abs(x).max())
            assert np.array_equal(got, off.forwardBatch(x, level)), (n, level)
            back = np.stack([co.transform_1d(co.FWT, co.REVERSE, cls, want[b], level) for b in range(batch)])
            close(on.reverseBatch(want, level), back, np.abs(want).max())


@pytest.mark.parametrize("cls", ["Symlet8", "Daubechies20", "Haar1"])
def test_packet_transform_tma_stores(monkeypatch, cls):
    """The WPT tile kernels hand finished tiles to the copy engine (cp.async.bulk.tensor stores from a 128-byte-swizzled
    shared-memory image, jwc_wpt_rev.cu / jwc_wpt_fwd.cu).  Reverse: on by default; forward: JWC_TUNE
    wpt_tma_store_fwd=1.  Both must be bit-identical to the register stores (wpt_tma_store=0) and match the oracle, on
    widths with one and with several tiles per line and passes of 3, 2 and 1 levels; wpt_rev_m=6 runs the reverse as
    ONE pass of 6 levels (several tail steps per lane, short-packet staging)."""
    w = jw.WaveletBuilder.create(cls)
    ctxs = {t: jw.CudaWaveletPacketTransform(w, context=_tuned_context(monkeypatch, t))
            for t in ("", "wpt_tma_store=0", "wpt_tma_store_fwd=1", "wpt_rev_m=6")}
    for n, batch, level in ((1 << 16, 3, 6), (1 << 13, 5, 5), (2048, 9, 3), (4096, 2, 1), (1 << 14, 2, 14)):
        x = rng_signal(n + level, batch, n)
        want = np.stack([co.transform_1d(co.WPT, co.FORWARD, cls, x[b], level) for b in range(batch)])
        back = np.stack([co.transform_1d(co.WPT, co.REVERSE, cls, want[b], level) for b in range(batch)])
        f0 = ctxs["wpt_tma_store=0"].forwardBatch(x, level)
        r0 = ctxs["wpt_tma_store=0"].reverseBatch(want, level)
        close(f0, want, np.abs(x).max())
        close(r0, back, np.abs(want).max())
        for t in ("", "wpt_tma_store_fwd=1"):
            assert np.array_equal(ctxs[t].forwardBatch(x, level), f0), (t, n, level)
            assert np.array_equal(ctxs[t].reverseBatch(want, level), r0), (t, n, level)
        close(ctxs["wpt_rev_m=6"].reverseBatch(want, level), back, np.abs(want).max())
