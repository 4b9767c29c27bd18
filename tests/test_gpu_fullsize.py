"""GPU tests at BASELINE.json's full sizes, through size-independent properties (round trip,
energy, linearity, separability) plus oracle parity on a fixed subset of signals / pencils -
the C oracle at full size would take minutes.  Device-resident (torch tensors) to stay off PCIe."""
import numpy as np
import pytest
import torch

import jwave_b200 as jw
from jwave_b200 import _lib
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _dev(cls):
    from jwave_b200.device import DeviceTransforms
    return DeviceTransforms(jw.WaveletBuilder.create(cls))


def _randn(*shape, seed=42):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.randn(*shape, dtype=torch.float64, device="cuda", generator=g)


def _need(gib):
    free, _ = torch.cuda.mem_get_info()
    if free < gib * (1 << 30):
        pytest.skip(f"needs {gib} GiB of free device memory")


def test_config2_daub4_fwt_65536_x_16384():
    """Config 2: Daubechies4 FWT, full depth, 65,536 signals x 2^14."""
    _need(40)
    batch, n, level = 65536, 1 << 14, 14
    dev = _dev("Daubechies4")
    x = _randn(batch, n)
    c = dev.transform1d(_lib.FWT, _lib.FORWARD, x, level)
    # parity on the first / last signals against the oracle (1e-12 * ||x||_inf)
    idx = list(range(8)) + list(range(batch - 8, batch))
    xs = x[idx].cpu().numpy()
    ref = co.batch_1d(co.FWT, co.FORWARD, "Daubechies4", xs, level)
    assert np.abs(c[idx].cpu().numpy() - ref).max() <= 1e-12 * np.abs(xs).max()
    # orthonormal filter: energy is conserved per signal
    ex, ec = (x * x).sum(dim=1), (c * c).sum(dim=1)
    assert float(((ex - ec).abs() / ex).max()) < 1e-9
    # forward -> reverse reconstructs to 1e-10 (the reference itself reaches 8.8e-12 here)
    r = dev.transform1d(_lib.FWT, _lib.REVERSE, c, level)
    assert float((r - x).abs().max()) <= 1e-10
    refr = co.batch_1d(co.FWT, co.REVERSE, "Daubechies4", ref, level)
    assert np.abs(r[idx].cpu().numpy() - refr).max() <= 1e-12 * np.abs(ref).max()
    del r
    # linearity: T(2x - 3y) == 2 T(x) - 3 T(y) with y = x rolled by one signal
    y = torch.roll(x, 1, dims=0)
    lhs = dev.transform1d(_lib.FWT, _lib.FORWARD, 2.0 * x - 3.0 * y, level)
    rhs = 2.0 * c - 3.0 * torch.roll(c, 1, dims=0)
    assert float((lhs - rhs).abs().max()) <= 1e-11 * float(x.abs().max())
    dev.close()


def test_config3_sym8_wpt6_4096_x_65536():
    """Config 3: Symlet8 WPT, 6 levels, 4,096 signals x 2^16."""
    _need(12)
    batch, n, level = 4096, 1 << 16, 6
    dev = _dev("Symlet8")
    x = _randn(batch, n, seed=7)
    c = dev.transform1d(_lib.WPT, _lib.FORWARD, x, level)
    idx = [0, 1, 2, batch - 2, batch - 1]
    xs = x[idx].cpu().numpy()
    ref = co.batch_1d(co.WPT, co.FORWARD, "Symlet8", xs, level)
    assert np.abs(c[idx].cpu().numpy() - ref).max() <= 1e-12 * np.abs(xs).max()
    ex, ec = (x * x).sum(dim=1), (c * c).sum(dim=1)
    assert float(((ex - ec).abs() / ex).max()) < 1e-9
    r = dev.transform1d(_lib.WPT, _lib.REVERSE, c, level)
    assert float((r - x).abs().max()) <= 1e-10
    # every one of the 64 leaf packets carries energy (natural order, no packet left empty)
    pe = (c.view(batch, 64, n // 64) ** 2).sum(dim=2)
    assert float(pe.min()) > 0.0
    dev.close()


def test_config4_daub20_fwt2d_one_8192_image():
    """Config 4 (one image of the 64): Daubechies20 2-D FWT, 13 + 13 levels, 8192 x 8192."""
    _need(6)
    n, level = 8192, 13
    dev = _dev("Daubechies20")
    img = _randn(1, n, n, seed=3)
    c = dev.transform2d(_lib.FWT, _lib.FORWARD, img, level, level)
    # separability: the 2-D transform is rows-then-columns of the 1-D transform (BasicTransform.java:361-399)
    rows = dev.axis(_lib.FWT, _lib.FORWARD, img, n, n, 1, level)
    cols = dev.axis(_lib.FWT, _lib.FORWARD, rows, 1, n, n, level)
    assert torch.equal(c, cols.view_as(c))
    # oracle parity of the row pass on a few rows, and of the column pass on a few columns
    pick = [0, 1, 4095, 8191]
    xr = img[0, pick].cpu().numpy()
    ref_rows = co.batch_1d(co.FWT, co.FORWARD, "Daubechies20", xr, level)
    assert np.abs(rows.view(n, n)[pick].cpu().numpy() - ref_rows).max() <= 1e-12 * np.abs(xr).max()
    xc = np.ascontiguousarray(rows.view(n, n)[:, pick].cpu().numpy().T)
    assert np.abs(np.ascontiguousarray(cols.view(n, n)[:, pick].cpu().numpy().T)
                  - co.batch_1d(co.FWT, co.FORWARD, "Daubechies20", xc, level)).max() <= 1e-12 * np.abs(xc).max()
    # round trip: the reference itself is borderline at 1e-10 for Daubechies20 at this size (F8)
    r = dev.transform2d(_lib.FWT, _lib.REVERSE, c, level, level)
    assert float((r - img).abs().max()) <= 1e-9
    dev.close()


def test_config5_coiflet5_fwt3d_1024_cubed():
    """Config 5 on one GPU: Coiflet5 3-D FWT of a 1024^3 volume, 10 levels per axis."""
    _need(40)
    n, level = 1024, 10
    dev = _dev("Coiflet5")
    vol = _randn(n, n, n, seed=5)
    c = dev.transform3d(_lib.FWT, _lib.FORWARD, vol, level, level, level)
    # separability: axis k, then axis j, then axis i (BasicTransform.java:509-566)
    t = dev.axis(_lib.FWT, _lib.FORWARD, vol, n * n, n, 1, level)
    t = dev.axis(_lib.FWT, _lib.FORWARD, t, n, n, n, level)
    pencils_in = t[:, [0, 17, 1023], [0, 511, 1023]].T.contiguous().cpu().numpy()   # 3 pencils along i
    t = dev.axis(_lib.FWT, _lib.FORWARD, t, 1, n, n * n, level)
    assert torch.equal(c, t)
    pencils_out = t[:, [0, 17, 1023], [0, 511, 1023]].T.contiguous().cpu().numpy()
    ref = co.batch_1d(co.FWT, co.FORWARD, "Coiflet5", pencils_in, level)
    assert np.abs(pencils_out - ref).max() <= 1e-12 * np.abs(pencils_in).max()
    del t
    # Coiflet5's table is not orthonormal to machine precision: the reference itself only
    # reconstructs to ~5e-8, so hold the GPU to that, not to 1e-10
    r = dev.transform3d(_lib.FWT, _lib.REVERSE, c, level, level, level)
    assert float((r - vol).abs().max()) <= 1e-6
    dev.close()
