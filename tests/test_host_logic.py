"""CPU tests of the host-side mirror: tap tables, MathToolKit, exceptions, facade policy."""
import numpy as np
import pytest

import jwave_b200 as jw
from jwave_b200.wavelets import WAVELET_CLASSES
from oracle import c_oracle as co


def test_wavelet_tables_match_oracle_bit_for_bit():
    for cls in WAVELET_CLASSES:
        w = jw.WaveletBuilder.create(cls)
        ow = co.wavelet(cls).contents
        assert w.getName() == ow.name.decode()
        assert w.getMotherWavelength() == ow.motherWavelength
        assert w.getTransformWavelength() == 2
        theirs = list(ow.taps())
        # Haar1Orthogonal: the host mirror folds the reverse step's factor into the reconstruction filters
        theirs[2], theirs[3] = ow.reconFactor * theirs[2], ow.reconFactor * theirs[3]
        for mine, ref in zip((w.getScalingDeComposition(), w.getWaveletDeComposition(),
                              w.getScalingReConstruction(), w.getWaveletReConstruction()), theirs):
            assert np.array_equal(mine, ref)


def test_getters_return_copies():
    w = jw.WaveletBuilder.create("Daubechies 4")
    w.getScalingDeComposition()[0] = 99.0
    assert w.getScalingDeComposition()[0] != 99.0  # Wavelet.java:178-219 hands out copies


def test_builder_names():
    assert jw.WaveletBuilder.create("Symlet 8").getMotherWavelength() == 16
    assert jw.WaveletBuilder.create("Symlet8").getName() == "Symlet 8"
    with pytest.raises(jw.JWaveFailure):
        jw.WaveletBuilder.create("Mexican Hat")
    assert len(jw.WaveletBuilder.create2arr()) == 52  # 44 orthonormal + 8 BiOrthogonal


def test_math_tool_kit():
    """MathToolKit.java:185-189, :202-208; SURVEY.md F14: exact on every 2^p, p = 0..30."""
    for p in range(31):
        assert jw.MathToolKit.isBinary(1 << p)
        assert jw.MathToolKit.getExponent(1 << p) == p
        assert co.lib().jwo_get_exponent(float(1 << p)) == p
    for bad in (0, -4, 3, 100, (1 << 20) + 1):
        assert not jw.MathToolKit.isBinary(bad)


def test_exception_hierarchy():
    assert issubclass(jw.JWaveFailure, jw.JWaveException)
    assert issubclass(jw.JWaveError, jw.JWaveException)
    assert jw.JWaveFailure("x").getMessage() == "x"


def test_facade_swallows_and_returns_none(capsys):
    class Broken(jw.BasicTransform):
        def _forward1(self, arr, level=None):
            raise jw.JWaveFailure("boom")

    assert jw.Transform(Broken()).forward(np.ones(4)) is None  # Transform.java:81-90
    assert "boom" in capsys.readouterr().out
    jw.Transform(None)  # Transform.java:62-70: prints, does not throw


def test_cuda_transform_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(jw.JWaveError):
        jw.CudaFastWaveletTransform(jw.WaveletBuilder.create("Haar"), context=jw.CudaContext(0))


def test_transform_builder_rejects_unknown_names():
    """TransformBuilder.java:62-65: unknown transform names (and the out-of-scope DFT) are a JWaveFailure;
    checked before any CUDA context is touched."""
    with pytest.raises(jw.JWaveFailure):
        jw.TransformBuilder.create("Discrete Fourier Transform", "Haar")
    with pytest.raises(jw.JWaveFailure):
        jw.TransformBuilder.create("Fast Wavelet Transform", "no such wavelet")


def test_qbench_tap_tables_are_the_library_taps():
    """tools/qbench_taps.h (the native A/B driver's filter tables, hex float literals) must be bit-identical to the
    four getter arrays of the Python wavelets it was generated from."""
    import os
    import re
    import jwave_b200 as jw
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "qbench_taps.h")
    text = open(path).read()
    entries = re.findall(r'\{"(\w+)", (\d+), \{(.*?)\}\},', text, flags=re.S)
    assert len(entries) >= 5
    for name, L, body in entries:
        w = jw.WaveletBuilder.create(name)
        rows = re.findall(r"\{([^{}]*)\}", body)
        want = (w.getScalingDeComposition(), w.getWaveletDeComposition(),
                w.getScalingReConstruction(), w.getWaveletReConstruction())
        assert len(rows) == 4
        for row, arr in zip(rows, want):
            got = [float.fromhex(t.strip()) for t in row.split(",") if t.strip()]
            assert len(got) == int(L) == len(arr)
            assert all(a == float(b) for a, b in zip(got, arr)), name
