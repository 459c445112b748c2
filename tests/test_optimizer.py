"""JpegOptimizer on the GPU.  The reference's own test (Optimizer/OptimizerTests.cs:27-60,
lake.jpg, strip in {true,false}) asserts: output is smaller and decodes to identical pixels."""
import numpy as np
import pytest

import jpeglibrary_b200 as J
import oracle_ffi as O
import synth
from conftest import golden_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("strip", [True, False])
def test_optimize_lake(strip):
    src = golden_bytes("lake.jpg")
    opt = J.JpegOptimizer()
    opt.SetInput(src)
    opt.Scan()
    out = bytearray()
    opt.SetOutput(out)
    opt.Optimize(strip)
    assert len(out) < len(src)
    a, b = O.decode(src), O.decode(bytes(out))
    for ca, cb in zip(a.coef, b.coef):
        assert np.array_equal(ca, cb)          # lossless transcode: identical coefficients ...
    assert np.array_equal(a.rgb, b.rgb)         # ... hence identical pixels (the reference's assertion)
    if not strip:
        assert b"JFIF" in bytes(out[:32])


@pytest.mark.parametrize("kw", [dict(subsampling="4:2:0"), dict(subsampling="4:4:4", quality=95), dict(gray=True)])
def test_optimize_synthetic_and_table_identity(kw):
    src = synth.synth_jpeg(12, 640, 400, **kw)
    opt = J.JpegOptimizer()
    opt.SetInput(src)
    opt.Scan()
    out = bytearray()
    opt.SetOutput(out)
    opt.Optimize()
    assert len(out) < len(src)
    a, b = O.decode(src, want_rgb=False), O.decode(bytes(out), want_rgb=False)
    for ca, cb in zip(a.coef, b.coef):
        assert np.array_equal(ca, cb)
    # libjpeg's own optimiser on the same pixels lands within a few bytes per table: same symbol statistics
    ref = synth.synth_jpeg(12, 640, 400, optimize=True, **kw)
    assert abs(len(out) - len(ref)) < 400
    # the tables written are exactly what the builder yields for the decoded symbol histogram:
    # re-optimising the optimised stream must be a fixed point
    opt2 = J.JpegOptimizer()
    opt2.SetInput(bytes(out))
    opt2.Scan()
    out2 = bytearray()
    opt2.SetOutput(out2)
    opt2.Optimize()
    assert bytes(out2) == bytes(out)


def test_optimizer_errors():
    opt = J.JpegOptimizer()
    with pytest.raises(J.InvalidOperationException):
        opt.Scan()
    with pytest.raises(J.InvalidOperationException):
        opt.Optimize()
    opt.SetInput(golden_bytes("progress.jpg"))
    with pytest.raises(J.InvalidDataException):
        opt.Scan()                              # "Progressive JPEG is not supported currently."
    opt.SetInput(synth.synth_jpeg(3, 64, 48, restart_rows=1))
    with pytest.raises(J.NotSupportedException):
        opt.Scan()
