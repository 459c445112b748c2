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


def test_optimize_with_most_optimal_coding():
    """JpegOptimizer.MostOptimalCoding (JpegOptimizer.cs:39): package-merge tables over the scan's symbol statistics.
    Lossless, and the tables written are what the oracle's package-merge builder yields for the same histograms."""
    src = golden_bytes("lake.jpg")
    opt = J.JpegOptimizer()
    opt.MostOptimalCoding = True
    opt.SetInput(src)
    opt.Scan()
    out = bytearray()
    opt.SetOutput(out)
    opt.Optimize()
    assert len(out) < len(src)
    a, b = O.decode(src, want_rgb=False), O.decode(bytes(out), want_rgb=False)
    for ca, cb in zip(a.coef, b.coef):
        assert np.array_equal(ca, cb)
    std = J.JpegOptimizer()
    std.SetInput(src)
    std.Scan()
    out_std = bytearray()
    std.SetOutput(out_std)
    std.Optimize()
    assert [list(t.bits) for t in opt.last_tables] != [list(t.bits) for t in std.last_tables] or bytes(out) == bytes(out_std)


def test_optimizer_errors():
    opt = J.JpegOptimizer()
    with pytest.raises(J.InvalidOperationException):
        opt.Scan()
    with pytest.raises(J.InvalidOperationException):
        opt.Optimize()
    opt.SetInput(golden_bytes("progress.jpg"))
    with pytest.raises(J.InvalidDataException):
        opt.Scan()                              # "Progressive JPEG is not supported currently."
    src = synth.synth_jpeg(3, 64, 48, subsampling="4:4:4")
    opt.SetInput(synth.resequence_scans(src, O.decode(src, want_rgb=False), [[0], [1], [2]]))
    with pytest.raises(J.NotSupportedException):
        opt.Scan()                              # several scans: refused, not transcoded wrongly


@pytest.mark.parametrize("kw", [dict(width=640, height=400, subsampling="4:2:0", restart_rows=1),
                                dict(width=333, height=211, subsampling="4:2:0", restart_blocks=7),   # partial last interval
                                dict(width=200, height=120, subsampling="4:2:2", restart_blocks=1),   # DRI = 1
                                dict(width=1920, height=1080, subsampling="4:4:4", restart_rows=2, quality=92),
                                dict(width=160, height=96, gray=True, restart_blocks=5)],
                         ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
@pytest.mark.parametrize("strip", [True, False])
def test_optimize_keeps_restart_intervals(kw, strip):
    """CopyScanBaseline (JpegOptimizer.cs:772-812): DC prediction restarts per interval, intervals are padded with
    1-bits and separated by RSTn.  The DRI segment is kept even with strip (documented deviation, quirk Q6)."""
    kw = dict(kw)
    w, h = kw.pop("width"), kw.pop("height")
    src = synth.synth_jpeg(14, w, h, **kw)
    opt = J.JpegOptimizer()
    opt.SetInput(src)
    opt.Scan()
    out = bytearray()
    opt.SetOutput(out)
    opt.Optimize(strip)
    out = bytes(out)
    assert len(out) < len(src)
    pa, pb = J.Parsed(src), J.Parsed(out)
    assert pb.desc.scans[0].restart_interval == pa.desc.scans[0].restart_interval != 0
    a, b = O.decode(src, want_rgb=False), O.decode(out, want_rgb=False)   # the oracle checks every RSTn on its way
    for ca, cb in zip(a.coef, b.coef):
        assert np.array_equal(ca, cb)
    sos = out.find(b"\xff\xda")
    nint = -(-(a.mcus_per_line * a.mcus_per_col) // pa.desc.scans[0].restart_interval)
    rst = [out[i + 1] for i in range(sos, len(out) - 1) if out[i] == 0xFF and 0xD0 <= out[i + 1] <= 0xD7]
    assert rst == [0xD0 + (k & 7) for k in range(len(rst))] and len(rst) > 0
    assert len(rst) == nint - 1
    # the GPU decoder takes the result too (restart-segment path), and a second pass is a fixed point
    lay, coef = J.decode_coefficients(out)
    assert np.array_equal(coef, O.scan_order_coefficients(b).reshape(-1, 64))
    opt2 = J.JpegOptimizer()
    opt2.SetInput(out)
    opt2.Scan()
    out2 = bytearray()
    opt2.SetOutput(out2)
    opt2.Optimize(strip)
    assert bytes(out2) == out


def _optimize_one(src, strip=True, most_optimal=False):
    opt = J.JpegOptimizer()
    opt.MostOptimalCoding = most_optimal
    opt.SetInput(src)
    opt.Scan()
    out = bytearray()
    opt.SetOutput(out)
    opt.Optimize(strip)
    return bytes(out)


@pytest.mark.parametrize("most_optimal", [False, True], ids=["standard", "package-merge"])
def test_batch_optimizer_equals_one_optimizer_per_stream(most_optimal):
    """JpegBatchOptimizer: every stream of a mixed batch (sizes, sampling, grey, restart intervals, a real photo) comes
    out byte for byte as JpegOptimizer writes it alone, and decodes to the coefficients it went in with."""
    srcs = [synth.synth_jpeg(60, 320, 200, subsampling="4:2:0"),
            synth.synth_jpeg(61, 200, 136, subsampling="4:4:4", quality=93),
            synth.synth_jpeg(62, 256, 144, subsampling="4:2:0", restart_rows=1),
            synth.synth_jpeg(63, 96, 64, gray=True),
            synth.synth_jpeg(64, 200, 120, subsampling="4:2:2", restart_blocks=7),
            golden_bytes("lake.jpg")]
    with J.JpegBatchOptimizer(srcs, most_optimal=most_optimal) as b:
        outs = b.run()
        assert b.launch_count() > 0
        again = b.run(strip=False)  # the batch object can be re-run
    for i, (src, out) in enumerate(zip(srcs, outs)):
        assert out == _optimize_one(src, True, most_optimal), f"stream {i}"
        assert again[i] == _optimize_one(src, False, most_optimal), f"stream {i} (strip=False)"
        assert len(out) < len(src)
        a, c = O.decode(src, want_rgb=False), O.decode(out, want_rgb=False)
        assert all(np.array_equal(x, y) for x, y in zip(a.coef, c.coef))


def test_batch_optimizer_errors():
    good = synth.synth_jpeg(65, 128, 96, restart_rows=1)
    with pytest.raises(J.InvalidDataException, match="Progressive"):
        J.JpegBatchOptimizer([good, synth.synth_jpeg(66, 64, 64, progressive=True)])
    with pytest.raises(J.InvalidDataException):
        J.JpegBatchOptimizer([good, b"\x00\x01\x02"])
    # a damaged scan: the decoder's verdict is what Scan() raises
    p0, p1 = good.index(b"\xff\xd0"), good.index(b"\xff\xd1")
    bad = good[:p0 + 2] + good[p1:]  # the second restart interval is empty: "Invalid Huffman code encountered."
    with pytest.raises(O.OracleError):
        O.decode(bad)
    with J.JpegBatchOptimizer([good, bad]) as b:
        with pytest.raises(J.InvalidDataException):
            b.run()
    with pytest.raises(J.InvalidDataException):
        _optimize_one(bad)
