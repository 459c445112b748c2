"""The oracle against the reference's own golden vectors -- CPU only.

When /root/reference is present (build container) the golden PNG pairs are loaded exactly like
tests/JpegLibrary.Tests/Utils/ImageHelper.cs and compared sample by sample; everywhere (also on
the GPU box, where the reference tree does not exist) the committed sha256 of those buffers
(tests/golden/golden.json, made by tests/golden/make_golden.py) is checked.
"""
import hashlib
import os

import numpy as np
import pytest

import oracle_ffi as O
import synth
from conftest import GOLDEN_DIR, REFERENCE_ASSETS, golden_bytes

ASSETS = ["cramps.jpg", "lake.jpg", "testorig12.jpg", "progress.jpg", "yellowcat_progressive_restart.jpg"] + \
    ["lossless%d_s22.jpg" % i for i in range(1, 8)]


def expected16(planes, precision):
    s = planes.astype(np.int16).view(np.uint16).astype(np.uint32)  # (ushort) cast of the test writer
    s = np.minimum(s, (1 << precision) - 1)
    rem = 16 - precision
    return ((s << rem) | (s & ((1 << rem) - 1))).astype(np.uint16).transpose(1, 2, 0)


@pytest.mark.parametrize("name", ASSETS)
def test_oracle_matches_committed_golden_hash(name, golden):
    d = O.decode(golden_bytes(name))
    g = golden["assets"][name]
    assert (d.width, d.height, d.ncomp, d.precision, d.sof) == (g["width"], g["height"], g["ncomp"], g["precision"], g["sof"])
    e = np.ascontiguousarray(expected16(d.planes, d.precision))
    assert hashlib.sha256(e.tobytes()).hexdigest() == g["golden16_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(d.planes).tobytes()).hexdigest() == g["planes_i16_sha256"]
    assert int((d.planes < 0).sum()) == g["negative_samples"]


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ASSETS), reason="reference tree not present")
@pytest.mark.parametrize("name", ASSETS)
def test_oracle_matches_reference_png_goldens(name, golden):
    from PIL import Image
    src = os.path.join("/root/reference", golden["assets"][name]["source"])
    assert open(src, "rb").read() == golden_bytes(name)
    d = O.decode(golden_bytes(name))
    hi = np.array(Image.open(src + ".high.png").convert("RGBA")).astype(np.uint16)
    lo = np.array(Image.open(src + ".low-diff.png").convert("RGBA")).astype(np.uint16)
    gold = ((hi << 8) | (hi ^ lo))[..., :d.ncomp]
    assert np.array_equal(gold, expected16(d.planes, d.precision))


def test_oracle_rgb_close_to_independent_decoder():
    """Sanity anchor for D10/D11 (untested in the reference): libjpeg-turbo's RGB of a 4:4:4 image
    differs only by IDCT rounding + the reference's d4 = 22553 constant."""
    import io
    from PIL import Image
    import synth
    rgb = synth.synth_rgb(3, 96, 64)
    blob = synth.encode_jpeg(rgb, quality=90, subsampling="4:4:4")
    d = O.decode(blob)
    ref = np.array(Image.open(io.BytesIO(blob)).convert("RGB")).astype(int)
    assert np.abs(d.rgb.astype(int) - ref).max() <= 3


def test_oracle_restart_interval_equals_plain_stream():
    """Baseline DRI handling is not pinned by any reference asset: the same pixels coded with and
    without restart markers must decode to identical coefficients."""
    import synth
    rgb = synth.synth_rgb(5, 200, 120)
    a = O.decode(synth.encode_jpeg(rgb, restart_blocks=3))
    b = O.decode(synth.encode_jpeg(rgb))
    assert a.restart_interval == 3 and b.restart_interval == 0
    for ca, cb in zip(a.coef, b.coef):
        assert np.array_equal(ca, cb)
    assert np.array_equal(a.planes, b.planes)


def test_oracle_errors():
    import synth
    blob = bytearray(synth.synth_jpeg(1, 64, 48, restart_blocks=2))
    with pytest.raises(O.OracleError):
        O.decode(bytes(blob[:2]) + b"\x00\x00")
    # destroy the first restart marker -> "Expect restart marker." (InvalidOperationException)
    i = blob.find(b"\xff\xd0")
    blob[i + 1] = 0x00
    with pytest.raises(O.OracleError) as e:
        O.decode(bytes(blob))
    assert e.value.code in (-1, -2)


@pytest.mark.parametrize("kw", [
    dict(width=40, height=24, predictor=1),
    dict(width=33, height=17, predictor=4, restart=5, ncomp=1),
    dict(width=48, height=32, predictor=7, sampling=[(2, 2), (1, 1), (1, 1)], restart=6),
    dict(width=30, height=20, predictor=5, precision=16),            # category 16 / (short) wrap-around
    dict(width=30, height=20, predictor=6, precision=12, point_transform=2),
    dict(width=48, height=32, predictor=5, sampling=[(2, 2), (1, 1), (1, 1)], scan_components=[0, 2, 2]),
], ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_oracle_decodes_synthetic_lossless_streams(kw):
    """The test-only SOF3 generator (tests/synth.py) and the oracle (pinned by the 7 lossless goldens) agree:
    decoding returns the samples that were coded, replicated to full resolution."""
    import synth
    kw = dict(kw)
    w, h = kw.pop("width"), kw.pop("height")
    blob, coded = synth.synth_lossless(3, w, h, **kw)
    d = O.decode(blob, want_rgb=False)
    assert (d.sof, d.width, d.height) == (3, w, h)
    assert np.array_equal(d.planes, coded)


@pytest.mark.parametrize("precision", [2, 3, 4, 5, 6, 7])
def test_oracle_less_than_8_bit_writer(precision):
    """JpegBufferOutputWriterLessThan8Bit.cs:35-92 in closed form: floor(8/P) copies of the P-bit value, then
    the value's LOW (8 mod P) bits -- e.g. P=3, 0b101 -> 0b10110101."""
    blob, coded = synth.synth_lossless(4, 48, 32, precision=precision, predictor=1, ncomp=1)
    o = O.decode(blob)
    v = coded[0].astype(np.int64)
    assert v.min() >= 0 and v.max() < (1 << precision)
    want = np.zeros_like(v)
    for _ in range(8 // precision):
        want = (want << precision) | v
    rem = 8 % precision
    want = (want << rem) | (v & ((1 << rem) - 1))
    assert np.array_equal(o.ycbcr[..., 0], want.astype(np.uint8))
    assert (o.ycbcr[..., 1:] == 128).all()
