"""GPU parity tests: the CUDA path through the C-ABI against the oracle and the reference goldens.

Bar (north_star): entropy-decoded coefficients bit-exact; int16 sample planes bit-exact (the IDCT
is written with non-contractable fp32 ops); 8-bit RGB within +-1 LSB of the oracle (in practice 0:
the colour arithmetic is integer) with the max-abs-diff histogram printed.
"""
import hashlib

import numpy as np
import pytest

import jpeglibrary_b200 as J
import oracle_ffi as O
import synth
from conftest import golden_bytes

pytestmark = pytest.mark.gpu

BASELINE_ASSETS = ["lake.jpg", "cramps.jpg", "testorig12.jpg"]


def gpu_planes(blob):
    dec = J.JpegDecoder()
    dec.SetInput(blob)
    dec.Identify()
    planes = np.zeros((dec.NumberOfComponents, dec.Height, dec.Width), dtype=np.int16)
    dec.SetOutputWriter(J.CudaOutputWriter(planes, J.JB_OUT_PLANAR_I16))
    dec.Decode()
    return planes


def gpu_pixels(blob, fmt=J.JB_OUT_RGB24, bpp=3):
    dec = J.JpegDecoder()
    dec.SetInput(blob)
    dec.Identify()
    out = np.zeros((dec.Height, dec.Width, bpp), dtype=np.uint8)
    dec.SetOutputWriter(J.CudaOutputWriter(out, fmt))
    dec.Decode()
    return out


def check_coefficients(blob):
    o = O.decode(blob, want_rgb=False)
    lay, coef = J.decode_coefficients(blob)
    assert lay.interleaved == 1
    want = O.scan_order_coefficients(o).reshape(-1, 64)
    assert coef.shape == want.shape
    assert np.array_equal(coef, want), f"{int((coef != want).any(axis=1).sum())} blocks differ"
    return o


@pytest.mark.parametrize("name", BASELINE_ASSETS)
def test_golden_assets_coefficients_bit_exact(name):
    check_coefficients(golden_bytes(name))


def reference_golden16(planes, precision):
    """What the reference's test writer stores for unclamped int16 samples (Utils/JpegExtendingOutputWriter.cs:57,77-80:
    (ushort) cast, clamp to 2^P - 1, expand to 16 bits): the bytes of <asset>.jpg.high.png / .low-diff.png."""
    s = np.minimum(planes.astype(np.int16).view(np.uint16).astype(np.uint32), (1 << precision) - 1)
    rem = 16 - precision
    return np.ascontiguousarray(((s << rem) | (s & ((1 << rem) - 1))).astype(np.uint16).transpose(1, 2, 0))


@pytest.mark.parametrize("name", BASELINE_ASSETS)
def test_golden_assets_planes_match_reference_goldens(name, golden):
    planes = gpu_planes(golden_bytes(name))
    g = golden["assets"][name]
    # the reference's own golden vector (sha256 of the buffer its test compares against) ...
    assert hashlib.sha256(reference_golden16(planes, g["precision"]).tobytes()).hexdigest() == g["golden16_sha256"]
    # ... and the oracle's unclamped planes, which carry more information than the clamped golden
    assert hashlib.sha256(np.ascontiguousarray(planes).tobytes()).hexdigest() == g["planes_i16_sha256"]
    o = O.decode(golden_bytes(name), want_rgb=False)
    assert np.array_equal(planes, o.planes)


@pytest.mark.parametrize("name", BASELINE_ASSETS)
def test_golden_assets_rgb(name):
    blob = golden_bytes(name)
    o = O.decode(blob)
    rgb = gpu_pixels(blob)
    diff = np.abs(rgb.astype(int) - o.rgb.astype(int))
    hist = np.bincount(diff.ravel(), minlength=3)
    print(f"{name}: max|diff|={diff.max()} histogram={hist[:4].tolist()}")
    assert diff.max() <= 1
    assert np.array_equal(gpu_pixels(blob, J.JB_OUT_YCBCR888), o.ycbcr)
    rgba = gpu_pixels(blob, J.JB_OUT_RGBA32, 4)
    assert np.array_equal(rgba[..., :3], rgb) and (rgba[..., 3] == 255).all()


SHAPES = [
    dict(width=64, height=48, subsampling="4:2:0", restart_rows=1),
    dict(width=333, height=211, subsampling="4:2:0", restart_blocks=7),      # ragged edges, odd DRI
    dict(width=200, height=120, subsampling="4:2:2", restart_blocks=3),
    dict(width=97, height=131, subsampling="4:4:4", restart_rows=2),
    dict(width=640, height=360, subsampling="4:2:0"),                        # no restart markers
    dict(width=8, height=8, subsampling="4:4:4"),                            # single MCU
    dict(width=17, height=9, subsampling="4:2:0", restart_blocks=1),         # DRI = 1
    dict(width=256, height=256, subsampling="4:2:0", restart_rows=1, optimize=True),  # optimised tables (long codes)
    dict(width=160, height=96, gray=True, restart_blocks=5),
    dict(width=1920, height=1080, subsampling="4:2:0", restart_rows=1, quality=95),
    # 16-byte aligned rows whose width/height are not multiples of the renderer's unit: the 2-D TMA tensor store
    # clips the partial units at the right and bottom edges
    dict(width=336, height=200, subsampling="4:2:0", restart_rows=1),
    dict(width=352, height=120, subsampling="4:4:4", restart_blocks=9),
    dict(width=400, height=104, subsampling="4:2:2", restart_rows=1),
    dict(width=272, height=100, gray=True, restart_blocks=11),
]


@pytest.mark.parametrize("kw", SHAPES, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_synthetic_streams(kw):
    kw = dict(kw)
    w, h = kw.pop("width"), kw.pop("height")
    blob = synth.encode_jpeg(synth.synth_rgb(7, w, h), **kw)
    o = check_coefficients(blob)
    o = O.decode(blob)
    assert np.array_equal(gpu_planes(blob), o.planes)
    rgb = gpu_pixels(blob)
    diff = np.abs(rgb.astype(int) - o.rgb.astype(int))
    print("max|diff|", diff.max(), np.bincount(diff.ravel(), minlength=2)[:3].tolist())
    assert diff.max() <= 1
    rgba = gpu_pixels(blob, J.JB_OUT_RGBA32, 4)
    assert np.array_equal(rgba[..., :3], rgb) and (rgba[..., 3] == 255).all()
    assert np.array_equal(gpu_pixels(blob, J.JB_OUT_YCBCR888), o.ycbcr)


@pytest.mark.parametrize("shape", [(96, 112), (104, 80), (1024, 64)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_vertically_subsampled_440_streams(shape):
    """Luma 1x2 (4:4:0): Pillow cannot write it, so the stream comes from the oracle's encoder.  Exercises the
    renderer's <HS=1, VS=2> instantiation, with (96, 1024 px) and without (104 px) the TMA tensor store."""
    w, h = shape
    blob = O.encode_ycbcr(O.rgb_to_ycbcr(synth.synth_rgb(11, w, h)), quality=88, subsampling=(1, 2)).bytes
    check_coefficients(blob)
    o = O.decode(blob)
    assert np.array_equal(gpu_planes(blob), o.planes)
    rgb = gpu_pixels(blob)
    assert np.abs(rgb.astype(int) - o.rgb.astype(int)).max() <= 1
    rgba = gpu_pixels(blob, J.JB_OUT_RGBA32, 4)
    assert np.array_equal(rgba[..., :3], rgb) and (rgba[..., 3] == 255).all()
    assert np.array_equal(gpu_pixels(blob, J.JB_OUT_YCBCR888), o.ycbcr)


@pytest.mark.parametrize("ss,shape", [((4, 1), (128, 40)), ((4, 2), (128, 48)), ((1, 4), (40, 64))],
                         ids=lambda v: "x".join(map(str, v)))
def test_sampling_factor_four_goes_through_the_generic_renderer(ss, shape):
    """Factors of 4 (JpegEncoder.AddComponent accepts 1, 2 or 4) have no specialised renderer instance."""
    w, h = shape
    blob = O.encode_ycbcr(O.rgb_to_ycbcr(synth.synth_rgb(12, w, h)), quality=88, subsampling=ss).bytes
    check_coefficients(blob)
    o = O.decode(blob)
    assert np.array_equal(gpu_planes(blob), o.planes)
    assert np.abs(gpu_pixels(blob).astype(int) - o.rgb.astype(int)).max() <= 1
    assert np.array_equal(gpu_pixels(blob, J.JB_OUT_YCBCR888), o.ycbcr)


NO_RESTART_SHAPES = [
    dict(width=1920, height=1080, subsampling="4:2:0", quality=85),   # ~800 sub-sequences, several CTAs
    dict(width=1280, height=720, subsampling="4:4:4", quality=92),
    dict(width=1000, height=700, subsampling="4:2:2", quality=75, optimize=True),
    dict(width=900, height=600, gray=True, quality=90),
    dict(width=512, height=512, subsampling="4:2:0", quality=30),     # long zero runs, EOB-heavy blocks
    dict(width=700, height=500, subsampling="4:2:0", quality=100),    # quant tables of ones: long blocks
]


@pytest.mark.parametrize("kw", NO_RESTART_SHAPES, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_self_synchronising_decode_without_restart_markers(kw):
    """configs[2]: speculative sub-sequence decode + sync rounds + DC prefix must reproduce the
    sequential decoder bit for bit."""
    kw = dict(kw)
    w, h = kw.pop("width"), kw.pop("height")
    blob = synth.encode_jpeg(synth.synth_rgb(21, w, h), **kw)
    assert J.Parsed(blob).desc.scans[0].restart_interval == 0
    check_coefficients(blob)
    o = O.decode(blob)
    assert np.array_equal(gpu_planes(blob), o.planes)
    assert np.abs(gpu_pixels(blob).astype(int) - o.rgb.astype(int)).max() <= 1


@pytest.mark.parametrize("kw", [dict(subsampling="4:2:0", restart_rows=1), dict(subsampling="4:4:4"), dict(gray=True),
                                dict(subsampling="4:2:2", progressive=True)], ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_host_destination_with_a_padded_pitch_keeps_its_padding(kw):
    """A pitch wider than a row of pixels: the bytes between the rows are the caller's (the reference's writer touches
    pixels only); the result leaves the staging buffer row by row, not as one block."""
    blob = synth.synth_jpeg(21, 77, 45, **kw)
    want = gpu_pixels(blob)
    pitch = 3 * 77 + 29
    padded = np.full((45, pitch), 0xA5, np.uint8)
    dec = J.JpegDecoder()
    dec.SetInput(blob)
    dec.SetOutputWriter(J.CudaOutputWriter(padded, J.JB_OUT_RGB24, pitch=pitch))
    dec.Decode()
    assert np.array_equal(padded[:, :3 * 77].reshape(45, 77, 3), want)
    assert (padded[:, 3 * 77:] == 0xA5).all()


def test_self_sync_stream_that_needs_more_rounds_than_the_launch_runs():
    """A valid noise frame at quality 96 (long codes, little to synchronise on): after the five re-sync rounds of the
    launch some entry states still move, the host iterates to convergence and redoes the write pass -- whose verdict
    counts, not what the first write pass flagged while it decoded from wrong entry states (tests/campaigns/fuzz_shapes.py
    found the stale InvalidDataException)."""
    import os
    blob = open(os.path.join(os.path.dirname(__file__), "fixtures", "valid_420_no_restart_slow_to_synchronise.jpg"), "rb").read()
    o = O.decode(blob)
    for rounds in ("2", "5", None):   # JB_SS_ROUNDS_NOW: fewer re-sync rounds per launch than the library's dozen
        if rounds:
            os.environ["JB_SS_ROUNDS_NOW"] = rounds
        try:
            check_coefficients(blob)
            assert np.array_equal(gpu_planes(blob), o.planes)
            assert np.abs(gpu_pixels(blob).astype(int) - o.rgb.astype(int)).max() <= 1
            with J.JpegBatchDecoder([blob, synth.synth_jpeg(3, 320, 240), blob], J.JB_OUT_RGB24, device_output=True) as b:
                b.run()
                assert b.status() == [0, 0, 0]
                assert np.abs(b.read_output(2).astype(int) - o.rgb.astype(int)).max() <= 1
        finally:
            os.environ.pop("JB_SS_ROUNDS_NOW", None)


def test_self_sync_equals_restart_decode_on_same_pixels():
    """Size-independent property at the bench's full size: the same 4K pixels coded with and without
    restart markers must give identical coefficients through the two different GPU entropy paths."""
    rgb = synth.synth_rgb(1, 3840, 2160)
    a = synth.encode_jpeg(rgb, restart_rows=1)
    b = synth.encode_jpeg(rgb)
    _, ca = J.decode_coefficients(a)
    _, cb = J.decode_coefficients(b)
    assert np.array_equal(ca, cb)
    want = O.scan_order_coefficients(O.decode(b, want_rgb=False)).reshape(-1, 64)
    assert np.array_equal(cb, want)


def test_batch_of_mixed_images_device_resident():
    blobs = [synth.synth_jpeg(i, 320 + 16 * i, 240 - 8 * i, restart_rows=1, subsampling="4:2:0" if i % 2 else "4:4:4")
             for i in range(6)]
    blobs.append(golden_bytes("lake.jpg"))
    with J.JpegBatchDecoder(blobs, J.JB_OUT_RGB24, device_output=True) as b:
        b.run()
        assert b.status() == [0] * len(blobs)
        # restart scan + segment descriptors + absent-interval clear + segment Huffman + self-sync chain for lake.jpg (un-stuff count and
        # copy, guess round, 12 sync rounds, prefix sums, descriptors, write) + one IDCT/colour launch per layout
        # (+ the status, MCU-limit and first-error clears at the start and the status mailbox post at the end)
        assert b.launch_count() == 3 + 1 + 3 + (6 + 12) + 2 + 1
        for i, blob in enumerate(blobs):
            assert np.array_equal(b.read_output(i), O.decode(blob).rgb)
        b.upload(); b.launch(); b.finish()   # a batch object can be re-run
        assert np.array_equal(b.read_output(0), O.decode(blobs[0]).rgb)


def test_compatibility_path_replays_write_block_calls():
    """Arbitrary JpegBlockOutputWriter: same blocks, same order as the reference decoder would issue."""
    blob = synth.synth_jpeg(11, 40, 24, restart_rows=1)

    class Recorder(J.JpegBlockOutputWriter):
        def __init__(self):
            self.calls = []

        def WriteBlock(self, blockRef, componentIndex, x, y):
            self.calls.append((componentIndex, x, y, blockRef.copy()))

    rec = Recorder()
    dec = J.JpegDecoder()
    dec.SetInput(blob)
    dec.Identify()
    dec.SetOutputWriter(rec)
    dec.Decode()
    o = O.decode(blob)
    # 40x24 4:2:0 -> 3x2 MCUs, per MCU 4 Y blocks + 4 replicated Cb + 4 replicated Cr
    assert len(rec.calls) == 3 * 2 * 12
    assert [c[:3] for c in rec.calls[:6]] == [(0, 0, 0), (0, 8, 0), (0, 0, 8), (0, 8, 8), (1, 0, 0), (1, 8, 0)]
    for ci, x, y, blk in rec.calls:
        hh, ww = min(8, o.height - y), min(8, o.width - x)
        if hh > 0 and ww > 0:
            assert np.array_equal(blk.reshape(8, 8)[:hh, :ww], o.planes[ci, y:y + hh, x:x + ww])


def test_error_codes_follow_reference_exceptions():
    blob = bytearray(synth.synth_jpeg(1, 64, 48, restart_blocks=2))
    i = blob.find(b"\xff\xd1")
    bad = bytearray(blob)
    bad[i + 1] = 0x00  # RST1 becomes a stuffed FF: the segment count no longer matches
    out = np.zeros((48, 64, 3), np.uint8)
    dec = J.JpegDecoder()
    dec.SetInput(bytes(bad))
    dec.SetOutputWriter(J.CudaOutputWriter(out))
    with pytest.raises((J.InvalidOperationException, J.InvalidDataException)):
        dec.Decode()
    with pytest.raises(O.OracleError):
        O.decode(bytes(bad))
    # destination too small -> ArgumentException("Destination buffer is too small.")
    dec = J.JpegDecoder()
    dec.SetInput(bytes(blob))
    dec.SetOutputWriter(J.CudaOutputWriter(np.zeros(100, np.uint8)))
    with pytest.raises(J.ArgumentException):
        dec.Decode()


@pytest.mark.parametrize("kind", ["baseline", "lossless"])
def test_eoi_at_a_restart_boundary_ends_the_scan_quietly(kind):
    """A frame header that promises more MCU rows than the stream holds, with the EOI sitting exactly where the next
    RSTn would be: the reference stops decoding without an error (JpegHuffmanBaselineScanDecoder.cs:144-150,
    JpegHuffmanLosslessScanDecoder.cs:172-176).  (Progressive scans are followed by DHT/SOS, not EOI: there the same
    situation is "Expect restart marker.", see test_progressive_complete_last_interval_quirk.)"""
    if kind == "lossless":
        blob = bytearray(synth.synth_lossless(6, 64, 40, predictor=2, restart=64)[0])   # one interval per row
    else:
        blob = bytearray(synth.synth_jpeg(6, 160, 112, subsampling="4:4:4", restart_rows=1))
    sof = next(i for i in range(2, len(blob)) if blob[i] == 0xFF and blob[i + 1] in (0xC0, 0xC3))
    height = int.from_bytes(blob[sof + 5:sof + 7], "big")
    blob[sof + 5:sof + 7] = (height + 48).to_bytes(2, "big")
    blob = bytes(blob)
    o = O.decode(blob, want_rgb=False)          # no error
    assert o.height == height + 48
    planes = gpu_planes(blob)
    assert np.array_equal(planes[:, :height], o.planes[:, :height])   # what was decoded is identical ...
    assert np.array_equal(planes, o.planes)     # ... and the rest was never written (a fresh buffer's zeros)


def _decode_into(blob, out, fmt, ctx=None, on_device=False):
    dec = J.JpegDecoder(ctx)
    dec.SetInput(blob)
    dec.Identify()
    if on_device:
        dev = ctx.device_alloc(out.nbytes)
        ctx.h2d(dev, out)
        dec.SetOutputWriter(J.CudaOutputWriter(dev, fmt, on_device=True, capacity=out.nbytes))
        dec.Decode()
        ctx.d2h(out, dev)
        ctx.device_free(dev)
    else:
        dec.SetOutputWriter(J.CudaOutputWriter(out, fmt))
        dec.Decode()


@pytest.mark.parametrize("restart", [dict(restart_rows=1), dict(restart_blocks=7)], ids=["rows", "blocks7"])
@pytest.mark.parametrize("subsampling", ["4:2:0", "4:4:4"])
@pytest.mark.parametrize("on_device", [False, True], ids=["host", "device"])
def test_eoi_at_a_restart_boundary_leaves_the_missing_intervals_unwritten(restart, subsampling, on_device):
    """JpegHuffmanBaselineScanDecoder.cs:144-150: the scan returns at an EOI that sits where an RSTn would be; WriteBlock is
    never called for the MCUs of the missing intervals, so whatever the destination held stays there."""
    blob = bytearray(synth.synth_jpeg(9, 200, 136, subsampling=subsampling, **restart))
    rst = [i for i in range(len(blob) - 1) if blob[i] == 0xFF and 0xD0 <= blob[i + 1] <= 0xD7]
    cut = rst[len(rst) * 2 // 3]
    blob[cut:] = b"\xff\xd9"                     # ends exactly where a restart marker was
    blob = bytes(blob)
    o = O.decode(blob)                            # no error
    wr = O.written_samples(o)
    assert 0 < wr[0].sum() < wr[0].size and not wr[0, -1].any()
    assert not o.planes[~wr].any()
    ctx = J.Context.default()
    planes = np.full((3, o.height, o.width), 0x5A5A, np.int16)
    _decode_into(blob, planes, J.JB_OUT_PLANAR_I16, ctx, on_device)
    assert np.array_equal(planes[wr], o.planes[wr])
    assert (planes[~wr] == 0x5A5A).all()
    px = wr.all(axis=0)                           # pixels the application's converter gets all three samples of
    assert np.array_equal(px, wr.any(axis=0))
    for fmt, bpp in ((J.JB_OUT_RGB24, 3), (J.JB_OUT_RGBA32, 4), (J.JB_OUT_YCBCR888, 3)):
        out = np.full((o.height, o.width, bpp), 0xA5, np.uint8)
        _decode_into(blob, out, fmt, ctx, on_device)
        want = o.ycbcr if fmt == J.JB_OUT_YCBCR888 else o.rgb
        assert np.array_equal(out[px][:, :3], want[px])
        assert (out[~px] == 0xA5).all()


# ------------------------------------------------------------------------------------------ progressive
def check_progressive_coefficients(blob):
    o = O.decode(blob, want_rgb=False)
    lay, coef = J.decode_coefficients(blob)
    assert lay.interleaved == 0
    for c in range(o.ncomp):
        w, h = lay.comp_blocks_w[c], lay.comp_blocks_h[c]
        assert (w, h) == (o.coef_w[c], o.coef_h[c])
        plane = coef[lay.comp_block_offset[c]:lay.comp_block_offset[c] + w * h].reshape(h, w, 64)
        aw, ah = o.alloc_w[c], o.alloc_h[c]  # blocks outside the reference allocator's grid hit its dummy block
        assert np.array_equal(plane[:ah, :aw], o.coef[c][:ah, :aw]), f"component {c}"
    return o


@pytest.mark.parametrize("name", ["progress.jpg", "yellowcat_progressive_restart.jpg"])
def test_progressive_golden_assets(name, golden):
    blob = golden_bytes(name)
    o = check_progressive_coefficients(blob)
    planes = gpu_planes(blob)
    assert hashlib.sha256(reference_golden16(planes, golden["assets"][name]["precision"]).tobytes()).hexdigest() == golden["assets"][name]["golden16_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(planes).tobytes()).hexdigest() == golden["assets"][name]["planes_i16_sha256"]
    assert np.array_equal(planes, o.planes)
    o = O.decode(blob)
    assert np.abs(gpu_pixels(blob).astype(int) - o.rgb.astype(int)).max() <= 1


PROGRESSIVE_SHAPES = [
    dict(width=1920, height=1080, subsampling="4:4:4", quality=85),          # configs[3] shape
    dict(width=333, height=211, subsampling="4:2:0", quality=85),            # ragged, MCU padding blocks
    dict(width=640, height=480, subsampling="4:2:0", quality=90, restart_blocks=7),
    dict(width=200, height=120, subsampling="4:2:2", quality=75, restart_blocks=4),
    dict(width=160, height=96, gray=True, quality=90),
    dict(width=256, height=256, subsampling="4:4:4", quality=95, optimize=True),
]


@pytest.mark.parametrize("kw", PROGRESSIVE_SHAPES, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_progressive_synthetic(kw):
    kw = dict(kw)
    w, h = kw.pop("width"), kw.pop("height")
    blob = synth.encode_jpeg(synth.synth_rgb(9, w, h), progressive=True, **kw)
    assert J.Parsed(blob).desc.sof == 2
    check_progressive_coefficients(blob)
    o = O.decode(blob)
    assert np.array_equal(gpu_planes(blob), o.planes)
    assert np.abs(gpu_pixels(blob).astype(int) - o.rgb.astype(int)).max() <= 1


@pytest.mark.parametrize("order", [
    [0, 1, 4, 5, 9, 2, 3, 7, 8, 6],          # luma chain first: a legal progression in another order
    [0, 2, 3, 7, 8, 1, 4, 5, 9, 6],          # chroma first
    [0, 1, 4, 5, 9, 1, 2, 3, 7, 8, 6],       # an AC first scan BEHIND the refinement of its band
    [0, 1, 2, 3, 5, 4, 6, 7, 8, 9],          # refinement in front of the band's first scan: the stream does not decode
    [0, 1, 2, 3, 4, 5, 6, 7, 8, 5, 9],       # a refinement scan twice: the second pass does not decode
], ids=lambda o: "".join(map(str, o)))
def test_progressive_scan_scripts_out_of_order(order):
    """The reference decodes the scans of a progressive frame in file order whatever the progression
    (JpegHuffmanProgressiveScanDecoder.ProcessScan :57-90).  K1c runs the scans concurrently behind their producers (scans
    that share a component and overlap in band): any script must give what the oracle gives, coefficient by coefficient,
    or the same error."""
    base = synth.synth_jpeg(9, 320, 240, progressive=True, subsampling="4:4:4", quality=85)
    blob = synth.reorder_progressive_scans(base, order)
    try:
        o = O.decode(blob)
    except O.OracleError as e:
        with pytest.raises(J.InvalidDataException if e.code == -1 else J.InvalidOperationException):
            gpu_planes(blob)
        return
    check_progressive_coefficients(blob)
    assert np.array_equal(gpu_planes(blob), o.planes)


def test_progressive_complete_last_interval_quirk():
    """Reference quirk: when a progressive scan ends exactly on a restart-interval boundary, HandleRestart
    (JpegHuffmanProgressiveScanDecoder.cs:196-224) demands RSTn or EOI, but the next marker is DHT/SOS ->
    InvalidOperationException("Expect restart marker.").  The GPU path reports the same error."""
    blob = synth.synth_jpeg(9, 640, 480, progressive=True, restart_rows=2)
    with pytest.raises(O.OracleError) as e:
        O.decode(blob)
    assert e.value.code == -2
    dec = J.JpegDecoder()
    dec.SetInput(blob)
    dec.SetOutputWriter(J.CudaOutputWriter(np.zeros((480, 640, 3), np.uint8)))
    with pytest.raises(J.InvalidOperationException):
        dec.Decode()


def test_progressive_and_sequential_in_one_batch():
    blobs = [synth.synth_jpeg(3, 320, 240, progressive=True, subsampling="4:4:4"),
             synth.synth_jpeg(4, 320, 240, restart_rows=1),
             golden_bytes("progress.jpg"),
             synth.synth_jpeg(5, 640, 360),
             golden_bytes("lossless4_s22.jpg"),
             golden_bytes("cramps.jpg")]
    with J.JpegBatchDecoder(blobs, J.JB_OUT_RGB24, device_output=True) as b:
        b.run()
        assert b.status() == [0] * len(blobs)
        for i, blob in enumerate(blobs):
            assert np.array_equal(b.read_output(i), O.decode(blob).rgb)


def test_progressive_batch_with_concurrent_dependent_scans():
    """A few thousand K1c warps at once: every scan of every frame runs in one launch, refinement scans following
    their producers block by block (more warps than the GPU holds at a time, so late jobs start behind early ones)."""
    kinds = [dict(subsampling="4:4:4", quality=85), dict(subsampling="4:2:0", quality=92),
             dict(subsampling="4:2:2", quality=70, restart_blocks=9), dict(gray=True, quality=88),
             dict(subsampling="4:2:0", quality=80, restart_blocks=37)]
    distinct = [synth.encode_jpeg(synth.synth_rgb(60 + i, 400 + 24 * i, 296 - 16 * i), progressive=True, **kinds[i % len(kinds)])
                for i in range(10)]
    want = []
    for blob in distinct:  # (one of them ends a scan on a restart-interval boundary: the reference's quirk above)
        try:
            want.append(O.decode(blob).rgb)
        except O.OracleError as e:
            want.append(e.code)
    assert sum(isinstance(w, int) for w in want) <= 2
    blobs = [distinct[i % len(distinct)] for i in range(600)]
    with J.JpegBatchDecoder(blobs, J.JB_OUT_RGB24, device_output=True) as b:
        try:
            b.run()
        except (J.InvalidOperationException, J.InvalidDataException):
            pass
        st = b.status()
        for i in range(600):
            w = want[i % len(distinct)]
            assert st[i] == (w if isinstance(w, int) else 0), i
        for i in list(range(0, 600, 7)) + [599]:
            w = want[i % len(distinct)]
            if not isinstance(w, int):
                assert np.array_equal(b.read_output(i), w), i


def test_progressive_large_batch_packed_scans_with_end_of_band_runs():
    """More jobs than the GPU holds, the cheap scans packed 32 frames to a warp, low quality (long end-of-band runs: an
    AC first scan that follows another one must not report skipped blocks as done before its producer is past them)."""
    distinct = [synth.synth_jpeg(90 + i, 256, 192, progressive=True, subsampling="4:4:4", quality=35 + 10 * i) for i in range(4)]
    want = [O.decode(b).rgb for b in distinct]
    n = 2048
    with J.JpegBatchDecoder([distinct[i % 4] for i in range(n)], J.JB_OUT_RGB24, device_output=True) as b:
        for _ in range(3):
            b.run()
            assert b.status() == [0] * n
        for i in list(range(0, n, 97)) + [n - 1]:
            assert np.array_equal(b.read_output(i), want[i % 4]), i


def test_progressive_schedule_trace():
    """jb_decode_batch_scan_trace: one record per K1c job; a consumer scan never ends before its producer."""
    blobs = [synth.synth_jpeg(80 + i, 320, 240, progressive=True, subsampling="4:4:4") for i in range(3)]
    with J.JpegBatchDecoder(blobs, J.JB_OUT_RGB24, device_output=True) as b:
        b.set_profiling(True, trace=True)
        b.run()
        trace = b.scan_trace()
        assert np.array_equal(b.read_output(1), O.decode(blobs[1]).rgb)     # the instrumented kernel decodes the same
    assert len(trace) == 3 * 10                                             # libjpeg's script: 10 scans, one segment each
    for image, scan, seg, start, end, waited in trace:
        assert 0 <= image < 3 and 0 <= scan < 10 and seg == 0 and end >= start and waited <= end - start
    for image in range(3):
        ends = {scan: end for im, scan, _, _, end, _ in trace if im == image}
        assert ends[9] >= ends[5] >= max(ends[1], ends[4])                  # Y refine 2 follows Y refine 1 follows Y first scans
        assert ends[6] >= ends[0] and ends[7] >= ends[2] and ends[8] >= ends[3]


# ------------------------------------------------------------------------------------------ sequential, several scans
SCAN_SCRIPTS = [
    dict(subsampling="4:4:4", scans=[[0], [1], [2]]),                      # the usual non-interleaved file
    dict(subsampling="4:2:0", scans=[[0], [1], [2]]),                      # quirk Q2: luma walked MCU by MCU, 2x2 each
    dict(subsampling="4:2:0", scans=[[0, 1], [2]], restart=5),
    dict(subsampling="4:2:2", scans=[[2], [0], [1]], restart=3),
    dict(subsampling="4:2:0", scans=[[0, 2, 2]]),                          # a component named twice, one never
    dict(subsampling="4:4:4", scans=[[1]]),                                # two components never written
    dict(subsampling="4:2:0", scans=[[0, 1, 2], [0]]),                     # a second pass over the luma blocks
]


@pytest.mark.parametrize("kw", SCAN_SCRIPTS, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()).replace(" ", ""))
def test_sequential_frames_with_several_scans(kw):
    """SOF0 frames that are not one interleaved scan over every component go through the scan list (K1c): every
    scan is an MCU walk with the component's own h x v blocks like the reference's baseline decoder does it
    (JpegHuffmanBaselineScanDecoder.cs:99-137), components no scan names keep zero samples."""
    src = synth.synth_jpeg(33, 200, 136, subsampling=kw["subsampling"], quality=88)
    blob = synth.resequence_scans(src, O.decode(src, want_rgb=False), kw["scans"], kw.get("restart", 0))
    o = check_progressive_coefficients(blob)      # planar store; the oracle walks the scans like the reference
    o = O.decode(blob)
    covered = sorted({c for sc in kw["scans"] for c in sc})
    if covered == [0, 1, 2]:
        assert np.array_equal(o.planes, O.decode(src, want_rgb=False).planes)   # the generator is sound
    planes = gpu_planes(blob)
    assert np.array_equal(planes, o.planes)
    assert np.array_equal(gpu_pixels(blob, J.JB_OUT_YCBCR888), o.ycbcr)
    assert np.abs(gpu_pixels(blob).astype(int) - o.rgb.astype(int)).max() <= 1


@pytest.mark.parametrize("which", [0, 1, 2])
@pytest.mark.parametrize("on_device", [False, True])
def test_sequential_scan_list_cut_at_a_restart_marker(which, on_device):
    """A multi-scan sequential frame whose scan `which` stops at one of its restart markers with EOI behind it: the
    reference's baseline decoder ends that scan quietly and never calls WriteBlock for the intervals behind the cut
    (JpegHuffmanBaselineScanDecoder.cs:144-150), the scans behind it do not exist.  Components are therefore written up
    to different MCUs: samples of a component behind its cut read 0 as far as any component was written (the writer's
    untouched buffer), and nothing is written behind the furthest cut -- checked with a pre-filled destination."""
    src = synth.synth_jpeg(34, 176, 120, subsampling="4:2:0", quality=88)
    # (88 MCUs in intervals of 5: a last interval that is complete would trip the reference's "Expect restart marker." quirk)
    blob = synth.resequence_scans(src, O.decode(src, want_rgb=False), [[0], [1], [2]], 5)
    sos = [i for i in range(len(blob) - 1) if blob[i] == 0xFF and blob[i + 1] == 0xDA]
    assert len(sos) == 3
    lo, hi = sos[which], (sos[which + 1] if which < 2 else len(blob) - 2)
    rst = [i for i in range(lo, hi - 1) if blob[i] == 0xFF and 0xD0 <= blob[i + 1] <= 0xD7]
    cut = blob[:rst[len(rst) // 2]] + b"\xff\xd9"
    want = O.decode(cut, want_rgb=False)
    wr = O.written_samples(want)
    assert wr.any() and not wr.all()
    dec = J.JpegDecoder()
    dec.SetInput(cut)
    dec.Identify()
    if on_device:
        ctx = J.Context.default()
        nbytes = 3 * dec.Height * dec.Width * 2
        dev = ctx.device_alloc(nbytes)
        fill = np.full((3, dec.Height, dec.Width), 0x5A5A, dtype=np.int16)
        ctx.check(J._native.cuda.jb_memcpy_h2d(ctx.handle, dev, fill.ctypes.data, nbytes))
        dec.SetOutputWriter(J.CudaOutputWriter(dev, J.JB_OUT_PLANAR_I16, on_device=True, capacity=nbytes))
        dec.Decode()
        planes = np.empty_like(fill)
        ctx.check(J._native.cuda.jb_memcpy_d2h(ctx.handle, planes.ctypes.data, dev, nbytes))
        ctx.device_free(dev)
    else:
        planes = np.full((3, dec.Height, dec.Width), 0x5A5A, dtype=np.int16)
        dec.SetOutputWriter(J.CudaOutputWriter(planes, J.JB_OUT_PLANAR_I16))
        dec.Decode()
    assert np.array_equal(planes[wr], want.planes[wr])
    some = wr.any(axis=0)                       # pixels some component was written for
    rows = np.nonzero(some.any(axis=1))[0]
    assert (planes[:, rows.max() + 1:, :] == 0x5A5A).all()          # whole MCU rows behind the furthest cut: untouched
    inside = np.broadcast_to(some, wr.shape) & ~wr
    assert (planes[inside] == 0).all()                              # behind a component's own cut, inside the written part


# ------------------------------------------------------------------------------------------ lossless (SOF3)
LOSSLESS_ASSETS = ["lossless%d_s22.jpg" % i for i in range(1, 8)]


@pytest.mark.parametrize("name", LOSSLESS_ASSETS)
def test_lossless_golden_assets_bit_exact(name, golden):
    """SOF3 output must be bit-exact (north_star): predictors 1..7 of the reference's own assets."""
    blob = golden_bytes(name)
    planes = gpu_planes(blob)
    assert hashlib.sha256(reference_golden16(planes, golden["assets"][name]["precision"]).tobytes()).hexdigest() == golden["assets"][name]["golden16_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(planes).tobytes()).hexdigest() == golden["assets"][name]["planes_i16_sha256"]
    o = O.decode(blob)
    assert np.array_equal(planes, o.planes)
    assert np.array_equal(gpu_pixels(blob, J.JB_OUT_YCBCR888), o.ycbcr)
    assert np.array_equal(gpu_pixels(blob), o.rgb)


LOSSLESS_SHAPES = [
    dict(width=40, height=24, predictor=1),
    dict(width=33, height=17, predictor=4, restart=5, ncomp=1),
    dict(width=48, height=32, predictor=7, sampling=[(2, 2), (1, 1), (1, 1)], restart=6),
    dict(width=30, height=20, predictor=5, precision=16),
    dict(width=30, height=20, predictor=6, precision=12, point_transform=2),
    dict(width=200, height=150, predictor=2, restart=200),              # several 32-row bands, one interval per row
    dict(width=256, height=96, predictor=3, sampling=[(2, 1), (1, 1), (1, 1)], restart=64),
    dict(width=1, height=1, predictor=1, ncomp=1),
    dict(width=48, height=32, predictor=4, sampling=[(2, 2), (1, 1), (1, 1)], restart=4, scan_components=[0, 2]),  # partial scan
    dict(width=48, height=32, predictor=5, sampling=[(2, 2), (1, 1), (1, 1)], scan_components=[0, 2, 2]),  # a component coded twice
]


@pytest.mark.parametrize("kw", LOSSLESS_SHAPES, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_lossless_synthetic(kw):
    kw = dict(kw)
    w, h = kw.pop("width"), kw.pop("height")
    blob, coded = synth.synth_lossless(5, w, h, **kw)
    assert J.Parsed(blob).desc.sof == 3
    planes = gpu_planes(blob)
    assert np.array_equal(planes, coded)                      # round trip: what was coded comes back
    assert np.array_equal(planes, O.decode(blob, want_rgb=False).planes)


LOSSLESS_SCAN_LISTS = [
    dict(scans=[dict(components=[0]), dict(components=[1], predictor=4), dict(components=[2], predictor=7)]),  # one scan per component
    dict(scans=[dict(components=[0, 1], predictor=2), dict(components=[2], predictor=5)], restart=20),
    dict(scans=[dict(components=[0, 1, 2]), dict(components=[1], predictor=3)], restart=7),       # the later scan over a component wins
    dict(scans=[dict(components=[0], predictor=6), dict(components=[1, 2], predictor=1)], sampling=[(2, 2), (1, 1), (1, 1)]),
    dict(scans=[dict(components=[2]), dict(components=[0])]),                                       # component 1 is never coded: zeros
    dict(scans=[dict(components=[0]), dict(components=[1])], precision=12, restart=100, ncomp=2),
]


@pytest.mark.parametrize("kw", LOSSLESS_SCAN_LISTS, ids=lambda k: "+".join("".join(map(str, s["components"])) for s in k["scans"]))
def test_lossless_frames_with_several_scans(kw):
    """JpegHuffmanLosslessScanDecoder.ProcessScan (:52-205) runs once per SOS over one scanline store."""
    blob, coded = synth.synth_lossless_scans(5, 64, 48, **kw)
    o = O.decode(blob, want_rgb=False)
    assert o.nscans == len(kw["scans"]) and np.array_equal(o.planes, coded)
    planes = gpu_planes(blob)
    assert np.array_equal(planes, coded)
    if kw.get("ncomp", 3) == 3 and kw.get("precision", 8) == 8:
        assert np.array_equal(gpu_pixels(blob), O.decode(blob).rgb)


def test_lossless_scan_list_keeps_the_complete_last_interval_quirk():
    """A scan whose last restart interval is complete looks for RSTn or EOI behind it (JpegHuffmanLosslessScanDecoder.cs
    :164-178); the SOS of the next scan there is "Expect restart marker." -- in the reference and here."""
    blob, _ = synth.synth_lossless_scans(5, 64, 48, scans=[dict(components=[0, 1]), dict(components=[2])], restart=16)
    with pytest.raises(O.OracleError):
        O.decode(blob, want_rgb=False)
    with pytest.raises(J.InvalidOperationException):
        gpu_planes(blob)


@pytest.mark.parametrize("precision", [2, 3, 5, 7, 12, 16])
def test_lossless_pixel_writers_of_other_precisions(precision):
    """8-bit pixel output of P-bit frames follows the reference application's writers: P > 8 shifts down
    (JpegBufferOutputWriterGreaterThan8Bit.cs:34-68), P < 8 repeats the bit pattern
    (JpegBufferOutputWriterLessThan8Bit.cs:35-92)."""
    blob, coded = synth.synth_lossless(9, 64, 40, precision=precision, predictor=4)
    o = O.decode(blob)
    assert np.array_equal(gpu_planes(blob), coded)
    assert np.array_equal(gpu_pixels(blob, J.JB_OUT_YCBCR888), o.ycbcr)
    assert np.array_equal(gpu_pixels(blob), o.rgb)
    rgba = gpu_pixels(blob, J.JB_OUT_RGBA32, 4)
    assert np.array_equal(rgba[..., :3], o.rgb) and (rgba[..., 3] == 255).all()


def test_lossless_missing_restart_marker():
    blob = bytearray(synth.synth_lossless(5, 64, 32, restart=16)[0])
    i = blob.find(b"\xff\xd1")
    blob[i + 1] = 0x00
    with pytest.raises(O.OracleError):
        O.decode(bytes(blob))
    dec = J.JpegDecoder()
    dec.SetInput(bytes(blob))
    dec.SetOutputWriter(J.CudaOutputWriter(np.zeros((3, 32, 64), np.int16), J.JB_OUT_PLANAR_I16))
    with pytest.raises((J.InvalidOperationException, J.InvalidDataException)):
        dec.Decode()


def test_pipelined_host_decode_matches_batch_decode():
    """JpegPipelinedBatchDecoder (two streams, chunked) must fill the host buffer exactly like one batch."""
    blobs = [synth.synth_jpeg(30 + i, 256 + 16 * (i % 3), 160, restart_rows=1 if i % 2 else 0) for i in range(7)]
    ctx = J.Context.default()
    out = ctx.pinned_array(8 * 1024 * 1024)
    out[:] = 0
    pipe = J.JpegPipelinedBatchDecoder([ctx, J.Context(0)], chunk=2, parse_threads=2)
    offs = pipe.decode(blobs, out)
    for i, blob in enumerate(blobs):
        o = O.decode(blob)
        got = out[offs[i]:offs[i] + o.rgb.size].reshape(o.rgb.shape)
        assert np.array_equal(got, o.rgb)
    ctx.pinned_free(out.ctypes.data)


@pytest.mark.parametrize("restart", [1, 0], ids=["restart", "selfsync"])
def test_batch_with_more_huffman_tables_than_the_shared_memory_cache(restart):
    """Every image brings its own optimised Huffman tables, and the images are so small that one CTA of the
    Huffman kernel spans all of them: only the first four tables fit its shared-memory cache, the other images
    decode through the global-memory fallback of the look-up (k_entropy_flat.cuh, JB_K1F_NOTAB)."""
    blobs = [synth.encode_jpeg(synth.synth_rgb(60 + i, 96 + 16 * (i % 2), 64 if restart else 160), quality=50 + 8 * i,
                               subsampling="4:2:0" if i % 2 else "4:4:4", optimize=True,
                               restart_rows=restart) for i in range(6)]
    tables = {bytes(J.Parsed(b).desc.tables[t].values[:16]) + bytes(J.Parsed(b).desc.tables[t].bits) for b in blobs for t in range(4)}
    assert len(tables) > 8
    with J.JpegBatchDecoder(blobs, J.JB_OUT_RGB24, device_output=True) as b:
        b.run()
        assert b.status() == [0] * len(blobs)
        for i, blob in enumerate(blobs):
            assert np.array_equal(b.read_output(i), O.decode(blob).rgb)
    for blob in blobs[:2]:
        check_coefficients(blob)


def _with_fill_bytes(blob, every=1):
    """Insert 0xFF fill bytes in front of RSTn / EOI markers (legal: B.1.1.2; FillBuffer skips them,
    JpegBitReader.cs:108-128)."""
    out = bytearray()
    i, n, seen = 0, len(blob), 0
    sos = blob.find(b"\xff\xda")
    while i < n:
        if i > sos and blob[i] == 0xFF and i + 1 < n and (0xD0 <= blob[i + 1] <= 0xD7 or blob[i + 1] == 0xD9):
            seen += 1
            if seen % every == 0:
                out += b"\xff" * (1 + seen % 3)
        out.append(blob[i])
        i += 1
    return bytes(out)


@pytest.mark.parametrize("kw", [dict(restart_rows=1), dict(restart_blocks=5), dict()], ids=["rows", "blocks5", "norestart"])
def test_fill_bytes_before_markers(kw):
    rgb = synth.synth_rgb(41, 208, 144)
    plain = synth.encode_jpeg(rgb, subsampling="4:2:0", **kw)
    filled = _with_fill_bytes(plain, every=2 if kw else 1)
    assert len(filled) > len(plain)
    a = O.decode(plain)
    b = O.decode(filled)
    assert np.array_equal(a.planes, b.planes)
    _, c0 = J.decode_coefficients(plain)
    _, c1 = J.decode_coefficients(filled)
    assert np.array_equal(c0, c1)
    assert np.array_equal(gpu_pixels(filled), a.rgb)


@pytest.mark.parametrize("kw", [dict(restart_rows=1), dict()], ids=["restart", "norestart"])
def test_truncated_entropy_data_is_an_error(kw):
    """A scan that ends early must not decode silently: ReceiveAndExtend / DecodeHuffmanCode throw
    InvalidDataException, a missing RSTn gives InvalidOperationException (JpegHuffmanScanDecoder.cs:81-115,
    JpegHuffmanBaselineScanDecoder.cs:139-154)."""
    blob = synth.synth_jpeg(43, 320, 240, **kw)
    assert blob.endswith(b"\xff\xd9")
    cut = blob[:len(blob) - 2 - 600] + b"\xff\xd9"
    with pytest.raises(O.OracleError):
        O.decode(cut)
    dec = J.JpegDecoder()
    dec.SetInput(cut)
    dec.SetOutputWriter(J.CudaOutputWriter(np.zeros((240, 320, 3), np.uint8)))
    with pytest.raises((J.InvalidDataException, J.InvalidOperationException)):
        dec.Decode()


# ---- abbreviated streams behind JpegDecoder.LoadTables (JpegDecoder.cs:313-360): the strips of a TIFF file ----------
@pytest.mark.parametrize("kw", [dict(restart_rows=1), dict(), dict(subsampling="4:4:4", restart_blocks=5), dict(progressive=True)], ids=str)
def test_abbreviated_stream_behind_load_tables(kw):
    blob = synth.synth_jpeg(40, 200, 136, **kw)
    tables, rest = synth.split_tables(blob, move=(0xC4, 0xDB, 0xDD))
    want = O.decode(rest, tables=tables)
    assert np.array_equal(want.rgb, O.decode(blob).rgb)  # the oracle's two walks agree with each other
    dec = J.JpegDecoder()
    dec.LoadTables(tables)
    dec.SetInput(rest)
    dec.Identify()
    planes = np.zeros((3, dec.Height, dec.Width), dtype=np.int16)
    dec.SetOutputWriter(J.CudaOutputWriter(planes, J.JB_OUT_PLANAR_I16))
    dec.Decode()
    assert np.array_equal(planes, want.planes)
    rgb = np.zeros((dec.Height, dec.Width, 3), dtype=np.uint8)
    dec.SetOutputWriter(J.CudaOutputWriter(rgb, J.JB_OUT_RGB24))
    dec.Decode()
    assert np.array_equal(rgb, want.rgb)
    # without the tables the same stream is refused like the reference refuses it
    bare = J.JpegDecoder()
    bare.SetInput(rest)
    if not kw.get("progressive"):
        with pytest.raises(J.InvalidDataException):
            bare.Identify()


def test_batch_of_strips_sharing_one_tables_stream():
    """Strips of one TIFF image: equal tables (libjpeg's standard ones), one JPEGTables stream, many abbreviated
    streams -- through the batch facade and the host-to-host pipeline."""
    blobs = [synth.synth_jpeg(50 + i, 256, 64 + 16 * (i % 2), restart_rows=i % 2) for i in range(6)]
    split = [synth.split_tables(b) for b in blobs]
    tables = split[0][0]
    assert all(t == tables for t, _ in split)  # same quality, standard Huffman tables
    strips = [r for _, r in split]
    with J.JpegBatchDecoder(strips, J.JB_OUT_RGB24, device_output=True, tables=tables) as b:
        b.run()
        assert b.status() == [0] * len(strips)
        for i, blob in enumerate(blobs):
            assert np.array_equal(b.read_output(i), O.decode(strips[i], tables=tables).rgb)
            assert np.array_equal(b.read_output(i), O.decode(blob).rgb)
    with pytest.raises(J.InvalidDataException):
        J.JpegBatchDecoder(strips, J.JB_OUT_RGB24, device_output=True)
    ctx = J.Context.default()
    out = ctx.pinned_array(2 * 1024 * 1024)
    out[:] = 0
    offs = J.JpegPipelinedBatchDecoder([ctx], chunk=4, parse_threads=2).decode(strips, out, tables=tables)
    for i, blob in enumerate(blobs):
        o = O.decode(blob)
        assert np.array_equal(out[offs[i]:offs[i] + o.rgb.size].reshape(o.rgb.shape), o.rgb)
    ctx.pinned_free(out.ctypes.data)
