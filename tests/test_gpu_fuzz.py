"""Corrupted entropy data: the CUDA path must do what the reference does with the same bytes.

A damaged scan is still a decodable bit stream most of the time: the reference (restated by the oracle) walks on
with wrong symbols and produces deterministic garbage, or it throws (bad code, bits running out, a restart
marker that is not where it should be).  For every mutated stream the GPU result has to be that same garbage,
bit for bit, or an error where the oracle reports one -- and never a hang or an out-of-bounds access
(profiles/*sanitizer* records compute-sanitizer over these tests).
"""
import numpy as np
import pytest

import jpeglibrary_b200 as J
import oracle_ffi as O
import synth

pytestmark = pytest.mark.gpu

GPU_ERRORS = (J.InvalidDataException, J.InvalidOperationException, J.NotSupportedException)


def entropy_ranges(blob):
    """[(first, last+1)] byte ranges of entropy-coded data (behind every SOS header up to the next non-RST marker)"""
    out, i, n = [], 2, len(blob)
    while i + 4 <= n:
        assert blob[i] == 0xFF
        m = blob[i + 1]
        ln = int.from_bytes(blob[i + 2:i + 4], "big")
        i += 2 + ln
        if m != 0xDA:
            continue
        j = i
        while j + 1 < n and not (blob[j] == 0xFF and blob[j + 1] not in (0x00, 0xFF) and not 0xD0 <= blob[j + 1] <= 0xD7):
            j += 1
        out.append((i, j))
        i = j
        if blob[i + 1] == 0xD9:
            break
    return out


def mutate(blob, rng, kind):
    b = bytearray(blob)
    ranges = entropy_ranges(blob)
    lo, hi = ranges[int(rng.integers(len(ranges)))]
    if kind == "flip":          # a few bit flips
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(lo, hi))] ^= 1 << int(rng.integers(8))
    elif kind == "bytes":       # random byte values (may create or destroy stuffing and markers)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(lo, hi))] = int(rng.integers(256))
    elif kind == "drop":        # remove a short run
        p = int(rng.integers(lo, hi))
        del b[p:p + int(rng.integers(1, 6))]
    elif kind == "dup":         # repeat a short run
        p = int(rng.integers(lo, hi))
        b[p:p] = b[p:p + int(rng.integers(1, 6))]
    elif kind == "ones":        # a run of 0xFF fill bytes / 1-bits
        p = int(rng.integers(lo, hi))
        b[p:p + int(rng.integers(1, 4))] = b"\xff" * 3
    elif kind == "rows":        # the frame header promises more or fewer lines than the scans hold
        sof = next(i for i in range(2, len(b)) if b[i] == 0xFF and b[i + 1] in (0xC0, 0xC1, 0xC2, 0xC3))
        h = int.from_bytes(b[sof + 5:sof + 7], "big")
        b[sof + 5:sof + 7] = max(1, h + int(rng.integers(-40, 41))).to_bytes(2, "big")
    elif kind == "cut":         # the scan stops at one of its restart markers, EOI follows
        rst = [i for i in range(lo, hi - 1) if b[i] == 0xFF and 0xD0 <= b[i + 1] <= 0xD7]
        p = rst[int(rng.integers(len(rst)))] if rst else int(rng.integers(lo, hi))
        b[p:] = b"\xff\xd9"
    return bytes(b)


def run_oracle(blob):
    try:
        return O.decode(blob, want_rgb=False), None
    except O.OracleError as e:
        return None, e


def run_gpu(blob):
    try:
        dec = J.JpegDecoder()
        dec.SetInput(blob)
        dec.Identify()
        planes = np.zeros((dec.NumberOfComponents, dec.Height, dec.Width), dtype=np.int16)
        dec.SetOutputWriter(J.CudaOutputWriter(planes, J.JB_OUT_PLANAR_I16))
        dec.Decode()
        return planes, None
    except GPU_ERRORS as e:
        return None, e


def base_streams():
    rgb = synth.synth_rgb(40, 160, 112)
    return {
        "restart": synth.encode_jpeg(rgb, quality=85, subsampling="4:2:0", restart_blocks=4),
        "restart_rows_444": synth.encode_jpeg(rgb, quality=92, subsampling="4:4:4", restart_rows=1, optimize=True),
        "plain": synth.encode_jpeg(synth.synth_rgb(41, 640, 360), quality=85, subsampling="4:2:0"),
        "progressive": synth.encode_jpeg(rgb, quality=85, subsampling="4:2:0", progressive=True),
        "progressive_restart": synth.encode_jpeg(rgb, quality=85, subsampling="4:4:4", progressive=True, restart_blocks=6),
        "lossless": synth.synth_lossless(42, 96, 64, predictor=4, restart=24)[0],
        "lossless_plain": synth.synth_lossless(43, 64, 48, predictor=7)[0],
    }


KINDS = ["flip", "bytes", "drop", "dup", "ones", "rows", "cut"]


@pytest.mark.parametrize("name", list(base_streams()))
def test_corrupted_scans_decode_like_the_reference(name):
    blob = base_streams()[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    problems, agree_ok, agree_err, same_class = [], 0, 0, 0
    for trial in range(140):
        kind = KINDS[trial % len(KINDS)]
        bad = mutate(blob, rng, kind)
        want, werr = run_oracle(bad)
        got, gerr = run_gpu(bad)
        if werr is not None and gerr is not None:
            agree_err += 1
            # several damaged intervals can fail in different ways: the reference reports the first one in stream
            # order, the batch status word is the union -- the exception class is compared as a statistic only
            same_class += isinstance(gerr, {-1: J.InvalidDataException, -2: J.InvalidOperationException,
                                            -3: J.NotSupportedException}[werr.code])
        elif werr is None and gerr is None:
            if np.array_equal(got, want.planes):
                agree_ok += 1
            else:
                problems.append(f"trial {trial} ({kind}): planes differ in {int((got != want.planes).sum())} samples")
        elif werr is not None:
            problems.append(f"trial {trial} ({kind}): oracle raised [{werr}] but the GPU decoded")
        else:
            problems.append(f"trial {trial} ({kind}): GPU raised [{type(gerr).__name__}: {gerr}] but the oracle decoded")
    print(f"{name}: {agree_ok} identical decodes, {agree_err} errors on both sides ({same_class} of the same "
          f"exception class), {len(problems)} disagreements")
    assert not problems, "\n".join(problems)


def run_gpu_rgb(blob):
    try:
        dec = J.JpegDecoder()
        dec.SetInput(blob)
        dec.Identify()
        out = np.zeros((dec.Height, dec.Width, 3), dtype=np.uint8)
        dec.SetOutputWriter(J.CudaOutputWriter(out, J.JB_OUT_RGB24))
        dec.Decode()
        return out, None
    except GPU_ERRORS as e:
        return None, e


@pytest.mark.parametrize("name", ["restart", "restart_rows_444", "plain"])
def test_corrupted_scans_render_the_same_pixels(name):
    """The same damage through the RGB24 sink, i.e. through the fast renderer (packed IDCT, 16-bit clamp, colour): wrong
    symbols give samples far outside 0..255 -- the reference stores them as int16 (wrap-around included) and its 8-bit
    writer clamps that; the pixels have to come out the same, byte for byte."""
    blob = base_streams()[name]
    rng = np.random.default_rng(1000 + sum(map(ord, name)))
    problems, same, wild = [], 0, 0
    for trial in range(70):
        bad = mutate(blob, rng, KINDS[trial % len(KINDS)])
        try:
            want = O.decode(bad, want_rgb=True)
        except O.OracleError:
            continue
        got, gerr = run_gpu_rgb(bad)
        if gerr is not None:
            problems.append(f"trial {trial}: GPU raised [{type(gerr).__name__}: {gerr}] but the oracle decoded")
            continue
        wild += int((np.abs(want.planes.astype(np.int32) - 128) > 300).any())
        wr = O.written_samples(want).all(axis=0)  # (a scan cut at a restart marker leaves the rest of the frame alone)
        if np.array_equal(got[wr], want.rgb[wr]) and not got[~wr].any():
            same += 1
        else:
            d = np.abs(got.astype(int) - want.rgb.astype(int)) * wr[:, :, None]
            problems.append(f"trial {trial}: {int((d > 0).sum())} bytes differ, max {int(d.max())}")
    print(f"{name}: {same} identical renderings ({wild} with samples far outside 0..255), {len(problems)} disagreements")
    assert same > 20 and not problems, "\n".join(problems)


def mutate_header(blob, rng, kind):
    """damage somewhere between SOI and the first bytes of the first scan: marker codes, segment lengths, frame and
    scan header fields, table definitions"""
    b = bytearray(blob)
    p = int(rng.integers(2, blob.find(b"\xff\xda") + 14))
    if kind == 0:
        b[p] ^= 1 << int(rng.integers(8))
    elif kind == 1:
        b[p] = int(rng.integers(256))
    elif kind == 2:
        del b[p:p + int(rng.integers(1, 4))]
    else:
        b[p:p] = bytes(rng.integers(0, 256, int(rng.integers(1, 4))).astype(np.uint8))
    return bytes(b)


@pytest.mark.parametrize("name", ["restart", "plain", "progressive", "lossless"])
def test_corrupted_headers_decode_like_the_reference(name):
    """The marker walk (JpegDecoder.cs:509-617, host side) and the device-side validation together must accept,
    reject and decode damaged headers like the reference algorithm."""
    blob = base_streams()[name]
    rng = np.random.default_rng(7 + sum(map(ord, name)))
    problems, agree_ok, agree_err = [], 0, 0
    for trial in range(150):
        bad = mutate_header(blob, rng, trial % 4)
        want, werr = run_oracle(bad)
        if werr is not None and "outside the oracle's scope" in str(werr):
            continue  # arithmetic-coded frame types: the oracle stops where the reference would go on
        try:
            got, gerr = run_gpu(bad)
        except (J.ArgumentException, MemoryError, ValueError) as e:  # absurd frame sizes: refused before any decode
            got, gerr = None, e
        if werr is not None and gerr is not None:
            agree_err += 1
        elif werr is None and gerr is None:
            if got.shape == want.planes.shape and np.array_equal(got, want.planes):
                agree_ok += 1
            else:
                problems.append(f"trial {trial}: planes differ")
        elif werr is not None:
            problems.append(f"trial {trial}: oracle raised [{werr}] but the GPU decoded")
        else:
            problems.append(f"trial {trial}: GPU raised [{type(gerr).__name__}: {gerr}] but the oracle decoded")
    print(f"{name}: {agree_ok} identical decodes, {agree_err} errors on both sides, {len(problems)} disagreements")
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("shape", [dict(blocks_w=6, blocks_h=4), dict(blocks_w=9, blocks_h=7, dri=5), dict(blocks_w=11, blocks_h=6, dri=11),
                                   dict(blocks_w=64, blocks_h=48)], ids=["one-segment", "dri5", "dri-rows", "self-sync"])
def test_final_code_in_the_padding_is_accepted_like_the_reference(shape):
    """DecodeHuffmanCode advances min(code size, bits available) and never fails; only magnitude bits that are not there
    are "The bit stream ended prematurely." (JpegHuffmanScanDecoder.cs:81-110).  Streams whose tables use the all-ones
    codes, cut short by 0..12 bytes at the end of the scan or of one restart interval: every verdict and every decoded
    sample must be the oracle's (restart segments: K1; no restart markers and >= 1 KiB: the self-synchronising chain)."""
    accepted = rejected = 0
    intervals = [None] if not shape.get("dri") else [None, 0, 3]
    for seed in range(3):
        for iv in intervals:
            for cut in range(0, 13):
                blob, coef = synth.handmade_grey(seed=seed, cut=cut, cut_interval=iv, **shape)
                want, werr = run_oracle(blob)
                got, gerr = run_gpu(blob)
                where = f"seed {seed} interval {iv} cut {cut}"
                if werr is None:
                    assert gerr is None, f"{where}: GPU raised [{type(gerr).__name__}: {gerr}] but the oracle decoded"
                    assert np.array_equal(got, want.planes), where
                    if cut == 0:
                        assert np.array_equal(O.scan_order_coefficients(want).reshape(-1, 64), coef)
                    accepted += 1
                else:
                    assert gerr is not None, f"{where}: oracle raised [{werr}] but the GPU decoded"
                    assert isinstance(gerr, {-1: J.InvalidDataException, -2: J.InvalidOperationException}[werr.code]), where
                    rejected += 1
    print(f"{accepted} streams accepted, {rejected} rejected, on both sides")
    assert accepted > 13 * len(intervals)  # cut streams among them


def test_streams_the_campaign_found():
    """profiles/r2_fuzz_campaign.txt: a gray progressive frame whose first scan says Ss 0 / Se 119 (a DC scan all the same:
    the reference never looks at Se there) and one whose DC table definition was swallowed by a damaged segment length
    (the reference asks for the DC table of every single-component scan with Ss = 0, the DC refinement scan included)."""
    import os
    here = os.path.join(os.path.dirname(__file__), "fixtures")
    blob = open(os.path.join(here, "fuzz_gray_progressive_dc_scan_with_se.jpg"), "rb").read()
    want = O.decode(blob, want_rgb=False)
    got, err = run_gpu(blob)
    assert err is None and np.array_equal(got, want.planes)
    blob = open(os.path.join(here, "fuzz_gray_progressive_dc_refine_without_table.jpg"), "rb").read()
    with pytest.raises(O.OracleError):
        O.decode(blob, want_rgb=False)
    got, err = run_gpu(blob)
    assert isinstance(err, J.InvalidDataException)
    # the file ends "... FF FF FF" instead of "... 5F FF D9": the reference drops an FF that has no next byte (JpegBitReader.cs
    # :113-117), the last DC refinement bits are missing and ReadBlockProgressiveDC fails ("Unexpected end of JPEG data
    # stream."); a zero behind the stream in the device arena had turned that FF into a stuffed data byte
    blob = open(os.path.join(here, "fuzz_progressive_trailing_ff_without_eoi.jpg"), "rb").read()
    with pytest.raises(O.OracleError):
        O.decode(blob, want_rgb=False)
    got, err = run_gpu(blob)
    assert isinstance(err, J.InvalidDataException)


def test_a_stream_with_two_defects_raises_what_the_reference_meets_first():
    """Restart-coded baseline frame with one truncated interval (the bit stream ends prematurely: InvalidDataException) and
    one restart marker overwritten by an SOI (InvalidOperationException, "Expect restart marker."): the reference
    stops at whichever comes first in the stream.  The kernels decode all intervals at once; every failure takes part in an
    atomicMin on its place in the stream (jb_report_error) and the host reports the class of the smallest."""
    base = synth.encode_jpeg(synth.synth_rgb(44, 320, 240), quality=85, subsampling="4:2:0", restart_rows=1)
    sos = base.find(b"\xff\xda")
    rst = [i for i in range(sos, len(base) - 1) if base[i] == 0xFF and 0xD0 <= base[i + 1] <= 0xD7]
    assert len(rst) >= 13
    seen = []
    for marker_at, short_at in ((3, 8), (8, 3), (5, 5), (2, 12), (12, 2), (6, 7), (7, 6)):
        b = bytearray(base)
        b[rst[marker_at] + 1] = 0xD8                                   # an SOI where an RSTn is due
        lo, hi = rst[short_at] + 2, rst[short_at + 1]
        del b[lo + 3:hi]                                               # (behind the marker edit: its position stays valid
        blob = bytes(b)                                                #  only when short_at > marker_at, so redo it)
        if short_at < marker_at:
            b = bytearray(base)
            del b[lo + 3:hi]
            b[rst[marker_at] + 1 - (hi - lo - 3)] = 0xD8
            blob = bytes(b)
        want, werr = run_oracle(blob)
        got, gerr = run_gpu(blob)
        assert werr is not None and gerr is not None, (marker_at, short_at)
        assert isinstance(gerr, J.InvalidDataException if werr.code == -1 else J.InvalidOperationException), (marker_at, short_at, werr)
        seen.append(werr.code)
    assert set(seen) == {-1, -2}, seen
