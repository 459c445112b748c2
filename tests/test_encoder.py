"""Encoder path (SURVEY 8a E1-E9, configs[4]).  The reference has NO encoder tests: parity is
oracle-vs-GPU (bit-exact coefficients, histograms, tables and scan bytes) with the oracle anchored on
round trips through the pinned decoder and libjpeg-turbo."""
import ctypes as C
import io

import numpy as np
import pytest

import jpeglibrary_b200 as J
import oracle_ffi as O
import synth


# ----------------------------------------------------------------------------- CPU: oracle + host builder
def test_oracle_encoder_round_trips_through_the_pinned_decoder():
    rgb = synth.synth_rgb(2, 256, 192)
    ycc = O.rgb_to_ycbcr(rgb)
    e = O.encode_ycbcr(ycc, quality=75)
    d = O.decode(e.bytes)
    assert (d.width, d.height, d.ncomp) == (256, 192, 3)
    for c in range(3):
        assert np.array_equal(d.coef[c][:d.alloc_h[c], :d.alloc_w[c]], e.coef[c])  # entropy coding is lossless
    psnr = 10 * np.log10(255 ** 2 / np.mean((d.rgb.astype(float) - rgb) ** 2))
    assert psnr > 27
    from PIL import Image
    img = Image.open(io.BytesIO(e.bytes))  # libjpeg-turbo accepts the stream; its raw Y plane differs by IDCT rounding only
    img.draft("YCbCr", img.size)
    im = np.array(img)[..., 0]
    assert np.abs(im.astype(int) - d.ycbcr[..., 0].astype(int)).max() <= 2


def test_rgb_to_ycbcr_constants():
    """E1 evaluated in fp32 like the C#: 19595/38470/7471, 11058/21710 (not libjpeg's 11059/21709)."""
    px = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 255], [0, 0, 0], [12, 200, 77]], dtype=np.uint8)
    out = O.rgb_to_ycbcr(px)
    for (r, g, b), (y, cb, cr) in zip(px.astype(int), out.astype(int)):
        assert y == (19595 * r + 38470 * g + 7471 * b + 32768) >> 16
        assert cb == ((-11058 * r - 21710 * g + 32768 * b + (128 << 16) + 32767) >> 16) & 255
        assert cr == ((32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16) & 255


def kraft(bits):
    return sum(int(n) * 2.0 ** -(l + 1) for l, n in enumerate(bits))


@pytest.mark.parametrize("seed", range(6))
def test_host_table_builder_equals_oracle_builder(seed):
    """Same histogram -> identical DHT (code lengths AND symbol order) from the product's builder
    (jb_build_huffman_table, also the GPU kernel's code) and the oracle's independent restatement."""
    rng = np.random.default_rng(seed)
    n = [3, 12, 40, 120, 200, 256][seed]
    freq = np.zeros(256, dtype=np.uint32)
    idx = rng.choice(256, size=n, replace=False)
    freq[idx] = (rng.pareto(0.7, size=n) * 10 + 1).astype(np.uint32)  # heavy tail -> lengths > 16 get limited
    if seed == 3:
        freq[idx] = 1  # all ties: exercises the lowest-index tie rule and the unstable sort
    bits, vals = O.build_huffman_table(freq)
    spec = J._native.HuffSpec()
    assert J._native.cuda.jb_build_huffman_table(freq.ctypes.data, 1, 0, C.byref(spec)) == 0
    assert list(spec.bits) == bits.tolist()
    assert list(spec.values[:spec.value_count]) == vals.tolist()
    assert sorted(vals.tolist()) == sorted(np.nonzero(freq)[0].tolist())
    assert kraft(bits) < 1.0  # one code point is reserved (no all-ones code)
    assert max(l + 1 for l, c in enumerate(bits) if c) <= 16


def _cost(freq, bits, vals):
    lens, k = {}, 0
    for l, c in enumerate(bits):
        for _ in range(int(c)):
            lens[int(vals[k])] = l + 1
            k += 1
    return sum(int(freq[v]) * l for v, l in lens.items())


@pytest.mark.parametrize("seed", range(10))
def test_package_merge_builder_equals_oracle_builder(seed):
    """MostOptimalCoding = true (JpegHuffmanEncodingTableBuilder.BuildUsingPackageMerge :287-413): the product's host
    builder (jb_build_huffman_table_optimal) and the oracle's restatement give the same DHT, code lengths and symbol
    order (four unstable sorts decide it); the code is complete minus the sentinel's code point, at most 16 bits long
    and never costs more than the standard method's."""
    rng = np.random.default_rng(100 + seed)
    n = [1, 2, 3, 12, 17, 40, 120, 200, 256, 256][seed]
    freq = np.zeros(256, dtype=np.uint32)
    idx = rng.choice(256, size=n, replace=False)
    freq[idx] = (rng.pareto(0.7, size=n) * 10 + 1).astype(np.uint32)
    if seed in (4, 8):
        freq[idx] = 1 + (np.arange(n) % 3)  # many ties
    bits, vals = O.build_huffman_table(freq, optimal=True)
    spec = J._native.HuffSpec()
    assert J._native.cuda.jb_build_huffman_table_optimal(freq.ctypes.data, 1, 2, C.byref(spec)) == 0
    assert (spec.table_class, spec.identifier) == (1, 2)
    assert list(spec.bits) == bits.tolist()
    assert list(spec.values[:spec.value_count]) == vals.tolist()
    assert sorted(vals.tolist()) == sorted(np.nonzero(freq)[0].tolist())
    assert int(bits.sum()) == n and kraft(bits) < 1.0
    assert max(l + 1 for l, c in enumerate(bits) if c) <= 16
    sbits, svals = O.build_huffman_table(freq)
    assert _cost(freq, bits, vals) <= _cost(freq, sbits, svals)
    f = [int(freq[v]) for v in vals]  # canonical order: code size ascending, frequency descending inside a size
    k = 0
    for c in bits:
        grp = f[k:k + int(c)]
        assert grp == sorted(grp, reverse=True)
        k += int(c)


def test_oracle_encoder_with_package_merge_tables_round_trips():
    rgb = synth.synth_rgb(3, 200, 136)
    e = O.encode_ycbcr(O.rgb_to_ycbcr(rgb), quality=75, optimal=True)
    d = O.decode(e.bytes)  # (not necessarily a shorter FILE than the standard method's: byte stuffing depends on the codes)
    for c in range(3):
        assert np.array_equal(d.coef[c][:d.alloc_h[c], :d.alloc_w[c]], e.coef[c])


def test_builder_rejects_empty_histogram():
    spec = J._native.HuffSpec()
    assert J._native.cuda.jb_build_huffman_table(np.zeros(256, np.uint32).ctypes.data, 0, 0, C.byref(spec)) == J._native.JB_ERR_INVALID_OPERATION


def test_encoder_argument_errors():
    enc = J.JpegEncoder()
    with pytest.raises(J.InvalidOperationException):
        enc.Encode()
    enc.SetQuantizationTable(J.JpegStandardQuantizationTable.GetLuminanceTable(0, 0))
    enc.SetHuffmanTable(True, 0)
    enc.SetHuffmanTable(False, 0)
    with pytest.raises(J.ArgumentException):
        enc.AddComponent(1, 0, 0, 0, 3, 1)       # "Subsampling factor can only be 1, 2 or 4."
    with pytest.raises(J.ArgumentException):
        enc.AddComponent(1, 1, 0, 0, 1, 1)       # "Quantization table is not defined."
    with pytest.raises(J.ArgumentException):
        enc.AddComponent(1, 0, 1, 0, 1, 1)       # "Huffman table is not defined."
    enc.AddComponent(1, 0, 0, 0, 1, 1)
    with pytest.raises(J.ArgumentException):
        enc.AddComponent(1, 0, 0, 0, 1, 1)       # component index already used
    with pytest.raises(J.ArgumentException):
        J.JpegStandardQuantizationTable.ScaleByQuality(J.JpegStandardQuantizationTable.GetLuminanceTable(0, 0), 101)
    q = J.JpegStandardQuantizationTable.ScaleByQuality(J.JpegStandardQuantizationTable.GetChrominanceTable(0, 1), 75)
    assert q.Elements == O.std_quant_table(True, 75).tolist()


# ----------------------------------------------------------------------------- GPU parity
def scan_order(e, shape, subsampling):
    """oracle allocator planes -> MCU scan order (the device store layout)"""
    H, W = shape
    hs, vs = subsampling
    mx, my = (W + 8 * hs - 1) // (8 * hs), (H + 8 * vs - 1) // (8 * vs)
    parts = []
    for c, a in enumerate(e.coef):
        h, v = (hs, vs) if c == 0 else (1, 1)
        if a.shape[:2] != (my * v, mx * h):
            # MCU-padding blocks alias the allocator's dummy block (JpegBlockAllocator.cs:108-111): in scan order every
            # one of them reads what the last TransformBlocks call left there
            full = np.empty((my * v, mx * h, 64), np.int16)
            full[:] = e.dummy
            full[:a.shape[0], :a.shape[1]] = a
            a = full
        parts.append(a.reshape(my, v, mx, h, 64).transpose(0, 2, 1, 3, 4).reshape(mx * my, v * h, 64))
    return np.concatenate(parts, axis=1).reshape(-1, 64)


ENC_SHAPES = [
    dict(width=256, height=192, subsampling=(2, 2), quality=75),
    dict(width=333, height=211, subsampling=(1, 1), quality=90),   # ragged edge: zero padding inside blocks
    dict(width=208, height=120, subsampling=(2, 1), quality=60),
    dict(width=96, height=112, subsampling=(1, 2), quality=85),
    dict(width=640, height=480, subsampling=(2, 2), quality=30),
    dict(width=1920, height=1088, subsampling=(2, 2), quality=95),
    # frames whose luma block grid is not a whole number of MCUs: padding blocks alias the allocator's dummy block
    # (JpegBlockAllocator.cs:108-111, JpegEncoder.cs:458-470, 551-597, 640-647)
    dict(width=1920, height=1080, subsampling=(2, 2), quality=75),  # 135 block rows
    dict(width=24, height=16, subsampling=(2, 2), quality=75),     # 3 block columns
    dict(width=40, height=24, subsampling=(2, 2), quality=50),     # both
    dict(width=72, height=40, subsampling=(2, 1), quality=80),
    dict(width=48, height=56, subsampling=(1, 2), quality=80),
    dict(width=3, height=5, subsampling=(2, 2), quality=75),       # one MCU, three of four luma blocks are padding
]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", ENC_SHAPES, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_gpu_encoder_is_bit_identical_to_the_oracle(kw):
    rgb = synth.synth_rgb(31, kw["width"], kw["height"])
    want = O.encode_ycbcr(O.rgb_to_ycbcr(rgb), quality=kw["quality"], subsampling=kw["subsampling"])
    got, enc = J.encode_rgb(rgb, quality=kw["quality"], subsampling=kw["subsampling"])
    assert np.array_equal(enc.last_coefficients, scan_order(want, rgb.shape[:2], kw["subsampling"]))  # E1-E5
    for s in enc.last_tables:                                                                       # E6 + E7
        bits, vals = want.dht[(s.table_class, s.identifier)]
        assert list(s.bits) == bits.tolist() and list(s.values[:s.value_count]) == vals.tolist()
    assert got == want.bytes                                                                        # E8 + E9, headers
    d = O.decode(got)                                                                               # re-decodes under the reference restatement
    assert (d.width, d.height) == (kw["width"], kw["height"])
    from PIL import Image
    img = Image.open(io.BytesIO(got))                                                               # and under libjpeg-turbo
    img.draft("YCbCr", img.size)
    assert np.abs(np.array(img)[..., 0].astype(int) - d.ycbcr[..., 0].astype(int)).max() <= 2


@pytest.mark.gpu
def test_gpu_encoder_4k_frame_is_bit_identical_to_the_oracle():
    """configs[4] at its full frame size (3840x2160 RGB -> q75 4:2:0 with optimised tables)."""
    rgb = synth.synth_rgb(1004, 3840, 2160)
    want = O.encode_ycbcr(O.rgb_to_ycbcr(rgb), quality=75)
    got, enc = J.encode_rgb(rgb, quality=75)
    assert np.array_equal(enc.last_coefficients, scan_order(want, rgb.shape[:2], (2, 2)))
    assert got == want.bytes


@pytest.mark.gpu
def test_gpu_encoder_host_builder_path_gives_the_same_stream():
    """The C# integration keeps the table build on the host (reference builder fed by GPU histograms)."""
    rgb = synth.synth_rgb(8, 320, 240)
    a, _ = J.encode_rgb(rgb, quality=75, host_builder=False)
    b, _ = J.encode_rgb(rgb, quality=75, host_builder=True)
    assert a == b


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(320, 240), (1920, 1080)])
def test_gpu_encoder_most_optimal_coding_is_bit_identical_to_the_oracle(shape):
    """JpegEncoder.MostOptimalCoding = true: K3b's histograms -> package-merge tables on the host -> K4."""
    rgb = synth.synth_rgb(12, *shape)
    want = O.encode_ycbcr(O.rgb_to_ycbcr(rgb), quality=75, optimal=True)
    got, enc = J.encode_rgb(rgb, quality=75, most_optimal=True)
    for s in enc.last_tables:
        bits, vals = want.dht[(s.table_class, s.identifier)]
        assert list(s.bits) == bits.tolist() and list(s.values[:s.value_count]) == vals.tolist()
    assert got == want.bytes


@pytest.mark.gpu
def test_gpu_encode_then_gpu_decode_round_trip():
    rgb = synth.synth_rgb(5, 512, 384)
    blob, _ = J.encode_rgb(rgb, quality=85)
    out = np.zeros((384, 512, 3), np.uint8)
    dec = J.JpegDecoder()
    dec.SetInput(blob)
    dec.SetOutputWriter(J.CudaOutputWriter(out))
    dec.Decode()
    assert np.array_equal(out, O.decode(blob).rgb)
    assert 10 * np.log10(255 ** 2 / np.mean((out.astype(float) - rgb) ** 2)) > 30


@pytest.mark.gpu
def test_gpu_encoder_compatibility_reader():
    class Reader(J.JpegBlockInputReader):  # apps/JpegEncode/JpegBufferInputReader.cs
        def __init__(self, ycc):
            self.ycc, (self.Height, self.Width) = ycc, ycc.shape[:2]

        def ReadBlock(self, blockRef, componentIndex, x, y):
            blockRef[:] = 0
            t = self.ycc[y:y + 8, x:x + 8, componentIndex]
            blk = blockRef.reshape(8, 8)
            blk[:t.shape[0], :t.shape[1]] = t

    rgb = synth.synth_rgb(6, 64, 48)
    ycc = O.rgb_to_ycbcr(rgb)
    enc = J.JpegEncoder()
    enc.SetQuantizationTable(J.JpegStandardQuantizationTable.ScaleByQuality(J.JpegStandardQuantizationTable.GetLuminanceTable(0, 0), 75))
    enc.SetQuantizationTable(J.JpegStandardQuantizationTable.ScaleByQuality(J.JpegStandardQuantizationTable.GetChrominanceTable(0, 1), 75))
    for isdc, ident in ((True, 0), (False, 0), (True, 1), (False, 1)):
        enc.SetHuffmanTable(isdc, ident)
    enc.AddComponent(1, 0, 0, 0, 2, 2)
    enc.AddComponent(2, 1, 1, 1, 1, 1)
    enc.AddComponent(3, 1, 1, 1, 1, 1)
    enc.SetInputReader(Reader(ycc))
    out = bytearray()
    enc.SetOutput(out)
    enc.Encode()
    assert bytes(out) == O.encode_ycbcr(ycc, quality=75).bytes
    # 24 px wide 4:2:0: three luma block columns -> one MCU-padding block per MCU row, read through the same reader
    ycc = O.rgb_to_ycbcr(synth.synth_rgb(1, 24, 16))
    enc.SetInputReader(Reader(ycc))
    out = bytearray()
    enc.SetOutput(out)
    enc.Encode()
    assert bytes(out) == O.encode_ycbcr(ycc, quality=75).bytes

    class Odd(Reader):  # a reader that invents samples outside its own frame: not what the GPU path computes
        def ReadBlock(self, blockRef, componentIndex, x, y):
            super().ReadBlock(blockRef, componentIndex, x, y)
            if x >= self.Width:
                blockRef[:] = 7

    enc.SetInputReader(Odd(ycc))
    with pytest.raises(J.NotSupportedException):
        enc.Encode()


def test_fp64_reciprocal_quantisation_equals_ieee_fp32_division():
    """K3 quantises with RN32((double)F * RN64(1 / (8q))) instead of fp32 division (k_encode.cuh).  The identity
    RN32(F * 0.125 / q) == that expression is argued in the kernel's comment; this checks it numerically (numpy's
    float32 division is IEEE) on random, exactly divisible and near-half-integer dividends for every 8-bit
    quantiser and a few 16-bit ones."""
    rng = np.random.default_rng(7)
    for q in list(range(1, 256)) + [256, 257, 1000, 4095, 32769, 65535]:
        r8 = np.float64(0.125) / np.float64(q)
        x = (rng.standard_normal(40000) * rng.choice([1, 10, 100, 1000, 8000], 40000)).astype(np.float32)
        k = rng.integers(-40000, 40000, 10000)
        xs = np.concatenate([x, (k * q * 8).astype(np.float32), ((k + 0.5) * q * 8).astype(np.float32),
                             (k * q * 8).astype(np.float32) + np.float32(0.5)])
        ref = (xs * np.float32(0.125)) / np.float32(q)
        got = (xs.astype(np.float64) * r8).astype(np.float32)
        assert np.array_equal(ref, got), q


@pytest.mark.gpu
def test_caller_tables_on_noise_outgrow_the_reserved_stream_space():
    """Annex-K tables installed by the caller (JpegEncoder.SetHuffmanTable(isDc, id, table)) on noise at quality 100 need
    more than the 768 bits per block that are reserved up front: the pack stage reserves what the bit totals ask for
    and runs again, instead of failing a frame the reference encodes."""
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    parsed = J.Parsed(synth.encode_jpeg(rgb[:64, :64], quality=90, subsampling="4:4:4"))  # libjpeg writes the Annex-K tables
    std, specs = parsed.desc, {}
    for i in range(std.table_count):
        s = J._native.HuffSpec()
        C.memmove(C.byref(s), C.byref(std.tables[i]), C.sizeof(s))  # (the descriptor's memory belongs to `parsed`)
        specs[(s.table_class, s.identifier)] = s
    assert sorted(specs) == [(0, 0), (0, 1), (1, 0), (1, 1)]
    for ident in (0, 1):  # a caller may install any table: these give the most frequent symbols the 16-bit codes
        t = specs[(1, ident)]
        vals = list(t.values[:t.value_count])[::-1]
        for i, v in enumerate(vals):
            t.values[i] = v
    enc = J.JpegEncoder()
    enc.SetQuantizationTable(J.JpegStandardQuantizationTable.ScaleByQuality(J.JpegStandardQuantizationTable.GetLuminanceTable(0, 0), 100))
    enc.SetQuantizationTable(J.JpegStandardQuantizationTable.ScaleByQuality(J.JpegStandardQuantizationTable.GetChrominanceTable(0, 1), 100))
    for (cls, ident), spec in specs.items():
        enc.SetHuffmanTable(cls == 0, ident, spec)
    enc.AddComponent(1, 0, 0, 0, 1, 1)
    enc.AddComponent(2, 1, 1, 1, 1, 1)
    enc.AddComponent(3, 1, 1, 1, 1, 1)
    enc.SetInputReader(J.CudaInputReader(rgb, format=J.JB_IN_RGB24))
    out = bytearray()
    enc.SetOutput(out)
    enc.Encode()
    nblk = 3 * 64 * 64
    assert len(out) > nblk * 96 + 4096                       # more than was reserved at create time
    d = O.decode(bytes(out))
    assert np.array_equal(O.scan_order_coefficients(d).reshape(-1, 64), enc.last_coefficients)
    from PIL import Image
    assert Image.open(io.BytesIO(bytes(out))).size == (512, 512)


@pytest.mark.parametrize("optimal", [False, True], ids=["standard", "package-merge"])
def test_table_builders_agree_on_random_histograms(optimal):
    """Differential campaign on the CPU: 1500 histograms of every shape that came to mind (heavy tails that need the
    16-bit limit, long runs of ties for the unstable sorts, one / two symbols, counts next to 2^31, geometric decays) --
    the product's host builders and the oracle's restatements write the same DHT."""
    rng = np.random.default_rng(4242 + optimal)
    fn = J._native.cuda.jb_build_huffman_table_optimal if optimal else J._native.cuda.jb_build_huffman_table
    crashes = 0
    for trial in range(1500):
        n = int(rng.choice([1, 2, 3, 5, 9, 17, 33, 64, 100, 162, 255, 256]))
        freq = np.zeros(256, dtype=np.uint32)
        idx = rng.choice(256, size=n, replace=False)
        kind = trial % 6
        if kind == 0:
            f = rng.pareto(0.5, size=n) * 5 + 1
        elif kind == 1:
            f = rng.integers(1, 4, size=n)                      # ties everywhere
        elif kind == 2:
            f = 2.0 ** rng.integers(0, 31, size=n)              # powers of two up to 2^30
        elif kind == 3:
            f = np.maximum(1, (1 << 20) * 0.6 ** np.arange(n))  # geometric: code lengths want to exceed 16
        elif kind == 4:
            f = rng.integers(1, 1 << 16, size=n)
        else:
            f = np.full(n, int(rng.integers(1, 1000)))           # all equal
        freq[idx] = np.minimum(f, 2 ** 31 - 1).astype(np.uint32)
        spec = J._native.HuffSpec()
        try:
            bits, vals = O.build_huffman_table(freq, optimal=optimal)
        except O.OracleError:
            # 255 symbols of equal weight + the sentinel = 256 codes of 8 bits: the reference counts them in a byte, finds
            # no code at all and dies of an IndexOutOfRangeException (standard method only)
            assert not optimal and n == 255 and kind == 5, (trial, kind, n)
            assert fn(freq.ctypes.data, 0, 1, C.byref(spec)) == J._native.JB_ERR_INVALID_OPERATION
            crashes += 1
            continue
        assert fn(freq.ctypes.data, 0, 1, C.byref(spec)) == 0, trial
        assert list(spec.bits) == bits.tolist(), (trial, kind, n)
        assert list(spec.values[:spec.value_count]) == vals.tolist(), (trial, kind, n)
        assert int(bits.sum()) == n and kraft(bits) < 1.0 and max(l + 1 for l, c in enumerate(bits) if c) <= 16, (trial, kind, n)
    assert optimal or crashes > 0
