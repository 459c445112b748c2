"""Deterministic synthetic inputs (SURVEY section 8d / BASELINE.md section 2).

image i <- np.random.default_rng(1000+i): sum of low-frequency sinusoids (periods 29..97 px, random
phase) + 8 random filled rectangles + Gaussian noise sigma=6, clipped to uint8 RGB.  JPEGs are
produced with Pillow/libjpeg-turbo (data generation only -- never on a measured or checked path).
"""
import io

import numpy as np


def synth_rgb(i, width, height):
    rng = np.random.default_rng(1000 + i)
    xs = np.arange(width, dtype=np.float32)
    ys = np.arange(height, dtype=np.float32)
    img = np.empty((height, width, 3), dtype=np.float32)
    for c in range(3):
        acc = np.full((height, width), 128.0, dtype=np.float32)
        for _ in range(3):
            px, py = rng.uniform(29, 97, size=2)
            ph = rng.uniform(0, 2 * np.pi)
            amp = np.float32(rng.uniform(15, 40))
            ax = (2 * np.pi / px) * xs + ph
            by = (2 * np.pi / py) * ys
            # sin(ax + by) = sin(ax)cos(by) + cos(ax)sin(by): two rank-1 updates instead of W*H sines
            acc += amp * (np.outer(np.cos(by), np.sin(ax)) + np.outer(np.sin(by), np.cos(ax))).astype(np.float32)
        img[..., c] = acc
    for _ in range(8):
        x0 = int(rng.integers(0, max(1, width - 8)))
        y0 = int(rng.integers(0, max(1, height - 8)))
        w = int(rng.integers(8, max(9, width // 4)))
        h = int(rng.integers(8, max(9, height // 4)))
        col = rng.uniform(0, 255, size=3).astype(np.float32)
        img[y0:y0 + h, x0:x0 + w, :] = col
    img += rng.standard_normal(size=img.shape, dtype=np.float32) * np.float32(6)
    return np.clip(img, 0, 255).astype(np.uint8)


def encode_jpeg(rgb, quality=85, subsampling="4:2:0", restart_rows=0, restart_blocks=0, progressive=False,
                optimize=False, gray=False):
    from PIL import Image
    im = Image.fromarray(rgb if not gray else rgb[..., 0], "L" if gray else "RGB")
    buf = io.BytesIO()
    kw = dict(format="JPEG", quality=quality, progressive=progressive, optimize=optimize)
    if not gray:
        kw["subsampling"] = subsampling
    if restart_rows:
        kw["restart_marker_rows"] = restart_rows
    if restart_blocks:
        kw["restart_marker_blocks"] = restart_blocks
    im.save(buf, **kw)
    return buf.getvalue()


def synth_jpeg(i, width, height, **kw):
    return encode_jpeg(synth_rgb(i, width, height), **kw)


# --------------------------------------------------------------------------------------------------
# Lossless (SOF3) test streams.  Pillow cannot write them, so this is a small encoder that inverts the
# decoder rules of ScanDecoder/JpegHuffmanLosslessScanDecoder.cs:84-178 (prediction from already coded
# samples, first-row / first-column / after-restart special cases, modulo-2^16 differences, category 16
# = 32768 without extra bits).  Data generation only; the oracle is the checker.
# --------------------------------------------------------------------------------------------------
def _lossless_predict(predictor, ra, rb, rc):
    if predictor == 1:
        return ra
    if predictor == 2:
        return rb
    if predictor == 3:
        return rc
    if predictor == 4:
        return ra + rb - rc
    if predictor == 5:
        return ra + ((rb - rc) >> 1)
    if predictor == 6:
        return rb + ((ra - rc) >> 1)
    if predictor == 7:
        return (ra + rb) >> 1
    return 0


def _s16(v):
    v &= 0xFFFF
    return v - 0x10000 if v & 0x8000 else v


def encode_lossless(planes, precision=8, predictor=1, point_transform=0, sampling=None, restart=0, scan_components=None):
    """planes: list of 2-D integer arrays at COMPONENT resolution (already divided by 2^Pt), component c of
    size (mcus_y * v_c, mcus_x * h_c).  Returns a complete SOF3 stream with one interleaved scan (over
    `scan_components` when given: the other frame components are not coded)."""
    n = len(planes)
    in_scan = list(range(n)) if scan_components is None else list(scan_components)
    sampling = sampling or [(1, 1)] * n
    hmax = max(h for h, _ in sampling)
    vmax = max(v for _, v in sampling)
    mcus_y = planes[0].shape[0] // sampling[0][1]
    mcus_x = planes[0].shape[1] // sampling[0][0]
    width, height = mcus_x * hmax, mcus_y * vmax
    for p, (h, v) in zip(planes, sampling):
        assert p.shape == (mcus_y * v, mcus_x * h)
    # one table for every component: categories 0..16, all 5-bit codes except the last two (6 bits)
    bits = [0, 0, 0, 0, 15, 2] + [0] * 10
    vals = list(range(17))
    codes, code, k = {}, 0, 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            codes[vals[k]] = (code, ln)
            code += 1
            k += 1
        code <<= 1
    initial = 1 << (precision - point_transform - 1)
    rec = [np.zeros(p.shape, dtype=np.int64) for p in planes]  # what the decoder will hold (as int16)
    out = bytearray()
    acc = nbits = 0

    def put(value, length):
        nonlocal acc, nbits
        acc = (acc << length) | (value & ((1 << length) - 1))
        nbits += length
        while nbits >= 8:
            b = (acc >> (nbits - 8)) & 0xFF
            out.append(b)
            if b == 0xFF:
                out.append(0)
            nbits -= 8
        acc &= (1 << nbits) - 1

    def flush():
        nonlocal acc, nbits
        if nbits:
            put((1 << (8 - nbits)) - 1, 8 - nbits)

    before, rst = restart, 0
    for row in range(mcus_y):
        for col in range(mcus_x):
            for c in in_scan:
                h, v = sampling[c]
                for y in range(v):
                    cy = row * v + y
                    for x in range(h):
                        cx = col * h + x
                        r = rec[c]
                        if row == 0 or (restart > 0 and before == restart):
                            if col == 0 and x == 0:
                                pred = initial
                            else:
                                ra = r[cy, cx - 1]
                                rb = initial if y == 0 else r[cy - 1, cx]
                                rc = initial if y == 0 else r[cy - 1, cx - 1]
                                pred = _lossless_predict(predictor, ra, rb, rc)
                        elif col == 0:
                            pred = r[cy - 1, cx]
                        else:
                            pred = _lossless_predict(predictor, r[cy, cx - 1], r[cy - 1, cx], r[cy - 1, cx - 1])
                        want = _s16(int(planes[c][cy, cx]))
                        diff = (want - int(pred)) & 0xFFFF
                        if diff == 0x8000:
                            put(*codes[16])
                        else:
                            d = _s16(diff)
                            cat = abs(d).bit_length()
                            put(*codes[cat])
                            if cat:
                                put(d if d >= 0 else d - 1, cat)
                        r[cy, cx] = want
            if restart > 0:
                before -= 1
                if before == 0 and not (row == mcus_y - 1 and col == mcus_x - 1):
                    flush()
                    out += bytes([0xFF, 0xD0 + (rst & 7)])
                    rst += 1
                    before = restart
    flush()

    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload

    s = bytearray(b"\xff\xd8")
    sof = bytes([precision]) + height.to_bytes(2, "big") + width.to_bytes(2, "big") + bytes([n])
    for c, (h, v) in enumerate(sampling):
        sof += bytes([c + 1, (h << 4) | v, 0])
    s += seg(0xC3, sof)
    s += seg(0xC4, bytes([0x00]) + bytes(bits) + bytes(vals))
    if restart:
        s += seg(0xDD, restart.to_bytes(2, "big"))
    sos = bytes([len(in_scan)])
    for c in in_scan:
        sos += bytes([c + 1, 0x00])
    s += seg(0xDA, sos + bytes([predictor, 0, point_transform]))
    s += out + b"\xff\xd9"
    return bytes(s)


def synth_lossless(i, width, height, precision=8, predictor=1, point_transform=0, sampling=None, restart=0, ncomp=3,
                   scan_components=None):
    """SOF3 stream of synthetic content; width/height are at full resolution and must be MCU multiples."""
    rng = np.random.default_rng(2000 + i)
    sampling = sampling or [(1, 1)] * ncomp
    hmax = max(h for h, _ in sampling)
    vmax = max(v for _, v in sampling)
    assert width % hmax == 0 and height % vmax == 0
    base = synth_rgb(i, width, height).astype(np.int64)
    planes = []
    for c, (h, v) in enumerate(sampling):
        p = base[::vmax // v, ::hmax // h, c % 3]
        p = (p << max(0, precision - 8)) >> point_transform if precision >= 8 else p >> (8 - precision + point_transform)
        if precision > 8:
            p = p + rng.integers(0, 1 << (precision - 8 - point_transform), size=p.shape)
        planes.append(np.ascontiguousarray(p))
    blob = encode_lossless(planes, precision, predictor, point_transform, sampling, restart, scan_components)
    if scan_components is not None:   # components the scan does not name are never written: zeros
        planes = [p if c in scan_components else np.zeros_like(p) for c, p in enumerate(planes)]
    # what a decoder must return: the coded samples as int16, replicated to full resolution
    full = np.stack([np.repeat(np.repeat(p, vmax // v, axis=0), hmax // h, axis=1)
                     for p, (h, v) in zip(planes, sampling)]).astype(np.uint16).view(np.int16)
    return blob, full


def synth_lossless_scans(i, width, height, scans, precision=8, sampling=None, ncomp=3, restart=0):
    """SOF3 frame coded as SEVERAL scans (the reference decodes scan by scan into one scanline store,
    JpegHuffmanLosslessScanDecoder.ProcessScan :52-205).  scans: list of dicts with `components` (frame component
    indices) and optionally `predictor`.  A component named by two scans is coded twice, with different content; the
    later scan wins.  One restart interval for the whole frame (the reference's scan decoder reads it once, in its
    constructor: JpegHuffmanLosslessScanDecoder.cs:32).  Returns (stream, expected int16 planes at full resolution)."""
    sampling = sampling or [(1, 1)] * ncomp
    stream, final = None, {}
    for k, sc in enumerate(scans):
        blob, full = synth_lossless(i + 17 * k, width, height, precision=precision, predictor=sc.get("predictor", 1),
                                    sampling=sampling, restart=restart, ncomp=ncomp, scan_components=sc["components"])
        for c in sc["components"]:
            final[c] = full[c]
        if stream is None:
            stream = bytearray(blob[:-2])                     # SOI, SOF3, DHT, [DRI], SOS, data
        else:
            stream += blob[blob.index(b"\xff\xda"):-2]        # SOS header + entropy-coded data
    stream += b"\xff\xd9"
    want = np.stack([final.get(c, np.zeros((height, width), np.int16)) for c in range(ncomp)])
    return bytes(stream), want


# --------------------------------------------------------------------------------------------------
# Sequential frames with several scans (or scans over some of the components).  Pillow only writes one
# interleaved scan, so an existing stream is re-sequenced: the quantised coefficients (from the oracle)
# are entropy-coded again, scan by scan, with the stream's own Huffman tables, each scan in the order
# the REFERENCE walks it -- MCU by MCU over the frame's MCU grid with the component's own h x v blocks,
# whatever the number of components in the scan (JpegHuffmanBaselineScanDecoder.cs:99-137, quirk Q2;
# the same as the standard's order only when every component is sampled 1x1).  Data generation only.
# --------------------------------------------------------------------------------------------------
def resequence_scans(blob, decoded, scans, restart=0):
    """blob: a baseline JPEG with one interleaved scan; decoded: oracle_ffi.decode(blob); scans: list of lists
    of frame component indices.  Returns the same frame coded as len(scans) scans."""
    def segments(data):
        i, out = 2, []
        while data[i + 1] != 0xDA:
            ln = int.from_bytes(data[i + 2:i + 4], "big")
            out.append((data[i + 1], data[i + 4:i + 2 + ln]))
            i += 2 + ln
        ln = int.from_bytes(data[i + 2:i + 4], "big")
        return out, data[i + 4:i + 2 + ln]

    segs, sos = segments(blob)
    codes = {}   # (class, id) -> {symbol: (code, length)}
    comp_ids, sel = [], {}
    for m, p in segs:
        if m == 0xC4:
            q = 0
            while q < len(p):
                tc, th = p[q] >> 4, p[q] & 15
                bits = list(p[q + 1:q + 17])
                vals = list(p[q + 17:q + 17 + sum(bits)])
                q += 17 + sum(bits)
                tab, code, k = {}, 0, 0
                for ln in range(1, 17):
                    for _ in range(bits[ln - 1]):
                        tab[vals[k]] = (code, ln)
                        code += 1
                        k += 1
                    code <<= 1
                codes[(tc, th)] = tab
        elif m in (0xC0, 0xC1):
            n = p[5]
            comp_ids = [p[6 + 3 * c] for c in range(n)]
    for i in range(sos[0]):
        sel[sos[1 + 2 * i]] = sos[2 + 2 * i]          # component id -> Td/Ta byte of the original scan
    d = decoded
    out = bytearray(b"\xff\xd8")
    for m, p in segs:
        if m != 0xDD:
            out += bytes([0xFF, m]) + (len(p) + 2).to_bytes(2, "big") + p
    if restart:
        out += b"\xff\xdd\x00\x04" + restart.to_bytes(2, "big")

    def size_of(v):
        return abs(int(v)).bit_length()

    for comps in scans:
        out += b"\xff\xda" + (6 + 2 * len(comps)).to_bytes(2, "big") + bytes([len(comps)])
        for c in comps:
            out += bytes([comp_ids[c], sel[comp_ids[c]]])
        out += bytes([0, 63, 0])
        acc = nbits = 0
        data = bytearray()

        def put(value, length):
            nonlocal acc, nbits
            acc = (acc << length) | (value & ((1 << length) - 1))
            nbits += length
            while nbits >= 8:
                b = (acc >> (nbits - 8)) & 0xFF
                data.append(b)
                if b == 0xFF:
                    data.append(0)
                nbits -= 8
            acc &= (1 << nbits) - 1

        def flush():
            nonlocal acc, nbits
            if nbits:
                put((1 << (8 - nbits)) - 1, 8 - nbits)

        pred = {c: 0 for c in comps}
        before, rst = restart, 0
        total = d.mcus_per_line * d.mcus_per_col
        for mcu in range(total):
            row, col = divmod(mcu, d.mcus_per_line)
            for c in comps:
                h, v = d.comp_h[c], d.comp_v[c]
                dct, act = codes[(0, sel[comp_ids[c]] >> 4)], codes[(1, sel[comp_ids[c]] & 15)]
                for y in range(v):
                    for x in range(h):
                        blk = d.coef[c][row * v + y, col * h + x]
                        diff = int(blk[0]) - pred[c]
                        pred[c] = int(blk[0])
                        s = size_of(diff)
                        put(*dct[s])
                        if s:
                            put(diff if diff >= 0 else diff - 1, s)
                        run = 0
                        for k in range(1, 64):
                            t = int(blk[k])
                            if t == 0:
                                run += 1
                                continue
                            while run > 15:
                                put(*act[0xF0])
                                run -= 16
                            s = size_of(t)
                            put(*act[(run << 4) | s])
                            put(t if t >= 0 else t - 1, s)
                            run = 0
                        if run:
                            put(*act[0])
            if restart:
                before -= 1
                if before == 0 and mcu != total - 1:
                    flush()
                    data += bytes([0xFF, 0xD0 + (rst & 7)])
                    rst += 1
                    before = restart
                    pred = {c: 0 for c in comps}
        flush()
        out += data
    out += b"\xff\xd9"
    return bytes(out)


# ---------------------------------------------------------------------------------------------
# Hand-made grey baseline streams whose Huffman tables use the ALL-ONES codes (DC category 0 = '1', AC end-of-block =
# '11').  Behind the data the reference's bit reader supplies 1-bits (JpegBitReader.cs:166) and DecodeHuffmanCode
# advances min(code size, bits available) (JpegHuffmanScanDecoder.cs:81-88), so with these tables a stream that is cut
# short keeps decoding -- empty blocks -- without an error, unless a symbol with magnitude bits straddles the cut.
def handmade_grey(blocks_w, blocks_h, seed=0, dri=0, cut=0, cut_interval=None):
    """returns (jpeg bytes, expected zig-zag coefficient blocks [n][64] of the UNCUT stream).
    cut: bytes removed from the end of the entropy-coded data (of restart interval `cut_interval`, default the last)."""
    rng = np.random.default_rng(seed)
    n = blocks_w * blocks_h
    coef = np.zeros((n, 64), np.int16)
    intervals, bits, pred = [], [], 0

    def flush():
        nonlocal bits
        while len(bits) % 8:
            bits.append(1)
        by = np.packbits(np.array(bits, np.uint8)).tobytes() if bits else b""
        intervals.append(by.replace(b"\xff", b"\xff\x00"))
        bits = []

    for i in range(n):
        if dri and i and i % dri == 0:
            flush()
            pred = 0
        d = int(rng.integers(-1, 2))
        bits += [1] if d == 0 else [0, 1 if d > 0 else 0]            # DC: '1' = category 0, '0' = category 1 + 1 bit
        pred += d
        coef[i, 0] = pred
        k = 1
        for _ in range(int(rng.integers(0, 4))):
            if rng.integers(2):
                v = int(rng.choice([-1, 1]))
                bits += [0, 1 if v > 0 else 0]                       # AC 0x01: '0' + 1 bit
            else:
                v = int(rng.choice([-3, -2, 2, 3]))
                bits += [1, 0] + [int(x) for x in format(v if v > 0 else v + 3, "02b")]  # AC 0x02: '10' + 2 bits
            coef[i, k] = v
            k += 1
        bits += [1, 1]                                               # EOB: '11'
    flush()
    if cut:
        k = len(intervals) - 1 if cut_interval is None else cut_interval
        intervals[k] = intervals[k][:max(0, len(intervals[k]) - cut)]
    data = b"".join(iv + (bytes([0xFF, 0xD0 + (k & 7)]) if k + 1 < len(intervals) else b"") for k, iv in enumerate(intervals))

    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload

    out = b"\xff\xd8" + seg(0xDB, bytes([0]) + bytes([1] * 64))
    out += seg(0xC0, bytes([8]) + (blocks_h * 8).to_bytes(2, "big") + (blocks_w * 8).to_bytes(2, "big") + bytes([1, 1, 0x11, 0]))
    out += seg(0xC4, bytes([0x00]) + bytes([2] + [0] * 15) + bytes([1, 0]))          # DC: '0' -> 1, '1' -> 0
    out += seg(0xC4, bytes([0x10]) + bytes([1, 2] + [0] * 14) + bytes([0x01, 0x02, 0x00]))  # AC: '0' 01, '10' 02, '11' EOB
    if dri:
        out += seg(0xDD, dri.to_bytes(2, "big"))
    out += seg(0xDA, bytes([1, 1, 0x00, 0, 63, 0])) + data + b"\xff\xd9"
    return out, coef


def reorder_progressive_scans(blob, order):
    """blob: a progressive JPEG whose scans each come behind their own DHT segments (Pillow / libjpeg with optimised
    tables); order: a permutation (or selection) of scan indices.  Returns the same frame with its scans -- each together
    with the tables in front of it -- in that order: the scan script of a damaged or hand-made file, which the
    reference decodes in file order without checking the progression."""
    i, head, chunks, cur = 2, [blob[:2]], [], []
    seen_sos = False
    while i < len(blob):
        assert blob[i] == 0xFF
        m = blob[i + 1]
        if m == 0xD9:
            break
        ln = int.from_bytes(blob[i + 2:i + 4], "big")
        seg = blob[i:i + 2 + ln]
        i += 2 + ln
        if m == 0xDA:
            j = i
            while True:
                j = blob.index(b"\xff", j)
                if blob[j + 1] != 0 and not 0xD0 <= blob[j + 1] <= 0xD7:
                    break
                j += 2
            chunks.append(b"".join(cur) + seg + blob[i:j])
            cur, i, seen_sos = [], j, True
        elif m == 0xC4 or seen_sos:
            cur.append(seg)
        else:
            head.append(seg)
    return b"".join(head) + b"".join(chunks[k] for k in order) + b"".join(cur) + b"\xff\xd9"


def split_tables(blob, move=(0xC4, 0xDB), tables_dri=None):
    """blob: a JPEG whose table segments all precede the first SOS.  Returns (tables, abbreviated): the DHT / DQT (and,
    with 0xDD in `move`, DRI) segments as a stand-alone tables stream SOI .. EOI -- what a TIFF writer puts into the
    JPEGTables field and a caller hands to JpegDecoder.LoadTables -- and the image without them.  tables_dri: a DRI
    segment of that value added to the tables stream."""
    i, tables, rest = 2, [b"\xff\xd8"], [b"\xff\xd8"]
    while i < len(blob):
        assert blob[i] == 0xFF
        m = blob[i + 1]
        ln = int.from_bytes(blob[i + 2:i + 4], "big")
        seg = blob[i:i + 2 + ln]
        if m == 0xDA:
            rest.append(blob[i:])
            break
        (tables if m in move else rest).append(seg)
        i += 2 + ln
    if tables_dri is not None:
        tables.append(b"\xff\xdd\x00\x04" + int(tables_dri).to_bytes(2, "big"))
    return b"".join(tables) + b"\xff\xd9", b"".join(rest)
