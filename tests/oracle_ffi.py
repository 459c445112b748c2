"""ctypes binding of the CPU oracle (oracle/libjpeg_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = os.path.join(ORACLE_DIR, "libjpeg_oracle.so")

JO_MAX_COMP = 4
JO_MAX_SCANS = 64


class ScanInfo(C.Structure):
    _fields_ = [
        ("ncomp", C.c_int),
        ("comp_index", C.c_int * JO_MAX_COMP),
        ("td", C.c_int * JO_MAX_COMP),
        ("ta", C.c_int * JO_MAX_COMP),
        ("ss", C.c_int), ("se", C.c_int), ("ah", C.c_int), ("al", C.c_int),
        ("entropy_offset", C.c_size_t),
        ("restart_interval", C.c_int),
    ]


class Image(C.Structure):
    _fields_ = [
        ("sof", C.c_int),
        ("precision", C.c_int), ("width", C.c_int), ("height", C.c_int), ("ncomp", C.c_int),
        ("comp_id", C.c_int * JO_MAX_COMP), ("comp_h", C.c_int * JO_MAX_COMP),
        ("comp_v", C.c_int * JO_MAX_COMP), ("comp_tq", C.c_int * JO_MAX_COMP),
        ("hmax", C.c_int), ("vmax", C.c_int), ("mcus_per_line", C.c_int), ("mcus_per_col", C.c_int),
        ("qt", (C.c_uint16 * 64) * JO_MAX_COMP),
        ("nscans", C.c_int),
        ("scans", ScanInfo * JO_MAX_SCANS),
        ("coef_w", C.c_int * JO_MAX_COMP), ("coef_h", C.c_int * JO_MAX_COMP),
        ("alloc_w", C.c_int * JO_MAX_COMP), ("alloc_h", C.c_int * JO_MAX_COMP),
        ("coef", C.POINTER(C.c_int16) * JO_MAX_COMP),
        ("planes", C.POINTER(C.c_int16)),
        ("ycbcr", C.POINTER(C.c_uint8)),
        ("rgb", C.POINTER(C.c_uint8)),
        ("consumed", C.c_size_t),
        ("error", C.c_char * 160),
        ("written", C.POINTER(C.c_uint8) * JO_MAX_COMP),
    ]


class EncodeParams(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("ncomp", C.c_int),
        ("h", C.c_int * JO_MAX_COMP), ("v", C.c_int * JO_MAX_COMP),
        ("tq", C.c_int * JO_MAX_COMP), ("td", C.c_int * JO_MAX_COMP), ("ta", C.c_int * JO_MAX_COMP),
        ("qt", (C.c_uint16 * 64) * 4),
        ("qt_present", C.c_int * 4),
        ("optimize", C.c_int),
    ]


class Encoded(C.Structure):
    _fields_ = [
        ("bytes", C.POINTER(C.c_uint8)), ("len", C.c_size_t),
        ("alloc_w", C.c_int * JO_MAX_COMP), ("alloc_h", C.c_int * JO_MAX_COMP),
        ("coef", C.POINTER(C.c_int16) * JO_MAX_COMP),
        ("hist", ((C.c_uint32 * 256) * 4) * 2),
        ("dht_bits", ((C.c_uint8 * 16) * 4) * 2),
        ("dht_vals", ((C.c_uint8 * 256) * 4) * 2),
        ("dht_nvals", (C.c_int * 4) * 2),
        ("scan_offset", C.c_size_t), ("scan_len", C.c_size_t),
        ("error", C.c_char * 160),
        ("dummy", C.c_int16 * 64),
    ]


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.jo_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Image)]
        _lib.jo_decode.restype = C.c_int
        _lib.jo_decode_with_tables.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Image)]
        _lib.jo_decode_with_tables.restype = C.c_int
        _lib.jo_free.argtypes = [C.POINTER(Image)]
        _lib.jo_dequant_idct_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.jo_ycbcr_to_rgb.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.jo_decode_batch_rgb.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.jo_decode_batch_rgb.restype = C.c_int
        if hasattr(_lib, "jo_encode_ycbcr"):
            _lib.jo_rgb_to_ycbcr.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
            _lib.jo_std_quant_table.argtypes = [C.c_int, C.c_int, C.c_void_p]
            _lib.jo_encode_ycbcr.argtypes = [C.c_void_p, C.POINTER(EncodeParams), C.POINTER(Encoded)]
            _lib.jo_encode_ycbcr.restype = C.c_int
            _lib.jo_encoded_free.argtypes = [C.POINTER(Encoded)]
            _lib.jo_build_huffman_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            _lib.jo_build_huffman_table.restype = C.c_int
            _lib.jo_build_huffman_table_optimal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            _lib.jo_build_huffman_table_optimal.restype = C.c_int
    return _lib


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


class Decoded:
    """numpy view of a jo_image (copies, so the C memory is freed immediately)."""


def decode(data: bytes, want_rgb=True, tables: bytes = None) -> Decoded:
    """tables: what JpegDecoder.LoadTables is given before the (abbreviated) stream"""
    buf = np.frombuffer(data, dtype=np.uint8)
    img = Image()
    if tables:
        tbuf = np.frombuffer(tables, dtype=np.uint8)
        rc = lib().jo_decode_with_tables(tbuf.ctypes.data, tbuf.size, buf.ctypes.data, buf.size, 7 if want_rgb else 3, C.byref(img))
    else:
        rc = lib().jo_decode(buf.ctypes.data, buf.size, 7 if want_rgb else 3, C.byref(img))
    try:
        if rc != 0:
            raise OracleError(rc, img.error.decode())
        d = Decoded()
        for k in ("sof", "precision", "width", "height", "ncomp", "hmax", "vmax", "mcus_per_line",
                  "mcus_per_col", "nscans", "consumed"):
            setattr(d, k, getattr(img, k))
        n = img.ncomp
        d.comp_h = list(img.comp_h)[:n]
        d.comp_v = list(img.comp_v)[:n]
        d.comp_tq = list(img.comp_tq)[:n]
        d.qt = np.array([list(img.qt[i]) for i in range(n)], dtype=np.uint16)
        d.coef_w = list(img.coef_w)[:n]
        d.coef_h = list(img.coef_h)[:n]
        d.alloc_w = list(img.alloc_w)[:n]
        d.alloc_h = list(img.alloc_h)[:n]
        d.scans = [img.scans[i] for i in range(img.nscans)]
        d.restart_interval = img.scans[0].restart_interval if img.nscans else 0
        d.coef = []
        for i in range(n):
            cnt = d.coef_w[i] * d.coef_h[i] * 64
            a = np.ctypeslib.as_array(img.coef[i], shape=(cnt,)).copy()
            d.coef.append(a.reshape(d.coef_h[i], d.coef_w[i], 64))
        # sequential frames: which blocks some scan handed to WriteBlock (None: every block is written at the end)
        d.written = [np.ctypeslib.as_array(img.written[i], shape=(d.coef_h[i] * d.coef_w[i],)).copy()
                     .reshape(d.coef_h[i], d.coef_w[i]).astype(bool) if bool(img.written[i]) else None for i in range(n)]
        W, H = img.width, img.height
        d.planes = np.ctypeslib.as_array(img.planes, shape=(n * H * W,)).copy().reshape(n, H, W)
        if want_rgb and bool(img.rgb):
            d.ycbcr = np.ctypeslib.as_array(img.ycbcr, shape=(H * W * 3,)).copy().reshape(H, W, 3)
            d.rgb = np.ctypeslib.as_array(img.rgb, shape=(H * W * 3,)).copy().reshape(H, W, 3)
        else:
            d.ycbcr = d.rgb = None
        return d
    finally:
        lib().jo_free(C.byref(img))


def written_samples(d: Decoded) -> np.ndarray:
    """bool [ncomp][H][W]: samples the reference's decoder hands to WriteBlock (all of them unless a sequential scan
    ends early at an EOI on a restart boundary, or no scan names the component)"""
    out = np.ones((d.ncomp, d.height, d.width), dtype=bool)
    for ci in range(d.ncomp):
        if d.written[ci] is None:
            continue
        hs, vs = d.hmax // d.comp_h[ci], d.vmax // d.comp_v[ci]
        m = np.repeat(np.repeat(d.written[ci], 8 * vs, axis=0), 8 * hs, axis=1)
        out[ci] = m[:d.height, :d.width]
    return out


def scan_order_coefficients(d: Decoded) -> np.ndarray:
    """Re-order the planar coefficient store into MCU scan order
    [mcu][block-in-mcu][64] (the device store layout for interleaved baseline scans)."""
    mcus = d.mcus_per_line * d.mcus_per_col
    parts = []
    for ci in range(d.ncomp):
        h, v = d.comp_h[ci], d.comp_v[ci]
        a = d.coef[ci].reshape(d.mcus_per_col, v, d.mcus_per_line, h, 64)
        a = a.transpose(0, 2, 1, 3, 4).reshape(mcus, v * h, 64)
        parts.append(a)
    return np.concatenate(parts, axis=1)


def dequant_idct_block(coef_zz, q_zz, level_shift):
    c = np.ascontiguousarray(coef_zz, dtype=np.int16)
    q = np.ascontiguousarray(q_zz, dtype=np.uint16)
    out = np.empty(64, dtype=np.int16)
    lib().jo_dequant_idct_block(c.ctypes.data, q.ctypes.data, level_shift, out.ctypes.data)
    return out


def ycbcr_to_rgb(ycbcr):
    a = np.ascontiguousarray(ycbcr, dtype=np.uint8)
    out = np.empty_like(a)
    lib().jo_ycbcr_to_rgb(a.ctypes.data, out.ctypes.data, a.size // 3)
    return out


def decode_batch_rgb(blobs, threads, keep_output=False):
    """Time-able CPU baseline: decode every blob to RGB24 on `threads` host threads."""
    n = len(blobs)
    arrs = [np.frombuffer(b, dtype=np.uint8) for b in blobs]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = (C.c_size_t * n)(*[a.size for a in arrs])
    return lib().jo_decode_batch_rgb(ptrs, lens, n, threads, None)


class EncodedResult:
    pass


def std_quant_table(chroma, quality):
    out = np.empty(64, dtype=np.uint16)
    lib().jo_std_quant_table(1 if chroma else 0, quality, out.ctypes.data)
    return out


def rgb_to_ycbcr(rgb):
    a = np.ascontiguousarray(rgb, dtype=np.uint8)
    out = np.empty_like(a)
    lib().jo_rgb_to_ycbcr(a.ctypes.data, out.ctypes.data, a.size // 3)
    return out


def build_huffman_table(freq, optimal=False):
    f = np.ascontiguousarray(freq, dtype=np.uint32)
    bits = np.zeros(16, dtype=np.uint8)
    vals = np.zeros(256, dtype=np.uint8)
    fn = lib().jo_build_huffman_table_optimal if optimal else lib().jo_build_huffman_table
    n = fn(f.ctypes.data, bits.ctypes.data, vals.ctypes.data)
    if n < 0:  # 256 codes of one size: the reference's byte counters wrap (JpegHuffmanEncodingTableBuilder.cs:117-160)
        raise OracleError(n, "IndexOutOfRangeException in the reference's table builder")
    return bits, vals[:n].copy()


def encode_ycbcr(ycbcr, quality=75, subsampling=(2, 2), gray=False, optimal=False):
    """apps/JpegEncode/EncodeAction.cs:37-63 with --optimize-coding: Annex-K tables scaled by quality,
    Y h x v / Cb,Cr 1x1, four optimised Huffman tables."""
    a = np.ascontiguousarray(ycbcr, dtype=np.uint8)
    H, W = a.shape[:2]
    p = EncodeParams()
    p.width, p.height = W, H
    p.ncomp = 1 if gray else 3
    p.optimize = 2 if optimal else 1   # 2: MostOptimalCoding (package merge)
    for c in range(p.ncomp):
        p.h[c], p.v[c] = (subsampling if c == 0 and not gray else (1, 1))
        p.tq[c] = p.td[c] = p.ta[c] = 0 if c == 0 else 1
    for t in range(2):
        q = std_quant_table(t, quality)
        for i in range(64):
            p.qt[t][i] = int(q[i])
        p.qt_present[t] = 1
    e = Encoded()
    rc = lib().jo_encode_ycbcr(a.ctypes.data, C.byref(p), C.byref(e))
    try:
        if rc:
            raise OracleError(rc, e.error.decode())
        r = EncodedResult()
        r.bytes = bytes(np.ctypeslib.as_array(e.bytes, shape=(e.len,)))
        r.scan_offset, r.scan_len = e.scan_offset, e.scan_len
        r.coef = []
        for c in range(p.ncomp):
            n = e.alloc_w[c] * e.alloc_h[c] * 64
            r.coef.append(np.ctypeslib.as_array(e.coef[c], shape=(n,)).copy().reshape(e.alloc_h[c], e.alloc_w[c], 64))
        r.dummy = np.array(list(e.dummy), dtype=np.int16)
        r.hist = np.array([[list(e.hist[cls][t]) for t in range(4)] for cls in range(2)], dtype=np.uint32)
        r.dht = {}
        for cls in range(2):
            for t in range(4):
                n = e.dht_nvals[cls][t]
                if n:
                    r.dht[(cls, t)] = (np.array(list(e.dht_bits[cls][t]), dtype=np.uint8),
                                       np.array(list(e.dht_vals[cls][t])[:n], dtype=np.uint8))
        return r
    finally:
        lib().jo_encoded_free(C.byref(e))
