"""ctypes bindings of the two native libraries (include/jpegb200.h, include/jpegb200_host.h).

There is no Python or CPU fallback for the compute path: if libjpegb200.so is missing this
module raises at import, and every compute entry point returns JB_ERR_NO_DEVICE without a GPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.environ.get("JB_LIBDIR") or os.path.join(_HERE, "lib")  # (JB_LIBDIR: A/B builds, see build.py)

JB_OK = 0
JB_ERR_INVALID_DATA = -1
JB_ERR_INVALID_OPERATION = -2
JB_ERR_NOT_SUPPORTED = -3
JB_ERR_ARGUMENT = -4
JB_ERR_NO_DEVICE = -5
JB_ERR_CUDA = -6
JB_ERR_NOMEM = -7

JB_OUT_RGB24 = 0
JB_OUT_RGBA32 = 1
JB_OUT_YCBCR888 = 2
JB_OUT_PLANAR_I16 = 3
JB_OUT_COEFFICIENTS = 4

JB_MAX_COMPONENTS = 4


class HuffSpec(C.Structure):
    _fields_ = [("table_class", C.c_uint8), ("identifier", C.c_uint8), ("bits", C.c_uint8 * 16),
                ("values", C.c_uint8 * 256), ("value_count", C.c_uint16)]


class ScanDesc(C.Structure):
    _fields_ = [("component_count", C.c_uint8), ("component_index", C.c_uint8 * 4),
                ("dc_table", C.c_int16 * 4), ("ac_table", C.c_int16 * 4),
                ("ss", C.c_uint8), ("se", C.c_uint8), ("ah", C.c_uint8), ("al", C.c_uint8),
                ("restart_interval", C.c_uint32), ("entropy_offset", C.c_uint64),
                ("entropy_length", C.c_uint64)]


class ImageDesc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("length", C.c_uint64),
                ("sof", C.c_uint8), ("precision", C.c_uint8), ("component_count", C.c_uint8),
                ("reserved0", C.c_uint8), ("width", C.c_uint16), ("height", C.c_uint16),
                ("h", C.c_uint8 * 4), ("v", C.c_uint8 * 4), ("quant", (C.c_uint16 * 64) * 4),
                ("scan_count", C.c_uint32), ("scans", C.POINTER(ScanDesc)),
                ("table_count", C.c_uint32), ("tables", C.POINTER(HuffSpec))]


class OutputDesc(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("pitch", C.c_uint64), ("capacity", C.c_uint64),
                ("format", C.c_int32), ("on_device", C.c_int32)]


class CoefLayout(C.Structure):
    _fields_ = [("interleaved", C.c_int32), ("mcus_per_line", C.c_int32), ("mcus_per_column", C.c_int32),
                ("blocks_per_mcu", C.c_int32), ("comp_block_offset", C.c_int32 * 4),
                ("comp_blocks_w", C.c_int32 * 4), ("comp_blocks_h", C.c_int32 * 4),
                ("total_blocks", C.c_uint64)]


JB_IN_RGB24 = 0
JB_IN_YCBCR888 = 1
JB_IN_GRAY8 = 2
JB_IN_COEFFICIENTS = 3


class EncodeDesc(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("pitch", C.c_uint64), ("on_device", C.c_int32), ("format", C.c_int32),
                ("width", C.c_uint16), ("height", C.c_uint16), ("component_count", C.c_uint8),
                ("h", C.c_uint8 * 4), ("v", C.c_uint8 * 4), ("tq", C.c_uint8 * 4), ("td", C.c_uint8 * 4),
                ("ta", C.c_uint8 * 4), ("reserved", C.c_uint8), ("restart_interval", C.c_uint16), ("quant", (C.c_uint16 * 64) * 4),
                ("quant_present", C.c_uint8 * 4)]


def _load(name):
    path = os.path.join(_LIBDIR, name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m jpeglibrary_b200.build` "
            "(nvcc, sm_100a). There is no fallback implementation.")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


cuda = _load("libjpegb200.so")
host = _load("libjpegb200_host.so")

_vp = C.c_void_p
_sigs = {
    cuda: {
        "jb_version": (C.c_char_p, []),
        "jb_device_count": (C.c_int, []),
        "jb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
        "jb_ctx_destroy": (None, [_vp]),
        "jb_last_error": (C.c_char_p, [_vp]),
        "jb_ctx_stream": (_vp, [_vp]),
        "jb_ctx_synchronize": (C.c_int, [_vp]),
        "jb_ctx_trim": (C.c_int, [_vp]),
        "jb_pinned_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
        "jb_pinned_free": (C.c_int, [_vp, _vp]),
        "jb_device_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
        "jb_device_free": (C.c_int, [_vp, _vp]),
        "jb_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
        "jb_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
        "jb_decode_batch_create": (C.c_int, [_vp, C.POINTER(ImageDesc), C.POINTER(OutputDesc), C.c_int, C.POINTER(_vp)]),
        "jb_decode_batch_upload": (C.c_int, [_vp]),
        "jb_decode_batch_launch": (C.c_int, [_vp]),
        "jb_decode_batch_finish": (C.c_int, [_vp]),
        "jb_decode_batch_run": (C.c_int, [_vp]),
        "jb_decode_batch_status": (C.c_int, [_vp, C.POINTER(C.c_int32), C.c_int]),
        "jb_decode_batch_coef_layout": (C.c_int, [_vp, C.c_int, C.POINTER(CoefLayout)]),
        "jb_decode_batch_launch_count": (C.c_int, [_vp]),
        "jb_decode_batch_set_profiling": (C.c_int, [_vp, C.c_int]),
        "jb_decode_batch_profile": (C.c_int, [_vp, _vp, C.POINTER(C.c_float), C.c_int]),
        "jb_decode_batch_scan_trace": (C.c_int, [_vp, _vp, C.c_int]),
        "jb_plan_scans": (C.c_int, [C.POINTER(ImageDesc), C.POINTER(C.c_int32), C.c_int]),
        "jb_decode_batch_destroy": (None, [_vp]),
        "jb_decode": (C.c_int, [_vp, C.POINTER(ImageDesc), C.POINTER(OutputDesc), C.c_int, C.POINTER(C.c_int32)]),
        "jb_encode_batch_create": (C.c_int, [_vp, C.POINTER(EncodeDesc), C.c_int, C.POINTER(_vp)]),
        "jb_encode_batch_transform": (C.c_int, [_vp]),
        "jb_encode_batch_histograms": (C.c_int, [_vp, _vp, C.c_int]),
        "jb_encode_batch_build_tables": (C.c_int, [_vp]),
        "jb_encode_batch_set_table": (C.c_int, [_vp, C.c_int, C.POINTER(HuffSpec)]),
        "jb_encode_batch_pack": (C.c_int, [_vp]),
        "jb_encode_batch_finish": (C.c_int, [_vp]),
        "jb_encode_batch_get_table": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(HuffSpec)]),
        "jb_encode_batch_scan_length": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_uint64)]),
        "jb_encode_batch_read_scan": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64]),
        "jb_encode_batch_read_coefficients": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64]),
        "jb_encode_batch_launch_count": (C.c_int, [_vp]),
        "jb_encode_batch_destroy": (None, [_vp]),
        "jb_build_huffman_table": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(HuffSpec)]),
        "jb_build_huffman_table_optimal": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(HuffSpec)]),
        "jb_render_from_coefficients": (C.c_int, [_vp, C.POINTER(ImageDesc), _vp, C.POINTER(OutputDesc)]),
    },
    host: {
        "jbh_parse": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
        "jbh_parse_with_tables": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, C.POINTER(_vp)]),
        "jbh_check_tables": (C.c_int, [_vp, C.c_uint64, C.POINTER(C.c_uint64)]),
        "jbh_desc": (C.POINTER(ImageDesc), [_vp]),
        "jbh_consumed": (C.c_uint64, [_vp]),
        "jbh_sof_marker": (C.c_int, [_vp]),
        "jbh_free": (None, [_vp]),
        "jbh_last_parse_error": (C.c_char_p, []),
        "jbh_parse_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
        "jbh_parse_batch_with_tables": (C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_int, C.c_int, _vp]),
        "jbh_collect_descs": (C.c_int, [_vp, C.c_int, C.POINTER(ImageDesc)]),
    },
}
CUDA_SYMBOLS = sorted(_sigs[cuda])
HOST_SYMBOLS = sorted(_sigs[host])
for _lib, _table in _sigs.items():
    for _name, (_res, _args) in _table.items():
        _fn = getattr(_lib, _name)  # AttributeError here == the .so does not export a declared symbol
        _fn.restype = _res
        _fn.argtypes = _args
