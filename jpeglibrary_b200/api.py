"""Python mirror of the reference's public surface for the hot path, over the C-ABI.

Same names, argument meaning and error behaviour as the C# API (PascalCase kept on purpose so
the parity tests read like the reference's own tests):

    JpegDecoder.SetInput / Identify / SetOutputWriter / Decode      (src/JpegLibrary/JpegDecoder.cs:49,75,501,509)
    JpegBlockOutputWriter.WriteBlock                                (src/JpegLibrary/JpegBlockOutputWriter.cs:17)

The marker walk runs on the host (libjpegb200_host.so); everything per-block runs on the GPU
through libjpegb200.so.  Recognised GPU-aware writers (CudaOutputWriter) never see WriteBlock:
kernels store straight into their buffer.  Any other JpegBlockOutputWriter gets the
compatibility path: unclamped int16 planes come back from the GPU and the reference's exact
WriteBlock call sequence is replayed on the host.
"""
import ctypes as C

import numpy as np

from . import _native as N


class InvalidDataException(Exception):
    pass


class InvalidOperationException(Exception):
    pass


class NotSupportedException(Exception):
    pass


class ArgumentException(Exception):
    pass


class CudaRuntimeError(RuntimeError):
    pass


def _raise(code, msg):
    msg = msg.decode() if isinstance(msg, bytes) else msg
    if code == N.JB_ERR_INVALID_DATA:
        raise InvalidDataException(msg)
    if code == N.JB_ERR_INVALID_OPERATION:
        raise InvalidOperationException(msg)
    if code == N.JB_ERR_NOT_SUPPORTED:
        raise NotSupportedException(msg)
    if code == N.JB_ERR_ARGUMENT:
        raise ArgumentException(msg)
    if code == N.JB_ERR_NO_DEVICE:
        raise CudaRuntimeError("no CUDA device: jpeglibrary_b200 has no CPU fallback")
    raise CudaRuntimeError(f"jpegb200 error {code}: {msg}")


class Context:
    """jb_ctx wrapper: one per GPU."""

    _default = {}

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = N.cuda.jb_ctx_create(device, C.byref(h))
        if rc:
            _raise(rc, "jb_ctx_create failed")
        self.handle = h
        self.device = device

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def last_error(self):
        return N.cuda.jb_last_error(self.handle).decode()

    def check(self, rc):
        if rc:
            _raise(rc, self.last_error())

    @property
    def stream(self):
        return N.cuda.jb_ctx_stream(self.handle)

    def synchronize(self):
        self.check(N.cuda.jb_ctx_synchronize(self.handle))

    def trim(self):
        """give the device memory cached from destroyed batches back to the driver"""
        self.check(N.cuda.jb_ctx_trim(self.handle))

    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self.check(N.cuda.jb_device_alloc(self.handle, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr):
        N.cuda.jb_device_free(self.handle, C.c_void_p(ptr))

    def pinned_alloc(self, nbytes):
        p = C.c_void_p()
        self.check(N.cuda.jb_pinned_alloc(self.handle, nbytes, C.byref(p)))
        return p.value

    def pinned_free(self, ptr):
        N.cuda.jb_pinned_free(self.handle, C.c_void_p(ptr))

    def pinned_array(self, nbytes):
        """numpy uint8 view of a fresh pinned allocation (freed with pinned_free(arr.ctypes.data))."""
        p = self.pinned_alloc(nbytes)
        return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p))

    def d2h(self, dst: np.ndarray, src_ptr):
        self.check(N.cuda.jb_memcpy_d2h(self.handle, dst.ctypes.data, C.c_void_p(src_ptr), dst.nbytes))

    def h2d(self, dst_ptr, src: np.ndarray):
        self.check(N.cuda.jb_memcpy_h2d(self.handle, C.c_void_p(dst_ptr), src.ctypes.data, src.nbytes))


class Parsed:
    """Result of the host marker walk for one stream (owns the native descriptor)."""

    def __init__(self, data, tables=None):
        self._buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        if tables is not None and len(tables):
            # an abbreviated stream behind JpegDecoder.LoadTables (JpegDecoder.cs:313-360)
            t = tables if isinstance(tables, np.ndarray) else np.frombuffer(tables, dtype=np.uint8)
            rc = N.host.jbh_parse_with_tables(t.ctypes.data, t.size, self._buf.ctypes.data, self._buf.size, C.byref(h))
        else:
            rc = N.host.jbh_parse(self._buf.ctypes.data, self._buf.size, C.byref(h))
        if rc:
            _raise(rc, N.host.jbh_last_parse_error())
        self.handle = h
        self.desc = N.host.jbh_desc(h).contents

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            N.host.jbh_free(h)
            self.handle = None

    @property
    def consumed(self):
        return N.host.jbh_consumed(self.handle)


# ------------------------------------------------------------------------------------------
class JpegBlockOutputWriter:
    """src/JpegLibrary/JpegBlockOutputWriter.cs:8-18"""

    def WriteBlock(self, blockRef, componentIndex, x, y):
        raise NotImplementedError


class CudaOutputWriter(JpegBlockOutputWriter):
    """Recognised GPU-aware sink: the kernels write `format` pixels straight into `buffer`
    (a numpy array in host memory, or an int device pointer with on_device=True) with the
    semantics of apps/JpegDecode (JpegBufferOutputWriter8Bit + JpegYCbCrToRgbConverter).
    WriteBlock is never called."""

    def __init__(self, buffer, format=N.JB_OUT_RGB24, pitch=0, on_device=False, capacity=0):
        self.buffer = buffer
        self.format = format
        self.pitch = pitch
        self.on_device = on_device
        if on_device and capacity <= 0:  # (a host array knows its size; a bare device pointer does not)
            raise ArgumentException("capacity: the size of a device destination must be given")
        self.capacity = capacity if on_device else (buffer.nbytes if capacity == 0 else capacity)

    def _output_desc(self):
        o = N.OutputDesc()
        o.dst = self.buffer if self.on_device else self.buffer.ctypes.data
        o.pitch = self.pitch
        o.capacity = self.capacity
        o.format = self.format
        o.on_device = 1 if self.on_device else 0
        return o

    def WriteBlock(self, blockRef, componentIndex, x, y):  # pragma: no cover
        raise InvalidOperationException("CudaOutputWriter is filled by the GPU, not by WriteBlock calls")


class JpegDecoder:
    """Mirror of JpegLibrary.JpegDecoder for the Huffman DCT path."""

    def __init__(self, context=None):
        self._ctx = context
        self._input = None
        self._parsed = None
        self._writer = None
        self._tables = b""

    # JpegDecoder.cs:49-62
    def SetInput(self, data):
        self._input = data
        self._parsed = None

    # JpegDecoder.cs:313-360: DHT / DQT / DRI segments of a separate tables stream (abbreviated streams, e.g. the
    # JPEGTables field of a TIFF file) become the state the marker loop of the next streams starts from.  Successive
    # calls add up like they do in the reference's table registries (a later definition of an identifier wins).
    def LoadTables(self, content):
        content = np.frombuffer(bytes(content), dtype=np.uint8)
        used = C.c_uint64(0)
        rc = N.host.jbh_check_tables(content.ctypes.data, content.size, C.byref(used)) if content.size else 0
        if rc:
            _raise(rc, N.host.jbh_last_parse_error())
        self._tables += content[:used.value].tobytes()  # without the EOI that ends the walk: the next call's bytes follow
        self._parsed = None

    # JpegDecoder.cs:960-973 (the registries LoadTables filled are emptied)
    def ResetTables(self):
        self._tables = b""
        self._parsed = None

    # JpegDecoder.cs:75-105
    def Identify(self, loadQuantizationTables=False):
        if self._input is None or len(self._input) == 0:
            raise InvalidOperationException("Input buffer is not specified.")
        self._parsed = Parsed(self._input, self._tables)
        return self._parsed.consumed

    def _frame(self):
        if self._parsed is None:
            raise InvalidOperationException("Call Identify() before this operation.")
        return self._parsed.desc

    Width = property(lambda s: s._frame().width)
    Height = property(lambda s: s._frame().height)
    Precision = property(lambda s: s._frame().precision)
    NumberOfComponents = property(lambda s: s._frame().component_count)

    def GetMaximumHorizontalSampling(self):
        d = self._frame()
        return max(d.h[i] for i in range(d.component_count))

    def GetMaximumVerticalSampling(self):
        d = self._frame()
        return max(d.v[i] for i in range(d.component_count))

    def GetHorizontalSampling(self, componentIndex):
        d = self._frame()
        if not 0 <= componentIndex < d.component_count:
            raise ArgumentException("componentIndex")
        return d.h[componentIndex]

    def GetVerticalSampling(self, componentIndex):
        d = self._frame()
        if not 0 <= componentIndex < d.component_count:
            raise ArgumentException("componentIndex")
        return d.v[componentIndex]

    # JpegDecoder.cs:501
    def SetOutputWriter(self, outputWriter):
        if outputWriter is None:
            raise ArgumentException("outputWriter")
        self._writer = outputWriter

    # JpegDecoder.cs:509-550
    def Decode(self):
        if self._input is None or len(self._input) == 0:
            raise InvalidOperationException("Input buffer is not specified.")
        if self._writer is None:
            raise InvalidOperationException("The output buffer is not specified.")
        if self._parsed is None:
            self._parsed = Parsed(self._input, self._tables)
        ctx = self._ctx or Context.default()
        d = self._parsed.desc
        if d.scan_count == 0 and d.sof != 2:
            # the marker loop met no SOS: the reference's sequential / lossless decoders are created at the frame header
            # and never asked for anything, so no WriteBlock call happens and Decode returns (JpegDecoder.cs:509-550)
            return
        if isinstance(self._writer, CudaOutputWriter):
            out = self._writer._output_desc()
            ctx.check(N.cuda.jb_decode(ctx.handle, C.byref(d), C.byref(out), 1, None))
            return
        # compatibility path: unclamped planes from the GPU, WriteBlock replay on the host
        W, H, n = d.width, d.height, d.component_count
        planes = np.empty((n, H, W), dtype=np.int16)
        out = CudaOutputWriter(planes, N.JB_OUT_PLANAR_I16)._output_desc()
        ctx.check(N.cuda.jb_decode(ctx.handle, C.byref(d), C.byref(out), 1, None))
        _replay_write_blocks(d, planes, self._writer)


def _replay_write_blocks(d, planes, writer):
    """Issue the reference's WriteBlock sequence from full-resolution planes.

    Baseline/extended: MCU order, components in scan order, v rows x h cols of blocks, each block
    expanded into hs x vs replicated 8x8 blocks (JpegHuffmanBaselineScanDecoder.cs:99-137, 238-268).
    Progressive: per component, block rows then columns (JpegBlockAllocator.Flush :120-149)."""
    n, H, W = planes.shape
    hmax = max(d.h[i] for i in range(n))
    vmax = max(d.v[i] for i in range(n))
    mcus_x = (W + 8 * hmax - 1) // (8 * hmax)
    mcus_y = (H + 8 * vmax - 1) // (8 * vmax)
    padded = np.zeros((n, mcus_y * 8 * vmax, mcus_x * 8 * hmax), dtype=np.int16)
    padded[:, :H, :W] = planes
    # samples outside the image are never observable through a clipping writer; blocks that
    # start outside still get a call, like in the reference

    def emit(ci, x, y):
        blk = np.ascontiguousarray(padded[ci, y:y + 8, x:x + 8]).reshape(64)
        writer.WriteBlock(blk, ci, x, y)

    if d.sof != 2:
        for si in range(d.scan_count):   # (one scan normally; every scan is an MCU walk over its own components)
            sc = d.scans[si]
            order = [sc.component_index[i] for i in range(sc.component_count)]
            for my in range(mcus_y):
                for mx in range(mcus_x):
                    for ci in order:
                        h, v = d.h[ci], d.v[ci]
                        hs, vs = hmax // h, vmax // v
                        for by in range(v):
                            for bx in range(h):
                                x0 = (mx * hmax + bx) * 8  # :134; h is 1 or hmax on the GPU path
                                y0 = (my * vmax + by) * 8
                                for sv in range(vs):
                                    for sh in range(hs):
                                        emit(ci, x0 + 8 * sh, y0 + 8 * sv)
    else:
        wblk, hblk = (W + 7) // 8, (H + 7) // 8
        for ci in range(n):
            hs, vs = hmax // d.h[ci], vmax // d.v[ci]
            cw, ch = (wblk + hs - 1) // hs, (hblk + vs - 1) // vs
            for row in range(ch):
                for col in range(cw):
                    for sv in range(vs):
                        for sh in range(hs):
                            emit(ci, col * hs * 8 + 8 * sh, row * vs * 8 + 8 * sv)


# ------------------------------------------------------------------------------------------
class JpegBatchDecoder:
    """Batch facade: one Decode() exposes a single image of parallelism, a batch
    exposes thousands of restart segments.  Streams are parsed on host threads, staged into one
    device arena and decoded by three kernel launches for the whole batch."""

    def __init__(self, blobs, format=N.JB_OUT_RGB24, context=None, device_output=True, parse_threads=8,
                 host_outputs=None, parsed=None, tables=None):
        """parsed: marker-walk results (jbh_parse handles) of `blobs` when the caller has them already; ownership
        passes to this object.  tables: one tables stream every blob is an abbreviated stream of (JpegDecoder.LoadTables,
        JpegDecoder.cs:313-360: the strips / tiles of a TIFF file with its JPEGTables field)."""
        self.ctx = context or Context.default()
        n = len(blobs)
        self.count = n
        self._bufs = [b if isinstance(b, np.ndarray) else np.frombuffer(b, dtype=np.uint8) for b in blobs]
        if parsed is not None:
            self._parsed = parsed
        else:
            ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in self._bufs])
            lens = (C.c_uint64 * n)(*[b.size for b in self._bufs])
            self._parsed = (C.c_void_p * n)()
            self._tables = None if tables is None else np.frombuffer(bytes(tables), dtype=np.uint8)
            if self._tables is not None and self._tables.size:
                failed = N.host.jbh_parse_batch_with_tables(self._tables.ctypes.data, self._tables.size, ptrs, lens, n,
                                                            parse_threads, self._parsed)
            else:
                failed = N.host.jbh_parse_batch(ptrs, lens, n, parse_threads, self._parsed)
            if failed:
                self._free_parsed()
                raise InvalidDataException(f"{failed} of {n} streams failed the marker walk")
        self.descs = (N.ImageDesc * n)()
        N.host.jbh_collect_descs(self._parsed, n, self.descs)
        self.format = format
        bpp = {N.JB_OUT_RGB24: 3, N.JB_OUT_RGBA32: 4, N.JB_OUT_YCBCR888: 3}.get(format)
        self.outs = (N.OutputDesc * n)()
        self.sizes = []
        self._dev_out = None
        self.host_outputs = host_outputs
        total = 0
        offs = []
        for i in range(n):
            d = self.descs[i]
            if format == N.JB_OUT_PLANAR_I16:
                sz = d.width * d.height * 2 * d.component_count
            elif format == N.JB_OUT_COEFFICIENTS:
                raise ArgumentException("use JpegDecoder for coefficient output")
            else:
                sz = d.width * d.height * bpp
            offs.append(total)
            self.sizes.append(sz)
            total += (sz + 255) // 256 * 256
        self.total_out_bytes = total
        if device_output:
            self._dev_out = self.ctx.device_alloc(max(total, 256))
        elif host_outputs is None:
            self.host_outputs = self.ctx.pinned_array(max(total, 256))
            self._own_pinned = True
        self.offsets = offs
        for i in range(n):
            o = self.outs[i]
            o.dst = (self._dev_out + offs[i]) if device_output else (self.host_outputs.ctypes.data + offs[i])
            o.pitch = 0
            o.capacity = self.sizes[i]
            o.format = format
            o.on_device = 1 if device_output else 0
        h = C.c_void_p()
        rc = N.cuda.jb_decode_batch_create(self.ctx.handle, self.descs, self.outs, n, C.byref(h))
        if rc:
            self.close()
            _raise(rc, self.ctx.last_error())
        self.handle = h

    def upload(self):
        self.ctx.check(N.cuda.jb_decode_batch_upload(self.handle))

    def launch(self):
        self.ctx.check(N.cuda.jb_decode_batch_launch(self.handle))

    def finish(self):
        self.ctx.check(N.cuda.jb_decode_batch_finish(self.handle))

    def run(self):
        self.ctx.check(N.cuda.jb_decode_batch_run(self.handle))

    def launch_count(self):
        return N.cuda.jb_decode_batch_launch_count(self.handle)

    def set_profiling(self, on, trace=False):
        """on: time every kernel with events (profile()); trace: also record the K1c schedule (scan_trace())."""
        self.ctx.check(N.cuda.jb_decode_batch_set_profiling(self.handle, (2 if trace else 1) if on else 0))

    def profile(self):
        names = ((C.c_char * 48) * 8)()
        ms = (C.c_float * 8)()
        k = N.cuda.jb_decode_batch_profile(self.handle, names, ms, 8)
        if k < 0:
            _raise(k, self.ctx.last_error())
        return [(names[i].value.decode(), ms[i]) for i in range(k)]

    def scan_trace(self):
        """Progressive frames after set_profiling(True, trace=True): [(image, scan, segment, start_ns, end_ns, waited_ns)]
        per K1c job."""
        n = N.cuda.jb_decode_batch_scan_trace(self.handle, None, 0)
        if n <= 0:
            return []
        buf = np.zeros((n, 4), dtype=np.uint64)
        N.cuda.jb_decode_batch_scan_trace(self.handle, buf.ctypes.data, n)
        return [(int(k >> 32), int(k >> 16) & 0xFFFF, int(k) & 0xFFFF, int(a), int(e), int(w)) for k, a, e, w in buf.tolist()]

    def status(self):
        st = (C.c_int32 * self.count)()
        N.cuda.jb_decode_batch_status(self.handle, st, self.count)
        return list(st)

    def read_output(self, i):
        """Copy image i's result to a new numpy array (test helper)."""
        d = self.descs[i]
        a = np.empty(self.sizes[i], dtype=np.uint8)
        if self._dev_out is not None:
            self.ctx.d2h(a, self._dev_out + self.offsets[i])
        else:
            a[:] = self.host_outputs[self.offsets[i]:self.offsets[i] + self.sizes[i]]
        if self.format == N.JB_OUT_PLANAR_I16:
            return a.view(np.int16).reshape(d.component_count, d.height, d.width)
        return a.reshape(d.height, d.width, -1)

    def _free_parsed(self):
        for i in range(self.count):
            if self._parsed[i]:
                N.host.jbh_free(self._parsed[i])
                self._parsed[i] = None

    def close(self):
        if getattr(self, "handle", None):
            N.cuda.jb_decode_batch_destroy(self.handle)
            self.handle = None
        if getattr(self, "_dev_out", None):
            self.ctx.device_free(self._dev_out)
            self._dev_out = None
        if getattr(self, "_own_pinned", False) and self.host_outputs is not None:
            self.ctx.pinned_free(self.host_outputs.ctypes.data)
            self.host_outputs = None
            self._own_pinned = False
        self._free_parsed()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class JpegPipelinedBatchDecoder:
    """Host-buffer decode of a long list of streams as a software pipeline: the list is cut into chunks, and
    `len(contexts)` worker threads (one jb_ctx = one CUDA stream each) run JpegBatchDecoder over alternating
    chunks, so that one chunk's marker walk / planning / H2D / kernels overlap the other chunk's D2H of pixels.
    The D2H of 24.9 MB of RGB per 4K image is what bounds a host-to-host decode (PCIe), so keeping that copy
    engine busy is the whole point.  ctypes releases the GIL inside the native calls."""

    def __init__(self, contexts=None, chunk=16, parse_threads=8):
        self.contexts = contexts or [Context(0), Context(0)]
        self.chunk = chunk
        self.parse_threads = parse_threads
        self.reset_stats()

    def reset_stats(self):
        """per-phase wall-clock seconds, summed over chunks (and over the worker threads)"""
        self.stats = {"walk": 0.0, "parser_blocked": 0.0, "worker_idle": 0.0, "create": 0.0, "upload_launch": 0.0,
                      "finish_wait": 0.0, "destroy": 0.0, "chunks": 0}

    def stats_summary(self, wall_seconds):
        """phase times as fractions of the wall-clock time of the decode() calls since reset_stats(): walk and
        parser_blocked belong to the one parser thread, the others are averaged over the worker threads"""
        w = max(1, len(self.contexts))
        s = self.stats
        per = {k: s[k] / wall_seconds for k in ("walk", "parser_blocked")}
        per.update({k: s[k] / w / wall_seconds for k in ("worker_idle", "create", "upload_launch", "finish_wait", "destroy")})
        return {"fraction_of_wall": {k: round(v, 4) for k, v in per.items()}, "chunks": s["chunks"], "workers": w,
                "meaning": "walk: marker walk of the next chunks; parser_blocked: parsed chunks waiting for a free worker; "
                           "create: descriptors + jb_decode_batch_create (tables, plan, allocations); upload_launch: enqueue "
                           "H2D + kernels; finish_wait: waiting for H2D + kernels + D2H of the chunk (the GPU / PCIe time); "
                           "worker_idle: worker waiting for a parsed chunk"}

    def decode(self, blobs, host_out, format=N.JB_OUT_RGB24, tables=None):
        """Decode `blobs` into the (pinned) uint8 array `host_out`; image i lands at self.offsets[i].
        Returns the per-image offsets.  tables: see JpegBatchDecoder."""
        import queue
        import threading
        bpp = {N.JB_OUT_RGB24: 3, N.JB_OUT_RGBA32: 4, N.JB_OUT_YCBCR888: 3}[format]
        n = len(blobs)
        chunks = [(i, min(i + self.chunk, n)) for i in range(0, n, self.chunk)]
        errors = []
        lock = threading.Lock()
        bufs = [b if isinstance(b, np.ndarray) else np.frombuffer(b, dtype=np.uint8) for b in blobs]
        tbuf = None if tables is None else np.frombuffer(bytes(tables), dtype=np.uint8)
        offsets = [0] * n
        ready = queue.Queue(maxsize=2 * len(self.contexts))  # parsed chunks, in order: (first, last, handles, start, end)

        import time as _time
        st = self.stats

        def parser():
            # One marker walk per stream (host threads), running ahead of the GPU workers: a chunk's place in the
            # output follows from the frame sizes of everything in front of it.
            off = 0
            try:
                for a, b in chunks:
                    t0 = _time.perf_counter()
                    m = b - a
                    ptrs = (C.c_void_p * m)(*[x.ctypes.data for x in bufs[a:b]])
                    lens = (C.c_uint64 * m)(*[x.size for x in bufs[a:b]])
                    handles = (C.c_void_p * m)()
                    if tbuf is not None and tbuf.size:
                        failed = N.host.jbh_parse_batch_with_tables(tbuf.ctypes.data, tbuf.size, ptrs, lens, m, self.parse_threads, handles)
                    else:
                        failed = N.host.jbh_parse_batch(ptrs, lens, m, self.parse_threads, handles)
                    if failed:
                        for h in handles:
                            if h:
                                N.host.jbh_free(h)
                        raise InvalidDataException(f"{failed} of {m} streams failed the marker walk")
                    start = off
                    for i in range(m):
                        d = N.host.jbh_desc(handles[i]).contents
                        offsets[a + i] = off
                        off += (d.width * d.height * bpp + 255) // 256 * 256
                    if off > host_out.size:
                        for h in handles:
                            N.host.jbh_free(h)
                        raise ArgumentException("Destination buffer is too small.")
                    t1 = _time.perf_counter()
                    ready.put((a, b, handles, start, off))
                    t2 = _time.perf_counter()
                    st["walk"] += t1 - t0
                    st["parser_blocked"] += t2 - t1
                    if errors:
                        break
            except Exception as e:  # noqa: BLE001 - reported to the caller below
                with lock:
                    errors.append(e)
            finally:
                for _ in self.contexts:
                    ready.put(None)

        def worker(w):
            ctx = self.contexts[w]
            while True:
                t0 = _time.perf_counter()
                item = ready.get()
                t1 = _time.perf_counter()
                if item is None:
                    return
                a, b, handles, start, end = item
                if errors:                      # drain: a chunk failed somewhere
                    for h in handles:
                        N.host.jbh_free(h)
                    continue
                try:
                    d = JpegBatchDecoder(bufs[a:b], format, context=ctx, device_output=False,
                                         host_outputs=host_out[start:end], parsed=handles)
                    try:
                        t2 = _time.perf_counter()
                        d.upload()
                        d.launch()
                        t3 = _time.perf_counter()
                        d.finish()
                        t4 = _time.perf_counter()
                    finally:
                        t5 = _time.perf_counter()
                        d.close()
                    t6 = _time.perf_counter()
                    with lock:
                        st["worker_idle"] += t1 - t0
                        st["create"] += t2 - t1
                        st["upload_launch"] += t3 - t2
                        st["finish_wait"] += t4 - t3
                        st["destroy"] += t6 - t5
                        st["chunks"] += 1
                except Exception as e:  # noqa: BLE001 - reported to the caller below
                    with lock:
                        errors.append(e)

        threads = [threading.Thread(target=parser)] + [threading.Thread(target=worker, args=(w,)) for w in range(len(self.contexts))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        self.offsets = offsets
        return offsets


def decode_coefficients(data, context=None):
    """Entropy-decode one stream on the GPU and return (layout, int16 blocks[total, 64])."""
    ctx = context or Context.default()
    p = Parsed(data)
    d = p.desc
    hmax = max(d.h[i] for i in range(d.component_count))
    vmax = max(d.v[i] for i in range(d.component_count))
    mcus = ((d.width + 8 * hmax - 1) // (8 * hmax)) * ((d.height + 8 * vmax - 1) // (8 * vmax))
    bpm = sum(d.h[i] * d.v[i] for i in range(d.component_count))
    buf = np.zeros((mcus * bpm, 64), dtype=np.int16)
    out = CudaOutputWriter(buf, N.JB_OUT_COEFFICIENTS)._output_desc()
    h = C.c_void_p()
    ctx.check(N.cuda.jb_decode_batch_create(ctx.handle, C.byref(d), C.byref(out), 1, C.byref(h)))
    try:
        ctx.check(N.cuda.jb_decode_batch_run(h))
        lay = N.CoefLayout()
        N.cuda.jb_decode_batch_coef_layout(h, 0, C.byref(lay))
    finally:
        N.cuda.jb_decode_batch_destroy(h)
    return lay, buf


# ==========================================================================================
# Encoder mirror (src/JpegLibrary/JpegEncoder.cs, JpegBlockInputReader.cs)
# ==========================================================================================
# JpegStandardQuantizationTable.cs:9-31 (zig-zag order)
_STD_LUMA = [16, 11, 12, 14, 12, 10, 16, 14, 13, 14, 18, 17, 16, 19, 24, 40, 26, 24, 22, 22, 24, 49, 35, 37,
             29, 40, 58, 51, 61, 60, 57, 51, 56, 55, 64, 72, 92, 78, 64, 68, 87, 69, 55, 56, 80, 109, 81, 87,
             95, 98, 103, 104, 103, 62, 77, 113, 121, 112, 100, 120, 92, 101, 103, 99]
_STD_CHROMA = [17, 18, 18, 24, 21, 24, 47, 26, 26, 47, 99, 66, 56, 66, 99, 99] + [99] * 48


class JpegQuantizationTable:
    """(ElementPrecision, Identifier, Elements in zig-zag order) -- JpegQuantizationTable.cs"""

    def __init__(self, elementPrecision, identifier, elements):
        self.ElementPrecision, self.Identifier, self.Elements = elementPrecision, identifier, list(elements)


class JpegStandardQuantizationTable:
    @staticmethod
    def GetLuminanceTable(elementPrecision, identifier):
        return JpegQuantizationTable(elementPrecision, identifier, _STD_LUMA)

    @staticmethod
    def GetChrominanceTable(elementPrecision, identifier):
        return JpegQuantizationTable(elementPrecision, identifier, _STD_CHROMA)

    @staticmethod
    def ScaleByQuality(table, quality):  # JpegStandardQuantizationTable.cs:64-89
        if not 0 <= quality <= 100:
            raise ArgumentException("quality")
        scale = 5000 // quality if quality < 50 else 200 - quality * 2
        return JpegQuantizationTable(table.ElementPrecision, table.Identifier,
                                     [min(max((x * scale + 50) // 100, 1), 255) for x in table.Elements])


class JpegBlockInputReader:
    """src/JpegLibrary/JpegBlockInputReader.cs:8-28"""
    Width = 0
    Height = 0

    def ReadBlock(self, blockRef, componentIndex, x, y):
        raise NotImplementedError


class CudaInputReader(JpegBlockInputReader):
    """Recognised GPU-aware source: interleaved pixels (RGB24 converted like apps/JpegEncode, YCbCr888 as
    JpegBufferInputReader reads them, or grey) in a numpy array or device memory.  ReadBlock is never called."""

    def __init__(self, pixels, width=None, height=None, format=N.JB_IN_RGB24, on_device=False, pitch=0):
        self.pixels, self.format, self.on_device, self.pitch = pixels, format, on_device, pitch
        if not on_device:
            self.pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
            height, width = self.pixels.shape[:2]
        self.Width, self.Height = width, height


def _marker_segment(marker, payload):
    return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload


class JpegEncoder:
    """Mirror of JpegLibrary.JpegEncoder (baseline SOF0, optimised or caller-provided Huffman tables)."""

    def __init__(self, context=None):
        self._ctx = context
        self._input = None
        self._output = None
        self._quant = []          # SetQuantizationTable order
        self._tables = []         # (class, id, spec-or-None) in SetHuffmanTable order
        self._components = []
        self.MostOptimalCoding = False
        self.use_host_table_builder = False  # True: histograms -> jb_build_huffman_table on the host

    def SetInputReader(self, inputReader):   # :84
        if inputReader is None:
            raise ArgumentException("inputReader")
        self._input = inputReader

    def SetOutput(self, output):             # :93 (anything with .write / bytearray)
        if output is None:
            raise ArgumentException("output")
        self._output = output

    def SetQuantizationTable(self, table):   # :102-135
        if table is None or not table.Elements:
            raise ArgumentException("Quantization table is not initialized.")
        self._quant = [t for t in self._quant if t.Identifier != table.Identifier] + [table]

    def SetHuffmanTable(self, isDcTable, identifier, table=None):  # :137-147; None = build an optimised table
        cls = 0 if isDcTable else 1
        self._tables = [t for t in self._tables if (t[0], t[1]) != (cls, identifier)] + [(cls, identifier, table)]

    def AddComponent(self, componentIndex, quantizationTableIdentifier, huffmanDcTableIdentifier,
                     huffmanAcTableIdentifier, horizontalSubsampling, verticalSubsampling):  # :175-240
        if horizontalSubsampling not in (1, 2, 4):
            raise ArgumentException("Subsampling factor can only be 1, 2 or 4.")
        if verticalSubsampling not in (1, 2, 4):
            raise ArgumentException("Subsampling factor can only be 1, 2 or 4.")
        if any(c[0] == componentIndex for c in self._components):
            raise ArgumentException("The component index is already used by another component.")
        if not any(t.Identifier == quantizationTableIdentifier for t in self._quant):
            raise ArgumentException("Quantization table is not defined.")
        if not any((t[0], t[1]) == (0, huffmanDcTableIdentifier) for t in self._tables):
            raise ArgumentException("Huffman table is not defined.")
        if not any((t[0], t[1]) == (1, huffmanAcTableIdentifier) for t in self._tables):
            raise ArgumentException("Huffman table is not defined.")
        self._components.append((componentIndex, quantizationTableIdentifier, huffmanDcTableIdentifier,
                                 huffmanAcTableIdentifier, horizontalSubsampling, verticalSubsampling))

    def _reader_pixels(self):
        r = self._input
        if isinstance(r, CudaInputReader):
            return r
        # compatibility path: replay ReadBlock over the full-resolution grid (the reference asks for exactly
        # these blocks, JpegEncoder.cs:743-786) into an interleaved buffer
        n = len(self._components)
        W, H = r.Width, r.Height
        buf = np.zeros(((H + 7) // 8 * 8, (W + 7) // 8 * 8, n), dtype=np.uint8)
        blk = np.zeros(64, dtype=np.int16)
        for ci in range(n):
            for y in range(0, H, 8):
                for x in range(0, W, 8):
                    r.ReadBlock(blk, ci, x, y)
                    buf[y:y + 8, x:x + 8, ci] = blk.reshape(8, 8).astype(np.uint8)
        # MCU-padding blocks (JpegEncoder.cs:458-470 reads them into the allocator's dummy block): the GPU path
        # computes them from zero samples, which is what JpegBufferInputReader.cs:36-39 returns outside the frame
        hmax = max(c[4] for c in self._components)
        vmax = max(c[5] for c in self._components)
        for ci, (_, _, _, _, h, v) in enumerate(self._components):
            if h != hmax or v != vmax:
                continue
            pads = [(x, 0) for x in range((W + 7) // 8 * 8, (W + 8 * hmax - 1) // (8 * hmax) * 8 * hmax, 8)]
            pads += [(0, y) for y in range((H + 7) // 8 * 8, (H + 8 * vmax - 1) // (8 * vmax) * 8 * vmax, 8)]
            for x, y in pads:
                blk[:] = 0
                r.ReadBlock(blk, ci, x, y)
                if blk.any():
                    raise NotSupportedException("input reader returns samples outside its own frame (MCU-padding blocks)")
        pix = np.ascontiguousarray(buf[:H, :W] if n == 3 else buf[:H, :W, 0])
        return CudaInputReader(pix, format=N.JB_IN_YCBCR888 if n == 3 else N.JB_IN_GRAY8)

    def _desc(self, reader):
        d = N.EncodeDesc()
        d.pixels = reader.pixels if reader.on_device else reader.pixels.ctypes.data
        d.pitch = reader.pitch
        d.on_device = 1 if reader.on_device else 0
        d.format = reader.format
        d.width, d.height = reader.Width, reader.Height
        d.component_count = len(self._components)
        for i, (_, tq, td, ta, h, v) in enumerate(self._components):
            d.h[i], d.v[i], d.tq[i], d.td[i], d.ta[i] = h, v, tq, td, ta
        for t in self._quant:
            for k in range(64):
                d.quant[t.Identifier][k] = t.Elements[k]
            d.quant_present[t.Identifier] = 1
        return d

    # :255-290
    def Encode(self):
        if self._input is None:
            raise InvalidOperationException("Input is not specified.")
        if self._output is None:
            raise InvalidOperationException("Output is not specified.")
        if not self._components:
            raise InvalidOperationException("No component is specified.")
        ctx = self._ctx or Context.default()
        reader = self._reader_pixels()
        desc = self._desc(reader)
        h = C.c_void_p()
        ctx.check(N.cuda.jb_encode_batch_create(ctx.handle, C.byref(desc), 1, C.byref(h)))
        try:
            ctx.check(N.cuda.jb_encode_batch_transform(h))
            build = [t for t in self._tables if t[2] is None]
            # MostOptimalCoding (JpegEncoder.cs:43, builder.Build(optimal: true)): the package-merge builder runs on the host
            # over the histograms K3b left, like a caller-side builder would
            if build and not self.use_host_table_builder and not self.MostOptimalCoding:
                ctx.check(N.cuda.jb_encode_batch_build_tables(h))
            elif build:
                hist = np.zeros((8, 256), dtype=np.uint32)
                ctx.check(N.cuda.jb_encode_batch_histograms(h, hist.ctypes.data, 1))
                builder = N.cuda.jb_build_huffman_table_optimal if self.MostOptimalCoding else N.cuda.jb_build_huffman_table
                for cls, ident, _ in build:
                    spec = N.HuffSpec()
                    rc = builder(hist[cls * 4 + ident].ctypes.data, cls, ident, C.byref(spec))
                    if rc:
                        raise InvalidOperationException("No symbol is recorded.")
                    ctx.check(N.cuda.jb_encode_batch_set_table(h, 0, C.byref(spec)))
            for cls, ident, spec in self._tables:
                if spec is not None:
                    ctx.check(N.cuda.jb_encode_batch_set_table(h, 0, C.byref(spec)))
            ctx.check(N.cuda.jb_encode_batch_pack(h))
            ctx.check(N.cuda.jb_encode_batch_finish(h))
            n = C.c_uint64()
            N.cuda.jb_encode_batch_scan_length(h, 0, C.byref(n))
            scan = np.empty(n.value, dtype=np.uint8)
            ctx.check(N.cuda.jb_encode_batch_read_scan(h, 0, scan.ctypes.data, scan.size))
            specs = []
            for cls, ident, _ in self._tables:
                s = N.HuffSpec()
                ctx.check(N.cuda.jb_encode_batch_get_table(h, 0, cls, ident, C.byref(s)))
                if s.value_count == 0:
                    raise InvalidOperationException("No symbol is recorded.")
                specs.append(s)
            self.last_tables = specs
            nblk = 0
            W, H = reader.Width, reader.Height
            hmax = max(c[4] for c in self._components)
            vmax = max(c[5] for c in self._components)
            nblk = ((W + 8 * hmax - 1) // (8 * hmax)) * ((H + 8 * vmax - 1) // (8 * vmax)) * sum(c[4] * c[5] for c in self._components)
            coef = np.empty((nblk, 64), dtype=np.int16)
            ctx.check(N.cuda.jb_encode_batch_read_coefficients(h, 0, coef.ctypes.data, nblk))
            self.last_coefficients = coef
        finally:
            N.cuda.jb_encode_batch_destroy(h)
        self._write(self.assemble(reader.Width, reader.Height, specs, scan.tobytes()))

    def assemble(self, width, height, specs, scan):
        """Stream layout of JpegEncoder.Encode (:255-290, :305-408): SOI, DQT (all tables, one segment),
        SOF0, DHT (all tables, one segment, insertion order), SOS, entropy-coded data, EOI."""
        out = bytearray(b"\xff\xd8")
        dqt = bytearray()
        for t in self._quant:
            dqt += bytes([(t.ElementPrecision << 4) | (t.Identifier & 15)]) + bytes(t.Elements)
        out += _marker_segment(0xDB, bytes(dqt))
        sof = bytes([8]) + height.to_bytes(2, "big") + width.to_bytes(2, "big") + bytes([len(self._components)])
        for ci, tq, _, _, h, v in self._components:
            sof += bytes([ci, (h << 4) | v, tq])
        out += _marker_segment(0xC0, sof)
        dht = bytearray()
        for s in specs:
            dht += bytes([(s.table_class << 4) | (s.identifier & 15)]) + bytes(s.bits) + bytes(s.values[:s.value_count])
        out += _marker_segment(0xC4, bytes(dht))
        sos = bytes([len(self._components)])
        for ci, _, td, ta, _, _ in self._components:
            sos += bytes([ci, (td << 4) | ta])
        out += _marker_segment(0xDA, sos + bytes([0, 63, 0]))
        out += scan + b"\xff\xd9"
        return bytes(out)

    def _write(self, data):
        if hasattr(self._output, "write"):
            self._output.write(data)
        else:
            self._output += data


def encode_rgb(rgb, quality=75, subsampling=(2, 2), context=None, host_builder=False, most_optimal=False):
    """apps/JpegEncode/EncodeAction.cs:37-63 with --optimize-coding, on the GPU. Returns (bytes, encoder)."""
    enc = JpegEncoder(context)
    enc.use_host_table_builder = host_builder
    enc.MostOptimalCoding = most_optimal
    enc.SetQuantizationTable(JpegStandardQuantizationTable.ScaleByQuality(JpegStandardQuantizationTable.GetLuminanceTable(0, 0), quality))
    enc.SetQuantizationTable(JpegStandardQuantizationTable.ScaleByQuality(JpegStandardQuantizationTable.GetChrominanceTable(0, 1), quality))
    for isdc, ident in ((True, 0), (False, 0), (True, 1), (False, 1)):
        enc.SetHuffmanTable(isdc, ident)
    enc.AddComponent(1, 0, 0, 0, subsampling[0], subsampling[1])
    enc.AddComponent(2, 1, 1, 1, 1, 1)
    enc.AddComponent(3, 1, 1, 1, 1, 1)
    enc.SetInputReader(CudaInputReader(rgb, format=N.JB_IN_RGB24))
    out = bytearray()
    enc.SetOutput(out)
    enc.Encode()
    return bytes(out), enc


class JpegBatchEncoder:
    """Batch facade for the encoder path: N frames -> N baseline JPEG streams with optimised tables
    (apps/JpegEncode semantics).  Frames are numpy arrays (host) or (device_ptr, width, height) tuples."""

    def __init__(self, frames, quality=75, subsampling=(2, 2), context=None, format=N.JB_IN_RGB24):
        self.ctx = context or Context.default()
        n = len(frames)
        self.count = n
        self.enc = JpegEncoder(self.ctx)
        e = self.enc
        e.SetQuantizationTable(JpegStandardQuantizationTable.ScaleByQuality(JpegStandardQuantizationTable.GetLuminanceTable(0, 0), quality))
        e.SetQuantizationTable(JpegStandardQuantizationTable.ScaleByQuality(JpegStandardQuantizationTable.GetChrominanceTable(0, 1), quality))
        for isdc, ident in ((True, 0), (False, 0), (True, 1), (False, 1)):
            e.SetHuffmanTable(isdc, ident)
        e.AddComponent(1, 0, 0, 0, subsampling[0], subsampling[1])
        e.AddComponent(2, 1, 1, 1, 1, 1)
        e.AddComponent(3, 1, 1, 1, 1, 1)
        self.descs = (N.EncodeDesc * n)()
        self._keep = []
        self.sizes = []
        for i, f in enumerate(frames):
            if isinstance(f, tuple):
                r = CudaInputReader(f[0], width=f[1], height=f[2], format=format, on_device=True)
            else:
                r = CudaInputReader(f, format=format)
                self._keep.append(r)
            self.descs[i] = e._desc(r)
            self.sizes.append((r.Width, r.Height))
        h = C.c_void_p()
        self.ctx.check(N.cuda.jb_encode_batch_create(self.ctx.handle, self.descs, n, C.byref(h)))
        self.handle = h

    def launch(self):
        """transform + on-device table build + pack, all asynchronous on the context stream"""
        self.ctx.check(N.cuda.jb_encode_batch_transform(self.handle))
        self.ctx.check(N.cuda.jb_encode_batch_build_tables(self.handle))
        self.ctx.check(N.cuda.jb_encode_batch_pack(self.handle))

    def finish(self):
        self.ctx.check(N.cuda.jb_encode_batch_finish(self.handle))

    def launch_count(self):
        return N.cuda.jb_encode_batch_launch_count(self.handle)

    def scan_length(self, i):
        n = C.c_uint64()
        N.cuda.jb_encode_batch_scan_length(self.handle, i, C.byref(n))
        return n.value

    def stream(self, i):
        """Complete JPEG stream of frame i (headers written on the host like JpegEncoder.Encode)."""
        scan = np.empty(self.scan_length(i), dtype=np.uint8)
        self.ctx.check(N.cuda.jb_encode_batch_read_scan(self.handle, i, scan.ctypes.data, scan.size))
        specs = []
        for cls, ident, _ in self.enc._tables:
            s = N.HuffSpec()
            self.ctx.check(N.cuda.jb_encode_batch_get_table(self.handle, i, cls, ident, C.byref(s)))
            specs.append(s)
        return self.enc.assemble(self.sizes[i][0], self.sizes[i][1], specs, scan.tobytes())

    def read_all_scans(self, out: np.ndarray):
        """D2H of every frame's scan bytes into `out` (pinned), returns the byte offsets."""
        offs, pos = [], 0
        for i in range(self.count):
            n = self.scan_length(i)
            self.ctx.check(N.cuda.jb_encode_batch_read_scan(self.handle, i, out.ctypes.data + pos, out.size - pos))
            offs.append((pos, n))
            pos += n
        return offs

    def close(self):
        if getattr(self, "handle", None):
            N.cuda.jb_encode_batch_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ==========================================================================================
# Optimizer mirror (src/JpegLibrary/JpegOptimizer.cs): lossless transcode with optimised Huffman tables
# ==========================================================================================
def _transcode_plan(d):
    """What JpegOptimizer.Scan checks of a parsed stream (JpegOptimizer.cs:72-154); returns (blocks of the store, its scan)."""
    if d.sof > 1:
        raise InvalidDataException("Progressive JPEG is not supported currently.")
    if d.scan_count < 1:
        raise InvalidDataException("No image data is read.")
    sc = d.scans[0]
    named = [sc.component_index[i] for i in range(sc.component_count)]
    if d.scan_count != 1 or sorted(named) != list(range(d.component_count)):
        raise NotSupportedException("only frames coded as one interleaved scan over every component are transcoded on the GPU path")
    hmax = max(d.h[i] for i in range(d.component_count))
    vmax = max(d.v[i] for i in range(d.component_count))
    nblk = ((d.width + 8 * hmax - 1) // (8 * hmax)) * ((d.height + 8 * vmax - 1) // (8 * vmax)) * \
        sum(d.h[i] * d.v[i] for i in range(d.component_count))
    return nblk, sc


def _transcode_desc(e, d, sc, coef_dev):
    """Fill the jb_encode_desc that re-packs the device coefficient store of one decoded frame (JB_IN_COEFFICIENTS):
    components in SCAN order (that is the store's block order).  Returns the (class, identifier) pairs of the tables in
    GetOrCreateTableBuilder order (JpegOptimizer.cs:394-395)."""
    e.pixels, e.on_device, e.format = coef_dev, 1, N.JB_IN_COEFFICIENTS
    e.width, e.height, e.component_count = d.width, d.height, sc.component_count
    e.restart_interval = sc.restart_interval
    order = []
    for i in range(sc.component_count):
        c = sc.component_index[i]
        e.h[i], e.v[i] = d.h[c], d.v[c]
        td = d.tables[sc.dc_table[i]].identifier
        ta = d.tables[sc.ac_table[i]].identifier
        e.td[i], e.ta[i] = td, ta
        for key in ((0, td), (1, ta)):
            if key not in order:
                order.append(key)
    return order


def _optimal_tables_from_histograms(ctx, h, image, hist, order):
    """MostOptimalCoding (JpegOptimizer.cs:39): builder.Build(optimal: true) over the scan's symbol statistics"""
    for cls, ident in order:
        spec = N.HuffSpec()
        if N.cuda.jb_build_huffman_table_optimal(hist[cls * 4 + ident].ctypes.data, cls, ident, C.byref(spec)):
            raise InvalidOperationException("No symbol is recorded.")
        ctx.check(N.cuda.jb_encode_batch_set_table(h, image, C.byref(spec)))


def _read_transcoded(ctx, h, image, order):
    """(scan bytes, tables in `order`) of one image of a packed encode batch"""
    n = C.c_uint64()
    N.cuda.jb_encode_batch_scan_length(h, image, C.byref(n))
    scan = np.empty(n.value, dtype=np.uint8)
    ctx.check(N.cuda.jb_encode_batch_read_scan(h, image, scan.ctypes.data, scan.size))
    specs = []
    for cls, ident in order:
        s = N.HuffSpec()
        ctx.check(N.cuda.jb_encode_batch_get_table(h, image, cls, ident, C.byref(s)))
        specs.append(s)
    return scan, specs


def _rewrite_stream(data, sc, scan, specs, strip):
    """JpegOptimizer.Optimize (:546-647): SOI / APP0 / SOF copied, the first DHT / DQT replaced by all tables, SOS copied
    with the new scan data, other segments dropped when `strip` (the DRI segment of a frame with restart intervals is
    always kept: quirk Q6, see JpegOptimizer below)."""
    data = bytes(data)
    # quantisation tables in definition order, latest definition wins (ProcessDefineQuantizationTable)
    qts = {}
    pos = 2
    segs = []  # (marker, payload_start, payload_end)
    while pos + 2 <= len(data):
        if data[pos] != 0xFF:
            pos += 1
            continue
        m = data[pos + 1]
        if m == 0xFF:
            pos += 1
            continue
        if m == 0xD9:
            segs.append((m, pos + 2, pos + 2))
            break
        if m == 0x00 or 0xD0 <= m <= 0xD7 or m == 0xD8:
            pos += 2
            continue
        if pos + 4 > len(data):
            raise InvalidDataException("Unexpected end of input data when reading segment length.")
        ln = int.from_bytes(data[pos + 2:pos + 4], "big")
        segs.append((m, pos + 4, pos + 2 + ln))
        if m == 0xDB:
            q = data[pos + 4:pos + 2 + ln]
            i = 0
            while i < len(q):
                size = 129 if q[i] >> 4 else 65
                qts[q[i] & 15] = q[i:i + size]
                i += size
        pos += 2 + ln
        if m == 0xDA:
            pos = sc.entropy_offset + sc.entropy_length
    out = bytearray(b"\xff\xd8")
    dht_written = dqt_written = False
    for m, a, b in segs:
        payload = data[a:b]
        if m in (0xE0, 0xC0, 0xC1):
            out += _marker_segment(m, payload)
        elif m == 0xC4:
            if not dht_written:
                body = bytearray()
                for s in specs:
                    body += bytes([(s.table_class << 4) | (s.identifier & 15)]) + bytes(s.bits) + bytes(s.values[:s.value_count])
                out += _marker_segment(0xC4, bytes(body))
                dht_written = True
        elif m == 0xDB:
            if not dqt_written:
                out += _marker_segment(0xDB, b"".join(qts[k] for k in qts))
                dqt_written = True
        elif m == 0xDA:
            out += _marker_segment(0xDA, payload) + scan.tobytes()
        elif m == 0xD9:
            out += b"\xff\xd9"
        elif not strip or (m == 0xDD and sc.restart_interval != 0):
            out += _marker_segment(m, payload)   # (DRI is kept with strip too: see the class comment, quirk Q6)
    return out


class JpegOptimizer:
    """Scan(): entropy-decode on the GPU (K0/K1), histogram the symbols (K3b), build optimised tables (K3c).
    Optimize(strip): re-pack the same coefficients with the new tables (K4) and rewrite the stream like
    JpegOptimizer.Optimize (:546-647): SOI/APP0/SOF copied, the first DHT/DQT replaced by all tables, SOS
    copied with the new scan data, other segments dropped when `strip`.  No DCT is involved.
    Restart intervals are preserved like CopyScanBaseline does (:772-812): DC prediction restarts, every interval is
    padded with 1-bits and followed by its RSTn.  Deliberate deviation (quirk Q6): the reference drops the DRI segment
    with strip=True while keeping the RSTn markers, which leaves a stream no decoder accepts -- the DRI segment is
    always kept here."""

    def __init__(self, context=None):
        self._ctx = context
        self._input = None
        self._output = None
        self._parsed = None
        self._batch = None
        self._coef_dev = None
        self.MostOptimalCoding = False

    def SetInput(self, data):      # :54-70
        self._input = data
        self._close()

    def SetOutput(self, output):   # :537
        if output is None:
            raise ArgumentException("output")
        self._output = output

    def _close(self):
        ctx = self._ctx or (Context._default.get(0) if Context._default else None)
        if self._batch is not None:
            N.cuda.jb_encode_batch_destroy(self._batch)
            self._batch = None
        if self._coef_dev is not None and ctx is not None:
            ctx.device_free(self._coef_dev)
            self._coef_dev = None

    def __del__(self):
        try:
            self._close()
        except Exception:
            pass

    def Scan(self):                # :72-154
        if self._input is None or len(self._input) == 0:
            raise InvalidOperationException("Input buffer is not specified.")
        ctx = self._ctx or Context.default()
        self._close()
        p = Parsed(self._input)
        d = p.desc
        nblk, sc = _transcode_plan(d)
        self._nblk = nblk
        self._coef_dev = ctx.device_alloc(nblk * 128)
        out = CudaOutputWriter(self._coef_dev, N.JB_OUT_COEFFICIENTS, on_device=True, capacity=nblk * 128)._output_desc()
        ctx.check(N.cuda.jb_decode(ctx.handle, C.byref(d), C.byref(out), 1, None))
        e = N.EncodeDesc()
        self._table_order = _transcode_desc(e, d, sc, self._coef_dev)
        h = C.c_void_p()
        ctx.check(N.cuda.jb_encode_batch_create(ctx.handle, C.byref(e), 1, C.byref(h)))
        self._batch = h
        ctx.check(N.cuda.jb_encode_batch_transform(h))
        if self.MostOptimalCoding:  # JpegOptimizer.cs:39: builder.Build(optimal: true) over the scan's symbol statistics
            hist = np.zeros((8, 256), dtype=np.uint32)
            ctx.check(N.cuda.jb_encode_batch_histograms(h, hist.ctypes.data, 1))
            _optimal_tables_from_histograms(ctx, h, 0, hist, self._table_order)
        else:
            ctx.check(N.cuda.jb_encode_batch_build_tables(h))
        self._parsed = p

    def Optimize(self, strip=True):  # :546-647
        if self._batch is None:
            raise InvalidOperationException()
        if self._output is None:
            raise InvalidOperationException()
        ctx = self._ctx or Context.default()
        h = self._batch
        ctx.check(N.cuda.jb_encode_batch_pack(h))
        ctx.check(N.cuda.jb_encode_batch_finish(h))
        scan, specs = _read_transcoded(ctx, h, 0, self._table_order)
        self.last_tables = specs
        out = _rewrite_stream(self._input, self._parsed.desc.scans[0], scan, specs, strip)
        if hasattr(self._output, "write"):
            self._output.write(bytes(out))
        else:
            self._output += out


class JpegBatchOptimizer:
    """JpegOptimizer (Scan + Optimize) over a batch of baseline / extended streams in one pass on the device: K0/K1 decode
    every stream into zig-zag coefficient stores that stay in device memory (JB_OUT_COEFFICIENTS), K3b histograms the
    symbols, K3c builds the optimised tables (or, with most_optimal, the host builds package-merge tables from the
    histograms), K4 re-packs -- the same C-ABI calls as JpegOptimizer, with `count` images per call.  Everything between
    upload and finish is enqueued on the context's stream without a host synchronisation."""

    def __init__(self, blobs, context=None, most_optimal=False, parse_threads=8):
        self.ctx = ctx = context or Context.default()
        self.count = n = len(blobs)
        self.most_optimal = most_optimal
        self._dec = self._enc = self._coef_dev = None
        self._bufs = [b if isinstance(b, np.ndarray) else np.frombuffer(b, dtype=np.uint8) for b in blobs]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in self._bufs])
        lens = (C.c_uint64 * n)(*[b.size for b in self._bufs])
        self._parsed = (C.c_void_p * n)()
        failed = N.host.jbh_parse_batch(ptrs, lens, n, parse_threads, self._parsed)
        if failed:
            self.close()
            raise InvalidDataException(f"{failed} of {n} streams failed the marker walk")
        self.descs = (N.ImageDesc * n)()
        N.host.jbh_collect_descs(self._parsed, n, self.descs)
        try:
            plans = [_transcode_plan(self.descs[i]) for i in range(n)]
            offs, total = [], 0
            for nblk, _sc in plans:
                offs.append(total)
                total += (nblk * 128 + 255) // 256 * 256
            self._coef_dev = ctx.device_alloc(max(total, 256))
            outs = (N.OutputDesc * n)()
            self._edescs = (N.EncodeDesc * n)()
            self._orders = []
            for i, (nblk, sc) in enumerate(plans):
                o = outs[i]
                o.dst, o.pitch, o.capacity, o.format, o.on_device = self._coef_dev + offs[i], 0, nblk * 128, N.JB_OUT_COEFFICIENTS, 1
                self._orders.append(_transcode_desc(self._edescs[i], self.descs[i], sc, self._coef_dev + offs[i]))
            h = C.c_void_p()
            ctx.check(N.cuda.jb_decode_batch_create(ctx.handle, self.descs, outs, n, C.byref(h)))
            self._dec = h
            h = C.c_void_p()
            ctx.check(N.cuda.jb_encode_batch_create(ctx.handle, self._edescs, n, C.byref(h)))
            self._enc = h
        except Exception:
            self.close()
            raise
        self._uploaded = False

    def upload(self):
        self.ctx.check(N.cuda.jb_decode_batch_upload(self._dec))
        self._uploaded = True

    def launch(self):
        """Scan() and the device half of Optimize() for every stream"""
        ctx = self.ctx
        if not self._uploaded:
            self.upload()
        ctx.check(N.cuda.jb_decode_batch_launch(self._dec))
        # a stream the decoder refuses raises what JpegOptimizer.Scan would have raised for it; the coefficient stores
        # reach their destinations here (jb_decode_batch_finish delivers JB_OUT_COEFFICIENTS results)
        ctx.check(N.cuda.jb_decode_batch_finish(self._dec))
        ctx.check(N.cuda.jb_encode_batch_transform(self._enc))  # takes the stores over (device to device) + histograms
        if self.most_optimal:
            hist = np.zeros((self.count, 8, 256), dtype=np.uint32)
            ctx.check(N.cuda.jb_encode_batch_histograms(self._enc, hist.ctypes.data, self.count))  # synchronises
            for i in range(self.count):
                _optimal_tables_from_histograms(ctx, self._enc, i, hist[i], self._orders[i])
        else:
            ctx.check(N.cuda.jb_encode_batch_build_tables(self._enc))
        ctx.check(N.cuda.jb_encode_batch_pack(self._enc))

    def finish(self):
        self.ctx.check(N.cuda.jb_encode_batch_finish(self._enc))

    def launch_count(self):
        return N.cuda.jb_decode_batch_launch_count(self._dec) + N.cuda.jb_encode_batch_launch_count(self._enc)

    def scan_length(self, i):
        n = C.c_uint64()
        N.cuda.jb_encode_batch_scan_length(self._enc, i, C.byref(n))
        return n.value

    def stream(self, i, strip=True):
        """The optimised file of stream i (JpegOptimizer.Optimize, :546-647)"""
        scan, specs = _read_transcoded(self.ctx, self._enc, i, self._orders[i])
        return bytes(_rewrite_stream(self._bufs[i].tobytes(), self.descs[i].scans[0], scan, specs, strip))

    def run(self, strip=True):
        self.launch()
        self.finish()
        return [self.stream(i, strip) for i in range(self.count)]

    def close(self):
        if self._enc is not None:
            N.cuda.jb_encode_batch_destroy(self._enc)
            self._enc = None
        if self._dec is not None:
            N.cuda.jb_decode_batch_destroy(self._dec)
            self._dec = None
        if self._coef_dev is not None:
            self.ctx.device_free(self._coef_dev)
            self._coef_dev = None
        parsed = getattr(self, "_parsed", None)
        if parsed is not None:
            for i in range(self.count):
                if parsed[i]:
                    N.host.jbh_free(parsed[i])
                    parsed[i] = None
            self._parsed = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
