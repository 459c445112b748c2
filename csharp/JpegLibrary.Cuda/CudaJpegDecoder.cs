// JpegLibrary.Cuda -- the drop-in for the decode hot path, written against JpegLibrary's PUBLIC surface only.
//
// JpegDecoder keeps everything it does on the host: SetInput / Identify / the marker loop of Decode()
// (JpegDecoder.cs:509-550) and its DQT/DRI registries.  CudaJpegDecoder overrides the protected virtual hook
// ProcessMarkerForDecode (JpegDecoder.cs:558): it lets the base class handle every marker, keeps its own copy of
// what the base class hides (frame header, raw DHT bytes -- JpegHuffmanDecodingTable does not retain the code
// length counts), and at StartOfScan hands the entropy-coded segment to libjpegb200 instead of
// JpegScanDecoder.ProcessScan (:592-599).  Sequential frames are decoded at their SOS; progressive frames collect
// their scans and are decoded at EOI, exactly where JpegHuffmanProgressiveScanDecoder.Dispose renders (:421-470).
//
// NOT compiled in this repository's build image (no .NET toolchain); the same C-ABI calls, in the same order, are
// exercised by jpeglibrary_b200/api.py (ctypes), which the tests and bench.py run on the GPU.
using System;
using System.Buffers;
using System.Collections.Generic;
using System.IO;
using System.Runtime.InteropServices;

namespace JpegLibrary.Cuda
{
    /// <summary>Recognised GPU-aware sink: kernels store pixels straight into its (pinned or device) buffer with the
    /// semantics of apps/JpegDecode (JpegBufferOutputWriter8Bit + JpegYCbCrToRgbConverter); WriteBlock is never
    /// called on the fast path (SURVEY 8b).</summary>
    public sealed class CudaRgbOutputWriter : JpegBlockOutputWriter
    {
        public IntPtr Buffer; public long Pitch; public long Capacity; public bool OnDevice; public int Format = Native.JB_OUT_RGB24;
        public override void WriteBlock(ref short blockRef, int componentIndex, int x, int y) =>
            throw new InvalidOperationException("CudaRgbOutputWriter is filled by the GPU, not by WriteBlock calls");
    }

    public sealed unsafe class CudaJpegDecoder : JpegDecoder, IDisposable
    {
        private readonly IntPtr _ctx;
        private JpegBlockOutputWriter? _writer;
        private ReadOnlyMemory<byte> _input;
        private JpegFrameHeader _frame;
        private JpegMarker _sof;
        private readonly List<Native.HuffSpec> _tables = new List<Native.HuffSpec>();
        private readonly int[,] _latest = new int[2, 4];           // [class, id] -> index into _tables, -1 = undefined
        private readonly List<Native.ScanDesc> _scans = new List<Native.ScanDesc>();
        private readonly List<byte> _quantOrder = new List<byte>();
        // Quantisation tables as the reference's scan decoders capture them: a sequential component is rendered with the
        // table in force at ITS scan (JpegHuffmanBaselineScanDecoder via InitDecodeComponents, JpegHuffmanScanDecoder.cs:38-66);
        // a progressive frame is rendered at Dispose from the `_components` slots the LAST scans left behind -- slot i
        // belongs to the i-th component of a scan (JpegHuffmanProgressiveScanDecoder.cs:69, :431-462; quirk P6).
        private readonly ushort[] _componentQuant = new ushort[4 * 64];
        private readonly int[] _slotComponent = { -1, -1, -1, -1 };
        private readonly ushort[] _slotQuant = new ushort[4 * 64];

        private readonly bool _ownsContext;

        public CudaJpegDecoder(int device = 0)
        {
            Native.Check(IntPtr.Zero, Native.jb_ctx_create(device, out _ctx));
            _ownsContext = true;
            for (int c = 0; c < 2; c++) for (int i = 0; i < 4; i++) _latest[c, i] = -1;
        }

        /// <summary>A decoder on somebody else's context (the walkers of CudaJpegBatchDecoder.Create share the batch's).</summary>
        internal CudaJpegDecoder(IntPtr sharedContext)
        {
            _ctx = sharedContext;
            _ownsContext = false;
            for (int c = 0; c < 2; c++) for (int i = 0; i < 4; i++) _latest[c, i] = -1;
        }

        public void Dispose() { if (_ownsContext) Native.jb_ctx_destroy(_ctx); }

        /// <summary>The library context (one per GPU): shared with CudaJpegEncoder / CudaJpegOptimizer / PinnedMemoryPool.</summary>
        public IntPtr Context => _ctx;

        // Transcoding (CudaJpegOptimizer.Scan): the same marker walk, but the frame is only entropy-decoded -- zig-zag
        // int16 blocks in scan order stay in device memory (JB_OUT_COEFFICIENTS) -- and the headers are handed out.
        internal struct CoefficientResult
        {
            public IntPtr Coefficients; public ulong Blocks;
            public byte Sof; public ushort Width, Height;
            public byte[] H, V;                 // per frame component
            public Native.ScanDesc Scan;        // the (only) scan of a sequential frame; EntropyOffset/Length locate its bytes
            public Native.HuffSpec[] Tables;    // Scan.DcTable / AcTable index into this
            public JpegQuantizationTable[] QuantizationTables; // every table of the stream, in the order of first definition
        }
        private ushort _restartAtFrame;
        private bool _coefficientsOnly;
        private CoefficientResult _coefficientResult;
        internal CoefficientResult DecodeCoefficients()
        {
            _coefficientsOnly = true;
            EnsureWriterForWalk();
            try { Decode(); } finally { _coefficientsOnly = false; }
            return _coefficientResult;
        }

        // JpegDecoder.Decode() refuses to start without an output writer (JpegDecoder.cs:515-518); the walks that produce no
        // pixels (coefficients only, descriptor only) give the base class a sink that is never written to.
        private void EnsureWriterForWalk()
        {
            if (_writer is null) base.SetOutputWriter(new CudaRgbOutputWriter());
        }

        /// <summary>What one stream contributes to a batch: its descriptor and everything the descriptor points to, pinned
        /// until Dispose (after jb_decode_batch_finish).</summary>
        internal sealed class CollectedImage : IDisposable
        {
            public Native.ImageDesc Desc;
            internal GCHandle Scans, Tables;
            internal MemoryHandle Input;
            public void Dispose()
            {
                if (Scans.IsAllocated) Scans.Free();
                if (Tables.IsAllocated) Tables.Free();
                Input.Dispose();
            }
        }
        private bool _collectOnly;
        private CollectedImage? _collected;

        /// <summary>The marker loop of Decode() without decoding: the jb_image_desc CudaJpegBatchDecoder hands to
        /// jb_decode_batch_create.  Tables loaded before (LoadTables, SetHuffmanTable, SetQuantizationTable) take part like
        /// they do in Decode().</summary>
        internal CollectedImage Collect()
        {
            _collectOnly = true;
            _collected = null;
            EnsureWriterForWalk();
            try { Decode(); } finally { _collectOnly = false; }
            return _collected ?? throw new InvalidDataException("Failed to decode JPEG data. No image data is read.");
        }

        // JpegDecoder.SetInput / SetOutputWriter are not virtual: keep our own references next to the base class's.
        public new void SetInput(ReadOnlyMemory<byte> input) { _input = input; _quantOrder.Clear(); base.SetInput(input); }
        public new void SetOutputWriter(JpegBlockOutputWriter outputWriter) { _writer = outputWriter; base.SetOutputWriter(outputWriter); }

        protected override bool ProcessMarkerForDecode(JpegMarker marker, ref JpegReader reader)
        {
            switch (marker)
            {
                case JpegMarker.StartOfFrame0: case JpegMarker.StartOfFrame1: case JpegMarker.StartOfFrame2: case JpegMarker.StartOfFrame3:
                {
                    JpegReader peek = reader; // JpegReader is a struct: an independent cursor over the same bytes
                    if (peek.TryReadLength(out ushort len) && peek.TryReadBytes(len, out ReadOnlySequence<byte> body) &&
                        JpegFrameHeader.TryParse(body, false, out JpegFrameHeader fh, out _))
                    {
                        _frame = fh; _sof = marker; _scans.Clear();
                        for (int i = 0; i < 4; i++) _slotComponent[i] = -1;
                        // The reference's sequential and lossless scan decoders are constructed HERE, at the frame header, and
                        // read the restart interval once, in their constructors (JpegHuffmanBaselineScanDecoder.cs:38,
                        // JpegHuffmanLosslessScanDecoder.cs:32): every scan of the frame uses the value in force now -- which,
                        // unless a DRI segment precedes the SOF, is what Identify() left behind (the last DRI of the stream).
                        // Progressive scans read it per scan (JpegHuffmanProgressiveScanDecoder.cs:78).
                        _restartAtFrame = GetRestartInterval();
                    }
                    return base.ProcessMarkerForDecode(marker, ref reader); // validation + the public Width/Height/... properties
                }
                case JpegMarker.DefineHuffmanTable:
                {
                    JpegReader peek = reader;
                    if (peek.TryReadLength(out ushort len) && peek.TryReadBytes(len, out ReadOnlySequence<byte> body)) RememberTables(body);
                    return base.ProcessMarkerForDecode(marker, ref reader);
                }
                case JpegMarker.DefineQuantizationTable:
                {
                    // the order in which identifiers are FIRST defined: JpegOptimizer writes its tables in that order
                    // (a later definition replaces the table in place, JpegOptimizer.cs ProcessDefineQuantizationTable)
                    JpegReader peek = reader;
                    if (peek.TryReadLength(out ushort len) && peek.TryReadBytes(len, out ReadOnlySequence<byte> body))
                    {
                        byte[] b = body.ToArray();
                        for (int p = 0; p < b.Length; p += (b[p] >> 4) != 0 ? 129 : 65)
                        {
                            byte id = (byte)(b[p] & 15);
                            if (id < 4 && !_quantOrder.Contains(id)) _quantOrder.Add(id);
                        }
                    }
                    return base.ProcessMarkerForDecode(marker, ref reader);
                }
                case JpegMarker.StartOfScan:
                    ProcessScanOnGpu(ref reader);
                    return true;
                case JpegMarker.EndOfImage:
                    if (_scans.Count > 0) Submit(); // progressive frames and scan lists: decode + render after the last scan
                    return false;
                default:
                    return base.ProcessMarkerForDecode(marker, ref reader); // DQT, DRI, APPn, COM, RSTn stay managed
            }
        }

        // DHT: Tc/Th byte, 16 counts, then the symbols (JpegHuffmanDecodingTable.TryParse :122-291)
        private void RememberTables(ReadOnlySequence<byte> body)
        {
            byte[] b = body.ToArray();
            int p = 0;
            while (p + 17 <= b.Length)
            {
                var s = new Native.HuffSpec { TableClass = (byte)(b[p] >> 4), Identifier = (byte)(b[p] & 15) };
                int n = 0;
                for (int i = 0; i < 16; i++) { s.Bits[i] = b[p + 1 + i]; n += b[p + 1 + i]; }
                if (n > 256 || p + 17 + n > b.Length || s.TableClass > 1 || s.Identifier > 3) return; // the base class reports it
                for (int i = 0; i < n; i++) s.Values[i] = b[p + 17 + i];
                s.ValueCount = (ushort)n;
                _latest[s.TableClass, s.Identifier] = _tables.Count;
                _tables.Add(s);
                p += 17 + n;
            }
        }

        // Tables that did not come through this decoder's marker loop (JpegDecoder.LoadTables / SetHuffmanTable,
        // JpegDecoder.cs:313-319, :768-860, used for TIFF-style abbreviated streams) are only visible as
        // JpegHuffmanDecodingTable objects, which keep no code-length counts.  Their DHT form is recovered from the
        // public Lookup(code16) (JpegHuffmanDecodingTable.cs:73-85): walking the 16-bit code space in increasing order
        // visits the canonical codes in DHT order.
        private int TableIndex(int tableClass, int identifier)
        {
            if (_latest[tableClass, identifier] >= 0) return _latest[tableClass, identifier];
            JpegHuffmanDecodingTable? t = GetHuffmanTable(tableClass == 0, (byte)identifier);
            if (t is null) return -1;
            var s = new Native.HuffSpec { TableClass = (byte)tableClass, Identifier = (byte)identifier };
            int n = 0;
            for (int code = 0; code < 0x10000 && n < 256;)
            {
                JpegHuffmanDecodingTable.Entry e = t.Lookup(code);
                if (e.CodeSize == 0 || e.CodeSize > 16) break;          // past the last code: only the all-ones prefix is left
                s.Bits[e.CodeSize - 1]++;
                s.Values[n++] = e.SymbolValue;
                code += 1 << (16 - e.CodeSize);                          // first 16-bit pattern of the next canonical code
            }
            s.ValueCount = (ushort)n;
            _latest[tableClass, identifier] = _tables.Count;
            _tables.Add(s);
            return _latest[tableClass, identifier];
        }

        private void ProcessScanOnGpu(ref JpegReader reader)
        {
            if (_frame.Components is null) throw new InvalidDataException("Failed to decode JPEG data. Scan header appears before frame header.");
            if (!reader.TryReadLength(out ushort len) || !reader.TryReadBytes(len, out ReadOnlySequence<byte> body) ||
                !JpegScanHeader.TryParse(body, false, out JpegScanHeader sh, out _))
                throw new InvalidDataException("Failed to decode JPEG data. Failed to parse scan header.");
            var sd = new Native.ScanDesc
            {
                ComponentCount = sh.NumberOfComponents,
                Ss = sh.StartOfSpectralSelection, Se = sh.EndOfSpectralSelection,
                Ah = sh.SuccessiveApproximationBitPositionHigh, Al = sh.SuccessiveApproximationBitPositionLow,
                RestartInterval = _sof == JpegMarker.StartOfFrame2 ? GetRestartInterval() : _restartAtFrame,
                EntropyOffset = (ulong)reader.ConsumedByteCount,
            };
            for (int i = 0; i < sh.NumberOfComponents; i++)
            {
                JpegScanComponentSpecificationParameters sc = sh.Components![i];
                int found = -1;
                for (int j = 0; j < _frame.NumberOfComponents; j++)
                    if (_frame.Components[j].Identifier == sc.ScanComponentSelector) found = j; // InitDecodeComponents :38-48
                if (found < 0) throw new InvalidDataException("Failed to decode JPEG data. The specified component is missing.");
                sd.ComponentIndex[i] = (byte)found;
                if (_sof != JpegMarker.StartOfFrame3) // lossless frames carry no DQT
                {
                    JpegQuantizationTable q = GetQuantizationTable(_frame.Components[found].QuantizationTableSelector); // zig-zag order (JpegQuantizationTable.cs:21)
                    if (q.IsEmpty) throw new InvalidDataException("Failed to decode JPEG data. Quantization table of component is not defined.");
                    _slotComponent[i] = found;
                    for (int k = 0; k < 64; k++)
                    {
                        _slotQuant[i * 64 + k] = q.Elements[k];
                        if (_sof != JpegMarker.StartOfFrame2) _componentQuant[found * 64 + k] = q.Elements[k];
                    }
                }
                sd.DcTable[i] = (short)TableIndex(0, sc.DcEntropyCodingTableSelector & 3);
                sd.AcTable[i] = (short)TableIndex(1, sc.AcEntropyCodingTableSelector & 3);
            }
            // the entropy-coded segment ends at the first marker that is neither a stuffed zero, a fill byte nor RSTn
            int end = FindScanEnd(reader.RemainingBytes);
            sd.EntropyLength = (ulong)end;
            _scans.Add(sd);
            reader.TryAdvance(end);
            // A sequential or lossless frame coded as ONE scan over every component (what every mainstream encoder writes)
            // is decoded right here, like the reference does.  Anything else -- several scans, a scan over some of the
            // components -- is collected and handed over as one scan list at EOI, exactly what the Python mirror
            // (jpeglibrary_b200/api.py) passes: a scan submitted alone would render the frame without the components
            // of the other scans.  (A stream that ends without EOI loses such a scan list; the managed decoder would
            // have written the scans it met.)
            if (_sof != JpegMarker.StartOfFrame2 && _scans.Count == 1 && NamesEveryComponentOnce(sd)) { Submit(); _scans.Clear(); }
        }

        private bool NamesEveryComponentOnce(Native.ScanDesc sd)
        {
            if (sd.ComponentCount != _frame.NumberOfComponents) return false;
            int seen = 0;
            for (int i = 0; i < sd.ComponentCount; i++)
            {
                int bit = 1 << sd.ComponentIndex[i];
                if ((seen & bit) != 0) return false;
                seen |= bit;
            }
            return true;
        }

        private static int FindScanEnd(ReadOnlySequence<byte> data)
        {
            int pos = 0, prev = -1;
            foreach (ReadOnlyMemory<byte> seg in data)
            {
                ReadOnlySpan<byte> s = seg.Span;
                for (int i = 0; i < s.Length; i++, pos++)
                {
                    int b = s[i];
                    if (prev == 0xFF && b != 0x00 && b != 0xFF && (b < 0xD0 || b > 0xD7)) return pos - 1;
                    prev = b;
                }
            }
            return pos;
        }

        // One jb_decode call for the frame: jb_image_desc from the collected headers, jb_output_desc from the writer.
        private void Submit()
        {
            int n = _frame.NumberOfComponents;
            var img = new Native.ImageDesc
            {
                Length = (ulong)_input.Length,
                Sof = (byte)(_sof - JpegMarker.StartOfFrame0), Precision = _frame.SamplePrecision, ComponentCount = (byte)n,
                Width = _frame.SamplesPerLine, Height = _frame.NumberOfLines,
                ScanCount = (uint)_scans.Count, TableCount = (uint)_tables.Count,
            };
            for (int c = 0; c < n; c++)
            {
                JpegFrameComponentSpecificationParameters fc = _frame.Components![c];
                img.H[c] = fc.HorizontalSamplingFactor; img.V[c] = fc.VerticalSamplingFactor;
                if (_sof != JpegMarker.StartOfFrame3 && _sof != JpegMarker.StartOfFrame2)
                    for (int i = 0; i < 64; i++) img.Quant[c * 64 + i] = _componentQuant[c * 64 + i]; // (0 for a component no scan names: never rendered)
            }
            if (_sof == JpegMarker.StartOfFrame2)
            {
                // what Dispose renders with (the same rule as the host walk of this repository, jpeg_host.cpp): slot i -> its
                // component and the table captured with it; a scan order that leaves the slots inconsistent is refused
                int seen = 0;
                for (int i = 0; i < n; i++)
                {
                    int c = _slotComponent[i];
                    if (c < 0) throw new InvalidDataException("Failed to decode JPEG data. progressive frame leaves a component slot without scans");
                    if ((seen & (1 << c)) != 0) throw new NotSupportedException("progressive scan order leaves component slots inconsistent (reference quirk P6)");
                    seen |= 1 << c;
                    for (int k = 0; k < 64; k++) img.Quant[c * 64 + k] = _slotQuant[i * 64 + k];
                }
            }
            Native.ScanDesc[] scans = _scans.ToArray();
            Native.HuffSpec[] tables = _tables.ToArray();
            if (_collectOnly)
            {
                // descriptor only: the arrays and the input stay pinned with the CollectedImage (a frame whose scans are
                // submitted one by one -- there is none on this path: single-scan frames call Submit once, scan lists at EOI)
                var held = new CollectedImage { Scans = GCHandle.Alloc(scans, GCHandleType.Pinned), Tables = GCHandle.Alloc(tables, GCHandleType.Pinned), Input = _input.Pin() };
                img.Data = (byte*)held.Input.Pointer;
                img.Scans = (Native.ScanDesc*)held.Scans.AddrOfPinnedObject();
                img.Tables = (Native.HuffSpec*)held.Tables.AddrOfPinnedObject();
                held.Desc = img;
                _collected?.Dispose();
                _collected = held;
                return;
            }
            using MemoryHandle pin = _input.Pin();
            fixed (Native.ScanDesc* ps = scans)
            fixed (Native.HuffSpec* pt = tables)
            {
                img.Data = (byte*)pin.Pointer; img.Scans = ps; img.Tables = pt;
                if (_coefficientsOnly)
                {
                    int hmx = 1, vmx = 1, bpm = 0;
                    byte[] hh = new byte[n], vv = new byte[n];
                    for (int c = 0; c < n; c++) { hh[c] = img.H[c]; vv[c] = img.V[c]; hmx = Math.Max(hmx, hh[c]); vmx = Math.Max(vmx, vv[c]); bpm += hh[c] * vv[c]; }
                    ulong blocks = (ulong)((img.Width + 8 * hmx - 1) / (8 * hmx)) * (ulong)((img.Height + 8 * vmx - 1) / (8 * vmx)) * (ulong)bpm;
                    Native.Check(_ctx, Native.jb_device_alloc(_ctx, (UIntPtr)(blocks * 128), out IntPtr dev));
                    var oc = new Native.OutputDesc { Dst = (void*)dev, Capacity = blocks * 128, Format = Native.JB_OUT_COEFFICIENTS, OnDevice = 1 };
                    Native.Check(_ctx, Native.jb_decode(_ctx, &img, &oc, 1, null));               // K0 + K1 only
                    var qts = new List<JpegQuantizationTable>();
                    foreach (byte id in _quantOrder) { JpegQuantizationTable q = GetQuantizationTable(id); if (!q.IsEmpty) qts.Add(q); }
                    _coefficientResult = new CoefficientResult { Coefficients = dev, Blocks = blocks, Sof = img.Sof, Width = img.Width, Height = img.Height,
                                                                 H = hh, V = vv, Scan = scans[0], Tables = tables, QuantizationTables = qts.ToArray() };
                    return;
                }
                if (_writer is CudaRgbOutputWriter w)
                {
                    var o = new Native.OutputDesc { Dst = (void*)w.Buffer, Pitch = (ulong)w.Pitch, Capacity = (ulong)w.Capacity, Format = w.Format, OnDevice = w.OnDevice ? 1 : 0 };
                    Native.Check(_ctx, Native.jb_decode(_ctx, &img, &o, 1, null));
                    return;
                }
                // compatibility path: the exact samples WriteBlock receives come back as unclamped int16 planes and the
                // reference's call sequence is replayed on the host
                short[] planes = new short[n * img.Width * img.Height];
                fixed (short* pp = planes)
                {
                    var o = new Native.OutputDesc { Dst = pp, Pitch = 0, Capacity = (ulong)planes.Length * 2, Format = Native.JB_OUT_PLANAR_I16, OnDevice = 0 };
                    Native.Check(_ctx, Native.jb_decode(_ctx, &img, &o, 1, null));
                }
                // sequential frames: the reference hands over every scan's blocks as it decodes them, scan after scan
                if (_sof == JpegMarker.StartOfFrame2 || _sof == JpegMarker.StartOfFrame3) ReplayWriteBlocks(planes, scans[0]);
                else foreach (Native.ScanDesc s in scans) ReplayWriteBlocks(planes, s);
            }
        }

        // JpegHuffmanBaselineScanDecoder.ProcessScan :99-137 + WriteBlock :225-268 (sequential frames: MCU order, scan
        // component order, v x h blocks, each expanded to hs x vs replicated 8x8 blocks); JpegBlockAllocator.Flush
        // :120-149 (progressive: per component, block rows then columns).
        private void ReplayWriteBlocks(short[] planes, Native.ScanDesc firstScan)
        {
            JpegBlockOutputWriter writer = _writer ?? throw new InvalidOperationException("The output buffer is not specified.");
            int n = _frame.NumberOfComponents, W = _frame.SamplesPerLine, H = _frame.NumberOfLines;
            int hmax = 1, vmax = 1;
            foreach (JpegFrameComponentSpecificationParameters c in _frame.Components!) { hmax = Math.Max(hmax, c.HorizontalSamplingFactor); vmax = Math.Max(vmax, c.VerticalSamplingFactor); }
            // lossless frames: JpegPartialScanlineAllocator flushes 8-row bands of 8x8 tiles as they complete; a clipping
            // writer sees the same samples from the per-component order used here
            int unit = _sof == JpegMarker.StartOfFrame3 ? 1 : 8;
            Span<short> block = stackalloc short[64];
            void Emit(int ci, int x, int y, Span<short> blk)
            {
                for (int r = 0; r < 8; r++)
                    for (int c = 0; c < 8; c++)
                        blk[r * 8 + c] = (y + r < H && x + c < W) ? planes[(ci * H + y + r) * W + x + c] : (short)0;
                writer.WriteBlock(ref blk[0], ci, x, y);
            }
            if (_sof == JpegMarker.StartOfFrame2 || unit == 1)
            {
                int wblk = (W + 7) / 8, hblk = (H + 7) / 8;
                for (int ci = 0; ci < n; ci++)
                {
                    int hs = hmax / _frame.Components[ci].HorizontalSamplingFactor, vs = vmax / _frame.Components[ci].VerticalSamplingFactor;
                    int cw = (wblk + hs - 1) / hs, ch = (hblk + vs - 1) / vs;
                    for (int row = 0; row < ch; row++)
                        for (int col = 0; col < cw; col++)
                            for (int sv = 0; sv < vs; sv++)
                                for (int sh = 0; sh < hs; sh++)
                                    Emit(ci, col * hs * 8 + 8 * sh, row * vs * 8 + 8 * sv, block);
                }
                return;
            }
            int mcusX = (W + 8 * hmax - 1) / (8 * hmax), mcusY = (H + 8 * vmax - 1) / (8 * vmax);
            for (int my = 0; my < mcusY; my++)
                for (int mx = 0; mx < mcusX; mx++)
                    for (int k = 0; k < firstScan.ComponentCount; k++)
                    {
                        int ci = firstScan.ComponentIndex[k];
                        int h = _frame.Components[ci].HorizontalSamplingFactor, v = _frame.Components[ci].VerticalSamplingFactor;
                        int hs = hmax / h, vs = vmax / v;
                        for (int by = 0; by < v; by++)
                            for (int bx = 0; bx < h; bx++)
                                for (int sv = 0; sv < vs; sv++)
                                    for (int sh = 0; sh < hs; sh++)
                                        Emit(ci, (mx * hmax + bx) * 8 + 8 * sh, (my * vmax + by) * 8 + 8 * sv, block);
                    }
        }
    }
}
