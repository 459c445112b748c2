// GPU twin of JpegOptimizer (JpegOptimizer.cs): Scan() = entropy decode to coefficients on the device + symbol
// histograms + optimised tables (no IDCT), Optimize() = re-pack the same coefficients with the new codes and rewrite the
// marker stream like Optimize (:546-647).  Restart intervals are preserved as CopyScanBaseline does (:772-812).
// The marker loop below is the reference's own (SOI/APP0/SOFn copied, first DHT/DQT replaced by all tables, SOS copied,
// the rest dropped when strip) -- with one deliberate deviation: a DRI segment is kept even when strip is set, because
// the RSTn markers stay in the scan (reference quirk Q6 produces a stream no decoder accepts).
// Python twin with the tests (tests/test_optimizer.py): jpeglibrary_b200/api.py JpegOptimizer.
// NOT compiled in this repository's build image (no .NET toolchain).
using System;
using System.Buffers;

namespace JpegLibrary.Cuda
{
    public sealed unsafe class CudaJpegOptimizer : IDisposable
    {
        public const int JB_IN_COEFFICIENTS = 3;

        private readonly CudaJpegDecoder _walker;   // owns the library context; runs the reference's marker loop
        private readonly IntPtr _ctx;
        private ReadOnlyMemory<byte> _input;
        private IBufferWriter<byte>? _output;
        private IntPtr _batch;
        private CudaJpegDecoder.CoefficientResult _frame;

        public CudaJpegOptimizer(CudaJpegDecoder decoder) { _walker = decoder; _ctx = decoder.Context; }

        public void SetInput(ReadOnlyMemory<byte> input) { _input = input; Release(); }
        public void SetOutput(IBufferWriter<byte> output) => _output = output ?? throw new ArgumentNullException(nameof(output));

        public void Scan()
        {
            if (_input.IsEmpty) throw new InvalidOperationException("Input buffer is not specified.");
            Release();
            // marker walk of the reference (JpegOptimizer.Scan :72-154) through the decoder subclass: frame header,
            // tables, DRI and the position of the entropy-coded bytes; the scan is entropy-decoded on the device (K0 + K1)
            _walker.SetInput(_input);
            _frame = _walker.DecodeCoefficients();
            if (_frame.Sof > 1) throw new System.IO.InvalidDataException("Progressive JPEG is not supported currently.");

            Native.ScanDesc scan = _frame.Scan;
            Native.EncodeDesc e = default;
            e.Pixels = (void*)_frame.Coefficients; e.OnDevice = 1; e.Format = JB_IN_COEFFICIENTS;
            e.Width = _frame.Width; e.Height = _frame.Height; e.ComponentCount = scan.ComponentCount;
            e.RestartInterval = (ushort)scan.RestartInterval;
            for (int i = 0; i < scan.ComponentCount; i++)                                      // components in SCAN order
            {
                int c = scan.ComponentIndex[i];
                e.H[i] = _frame.H[c]; e.V[i] = _frame.V[c];
                e.Td[i] = _frame.Tables[scan.DcTable[i]].Identifier;
                e.Ta[i] = _frame.Tables[scan.AcTable[i]].Identifier;
            }
            Native.Check(_ctx, Native.jb_encode_batch_create(_ctx, &e, 1, out _batch));
            Native.Check(_ctx, Native.jb_encode_batch_transform(_batch));                     // histograms of the decoded symbols
            Native.Check(_ctx, Native.jb_encode_batch_build_tables(_batch));                  // BuildUsingStandardMethod on the device
        }

        public void Optimize(bool strip = true)
        {
            if (_batch == IntPtr.Zero) throw new InvalidOperationException();
            IBufferWriter<byte> output = _output ?? throw new InvalidOperationException();
            Native.Check(_ctx, Native.jb_encode_batch_pack(_batch));
            Native.Check(_ctx, Native.jb_encode_batch_finish(_batch));
            ulong length;
            Native.jb_encode_batch_scan_length(_batch, 0, &length);

            var reader = new JpegReader(_input);
            var writer = new JpegWriter(output, 4096);
            bool dht = false, dqt = false, eoi = false;
            while (!eoi && !reader.IsEmpty && reader.TryReadMarker(out JpegMarker marker))
            {
                switch (marker)
                {
                    case JpegMarker.StartOfImage: writer.WriteMarker(marker); break;
                    case JpegMarker.App0: case JpegMarker.StartOfFrame0: case JpegMarker.StartOfFrame1:
                        writer.WriteMarker(marker); CopySegment(ref reader, ref writer); break;
                    case JpegMarker.DefineHuffmanTable:
                        if (!dht) { WriteTables(ref writer); dht = true; }
                        SkipSegment(ref reader); break;
                    case JpegMarker.DefineQuantizationTable:
                        if (!dqt) { WriteQuantizationTables(ref writer); dqt = true; }
                        SkipSegment(ref reader); break;
                    case JpegMarker.DefineRestartInterval:                                  // deviation from the reference: kept
                        if (_frame.Scan.RestartInterval != 0 || !strip) { writer.WriteMarker(marker); CopySegment(ref reader, ref writer); }
                        else SkipSegment(ref reader);
                        break;
                    case JpegMarker.StartOfScan:
                        writer.WriteMarker(marker); CopySegment(ref reader, ref writer);
                        Span<byte> dst = writer.GetSpan((int)length);
                        fixed (byte* pd = dst) Native.Check(_ctx, Native.jb_encode_batch_read_scan(_batch, 0, pd, length));
                        writer.Advance((int)length);
                        reader.TryAdvance((int)_frame.Scan.EntropyLength);                          // the old scan data incl. its RSTn
                        break;
                    case JpegMarker.EndOfImage: writer.WriteMarker(marker); eoi = true; break;
                    default:
                        if (strip) SkipSegment(ref reader); else { writer.WriteMarker(marker); CopySegment(ref reader, ref writer); }
                        break;
                }
            }
            writer.Flush();
        }

        private void WriteTables(ref JpegWriter writer)
        {
            // one DHT segment with every table the scan uses, in GetOrCreateTableBuilder order (JpegOptimizer.cs:394-395)
            Span<byte> body = stackalloc byte[8 * (1 + 16 + 256)];
            int n = 0;
            Span<bool> seen = stackalloc bool[8];
            Native.ScanDesc scan = _frame.Scan;
            for (int i = 0; i < scan.ComponentCount; i++)
                foreach (int cls in new[] { 0, 1 })
                {
                    int id = cls == 0 ? _frame.Tables[scan.DcTable[i]].Identifier : _frame.Tables[scan.AcTable[i]].Identifier;
                    if (seen[cls * 4 + id]) continue;
                    seen[cls * 4 + id] = true;
                    Native.HuffSpec s;
                    Native.Check(_ctx, Native.jb_encode_batch_get_table(_batch, 0, cls, id, &s));
                    body[n++] = (byte)((cls << 4) | id);
                    for (int k = 0; k < 16; k++) body[n++] = s.Bits[k];
                    for (int k = 0; k < s.ValueCount; k++) body[n++] = s.Values[k];
                }
            writer.WriteMarker(JpegMarker.DefineHuffmanTable);
            writer.WriteLength((ushort)n);
            writer.WriteBytes(body.Slice(0, n));
        }

        // The three helpers below do what JpegOptimizer's private WriteQuantizationTables, CopyMarkerData and SkipMarkerData do
        // (JpegOptimizer.cs:649-716), on what the walker collected.
        private void WriteQuantizationTables(ref JpegWriter writer)
        {
            JpegQuantizationTable[] tables = _frame.QuantizationTables;
            if (tables is null || tables.Length == 0) throw new InvalidOperationException();
            int total = 0;
            foreach (JpegQuantizationTable t in tables) total += t.BytesRequired;
            writer.WriteMarker(JpegMarker.DefineQuantizationTable);
            writer.WriteLength((ushort)total);
            foreach (JpegQuantizationTable t in tables)
            {
                Span<byte> dst = writer.GetSpan(t.BytesRequired);
                t.TryWrite(dst, out int written);
                writer.Advance(written);
            }
        }

        private static void CopySegment(ref JpegReader reader, ref JpegWriter writer)
        {
            if (!reader.TryReadLength(out ushort length)) // (the payload length: TryReadLength takes the two length bytes off)
                throw new System.IO.InvalidDataException($"Failed to decode JPEG data at offset {reader.ConsumedByteCount}. Unexpected end of input data when reading segment length.");
            if (!reader.TryReadBytes(length, out ReadOnlySequence<byte> payload))
                throw new System.IO.InvalidDataException($"Failed to decode JPEG data at offset {reader.ConsumedByteCount}. Unexpected end of input data when reading segment content.");
            writer.WriteLength(length);
            writer.WriteBytes(payload);
        }

        private static void SkipSegment(ref JpegReader reader)
        {
            if (!reader.TryReadLength(out ushort length))
                throw new System.IO.InvalidDataException($"Failed to decode JPEG data at offset {reader.ConsumedByteCount}. Unexpected end of input data when reading segment length.");
            if (!reader.TryAdvance(length))
                throw new System.IO.InvalidDataException($"Failed to decode JPEG data at offset {reader.ConsumedByteCount}. Unexpected end of input data when reading segment content.");
        }

        private void Release()
        {
            if (_batch != IntPtr.Zero) { Native.jb_encode_batch_destroy(_batch); _batch = IntPtr.Zero; }
            if (_frame.Coefficients != IntPtr.Zero) { Native.jb_device_free(_ctx, _frame.Coefficients); _frame.Coefficients = IntPtr.Zero; }
        }
        public void Dispose() => Release();
    }
}
