// GPU twin of JpegEncoder.Encode (JpegEncoder.cs:255-338) for interleaved 8-bit frames whose pixels are available as
// one RGB24 / YCbCr888 / grey buffer (the reference pulls 8x8 blocks through JpegBlockInputReader.ReadBlock one by
// one, JpegEncoder.cs:458-470; a maintainer keeps that path for arbitrary readers).  Everything that depends on the
// samples runs on the device (TransformBlocks :414-485, GatherBlockStatistics :551-597, WritePreparedScanData
// :603-656); the Huffman tables come either from the device builder or -- shown here -- from the library's own
// JpegHuffmanEncodingTableBuilder fed with the device histograms, and the marker segments are written by the
// library's JpegWriter exactly as in WriteStartOfImage .. WriteEndOfImage (:340-412).
// Python twin, bit-identical to the oracle's byte stream in tests/test_encoder.py: jpeglibrary_b200/api.py JpegEncoder.
// NOT compiled in this repository's build image (no .NET toolchain).
using System;
using System.Buffers;

namespace JpegLibrary.Cuda
{
    public sealed unsafe class CudaJpegEncoder : JpegEncoder
    {
        public const int JB_IN_RGB24 = 0, JB_IN_YCBCR888 = 1, JB_IN_GRAY8 = 2;

        private readonly IntPtr _ctx;
        private ReadOnlyMemory<byte> _pixels;
        private int _width, _height, _format;
        private IBufferWriter<byte>? _output;
        // the base class keeps its quantisation tables, components and output private: they are recorded on the way in
        private readonly JpegQuantizationTable?[] _quant = new JpegQuantizationTable?[4];
        private readonly System.Collections.Generic.List<(byte index, byte tq, byte td, byte ta, byte h, byte v)> _components = new();

        public CudaJpegEncoder(IntPtr context) { _ctx = context; } // CudaJpegDecoder.Context

        /// <summary>Whole-frame input instead of SetInputReader: interleaved pixels, tightly packed.</summary>
        public void SetInput(ReadOnlyMemory<byte> pixels, int width, int height, int format)
        {
            _pixels = pixels; _width = width; _height = height; _format = format;
        }

        public new void SetOutput(IBufferWriter<byte> output) { _output = output; base.SetOutput(output); }
        public new void SetQuantizationTable(JpegQuantizationTable table)
        {
            base.SetQuantizationTable(table);              // validation and exceptions of the reference (:102-135)
            _quant[table.Identifier] = table;
        }
        public new void AddComponent(byte componentIndex, byte tq, byte td, byte ta, byte h, byte v)
        {
            base.AddComponent(componentIndex, tq, td, ta, h, v); // "Subsampling factor can only be 1, 2 or 4." etc. (:175-214)
            _components.Add((componentIndex, tq, td, ta, h, v));
        }

        public override void Encode()
        {
            if (_pixels.IsEmpty) { base.Encode(); return; } // block reader path: stays managed
            IBufferWriter<byte> output = _output ?? throw new InvalidOperationException("Output is not specified.");
            if (_components.Count == 0) throw new InvalidOperationException("No component is specified.");

            Native.EncodeDesc e = default;
            e.Width = (ushort)_width; e.Height = (ushort)_height; e.Format = _format;
            e.ComponentCount = (byte)_components.Count;
            for (int i = 0; i < _components.Count; i++)
            {
                var c = _components[i];
                e.H[i] = c.h; e.V[i] = c.v; e.Tq[i] = c.tq; e.Td[i] = c.td; e.Ta[i] = c.ta;
                JpegQuantizationTable q = _quant[c.tq] ?? throw new ArgumentException("Quantization table is not defined.");
                for (int k = 0; k < 64; k++) e.Quant[c.tq * 64 + k] = q.Elements[k]; // zig-zag order, as in the DQT segment
                e.QuantPresent[c.tq] = 1;
            }
            using var pin = _pixels.Pin();
            e.Pixels = pin.Pointer;

            Native.Check(_ctx, Native.jb_encode_batch_create(_ctx, &e, 1, out IntPtr batch));
            try
            {
                Native.Check(_ctx, Native.jb_encode_batch_transform(batch));            // E1-E6
                // E7 with the library's own builder: histograms back (8 KB), DHT specs forth
                uint[] hist = new uint[8 * 256];
                fixed (uint* ph = hist) Native.Check(_ctx, Native.jb_encode_batch_histograms(batch, ph, 1));
                var tables = new JpegHuffmanEncodingTableCollection();
                foreach (var c in _components)
                {
                    InstallTable(batch, ref tables, hist, 0, c.td);
                    InstallTable(batch, ref tables, hist, 1, c.ta);
                }
                Native.Check(_ctx, Native.jb_encode_batch_pack(batch));                 // E8 + E9
                Native.Check(_ctx, Native.jb_encode_batch_finish(batch));

                var writer = new JpegWriter(output, 4096);
                WriteStartOfImage(ref writer);                                          // protected helpers of the base class
                WriteQuantizationTables(ref writer);                                    //   (JpegEncoder.cs:296-387, :930)
                WriteStartOfFrame(ref writer);
                tables.Write(ref writer);
                WriteStartOfScan(ref writer);
                ulong length;
                Native.jb_encode_batch_scan_length(batch, 0, &length);
                Span<byte> dst = writer.GetSpan((int)length);                           // stuffed + padded scan bytes, device -> writer
                fixed (byte* pd = dst) Native.Check(_ctx, Native.jb_encode_batch_read_scan(batch, 0, pd, length));
                writer.Advance((int)length);
                WriteEndOfImage(ref writer);
                writer.Flush();
            }
            finally { Native.jb_encode_batch_destroy(batch); }
        }

        // (JpegHuffmanEncodingTableCollection is a mutable struct that creates its list on the first AddTable: by reference)
        private void InstallTable(IntPtr batch, ref JpegHuffmanEncodingTableCollection tables, uint[] hist, int tableClass, byte id)
        {
            if (tables.GetTable(tableClass == 0, id) is not null) return;
            var builder = new JpegHuffmanEncodingTableBuilder();
            for (int s = 0; s < 256; s++)
                for (uint n = hist[(tableClass * 4 + id) * 256 + s]; n != 0; n--) builder.IncrementCodeCount(s); // (or a SetFrequency overload)
            JpegHuffmanEncodingTable table = builder.Build(MostOptimalCoding);
            tables.AddTable((byte)tableClass, id, table);
            Native.HuffSpec spec = default;
            spec.TableClass = (byte)tableClass; spec.Identifier = id;
            Span<byte> dht = stackalloc byte[16 + 256];
            table.TryWrite(dht, out int written);                                      // 16 counts + the symbols (DHT body)
            for (int i = 0; i < 16; i++) spec.Bits[i] = dht[i];
            for (int i = 16; i < written; i++) spec.Values[i - 16] = dht[i];
            spec.ValueCount = (ushort)(written - 16);
            Native.Check(_ctx, Native.jb_encode_batch_set_table(batch, 0, &spec));
        }
    }
}
