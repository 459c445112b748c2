// P/Invoke binding of include/jpegb200.h.  NOT compiled in this repository's image (no .NET toolchain):
// it is the reference-side stub a JpegLibrary maintainer adds; struct layouts mirror the C header 1:1.
using System;
using System.Runtime.InteropServices;

namespace JpegLibrary.Cuda
{
    internal static unsafe class Native
    {
        private const string Lib = "jpegb200"; // libjpegb200.so

        public const int JB_OK = 0, JB_ERR_INVALID_DATA = -1, JB_ERR_INVALID_OPERATION = -2, JB_ERR_NOT_SUPPORTED = -3,
                         JB_ERR_ARGUMENT = -4, JB_ERR_NO_DEVICE = -5, JB_ERR_CUDA = -6, JB_ERR_NOMEM = -7;
        public const int JB_OUT_RGB24 = 0, JB_OUT_RGBA32 = 1, JB_OUT_YCBCR888 = 2, JB_OUT_PLANAR_I16 = 3, JB_OUT_COEFFICIENTS = 4;

        [StructLayout(LayoutKind.Sequential)]
        public struct HuffSpec { public byte TableClass, Identifier; public fixed byte Bits[16]; public fixed byte Values[256]; public ushort ValueCount; }

        [StructLayout(LayoutKind.Sequential)]
        public struct ScanDesc
        {
            public byte ComponentCount; public fixed byte ComponentIndex[4];
            public fixed short DcTable[4]; public fixed short AcTable[4];
            public byte Ss, Se, Ah, Al; public uint RestartInterval; public ulong EntropyOffset, EntropyLength;
        }

        [StructLayout(LayoutKind.Sequential)]
        public struct ImageDesc
        {
            public byte* Data; public ulong Length;
            public byte Sof, Precision, ComponentCount, Reserved0; public ushort Width, Height;
            public fixed byte H[4]; public fixed byte V[4]; public fixed ushort Quant[4 * 64];
            public uint ScanCount; public ScanDesc* Scans; public uint TableCount; public HuffSpec* Tables;
        }

        [StructLayout(LayoutKind.Sequential)]
        public struct OutputDesc { public void* Dst; public ulong Pitch, Capacity; public int Format, OnDevice; }

        [DllImport(Lib)] public static extern int jb_device_count();
        [DllImport(Lib)] public static extern int jb_ctx_create(int device, out IntPtr ctx);
        [DllImport(Lib)] public static extern void jb_ctx_destroy(IntPtr ctx);
        [DllImport(Lib)] public static extern IntPtr jb_last_error(IntPtr ctx);
        [DllImport(Lib)] public static extern int jb_pinned_alloc(IntPtr ctx, UIntPtr bytes, out IntPtr p);
        [DllImport(Lib)] public static extern int jb_pinned_free(IntPtr ctx, IntPtr p);
        [DllImport(Lib)] public static extern int jb_decode_batch_create(IntPtr ctx, ImageDesc* images, OutputDesc* outputs, int count, out IntPtr batch);
        [DllImport(Lib)] public static extern int jb_decode_batch_run(IntPtr batch);
        [DllImport(Lib)] public static extern int jb_decode_batch_status(IntPtr batch, int* status, int count);
        [DllImport(Lib)] public static extern void jb_decode_batch_destroy(IntPtr batch);
        [DllImport(Lib)] public static extern int jb_decode(IntPtr ctx, ImageDesc* images, OutputDesc* outputs, int count, int* status);
        [DllImport(Lib)] public static extern int jb_decode_batch_upload(IntPtr batch);
        [DllImport(Lib)] public static extern int jb_decode_batch_launch(IntPtr batch);
        [DllImport(Lib)] public static extern int jb_decode_batch_finish(IntPtr batch);
        [DllImport(Lib)] public static extern int jb_ctx_synchronize(IntPtr ctx);
        [DllImport(Lib)] public static extern int jb_ctx_trim(IntPtr ctx);
        [DllImport(Lib)] public static extern int jb_device_alloc(IntPtr ctx, UIntPtr bytes, out IntPtr p);
        [DllImport(Lib)] public static extern int jb_device_free(IntPtr ctx, IntPtr p);

        // encode path (JpegEncoder.TransformBlocks / BuildHuffmanTables / WritePreparedScanData, JpegEncoder.cs:414-656)
        [StructLayout(LayoutKind.Sequential)]
        public struct EncodeDesc
        {
            public void* Pixels; public ulong Pitch; public int OnDevice, Format;
            public ushort Width, Height; public byte ComponentCount;
            public fixed byte H[4]; public fixed byte V[4]; public fixed byte Tq[4]; public fixed byte Td[4]; public fixed byte Ta[4];
            public byte Reserved; public ushort RestartInterval /* transcoding (JB_IN_COEFFICIENTS) only */; public fixed ushort Quant[4 * 64]; public fixed byte QuantPresent[4];
        }
        [DllImport(Lib)] public static extern int jb_encode_batch_create(IntPtr ctx, EncodeDesc* images, int count, out IntPtr batch);
        [DllImport(Lib)] public static extern int jb_encode_batch_transform(IntPtr batch);
        [DllImport(Lib)] public static extern int jb_encode_batch_histograms(IntPtr batch, uint* hist, int count);
        [DllImport(Lib)] public static extern int jb_encode_batch_build_tables(IntPtr batch);
        [DllImport(Lib)] public static extern int jb_encode_batch_set_table(IntPtr batch, int image, HuffSpec* table);
        [DllImport(Lib)] public static extern int jb_encode_batch_pack(IntPtr batch);
        [DllImport(Lib)] public static extern int jb_encode_batch_finish(IntPtr batch);
        [DllImport(Lib)] public static extern int jb_encode_batch_get_table(IntPtr batch, int image, int tableClass, int identifier, HuffSpec* table);
        [DllImport(Lib)] public static extern int jb_encode_batch_scan_length(IntPtr batch, int image, ulong* length);
        [DllImport(Lib)] public static extern int jb_encode_batch_read_scan(IntPtr batch, int image, byte* dst, ulong capacity);
        [DllImport(Lib)] public static extern void jb_encode_batch_destroy(IntPtr batch);
        // table construction on the native side for callers without a managed builder at hand (the encoder subclass uses
        // JpegHuffmanEncodingTableBuilder.Build(MostOptimalCoding) itself)
        [DllImport(Lib)] public static extern int jb_build_huffman_table(uint* frequencies, int tableClass, int identifier, HuffSpec* table);
        [DllImport(Lib)] public static extern int jb_build_huffman_table_optimal(uint* frequencies, int tableClass, int identifier, HuffSpec* table);

        // status code -> the exception the managed decoder throws in the same situation
        public static void Check(IntPtr ctx, int rc)
        {
            if (rc == JB_OK) return;
            string msg = Marshal.PtrToStringAnsi(jb_last_error(ctx)) ?? "jpegb200 error";
            switch (rc)
            {
                case JB_ERR_INVALID_DATA: throw new System.IO.InvalidDataException(msg);      // JpegDecoder.cs:365-375
                case JB_ERR_INVALID_OPERATION: throw new InvalidOperationException(msg);      // "Expect restart marker."
                case JB_ERR_NOT_SUPPORTED: throw new NotSupportedException(msg);              // JpegDecoder.cs:627-630
                case JB_ERR_ARGUMENT: throw new ArgumentException(msg);
                default: throw new InvalidOperationException("CUDA: " + msg);
            }
        }
    }
}
