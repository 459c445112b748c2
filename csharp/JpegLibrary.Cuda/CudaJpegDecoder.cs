// Sketch of the drop-in: JpegDecoder keeps its marker loop (ProcessMarkerForDecode is a protected virtual
// hook, JpegDecoder.cs:558) and hands each SOS to the GPU instead of JpegScanDecoder.Create (:592-599).
// NOT compiled here (no .NET toolchain in the build image).
using System;
using System.Buffers;

namespace JpegLibrary.Cuda
{
    /// <summary>Recognised GPU-aware sink: kernels store RGB straight into its (pinned or device) buffer;
    /// WriteBlock is never called on the fast path (SURVEY 8b).</summary>
    public sealed class CudaRgbOutputWriter : JpegBlockOutputWriter
    {
        public IntPtr Buffer; public long Pitch; public long Capacity; public bool OnDevice; public int Format = Native.JB_OUT_RGB24;
        public override void WriteBlock(ref short blockRef, int componentIndex, int x, int y) =>
            throw new InvalidOperationException("filled by the GPU");
    }

    public sealed class CudaJpegDecoder : JpegDecoder
    {
        private readonly IntPtr _ctx;
        public CudaJpegDecoder(int device = 0) { Native.Check(IntPtr.Zero, Native.jb_ctx_create(device, out _ctx)); }

        protected override bool ProcessMarkerForDecode(JpegMarker marker, ref JpegReader reader)
        {
            if (marker != JpegMarker.StartOfScan) return base.ProcessMarkerForDecode(marker, ref reader); // DHT/DQT/DRI/SOF stay managed
            // 1. parse the scan header with the reference's own JpegScanHeader.TryParse
            // 2. fill Native.ImageDesc from GetFrameHeader(), GetHuffmanTable(), GetQuantizationTable(), GetRestartInterval()
            // 3. if the output writer is a CudaRgbOutputWriter: jb_decode(ctx, &desc, &out, 1, null)
            //    else: request JB_OUT_PLANAR_I16 into a pinned buffer and replay WriteBlock in the reference's
            //    order (JpegHuffmanBaselineScanDecoder.cs:99-137, 238-268) -- see jpeglibrary_b200/api.py
            //    _replay_write_blocks for the exact sequence
            // 4. advance `reader` past the entropy-coded segment (JpegReader.TryReadMarker skips it anyway)
            throw new NotImplementedException("illustrative");
        }
    }
}
