// Batch facade (SURVEY 8b): one JpegDecoder.Decode() exposes a single image of parallelism; a batch exposes
// hundreds of thousands of restart segments.  Every stream is walked by the reference's own marker loop (a
// CudaJpegDecoder in "collect" mode would fill the descriptors; shown here with the descriptors already built),
// then ONE jb_decode_batch_* sequence decodes all of them.  Python twin: jpeglibrary_b200/api.py JpegBatchDecoder /
// JpegPipelinedBatchDecoder.  NOT compiled in this repository's build image (no .NET toolchain).
using System;

namespace JpegLibrary.Cuda
{
    public sealed unsafe class CudaJpegBatchDecoder : IDisposable
    {
        private readonly IntPtr _ctx;
        private IntPtr _batch;
        private readonly int _count;

        /// <param name="images">descriptors filled from JpegDecoder.Identify + the marker loop (see CudaJpegDecoder.Submit)</param>
        /// <param name="outputs">one destination per image: device pointer (OnDevice = 1) or pinned host memory</param>
        internal CudaJpegBatchDecoder(IntPtr ctx, Native.ImageDesc[] images, Native.OutputDesc[] outputs)
        {
            _ctx = ctx; _count = images.Length;
            fixed (Native.ImageDesc* pi = images)
            fixed (Native.OutputDesc* po = outputs)
                Native.Check(ctx, Native.jb_decode_batch_create(ctx, pi, po, _count, out _batch));
        }

        /// <summary>cudaMemcpyAsync of the compressed bytes (pinned input avoids a staging copy).</summary>
        public void Upload() => Native.Check(_ctx, Native.jb_decode_batch_upload(_batch));
        /// <summary>K0..K2 on the context's stream; results stay in device memory.</summary>
        public void Launch() => Native.Check(_ctx, Native.jb_decode_batch_launch(_batch));
        /// <summary>D2H of pixel outputs that live on the host, then per-image status -> the reference's exceptions.</summary>
        public void Finish() => Native.Check(_ctx, Native.jb_decode_batch_finish(_batch));
        public void Run() => Native.Check(_ctx, Native.jb_decode_batch_run(_batch));

        public int[] Status()
        {
            int[] st = new int[_count];
            fixed (int* p = st) Native.jb_decode_batch_status(_batch, p, _count);
            return st;
        }

        public void Dispose()
        {
            if (_batch != IntPtr.Zero) { Native.jb_decode_batch_destroy(_batch); _batch = IntPtr.Zero; }
        }
    }
}
