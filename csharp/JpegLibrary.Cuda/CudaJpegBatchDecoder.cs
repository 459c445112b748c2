// Batch facade (SURVEY 8b): one JpegDecoder.Decode() exposes a single image of parallelism; a batch exposes
// hundreds of thousands of restart segments.  Every stream is walked by the reference's own marker loop
// (CudaJpegDecoder.Collect: Identify + the marker loop of Decode, nothing decoded), then ONE jb_decode_batch_*
// sequence decodes all of them.  Python twin: jpeglibrary_b200/api.py JpegBatchDecoder /
// JpegPipelinedBatchDecoder.  NOT compiled in this repository's build image (no .NET toolchain).
using System;
using System.Collections.Generic;

namespace JpegLibrary.Cuda
{
    public sealed unsafe class CudaJpegBatchDecoder : IDisposable
    {
        private readonly IntPtr _ctx;
        private IntPtr _batch;
        private readonly int _count;
        private CudaJpegDecoder.CollectedImage[]? _held; // descriptors' scan lists, tables and inputs, pinned until Dispose

        /// <summary>Walks every stream with the reference's marker loop and builds one batch.  `tables`: a tables stream
        /// every stream is an abbreviated stream of (JpegDecoder.LoadTables: the strips of a TIFF file), or empty.
        /// A stream the walk refuses throws what JpegDecoder.Decode() throws for it.</summary>
        public static CudaJpegBatchDecoder Create(IntPtr ctx, IReadOnlyList<ReadOnlyMemory<byte>> streams, IReadOnlyList<CudaRgbOutputWriter> outputs,
                                                  ReadOnlyMemory<byte> tables = default)
        {
            if (streams.Count != outputs.Count) throw new ArgumentException("one output per stream", nameof(outputs));
            int n = streams.Count;
            var held = new CudaJpegDecoder.CollectedImage[n];
            var images = new Native.ImageDesc[n];
            var outs = new Native.OutputDesc[n];
            try
            {
                for (int i = 0; i < n; i++) // (independent walks: a caller with many cores runs them on a Parallel.For)
                {
                    using var walker = new CudaJpegDecoder(ctx);
                    if (!tables.IsEmpty) walker.LoadTables(new System.Buffers.ReadOnlySequence<byte>(tables));
                    walker.SetInput(streams[i]);
                    walker.Identify();
                    held[i] = walker.Collect();
                    images[i] = held[i].Desc;
                    CudaRgbOutputWriter w = outputs[i];
                    outs[i] = new Native.OutputDesc { Dst = (void*)w.Buffer, Pitch = (ulong)w.Pitch, Capacity = (ulong)w.Capacity, Format = w.Format, OnDevice = w.OnDevice ? 1 : 0 };
                }
                return new CudaJpegBatchDecoder(ctx, images, outs) { _held = held };
            }
            catch
            {
                foreach (CudaJpegDecoder.CollectedImage? h in held) h?.Dispose();
                throw;
            }
        }

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
            if (_held != null) { foreach (CudaJpegDecoder.CollectedImage h in _held) h.Dispose(); _held = null; }
        }
    }
}
