// MemoryPool<byte> whose blocks are page-locked (jb_pinned_alloc): JpegDecoder.SetInput(ReadOnlyMemory<byte>)
// (JpegDecoder.cs:49) and the RGB destination of CudaRgbOutputWriter can then point straight into DMA-able memory, so
// H2D / D2H run at PCIe speed without a staging copy.  Python twin: Context.pinned_array (jpeglibrary_b200/api.py).
// NOT compiled in this repository's build image (no .NET toolchain).
using System;
using System.Buffers;

namespace JpegLibrary.Cuda
{
    public sealed unsafe class PinnedMemoryPool : MemoryPool<byte>
    {
        private readonly IntPtr _ctx;
        public PinnedMemoryPool(IntPtr ctx) { _ctx = ctx; }

        public override int MaxBufferSize => int.MaxValue;

        public override IMemoryOwner<byte> Rent(int minBufferSize = -1)
        {
            int size = minBufferSize <= 0 ? 1 << 20 : minBufferSize;
            Native.Check(_ctx, Native.jb_pinned_alloc(_ctx, (UIntPtr)(uint)size, out IntPtr p));
            return new Block(_ctx, p, size);
        }

        protected override void Dispose(bool disposing) { }

        private sealed class Block : MemoryManager<byte>
        {
            private readonly IntPtr _ctx;
            private IntPtr _p;
            private readonly int _size;
            public Block(IntPtr ctx, IntPtr p, int size) { _ctx = ctx; _p = p; _size = size; }
            public override Span<byte> GetSpan() => new Span<byte>((void*)_p, _size);
            public override MemoryHandle Pin(int elementIndex = 0) => new MemoryHandle((byte*)_p + elementIndex); // already pinned
            public override void Unpin() { }
            protected override void Dispose(bool disposing)
            {
                if (_p != IntPtr.Zero) { Native.jb_pinned_free(_ctx, _p); _p = IntPtr.Zero; }
            }
        }
    }
}
