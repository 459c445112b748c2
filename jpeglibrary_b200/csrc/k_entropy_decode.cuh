// k_entropy_decode.cuh -- K0 (restart-marker index) and the byte-stream bit reader / 16-bit table look-up used by the
// progressive (K1c) and lossless (K1d) scan kernels.  The baseline decoder proper is k_entropy_flat.cuh.
//
// Replaces, for a whole batch of images at once:
//   JpegBitReader.FillBuffer / PeekBits / TryReadBits        (JpegBitReader.cs:95-204)
//   JpegHuffmanDecodingTable.Lookup / LookupSlow             (JpegHuffmanDecodingTable.cs:73-113)
//   DecodeHuffmanCode / ReceiveAndExtend                     (ScanDecoder/JpegHuffmanScanDecoder.cs:81-115)
//   ReadBlockBaseline + MCU loop + restart handling          (ScanDecoder/JpegHuffmanBaselineScanDecoder.cs:99-222)
#pragma once
#include "jb_device.cuh"

// ---------------------------------------------------------------------------------------------
// K0: restart-marker index.  One CTA per image walks the entropy-coded bytes in 4 KB tiles and
// records, in stream order, every FF xx with xx not in {00, FF}: RSTn markers (xx = D0..D7) and
// the first other marker, which terminates the scan (JpegBitReader.cs:108-128 semantics).
// HBM-bound: reads the compressed bytes once, writes ~4 B per restart interval.
// The compressed bytes are BULK-LOADED BY TMA (round 2): one elected thread issues a 1-D cp.async.bulk of a whole tile
// (+16 look-ahead bytes) into shared memory, completion is counted in bytes on an mbarrier, and two tiles are always
// in flight (double buffer), so the CTA's serial tile loop -- load, test, barrier -- no longer waits a DRAM round trip
// per tile, at no cost in registers (a register prefetch of the next tile had made the batch slower in round 1).
// SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK.TRANS64.TRYWAIT.
// ---------------------------------------------------------------------------------------------
#define JB_K0_THREADS 256
#ifndef JB_K0_TMA
#define JB_K0_TMA 1
#endif
#define JB_K0_TILE (JB_K0_THREADS * 16)
#define JB_K0_TILE_STRIDE (JB_K0_TILE + 128) // tile + look-ahead, 128-byte aligned

__device__ __forceinline__ uint32_t jb_k0_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t jb_ff_bytes(uint32_t w)
{
    // 0x80 in every byte of w that equals 0xFF
    uint32_t x = ~w;
    return (x - 0x01010101u) & ~x & 0x80808080u;
}

// A "zero byte" detector has false positives above a true zero byte only (borrow propagation);
// callers re-check each flagged byte, so this is only used as a fast reject.

// exact SWAR byte classifiers: 0x80 in every byte of w that equals 0xFF / 0x00
__device__ __forceinline__ uint32_t jb_is_ff(uint32_t w)
{
    const uint32_t inv = ~w;
    return ~(((inv & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | inv) & 0x80808080u;
}
__device__ __forceinline__ uint32_t jb_is_00(uint32_t w) { return ~(((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u; }

// markers (FF followed by a byte that is neither 00 nor FF) among the 16 bytes w[0..3]; w[4] low byte = look-ahead byte.
// Branch-free: the per-byte loop below ran with a handful of active lanes in every warp that met an FF byte (one 16-byte
// group in five holds one), which made K0 issue-bound at 11 of 32 lanes per instruction.
__device__ __forceinline__ uint32_t jb_count_markers16(const uint32_t w[5])
{
    uint32_t ff[5], stop[5]; // stop: bytes that make a preceding FF a non-marker (00 or FF)
#pragma unroll
    for (int i = 0; i < 5; i++) { ff[i] = jb_is_ff(w[i]); stop[i] = ff[i] | jb_is_00(w[i]); }
    uint32_t n = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) n += __popc(ff[i] & ~((stop[i] >> 8) | (stop[i + 1] << 24)));
    return n;
}

template <typename F>
__device__ __forceinline__ void jb_foreach_marker(const uint32_t w[5], uint32_t pos0, uint32_t len, uint32_t skew, F f)
{
    // w[0..3] = 16 bytes at pos0, w[4] low byte = look-ahead byte
    if ((jb_ff_bytes(w[0]) | jb_ff_bytes(w[1]) | jb_ff_bytes(w[2]) | jb_ff_bytes(w[3])) == 0) return;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t b = (w[i >> 2] >> ((i & 3) * 8)) & 0xFF;
        uint32_t nb = (w[(i + 1) >> 2] >> (((i + 1) & 3) * 8)) & 0xFF;
        if (b == 0xFF && nb != 0 && nb != 0xFF && pos0 + i + 1 < len && pos0 + i >= skew) f(pos0 + i - skew, nb);
    }
}

__global__ void __launch_bounds__(JB_K0_THREADS)
jb_k0_restart_scan(const JbScanRange *__restrict__ ranges, const uint8_t *__restrict__ arena,
                   uint32_t *__restrict__ marks, JbScanResult *__restrict__ results)
{
    const JbScanRange &im = ranges[blockIdx.x];
    // ranges of progressive scans start at arbitrary bytes: index from the enclosing 16-byte line
    const uint32_t skew = (uint32_t)(im.data_off & 15u);
    const uint8_t *data = arena + (im.data_off - skew);
    const uint32_t len = im.data_len ? im.data_len + skew : 0;
    const uint32_t cap = im.mark_cap;
    uint32_t *out = marks + im.mark_base;

    __shared__ uint32_t s_warp[JB_K0_THREADS / 32];
    __shared__ uint32_t s_base, s_term_idx, s_term_pos, s_term_marker;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#if JB_K0_TMA
    __shared__ __align__(128) uint8_t s_tile[2][JB_K0_TILE_STRIDE];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const uint32_t ntiles = (len + JB_K0_TILE - 1) / JB_K0_TILE;
    // tile t -> buffer t & 1: the tile's bytes and 16 more (the look-ahead of its last thread), never past the image's
    // arena slot (64 spare bytes behind every image)
    auto issue = [&](uint32_t t) {
        const uint32_t at = t * JB_K0_TILE;
        const uint32_t bytes = min((uint32_t)JB_K0_TILE + 16u, ((len - at + 15u) & ~15u) + 16u);
        const uint32_t bar = jb_k0_smem(&s_bar[t & 1]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(jb_k0_smem(s_tile[t & 1])), "l"(data + at), "r"(bytes), "r"(bar) : "memory");
    };
    auto wait = [&](uint32_t t) {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(jb_k0_smem(&s_bar[t & 1])), "r"((t >> 1) & 1u) : "memory");
    };
#endif
    if (tid == 0) {
        s_base = 0;
        s_term_idx = 0xFFFFFFFFu;
        s_term_pos = 0xFFFFFFFFu;
        s_term_marker = 0;
#if JB_K0_TMA
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(jb_k0_smem(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(jb_k0_smem(&s_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (ntiles > 0) issue(0);
        if (ntiles > 1) issue(1);
#endif
    }
    __syncthreads();

#if JB_K0_TMA
    uint32_t tnext = 0; // first tile that has not been waited for
    for (uint32_t t = 0; t < ntiles; t++) {
        const uint32_t tile = t * JB_K0_TILE;
#else
    for (uint32_t tile = 0; tile < len; tile += JB_K0_THREADS * 16) {
#endif
        const uint32_t pos0 = tile + tid * 16;
        uint32_t w[5] = {0, 0, 0, 0, 0};
#if JB_K0_TMA
        wait(t);
        tnext = t + 1;
        if (pos0 < len) {
            const uint4 v = *reinterpret_cast<const uint4 *>(s_tile[t & 1] + tid * 16);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            w[4] = *reinterpret_cast<const uint32_t *>(s_tile[t & 1] + tid * 16 + 16);
        }
#else
        if (pos0 < len) {
            // the arena is zero-padded by >= 32 bytes after every image: the over-read is safe
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(data + pos0));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            w[4] = __ldg(reinterpret_cast<const uint32_t *>(data + pos0 + 16));
        }
#endif
        uint32_t cnt = 0;
        if (pos0 >= skew && pos0 + 17 < len) cnt = jb_count_markers16(w); // all 16 bytes and their successors lie inside the range
        else jb_foreach_marker(w, pos0, len, skew, [&](uint32_t, uint32_t) { cnt++; });
        const bool any = __syncthreads_or(cnt != 0);
#if JB_K0_TMA
        // every thread holds its bytes in registers now: the buffer takes the tile after next (a terminator found in
        // THIS tile stops the walk below; one bulk copy too many is waited for at the end)
        if (tid == 0 && t + 2 < ntiles) issue(t + 2);
#endif
        if (!any) continue;

        // block-wide exclusive scan of cnt (rare path: only tiles that contain a marker)
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t warp_off = 0, total = 0;
#pragma unroll
        for (int i = 0; i < JB_K0_THREADS / 32; i++) {
            uint32_t t = s_warp[i];
            if (i < wid) warp_off += t;
            total += t;
        }
        uint32_t idx = s_base + warp_off + incl - cnt;
        jb_foreach_marker(w, pos0, len, skew, [&](uint32_t pos, uint32_t m) {
            const bool rst = (m & 0xF8u) == 0xD0u;
            if (idx < cap) out[idx] = (pos << 4) | (rst ? (m & 7u) : 8u);
            if (!rst) {
                uint32_t old = atomicMin(&s_term_idx, idx);
                if (idx < old) { // this thread owns the earliest terminator so far in its own view
                    atomicMin(&s_term_pos, pos);
                }
            }
            idx++;
        });
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
        if (s_term_idx != 0xFFFFFFFFu) break;
    }
#if JB_K0_TMA
    // bulk copies still in flight (the walk stopped at a terminator) must land before the CTA gives its shared memory up
    if (tid == 0)
        for (uint32_t t = tnext; t < min(ntiles, tnext + 2); t++) wait(t);
#endif
    __syncthreads();
    if (tid == 0) {
        JbScanResult r;
        uint32_t n = s_base;
        r.end_pos = len - skew;
        r.end_marker = 0;
        if (s_term_idx != 0xFFFFFFFFu) {
            n = s_term_idx + 1;
            if (s_term_idx < cap) {
                r.end_pos = out[s_term_idx] >> 4;
                r.end_marker = data[r.end_pos + skew + 1];
            } else {
                r.end_pos = s_term_pos;
            }
        }
        r.nmarkers = n < cap ? n : cap;
        r.pad = 0;
        results[blockIdx.x] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// Bit reader: 64-bit MSB-first window (hi:lo) refilled 32 bits at a time from the stuffed byte
// stream.  The stream is read as aligned 32-bit words held in registers (w0 = word under the read
// position, w1 = next, w2 = software prefetch), so the common refill is: "position aligned, no
// 0xFF in w0" -> append the byte-swapped word.  Everything else (FF 00 -> FF, FF FF fill bytes,
// misalignment after a stuffed pair or at the segment start, the segment end) goes through a
// byte-wise path that works on the registers and returns as soon as the position is aligned again.
// At the segment end the window is padded with 1-bits exactly like PeekBits does
// (JpegBitReader.cs:166); `pad` counts them so that consuming padding as magnitude bits is
// reported like ReceiveAndExtend's failure (JpegHuffmanScanDecoder.cs:100-110).
// ---------------------------------------------------------------------------------------------
struct JbBitReader {
    const uint8_t *data; // 256-byte aligned
    uint32_t pos, end;
    uint32_t w0, w1, w2;
    uint32_t hi, lo;
    int n;   // valid bits in hi:lo
    int pad; // of which padding (always the last `pad` bits)

    __device__ __forceinline__ uint32_t ldw(uint32_t byte_off) const
    {
        // plain read-only load: the 128-byte line stays in L1 for the next 31 words of this lane.  (An
        // L1::no_allocate load is tagged evict-first in L2 as well: profiles/r1d showed every 4-byte word
        // coming from DRAM.)
        return __ldg(reinterpret_cast<const uint32_t *>(data + byte_off));
    }
    __device__ __forceinline__ void init(const uint8_t *d, uint32_t start, uint32_t stop)
    {
        data = d; pos = start; end = stop; hi = lo = 0; n = 0; pad = 0;
        const uint32_t wp = start & ~3u;
        w0 = ldw(wp); w1 = ldw(wp + 4); w2 = ldw(wp + 8);
    }
    __device__ __forceinline__ void put(uint32_t w, int bits)
    { // append `bits` (0..32) bits held left-aligned in w; requires n <= 32
        hi |= __funnelshift_rc(w, 0u, n);
        lo |= __funnelshift_rc(0u, w, n);
        n += bits;
    }
    __device__ __forceinline__ void next_word()
    { // pos has just become a multiple of 4
        w0 = w1; w1 = w2;
        w2 = ldw(pos + 8);
    }
    __device__ __forceinline__ void refill_slow()
    {
        uint32_t w = 0;
        int bits = 0;
        do {
            if (pos >= end) {
                w |= 0xFFFFFFFFu >> bits;
                pad += 32 - bits;
                bits = 32;
                break;
            }
            const uint32_t o = pos & 3u;
            const uint32_t b = (w0 >> (8 * o)) & 0xFF;
            if (b == 0xFF) {
                const uint32_t b2 = o < 3 ? (w0 >> (8 * o + 8)) & 0xFF : w1 & 0xFF;
                if (b2 != 0xFF && b2 != 0) { // a marker inside the segment: it ends here
                    end = pos;
                    continue;
                }
                if (b2 == 0) { // stuffed zero: FF is data, skip both bytes
                    w |= 0xFFu << (24 - bits);
                    bits += 8;
                    if (((++pos) & 3u) == 0) next_word();
                }
                // (fill byte: the first FF is dropped)
            } else {
                w |= b << (24 - bits);
                bits += 8;
            }
            if (((++pos) & 3u) == 0) next_word();
        } while (bits < 32 && (pos & 3u) != 0);
        put(w, bits);
    }
    __device__ __forceinline__ void refill()
    { // call when n <= 32
        if ((pos & 3u) == 0 && pos + 4 <= end && jb_ff_bytes(w0) == 0) {
            put(__byte_perm(w0, 0, 0x0123), 32);
            pos += 4;
            next_word();
            return;
        }
        refill_slow();
    }
    __device__ __forceinline__ void ensure32()
    {
        while (n <= 32) refill();
    }
    __device__ __forceinline__ uint32_t peek16() const { return hi >> 16; }
    __device__ __forceinline__ void skip(int k)
    { // k in 0..31
        hi = __funnelshift_l(lo, hi, k);
        lo <<= k;
        n -= k;
    }
    // DecodeHuffmanCode advances min(code size, bits available) (JpegHuffmanScanDecoder.cs:85-86): a code whose tail lies
    // in the 1-bit padding behind the data is accepted, only magnitude bits that are not there are an error
    __device__ __forceinline__ void skip_code(int k)
    {
        skip(k);
        n = max(n, pad);
    }
    __device__ __forceinline__ uint32_t take(int k)
    { // k in 1..16
        uint32_t v = hi >> (32 - k);
        skip(k);
        return v;
    }
};

__device__ __forceinline__ int jb_extend(int v, int nbits)
{ // JpegHuffmanScanDecoder.cs:114
    return v - ((((v + v) >> nbits) - 1) & ((1 << nbits) - 1));
}

// LookupSlow, JpegHuffmanDecodingTable.cs:88-113; returns (symbol << 8) | size or 0xFFFFFFFF
__device__ __noinline__ uint32_t jb_huff_lookup_slow(const JbHuffTable *t, uint32_t code16)
{
    int size = 9;
    while (code16 > __ldg(&t->maxcode[size])) size++;
    if (size > 16) return 0xFFFFFFFFu;
    uint32_t sym = __ldg(&t->values[(__ldg(&t->valoffset[size]) + (code16 >> (16 - size))) & 0xFF]);
    return (sym << 8) | (uint32_t)size;
}

// returns (symbol << 8) | size, or 0xFFFFFFFF for an invalid code.  The tables (a few KB, shared by
// every lane of every warp) are read through the read-only path and stay L1-resident; keeping them
// out of shared memory is what lets all restart segments of a 1024-image batch be resident at once.
__device__ __forceinline__ uint32_t jb_huff_lookup(const JbHuffTable *t, uint32_t code16)
{
    uint32_t e = __ldg(&t->lut[code16 >> (16 - JB_LUT_BITS)]);
    if ((e & 0xFF) != 0) return e;
    if (e != 0) {
        e = __ldg(&t->lut2[((e >> 8) - 1) * 64 + (code16 & 63)]);
        if (e != 0) return e;
    }
    return jb_huff_lookup_slow(t, code16);
}
