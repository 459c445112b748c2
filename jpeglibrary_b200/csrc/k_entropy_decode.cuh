// k_entropy_decode.cuh -- K0 (restart-marker index) and K1a (restart-segment-parallel Huffman decode).
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
// HBM-bound: reads the compressed bytes once (uint4 per thread), writes ~4 B per restart interval.
// ---------------------------------------------------------------------------------------------
#define JB_K0_THREADS 256

__device__ __forceinline__ uint32_t jb_ff_bytes(uint32_t w)
{
    // 0x80 in every byte of w that equals 0xFF
    uint32_t x = ~w;
    return (x - 0x01010101u) & ~x & 0x80808080u;
}

// A "zero byte" detector has false positives above a true zero byte only (borrow propagation);
// callers re-check each flagged byte, so this is only used as a fast reject.

template <typename F>
__device__ __forceinline__ void jb_foreach_marker(const uint32_t w[5], uint32_t pos0, uint32_t len, F f)
{
    // w[0..3] = 16 bytes at pos0, w[4] low byte = look-ahead byte
    if ((jb_ff_bytes(w[0]) | jb_ff_bytes(w[1]) | jb_ff_bytes(w[2]) | jb_ff_bytes(w[3])) == 0) return;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t b = (w[i >> 2] >> ((i & 3) * 8)) & 0xFF;
        uint32_t nb = (w[(i + 1) >> 2] >> (((i + 1) & 3) * 8)) & 0xFF;
        if (b == 0xFF && nb != 0 && nb != 0xFF && pos0 + i + 1 < len) f(pos0 + i, nb);
    }
}

__global__ void __launch_bounds__(JB_K0_THREADS)
jb_k0_restart_scan(const JbDevImage *__restrict__ images, const uint8_t *__restrict__ arena,
                   uint32_t *__restrict__ marks, JbScanResult *__restrict__ results)
{
    const JbDevImage &im = images[blockIdx.x];
    const uint8_t *data = arena + im.data_off;
    const uint32_t len = im.data_len;
    const uint32_t cap = im.mark_cap;
    uint32_t *out = marks + im.mark_base;

    __shared__ uint32_t s_warp[JB_K0_THREADS / 32];
    __shared__ uint32_t s_base, s_term_idx, s_term_pos, s_term_marker;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_base = 0;
        s_term_idx = 0xFFFFFFFFu;
        s_term_pos = 0xFFFFFFFFu;
        s_term_marker = 0;
    }
    __syncthreads();

    for (uint32_t tile = 0; tile < len; tile += JB_K0_THREADS * 16) {
        const uint32_t pos0 = tile + tid * 16;
        uint32_t w[5] = {0, 0, 0, 0, 0};
        if (pos0 < len) {
            // the arena is zero-padded by >= 32 bytes after every image: the over-read is safe
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(data + pos0));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            w[4] = __ldg(reinterpret_cast<const uint32_t *>(data + pos0 + 16));
        }
        uint32_t cnt = 0;
        jb_foreach_marker(w, pos0, len, [&](uint32_t, uint32_t) { cnt++; });
        if (!__syncthreads_or(cnt != 0)) continue;

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
        jb_foreach_marker(w, pos0, len, [&](uint32_t pos, uint32_t m) {
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
    __syncthreads();
    if (tid == 0) {
        JbScanResult r;
        uint32_t n = s_base;
        r.end_pos = len;
        r.end_marker = 0;
        if (s_term_idx != 0xFFFFFFFFu) {
            n = s_term_idx + 1;
            if (s_term_idx < cap) {
                r.end_pos = out[s_term_idx] >> 4;
                r.end_marker = data[r.end_pos + 1];
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
// Bit reader: 64-bit MSB-first window (hi:lo), refilled 32 bits at a time from the stuffed
// byte stream.  FF 00 -> FF, FF FF -> fill byte skipped; at the segment end the window is padded
// with 1-bits exactly like PeekBits does (JpegBitReader.cs:166) and `pad` counts them so that
// consuming padding as magnitude bits is reported like ReceiveAndExtend's failure.
// ---------------------------------------------------------------------------------------------
struct JbBitReader {
    const uint8_t *data;
    uint32_t pos, end;
    uint32_t hi, lo;
    int n;   // valid bits in hi:lo
    int pad; // of which padding (always the last `pad` bits)

    __device__ __forceinline__ void init(const uint8_t *d, uint32_t start, uint32_t stop)
    {
        data = d; pos = start; end = stop; hi = lo = 0; n = 0; pad = 0;
    }
    __device__ __forceinline__ void put(uint32_t w, int bits)
    { // append `bits` (8..32) bits held left-aligned in w; requires n <= 32
        hi |= __funnelshift_rc(w, 0u, n);
        lo |= __funnelshift_rc(0u, w, n);
        n += bits;
    }
    __device__ __noinline__ void refill_slow()
    {
        // byte-wise: handles stuffing, fill bytes, misalignment and the segment end
        uint32_t w = 0;
        int bits = 0;
        while (bits < 32) {
            if (pos >= end) {
                w |= 0xFFFFFFFFu >> bits;
                pad += 32 - bits;
                bits = 32;
                break;
            }
            uint32_t b = data[pos++];
            if (b == 0xFF) {
                uint32_t b2 = pos < end ? data[pos] : 0xD9u;
                if (b2 == 0xFF) continue; // fill byte
                if (b2 == 0) pos++;       // stuffed zero
                else {                    // a marker inside the segment: treat as its end
                    pos = end;
                    continue;
                }
            }
            w |= b << (24 - bits);
            bits += 8;
            if ((pos & 3u) == 0 && pos + 4 <= end) break; // aligned again: let the fast path go on
        }
        put(w, bits);
    }
    __device__ __forceinline__ void refill()
    { // call when n <= 32
        if ((pos & 3u) == 0 && pos + 4 <= end) {
            uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(data + pos));
            if (jb_ff_bytes(w) == 0) {
                pos += 4;
                put(__byte_perm(w, 0, 0x0123), 32);
                return;
            }
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

// returns (symbol << 8) | size, or 0xFFFFFFFF for an invalid code
__device__ __forceinline__ uint32_t jb_huff_lookup(const JbHuffTable *t, uint32_t code16)
{
    uint32_t e = t->lut[code16 >> (16 - JB_LUT_BITS)];
    if ((e & 0xFF) != 0) return e;
    // LookupSlow, JpegHuffmanDecodingTable.cs:88-113
    int size = 9;
    while (code16 > t->maxcode[size]) size++;
    if (size > 16) return 0xFFFFFFFFu;
    uint32_t sym = t->values[(t->valoffset[size] + (code16 >> (16 - size))) & 0xFF];
    return (sym << 8) | (uint32_t)size;
}

// ---------------------------------------------------------------------------------------------
// K1a: one thread per restart segment; a warp's 32 lanes decode 32 consecutive segments of the
// same image block-synchronously (all lanes are on the same block-in-MCU, so table selection is
// warp-uniform).  Each lane assembles its current 8x8 block in a skewed shared-memory staging tile;
// after every block the warp flushes the 32 blocks with coalesced 128-bit stores (8 lanes per
// block), so every coefficient block leaves the SM as one full 128-byte line.
// ---------------------------------------------------------------------------------------------
#define JB_K1_WARPS 4
#define JB_K1_THREADS (JB_K1_WARPS * 32)

__device__ __forceinline__ uint32_t jb_stage_chunk(int lane, int chunk)
{ // 16-byte slot of (block of `lane`, 16-byte chunk 0..7): skewed so that flushes are conflict-free
    return (uint32_t)(lane * 8 + ((chunk + lane) & 7));
}

__global__ void __launch_bounds__(JB_K1_THREADS)
jb_k1_huff_segments(const JbDevImage *__restrict__ images,
                    const JbHuffTable *__restrict__ tables, const uint8_t *__restrict__ arena,
                    const uint32_t *__restrict__ marks, const JbScanResult *__restrict__ scanres,
                    int16_t *__restrict__ coef, uint32_t *__restrict__ status)
{
    extern __shared__ uint4 jb_smem[];
    __shared__ JbDevImage s_im;
    JbHuffTable *s_tab = reinterpret_cast<JbHuffTable *>(jb_smem);
    // grid = (CTAs per image, images): a CTA decodes JB_K1_THREADS consecutive segments of one image
    struct { uint32_t image, first_seg; } wk = {blockIdx.y, blockIdx.x * JB_K1_THREADS};
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (wk.first_seg >= images[wk.image].nseg) return;

    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(images + wk.image);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_im);
        for (int i = tid; i < (int)(sizeof(JbDevImage) / 4); i += JB_K1_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const int ntab = s_im.ntables;
    for (int t = 0; t < ntab; t++) {
        const uint4 *src = reinterpret_cast<const uint4 *>(tables + s_im.table_index[t]);
        uint4 *dst = reinterpret_cast<uint4 *>(s_tab + t);
        for (int i = tid; i < (int)(sizeof(JbHuffTable) / 16); i += JB_K1_THREADS) dst[i] = __ldg(src + i);
    }
    uint4 *s_stage = jb_smem + (JB_MAX_TABLE_SLOTS * sizeof(JbHuffTable)) / 16 + wid * 256;
    for (int i = lane; i < 256; i += 32) s_stage[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    const uint32_t nseg = s_im.nseg;
    const uint32_t seg = wk.first_seg + tid;
    const uint32_t dri = s_im.dri ? s_im.dri : s_im.total_mcus;
    const int bpm = s_im.bpm;
    const JbScanResult sr = scanres[wk.image];
    const uint8_t *data = arena + s_im.data_off;
    const uint32_t *mk = marks + s_im.mark_base;

    // segment bounds from the marker index
    uint32_t my_nmcu = 0, start = 0, stop = 0;
    uint32_t err = 0;
    bool last_needs_marker = false;
    if (seg < nseg) {
        my_nmcu = min(dri, s_im.total_mcus - seg * dri);
        if (seg > 0) {
            if (seg - 1 < sr.nmarkers && (mk[seg - 1] & 8u) == 0) start = (mk[seg - 1] >> 4) + 2;
            else { err |= JB_ST_EXPECT_RST; my_nmcu = 0; }
        }
        stop = seg < sr.nmarkers ? (mk[seg] >> 4) : sr.end_pos;
        // the reference expects RSTn or EOI right after every *complete* interval
        // (JpegHuffmanBaselineScanDecoder.cs:139-154)
        last_needs_marker = s_im.dri != 0 && my_nmcu == dri;
    }
    const uint32_t warp_nmcu = __reduce_max_sync(0xFFFFFFFFu, my_nmcu);

    JbBitReader br;
    br.init(data, start, stop);
    int pred0 = 0, pred1 = 0, pred2 = 0, pred3 = 0;
    int16_t *stage16 = reinterpret_cast<int16_t *>(s_stage);

    // global address of this warp's first lane's first block, and the per-lane stride
    const uint64_t warp_blk0 = s_im.coef_off + (uint64_t)(wk.first_seg + wid * 32) * dri * bpm;
    const uint64_t lane_stride = (uint64_t)dri * bpm; // blocks between consecutive segments

    for (uint32_t mcu = 0; mcu < warp_nmcu; mcu++) {
        const bool active = mcu < my_nmcu;
        for (int b = 0; b < bpm; b++) {
            if (active) {
                const int comp = s_im.blk_comp[b];
                const JbHuffTable *dct = s_tab + s_im.blk_dc[b];
                const JbHuffTable *act = s_tab + s_im.blk_ac[b];
                // ---- DC (ReadBlockBaseline :187-196)
                br.ensure32();
                uint32_t e = jb_huff_lookup(dct, br.peek16());
                if (e == 0xFFFFFFFFu) { err |= JB_ST_BAD_CODE; e = 1; }
                br.skip(e & 0xFF);
                int t = (int)(e >> 8);
                if (t > 16) { err |= JB_ST_BAD_CODE; t = 0; }
                int diff = 0;
                if (t != 0) diff = jb_extend((int)br.take(t), t);
                int pred = comp == 0 ? pred0 : comp == 1 ? pred1 : comp == 2 ? pred2 : pred3;
                pred += diff;
                if (comp == 0) pred0 = pred; else if (comp == 1) pred1 = pred; else if (comp == 2) pred2 = pred; else pred3 = pred;
                stage16[jb_stage_chunk(lane, 0) * 8] = (int16_t)pred;
                // ---- AC (:199-221)
                for (int i = 1; i < 64;) {
                    br.ensure32();
                    uint32_t e2 = jb_huff_lookup(act, br.peek16());
                    if (e2 == 0xFFFFFFFFu) { err |= JB_ST_BAD_CODE; break; }
                    br.skip(e2 & 0xFF);
                    const int s = (e2 >> 8) & 15, r = (int)(e2 >> 12);
                    if (s != 0) {
                        i += r;
                        const int v = jb_extend((int)br.take(s), s);
                        const int k = min(i, 63);
                        stage16[jb_stage_chunk(lane, k >> 3) * 8 + (k & 7)] = (int16_t)v;
                        i++;
                    } else {
                        if (r == 0) break;
                        i += 16;
                    }
                }
            }
            __syncwarp();
            // ---- cooperative flush: 8 lanes per block, 4 blocks per instruction
            const uint32_t amask = __ballot_sync(0xFFFFFFFFu, active);
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int bl = it * 4 + (lane >> 3); // whose block
                const int ch = lane & 7;
                if ((amask >> bl) & 1u) {
                    const uint32_t slot = jb_stage_chunk(bl, ch);
                    uint4 v = s_stage[slot];
                    s_stage[slot] = make_uint4(0, 0, 0, 0);
                    uint64_t blk = warp_blk0 + (uint64_t)bl * lane_stride + (uint64_t)mcu * bpm + b;
                    reinterpret_cast<uint4 *>(coef + blk * 64)[ch] = v;
                }
            }
            __syncwarp();
        }
    }

    if (seg < nseg && my_nmcu > 0) {
        // bits consumed beyond the real data => "The bit stream ended prematurely."
        if (br.n < br.pad) err |= JB_ST_PREMATURE_END;
        if (last_needs_marker && !(err & JB_ST_PREMATURE_END)) {
            // AdvanceAlignByte + TryReadMarker: after dropping the partial byte no whole byte may
            // remain before the marker (fill bytes FF are skipped by FillBuffer)
            int real = br.n - br.pad;
            uint32_t p = br.pos;
            while (p < stop && data[p] == 0xFF) p++;
            bool marker_ok = seg < sr.nmarkers; // an RSTn or terminator follows this segment
            if (marker_ok && (mk[seg] & 8u) != 0) marker_ok = sr.end_marker == 0xD9u; // EOI ends the scan
            if (real >= 8 || p < stop || !marker_ok) err |= JB_ST_EXPECT_RST;
        }
    }
    if (err) atomicOr(status + wk.image, err);
}
