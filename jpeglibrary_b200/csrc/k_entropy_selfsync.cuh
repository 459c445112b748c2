// k_entropy_selfsync.cuh -- K1b: Huffman decode of scans WITHOUT restart markers by speculative,
// self-synchronising sub-sequence decoding (north_star item (1), second half).
//
// The sequential decoder of the reference (ReadBlockBaseline + MCU loop,
// ScanDecoder/JpegHuffmanBaselineScanDecoder.cs:99-222) is a state machine over the bit stream with
// state (bit position p, block-in-MCU b, zig-zag index k) -- SURVEY Appendix C.  Huffman streams
// re-synchronise: a decoder started at an arbitrary bit with a wrong state falls into step with the
// true decoder after a few hundred bits (measured on the bench images: mean 730 bits, 99.4 % within
// 4096).  Pipeline, all kernels batched over images:
//
//   U  jb_k1b_unstuff      raw scan bytes -> "clean" stream (FF00 -> FF, terminator cut, 1-padded)
//   S0 jb_k1b_sync<0>      every thread decodes its 4096-bit sub-sequence from the guess
//                          (p = start, b = 0, k = 0) and records its exit state
//   Sr jb_k1b_sync<1>      every thread whose predecessor's exit state changed re-decodes from that
//                          state (rounds until nothing changes; round 1 re-decodes everything)
//   P  jb_k1b_scan         per image exclusive prefix sums over sub-sequences: blocks started,
//                          DC-difference sums per component  ->  first block index + DC predictors
//   W  jb_k1b_write        final decode from the converged entry states, emitting DC-predicted
//                          zig-zag blocks exactly like K1a (staging tile + 128-byte line flushes)
//
// A block belongs to the sub-sequence in which its DC symbol starts; its owner finishes it even if it
// runs past the sub-sequence end, and the next owner first skips the tail of that block.
#pragma once
#include "jb_device.cuh"
#include "k_entropy_decode.cuh"

#define JB_SUBSEQ_BITS 4096u
#define JB_K1B_THREADS 128

struct __align__(8) JbSubState { // decoder state at a symbol boundary (always moved as one 64-bit word)
    uint32_t p;     // bit position in the clean stream
    uint32_t bk;    // (b << 8) | k ; k = 0: a DC symbol comes next
};

struct JbSubInfo {  // what one sub-sequence contributes (valid once the entry states converged)
    uint32_t nblk;  // blocks whose DC symbol starts inside it
    int32_t dc[4];  // sum of the DC differences of those blocks, per component
};

// ---------------------------------------------------------------------------------------------
// U: unstuff.  One CTA per image walks the scan in 4 KB tiles: a byte is dropped iff it is the 00 of
// an FF 00 pair; the stream ends at the first FF xx with xx not in {00, FF} (JpegBitReader.cs:108-128;
// FF FF fill bytes only occur in front of that marker and are cut with it).  The clean stream is
// padded with 0xFF bytes (PeekBits pads with 1-bits, JpegBitReader.cs:166).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
jb_k1b_unstuff(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
               const uint8_t *__restrict__ arena, uint8_t *__restrict__ clean, uint32_t *__restrict__ clean_len)
{
    const uint32_t image = image_list[blockIdx.x];
    const JbDevImage &im = images[image];
    const uint8_t *data = arena + im.data_off;
    uint8_t *out = clean + im.data_off;
    const uint32_t len = im.data_len;
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base, s_end;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_base = 0; s_end = 0xFFFFFFFFu; }
    __syncthreads();
    for (uint32_t tile = 0; tile < len; tile += 256 * 16) {
        const uint32_t pos0 = tile + tid * 16;
        uint32_t w[5] = {0, 0, 0, 0, 0};
        uint32_t prev = 0;
        if (pos0 < len) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(data + pos0));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            w[4] = __ldg(reinterpret_cast<const uint32_t *>(data + pos0 + 16));
            if (pos0 > 0) prev = data[pos0 - 1];
        }
        // terminator inside my 16 bytes?
        uint32_t term = 0xFFFFFFFFu;
        if (jb_ff_bytes(w[0]) | jb_ff_bytes(w[1]) | jb_ff_bytes(w[2]) | jb_ff_bytes(w[3])) {
#pragma unroll
            for (int i = 15; i >= 0; i--) {
                const uint32_t b = (w[i >> 2] >> ((i & 3) * 8)) & 0xFF;
                const uint32_t nb = (w[(i + 1) >> 2] >> (((i + 1) & 3) * 8)) & 0xFF;
                if (b == 0xFF && nb != 0 && pos0 + i + 1 < len) term = pos0 + i; // FF FF counts too (fill before a marker)
            }
        }
        if (term != 0xFFFFFFFFu) atomicMin(&s_end, term);
        __syncthreads();
        const uint32_t end = min(s_end, len);
        // keep mask
        uint32_t keep = 0, cnt = 0;
        uint32_t pb = prev;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t b = (w[i >> 2] >> ((i & 3) * 8)) & 0xFF;
            const bool k = pos0 + i < end && !(b == 0 && pb == 0xFF);
            if (k) { keep |= 1u << i; cnt++; }
            pb = b;
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t off = s_base + incl - cnt, total = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t t = s_warp[i];
            if (i < wid) off += t;
            total += t;
        }
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (keep & (1u << i)) out[off++] = (uint8_t)(w[i >> 2] >> ((i & 3) * 8));
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
        if (s_end != 0xFFFFFFFFu) break;
    }
    __syncthreads();
    const uint32_t n = s_base;
    if (tid < 64) out[n + tid] = 0xFF; // padding (the arena keeps 64 spare bytes per image)
    if (tid == 0) clean_len[image] = n;
}

// ---------------------------------------------------------------------------------------------
// Bit reader over the clean stream: aligned big-endian words, branch-free 32-bit refill.
// ---------------------------------------------------------------------------------------------
struct JbCleanReader {
    const uint8_t *data; // 256-byte aligned, 1-padded
    uint32_t wpos;       // byte offset of the next aligned word to load
    uint32_t hi, lo;
    int n;

    __device__ __forceinline__ uint32_t ldw(uint32_t off) const
    {
        return __byte_perm(__ldg(reinterpret_cast<const uint32_t *>(data + off)), 0, 0x0123);
    }
    __device__ __forceinline__ void seek(const uint8_t *d, uint32_t bitpos)
    {
        data = d;
        const uint32_t byte = bitpos >> 3;
        wpos = byte & ~3u;
        hi = ldw(wpos);
        lo = ldw(wpos + 4);
        wpos += 8;
        n = 64;
        skip_any((int)(bitpos - (byte & ~3u) * 8)); // 0..31
    }
    __device__ __forceinline__ uint32_t position() const { return wpos * 8 - (uint32_t)n; }
    __device__ __forceinline__ void refill()
    { // when n <= 32
        const uint32_t w = ldw(wpos);
        wpos += 4;
        hi |= __funnelshift_rc(w, 0u, n);
        lo |= __funnelshift_rc(0u, w, n);
        n += 32;
    }
    __device__ __forceinline__ void ensure32() { if (n <= 32) refill(); }
    __device__ __forceinline__ uint32_t peek16() const { return hi >> 16; }
    __device__ __forceinline__ void skip_any(int k)
    { // 0..31
        hi = __funnelshift_l(lo, hi, k);
        lo <<= k;
        n -= k;
    }
    __device__ __forceinline__ uint32_t take(int k)
    { // 1..16
        const uint32_t v = hi >> (32 - k);
        skip_any(k);
        return v;
    }
};

struct JbSubGeom {
    uint32_t image, sub, nsub, start_bit, end_bit, total_bits;
    bool active;
};

// ---------------------------------------------------------------------------------------------
// S: synchronisation rounds.  ROUND0: decode from the guess.  Otherwise: decode from the
// predecessor's exit state if it differs from the entry used last time.
// `exits`/`used` are indexed by sub_base + sub.  `changed` counts re-decodes in this round.
// ---------------------------------------------------------------------------------------------
template <bool ROUND0>
__global__ void __launch_bounds__(JB_K1B_THREADS)
jb_k1b_sync(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
            const JbHuffTable *__restrict__ tables, const uint8_t *__restrict__ clean,
            const uint32_t *__restrict__ clean_len, JbSubState *exits, JbSubState *__restrict__ used,
            JbSubInfo *__restrict__ info, uint32_t *__restrict__ changed)
{
    __shared__ JbDevImage s_im;
    __shared__ uint2 s_binfo[JB_MAX_BLOCKS_PER_MCU];
    const uint32_t image = image_list[blockIdx.y];
    const int tid = threadIdx.x;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(images + image);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_im);
        for (int i = tid; i < (int)(sizeof(JbDevImage) / 4); i += JB_K1B_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t sub = blockIdx.x * JB_K1B_THREADS + tid;
    const uint32_t total_bits = clean_len[image] * 8;
    if (blockIdx.x * JB_K1B_THREADS * JB_SUBSEQ_BITS >= total_bits && blockIdx.x > 0) return;
    if (tid < JB_MAX_BLOCKS_PER_MCU)
        s_binfo[tid] = make_uint2(((uint32_t)s_im.blk_comp[tid] << 28) |
                                      (uint32_t)(s_im.table_index[s_im.blk_dc[tid]] * (sizeof(JbHuffTable) / 16)),
                                  (uint32_t)(s_im.table_index[s_im.blk_ac[tid]] * (sizeof(JbHuffTable) / 16)));
    __syncthreads();
    const uint32_t start_bit = sub * JB_SUBSEQ_BITS;
    if (start_bit >= total_bits) return;
    const uint32_t end_bit = start_bit + JB_SUBSEQ_BITS; // the last one simply runs into the padding
    const uint32_t gi = s_im.sub_base + sub;

    JbSubState entry;
    if (sub == 0) { entry.p = 0; entry.bk = 0; }
    else if (ROUND0) { entry.p = start_bit; entry.bk = 0; }
    else {
        // one 64-bit load: a neighbour may be rewriting its exit state in this very round
        const unsigned long long raw = *reinterpret_cast<const volatile unsigned long long *>(&exits[gi - 1]);
        entry.p = (uint32_t)raw; entry.bk = (uint32_t)(raw >> 32);
    }
    if (!ROUND0) {
        const JbSubState u = used[gi];
        if (u.p == entry.p && u.bk == entry.bk) return; // nothing new: my exit state stands
        atomicAdd(changed, 1u);
    }
    *reinterpret_cast<uint2 *>(&used[gi]) = make_uint2(entry.p, entry.bk);

    const int bpm = s_im.bpm;
    const uint8_t *tab_base = reinterpret_cast<const uint8_t *>(tables);
    JbCleanReader br;
    br.seek(clean + s_im.data_off, entry.p);
    int b = (int)(entry.bk >> 8), k = (int)(entry.bk & 0xFF);
    uint2 binfo = s_binfo[b];
    uint32_t nblk = 0;
    int dc0 = 0, dc1 = 0, dc2 = 0, dc3 = 0;
    uint32_t p = entry.p;
    // a sub-sequence holds at most 4096 symbols (>= 1 bit each); the guard also bounds corrupt data
    for (int guard = 0; p < end_bit && guard < 2 * (int)JB_SUBSEQ_BITS; guard++) {
        br.ensure32();
        const bool is_dc = k == 0;
        const uint32_t toff = is_dc ? (binfo.x & 0x0FFFFFFFu) : binfo.y;
        uint32_t e = jb_huff_lookup(reinterpret_cast<const JbHuffTable *>(tab_base + (size_t)toff * 16), br.peek16());
        if (e == 0xFFFFFFFFu) e = 0x0001u; // invalid code while speculating: keep moving
        br.skip_any(e & 0xFF);
        const int sym = (int)(e >> 8);
        int s = is_dc ? sym : (sym & 15);
        const int r = is_dc ? 0 : (sym >> 4);
        if (s > 16) s = 0;
        if (is_dc) {
            int v = 0;
            if (s != 0) v = jb_extend((int)br.take(s), s);
            const int comp = binfo.x >> 28;
            if (comp == 0) dc0 += v; else if (comp == 1) dc1 += v; else if (comp == 2) dc2 += v; else dc3 += v;
            nblk++;
            k = 1;
        } else if (s != 0) {
            br.skip_any(s);
            k += r + 1;
        } else {
            k = r == 0 ? 64 : k + 16;
        }
        if (k >= 64) {
            k = 0;
            if (++b == bpm) b = 0;
            binfo = s_binfo[b];
        }
        p = br.position();
    }
    *reinterpret_cast<volatile unsigned long long *>(&exits[gi]) =
        (unsigned long long)p | ((unsigned long long)(((uint32_t)b << 8) | (uint32_t)k) << 32);
    JbSubInfo inf;
    inf.nblk = nblk;
    inf.dc[0] = dc0; inf.dc[1] = dc1; inf.dc[2] = dc2; inf.dc[3] = dc3;
    info[gi] = inf;
}

// ---------------------------------------------------------------------------------------------
// P: per-image exclusive prefix sums of JbSubInfo over the sub-sequences (in place).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
jb_k1b_scan(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
            const uint32_t *__restrict__ clean_len, JbSubInfo *__restrict__ info, uint32_t *__restrict__ status)
{
    const uint32_t image = image_list[blockIdx.x];
    const JbDevImage &im = images[image];
    const uint32_t total_bits = clean_len[image] * 8;
    const uint32_t nsub = (total_bits + JB_SUBSEQ_BITS - 1) / JB_SUBSEQ_BITS;
    JbSubInfo *a = info + im.sub_base;
    __shared__ int s_w[8][5];
    __shared__ int s_carry[5];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < 5) s_carry[tid] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nsub; base += 256) {
        const uint32_t i = base + tid;
        int v[5] = {0, 0, 0, 0, 0};
        if (i < nsub) {
            const JbSubInfo x = a[i];
            v[0] = (int)x.nblk; v[1] = x.dc[0]; v[2] = x.dc[1]; v[3] = x.dc[2]; v[4] = x.dc[3];
        }
        int inc[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            int x = v[q];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xFFFFFFFFu, x, d);
                if (lane >= d) x += t;
            }
            inc[q] = x;
            if (lane == 31) s_w[wid][q] = x;
        }
        __syncthreads();
        int tot[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            int off = s_carry[q];
            int t = 0;
            for (int w = 0; w < 8; w++) {
                if (w < wid) off += s_w[w][q];
                t += s_w[w][q];
            }
            tot[q] = t;
            inc[q] += off - v[q]; // exclusive
        }
        if (i < nsub) {
            JbSubInfo x;
            x.nblk = (uint32_t)inc[0]; x.dc[0] = inc[1]; x.dc[1] = inc[2]; x.dc[2] = inc[3]; x.dc[3] = inc[4];
            a[i] = x;
        }
        __syncthreads();
        if (tid < 5) s_carry[tid] += tot[tid];
        __syncthreads();
    }
    // fewer blocks in the stream than the frame needs => "The bit stream ended prematurely."
    if (tid == 0 && (uint32_t)s_carry[0] < im.total_mcus * im.bpm) atomicOr(status + image, JB_ST_PREMATURE_END);
}

// ---------------------------------------------------------------------------------------------
// W: final decode + coefficient output.  Same symbol semantics and staging/flush scheme as K1a.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JB_K1B_THREADS)
jb_k1b_write(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
             const JbHuffTable *__restrict__ tables, const uint8_t *__restrict__ clean,
             const uint32_t *__restrict__ clean_len, const JbSubState *__restrict__ exits,
             const JbSubInfo *__restrict__ info, int16_t *__restrict__ coef, uint32_t *__restrict__ status)
{
    extern __shared__ uint4 jb_smem[];
    __shared__ JbDevImage s_im;
    __shared__ uint2 s_binfo[JB_MAX_BLOCKS_PER_MCU];
    const uint32_t image = image_list[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(images + image);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_im);
        for (int i = tid; i < (int)(sizeof(JbDevImage) / 4); i += JB_K1B_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t total_bits = clean_len[image] * 8;
    if (blockIdx.x * JB_K1B_THREADS * JB_SUBSEQ_BITS >= total_bits && blockIdx.x > 0) return;
    if (tid < JB_MAX_BLOCKS_PER_MCU)
        s_binfo[tid] = make_uint2(((uint32_t)s_im.blk_comp[tid] << 28) |
                                      (uint32_t)(s_im.table_index[s_im.blk_dc[tid]] * (sizeof(JbHuffTable) / 16)),
                                  (uint32_t)(s_im.table_index[s_im.blk_ac[tid]] * (sizeof(JbHuffTable) / 16)));
    uint8_t *s_stage = reinterpret_cast<uint8_t *>(jb_smem) + wid * JB_K1_STAGE_BYTES;
    for (int i = lane; i < JB_K1_STAGE_BYTES / 16; i += 32) reinterpret_cast<uint4 *>(s_stage)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    const uint32_t sub = blockIdx.x * JB_K1B_THREADS + tid;
    const uint32_t start_bit = sub * JB_SUBSEQ_BITS;
    const uint32_t end_bit = start_bit + JB_SUBSEQ_BITS;
    const uint32_t gi = s_im.sub_base + sub;
    const uint32_t total_blocks = s_im.total_mcus * s_im.bpm;
    const int bpm = s_im.bpm;
    const uint8_t *tab_base = reinterpret_cast<const uint8_t *>(tables);

    bool active = start_bit < total_bits;
    JbCleanReader br;
    int b = 0, k = 0;
    uint32_t blk = 0; // index of the block being decoded (scan order)
    int pred[4] = {0, 0, 0, 0};
    bool skipping = false; // tail of a block owned by the previous sub-sequence
    uint32_t err = 0;
    if (active) {
        JbSubState entry;
        if (sub == 0) { entry.p = 0; entry.bk = 0; }
        else entry = exits[gi - 1];
        const JbSubInfo base = info[gi];
        blk = base.nblk;
        pred[0] = base.dc[0]; pred[1] = base.dc[1]; pred[2] = base.dc[2]; pred[3] = base.dc[3];
        b = (int)(entry.bk >> 8);
        k = (int)(entry.bk & 0xFF);
        skipping = k != 0;
        br.seek(clean + s_im.data_off, entry.p);
        if (entry.p >= end_bit && !skipping) active = false; // predecessor already covered my range
        if (blk >= total_blocks && !skipping) active = false;
    } else {
        br.data = clean; br.wpos = 0; br.hi = br.lo = 0; br.n = 64;
    }
    uint2 binfo = s_binfo[b];
    int pred_cur = 0;
    {
        const int comp = binfo.x >> 28;
        pred_cur = comp == 0 ? pred[0] : comp == 1 ? pred[1] : comp == 2 ? pred[2] : pred[3];
    }
    uint8_t *gptr = reinterpret_cast<uint8_t *>(coef) + (s_im.coef_off + (uint64_t)blk) * 128;
    const uint32_t lane8 = (lane & 15) * 8;
    int guard = 0;

    while (__any_sync(0xFFFFFFFFu, active)) {
        bool finished = false; // completed a block that this lane owns
        if (active) {
            br.ensure32();
            const bool is_dc = k == 0;
            const uint32_t toff = is_dc ? (binfo.x & 0x0FFFFFFFu) : binfo.y;
            uint32_t e = jb_huff_lookup(reinterpret_cast<const JbHuffTable *>(tab_base + (size_t)toff * 16), br.peek16());
            if (e == 0xFFFFFFFFu) { err |= JB_ST_BAD_CODE; e = 0x0001u; if (!is_dc) k = 64; }
            br.skip_any(e & 0xFF);
            const int sym = (int)(e >> 8);
            int s = is_dc ? sym : (sym & 15);
            const int r = is_dc ? 0 : (sym >> 4);
            if (s > 16) { err |= JB_ST_BAD_CODE; s = 0; }
            int v = 0;
            if (s != 0) v = jb_extend((int)br.take(s), s);
            if (is_dc) {
                v += pred_cur;
                pred_cur = v;
                *reinterpret_cast<int16_t *>(s_stage + jb_stage_off(lane, 0)) = (int16_t)v;
                k = 1;
            } else if (s != 0) {
                k += r;
                if (!skipping) *reinterpret_cast<int16_t *>(s_stage + jb_stage_off(lane, min(k, 63))) = (int16_t)v;
                k++;
            } else {
                k = r == 0 ? 64 : k + 16;
            }
            if (++guard > 4 * (int)JB_SUBSEQ_BITS) { err |= JB_ST_BAD_CODE; active = false; }
            finished = k >= 64 && !skipping;
        }
        uint32_t fin = __ballot_sync(0xFFFFFFFFu, finished);
        while (fin) {
            const uint32_t fin2 = fin & (fin - 1);
            const uint32_t pick = (lane & 16) ? fin2 : fin;
            const int L = __ffs(pick) - 1;
            const uint32_t glo = __shfl_sync(0xFFFFFFFFu, (uint32_t)reinterpret_cast<uint64_t>(gptr), L & 31);
            const uint32_t ghi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(reinterpret_cast<uint64_t>(gptr) >> 32), L & 31);
            if (L >= 0) {
                uint2 *sp = reinterpret_cast<uint2 *>(s_stage + L * 128 + ((((lane & 15) + L) & 15) << 3));
                const uint2 val = *sp;
                *sp = make_uint2(0, 0);
                *reinterpret_cast<uint2 *>((((uint64_t)ghi << 32) | glo) + lane8) = val;
            }
            fin = fin2 & (fin2 - 1);
        }
        if (active && k >= 64) {
            // block boundary
            k = 0;
            if (++b == bpm) b = 0;
            const uint2 ni = s_binfo[b];
            if (!skipping) {
                blk++;
                gptr += 128;
            }
            if ((ni.x ^ binfo.x) >> 28) {
                const int comp = binfo.x >> 28, nc = ni.x >> 28;
                if (!skipping) {
                    if (comp == 0) pred[0] = pred_cur; else if (comp == 1) pred[1] = pred_cur; else if (comp == 2) pred[2] = pred_cur; else pred[3] = pred_cur;
                }
                pred_cur = nc == 0 ? pred[0] : nc == 1 ? pred[1] : nc == 2 ? pred[2] : pred[3];
            }
            binfo = ni;
            skipping = false;
            const uint32_t p = br.position();
            if (blk == total_blocks && p > total_bits) err |= JB_ST_PREMATURE_END; // ran into the padding
            if (p >= end_bit || blk >= total_blocks) active = false;
        }
    }
    if (err) atomicOr(status + image, err);
}
