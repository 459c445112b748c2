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
//   U  jb_k1b_count/_copy  raw scan bytes -> "clean" stream (FF00 -> FF, fill bytes dropped, 1-padded), in 64 KB
//                          chunks: kept bytes per chunk, then every chunk compacts to its prefix offset
//   S0 jb_k1b_sync<0>      every thread decodes its sub-sequence (2^12..2^15 bits, chosen per batch) from the guess
//                          (p = start, b = 0, k = 0) and records its exit state and 16 checkpoints on the way
//   Sr jb_k1b_sync<1>      every thread whose predecessor's exit state changed re-decodes from that state and
//                          stops at the first checkpoint it reproduces (the trajectories have merged there)
//   P  jb_k1b_scan         per image exclusive prefix sums over sub-sequences: blocks started,
//                          DC-difference sums per component  ->  first block index + DC predictors
//   D  jb_k1b_descs        one segment descriptor (JbSegDesc: entry bit, block-in-MCU, zig-zag index, DC predictors,
//                          first block) per sub-sequence
//   W  jb_k1_huff_flat<1>  the segment decoder of the restart path (k_entropy_flat.cuh) over those descriptors
//
// A block belongs to the sub-sequence in which its DC symbol starts; its owner finishes it even if it
// runs past the sub-sequence end, and the next owner first skips the tail of that block.
#pragma once
#include "jb_device.cuh"
#include "k_entropy_decode.cuh"
#include "k_entropy_flat.cuh"

#define JB_SUBSEQ_MIN_SHIFT 12  // sub-sequences are 2^shift bits long; the host picks the shift per batch (12..15):
#define JB_SUBSEQ_MAX_SHIFT 15  // long ones make the write pass as efficient as the restart-segment decoder
#define JB_SUBSEQ_CHECKS 16     // checkpoints per sub-sequence (re-decodes stop at the first one they reproduce)
#define JB_K1B_THREADS 256
#define JB_K1B_CHUNK 65536u   // bytes of stuffed stream per un-stuff CTA
#define JB_K1B_TABLES 4       // Huffman tables cached in shared memory per CTA (one image per CTA)

struct __align__(8) JbSubState { // decoder state at a symbol boundary (always moved as one 64-bit word)
    uint32_t p;     // bit position in the clean stream
    uint32_t bk;    // (b << 8) | k ; k = 0: a DC symbol comes next
};

struct JbSubInfo {  // what one sub-sequence contributes (valid once the entry states converged)
    uint32_t nblk;  // blocks whose DC symbol starts inside it
    int32_t dc[4];  // sum of the DC differences of those blocks, per component
};

// Decoder state at the first symbol boundary at or behind a checkpoint position, with what the sub-sequence
// contributes FROM there to its end (a function of that state alone, whatever happened in front of it).  A later
// round that arrives at the same (p, bk) is on the same trajectory from there on.
struct __align__(16) JbSubCheck {
    uint32_t p, bk, nblk;
    int32_t dc[4];
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------
// U: unstuff, chunk-parallel.  A byte is dropped iff it is the 00 of an FF 00 pair or an FF followed by
// another FF (fill byte); the stream ends at the marker K0 found (JpegBitReader.cs:108-128).  Pass 1
// counts the kept bytes of every 64 KB chunk, pass 2 compacts each chunk to the sum of the counts in
// front of it.  The clean stream is padded with 0xFF bytes (PeekBits pads with 1-bits, :166).
// ---------------------------------------------------------------------------------------------
// The stream ends at its FIRST marker of any kind: the bit reader stops feeding at FF xx whatever xx is
// (JpegBitReader.cs:108-128), and a scan without restart interval never reads an RSTn -- one that turns up in a damaged
// stream ends the data like EOI would.  K0 lists RSTn markers in front of the terminator, so the first entry is it.
__device__ __forceinline__ uint32_t jb_k1b_stream_end(const JbDevImage &im, const JbScanResult &sr, const uint32_t *marks)
{
    uint32_t end = sr.end_pos;
    if (sr.nmarkers) end = min(end, marks[im.mark_base] >> 4);
    return min(end, im.data_len);
}

// byte-permute selectors that pack the kept bytes of a little-endian word at its low end, in memory order:
// index = 4-bit mask of kept bytes (bit i = memory byte i), unused result bytes select the zero operand
__constant__ uint16_t jb_c_keepsel_le[16] = {0x4444, 0x4440, 0x4441, 0x4410, 0x4442, 0x4420, 0x4421, 0x4210,
                                             0x4443, 0x4430, 0x4431, 0x4310, 0x4432, 0x4320, 0x4321, 0x3210};

// One lane's 64 bytes of the stuffed stream (pos0 is 64-byte aligned): loads them and returns, per 32-bit word,
// the 4-bit mask of bytes that stay (nib[i] in bits 4i..4i+3 of the two result words).  Bytes at or behind `end`
// do not stay.  All four bytes of a word are classified at once with SWAR flags.
__device__ __forceinline__ void jb_k1b_keep64(const uint8_t *data, uint32_t pos0, uint32_t end, uint32_t (&w)[16], uint64_t &nibs)
{
    nibs = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
    if (pos0 >= end) return;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(data + pos0) + j); // the arena is padded: safe
        w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
    const uint32_t nextb = __ldg(data + pos0 + 64);
    const uint32_t prevb = pos0 ? __ldg(data + pos0 - 1) : 0u;
    uint32_t F[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t inv = ~w[i];
        F[i] = ~(((inv & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | inv) & 0x80808080u; // bytes that are 0xFF
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t Z = ~(((w[i] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w[i]) & 0x80808080u; // bytes that are 0x00
        const uint32_t nf = i < 15 ? (F[i + 1] << 24) & 0x80000000u : (nextb == 0xFFu ? 0x80000000u : 0u);
        const uint32_t pf = i > 0 ? (F[i - 1] >> 24) & 0x80u : (prevb == 0xFFu ? 0x80u : 0u);
        const uint32_t NF = (F[i] >> 8) | nf, PF = (F[i] << 8) | pf;
        // FF FF: the first FF is a fill byte; FF 00: the zero is stuffing
        uint32_t K = ~((F[i] & NF) | (Z & PF)) & 0x80808080u;
        const uint32_t p = pos0 + 4 * i;
        if (p + 4 > end) K &= p >= end ? 0u : (0xFFFFFFFFu >> (8 * (p + 4 - end)));
        nibs |= (uint64_t)(((K >> 7) * 0x01020408u) >> 24) << (4 * i);
    }
}

#define JB_K1B_UTHREADS 256
#define JB_K1B_UTILE (JB_K1B_UTHREADS * 64) // bytes of stuffed stream per CTA iteration

__global__ void __launch_bounds__(JB_K1B_UTHREADS)
jb_k1b_count(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
             const uint8_t *__restrict__ arena, const JbScanResult *__restrict__ scanres,
             const uint32_t *__restrict__ marks, uint32_t *__restrict__ chunk_kept)
{
    const uint32_t image = image_list[blockIdx.y];
    const JbDevImage &im = images[image];
    const uint32_t end = jb_k1b_stream_end(im, scanres[image], marks);
    const uint32_t c0 = blockIdx.x * JB_K1B_CHUNK;
    if (c0 >= end && blockIdx.x > 0) return;
    const uint8_t *data = arena + im.data_off;
    uint32_t cnt = 0, w[16];
    uint64_t nibs;
    for (uint32_t pos0 = c0 + threadIdx.x * 64; pos0 < c0 + JB_K1B_CHUNK; pos0 += JB_K1B_UTILE) {
        jb_k1b_keep64(data, pos0, end, w, nibs);
        cnt += __popcll(nibs);
    }
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) chunk_kept[im.chunk_base + blockIdx.x] = s_cnt;
}

// Pass 2: every lane compacts its 64 bytes through a 64-bit byte accumulator that emits aligned 32-bit words -- into a
// shared-memory image of the tile's output, laid out so that shared offset i is global byte (tile start & ~15) + i; the
// CTA then writes the image out as 128-bit stores (only the partial 16-byte lines at its two ends go byte by byte).
// Round 2: the lanes used to store their words straight to global memory, 32 lanes 64 bytes apart -- 32 sectors per
// store instruction, 8x the sector writes the data needs, and the kernel ran at 34 % issue rate behind the LSU.
#ifndef JB_K1B_STAGED_COPY
#define JB_K1B_STAGED_COPY 1
#endif
__global__ void __launch_bounds__(JB_K1B_UTHREADS)
jb_k1b_copy(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
            const uint8_t *__restrict__ arena, const JbScanResult *__restrict__ scanres,
            const uint32_t *__restrict__ marks, const uint32_t *__restrict__ chunk_kept, uint8_t *__restrict__ clean,
            uint32_t *__restrict__ clean_len)
{
    const uint32_t image = image_list[blockIdx.y];
    const JbDevImage &im = images[image];
    const uint32_t end = jb_k1b_stream_end(im, scanres[image], marks);
    const uint32_t c0 = blockIdx.x * JB_K1B_CHUNK;
    if (c0 >= end && blockIdx.x > 0) return;
    const uint8_t *data = arena + im.data_off;
    uint8_t *gout = clean + im.data_off; // (256-byte aligned)
    __shared__ uint32_t s_warp[JB_K1B_UTHREADS / 32];
    __shared__ uint32_t s_base;
#if JB_K1B_STAGED_COPY
    __shared__ __align__(16) uint8_t s_img[JB_K1B_UTILE + 32];
#endif
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        uint32_t base = 0;
        for (uint32_t i = 0; i < blockIdx.x; i++) base += chunk_kept[im.chunk_base + i];
        s_base = base;
    }
    __syncthreads();
    for (uint32_t tile = c0; tile < c0 + JB_K1B_CHUNK && tile < end; tile += JB_K1B_UTILE) {
        uint32_t w[16];
        uint64_t nibs;
        jb_k1b_keep64(data, tile + tid * 64, end, w, nibs);
        const uint32_t cnt = __popcll(nibs);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        const uint32_t tile_out = s_base; // first byte of the clean stream this tile writes
        uint32_t off = tile_out + incl - cnt, total = 0;
#pragma unroll
        for (int i = 0; i < JB_K1B_UTHREADS / 32; i++) {
            const uint32_t t = s_warp[i];
            if (i < wid) off += t;
            total += t;
        }
        // bytes [off, off + cnt) of the clean stream are this lane's
#if JB_K1B_STAGED_COPY
        uint8_t *const out = s_img;
        const uint32_t adj = tile_out & ~15u; // global byte offset - adj = its place in the image
#else
        uint8_t *const out = gout;
        const uint32_t adj = 0;
#endif
        uint64_t acc = 0;  // pending output bytes, memory order from bit 0
        uint32_t na = 0;   // how many
        uint32_t o = off;  // where the next pending byte goes
        const uint32_t stop = off + cnt;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t nib = (uint32_t)(nibs >> (4 * i)) & 0xFu;
            const uint32_t piece = __byte_perm(w[i], 0, jb_c_keepsel_le[nib]);
            acc |= (uint64_t)piece << (8 * na);
            na += __popc(nib);
            // bytes in front of the first aligned word leave one by one (at most three per lane)
            while (na != 0 && (o & 3u) != 0) {
                out[o++ - adj] = (uint8_t)acc;
                acc >>= 8;
                na--;
            }
            if (na >= 4) {
                *reinterpret_cast<uint32_t *>(out + (o - adj)) = (uint32_t)acc;
                acc >>= 32;
                na -= 4;
                o += 4;
            }
        }
        while (o < stop) { // at most three bytes are left
            out[o++ - adj] = (uint8_t)acc;
            acc >>= 8;
        }
        __syncthreads();
#if JB_K1B_STAGED_COPY
        {
            const uint32_t g0 = tile_out & ~15u;                   // global offset of image byte 0
            const uint32_t first = tile_out - g0, last = first + total; // the image's valid bytes
            const uint32_t body0 = (first + 15u) & ~15u, body1 = last & ~15u; // whole 16-byte lines
            for (uint32_t i = body0 + tid * 16; i < body1; i += JB_K1B_UTHREADS * 16)
                *reinterpret_cast<uint4 *>(gout + g0 + i) = *reinterpret_cast<const uint4 *>(s_img + i);
            // the partial lines at the two ends (a neighbouring tile or chunk owns their other bytes)
            if (body0 > body1) { // everything lies inside one line
                if (first + tid < last) gout[g0 + first + tid] = s_img[first + tid];
            } else {
                if (first + tid < body0) gout[g0 + first + tid] = s_img[first + tid];
                if (body1 + tid < last) gout[g0 + body1 + tid] = s_img[body1 + tid];
            }
        }
#endif
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    // the chunk that holds the end of the stream pads it and publishes the clean length
    if (c0 + JB_K1B_CHUNK >= end) {
        const uint32_t n = s_base;
        if (tid < 64) gout[n + tid] = 0xFF; // padding (the arena keeps 64 spare bytes per image)
        if (tid == 0) clean_len[image] = n;
    }
}

// ---------------------------------------------------------------------------------------------
// Shared set-up of the sync and write kernels: a CTA works on consecutive sub-sequences of ONE image, so the
// image's Huffman look-up tables (32-bit entries, k_entropy_flat.cuh) sit in shared memory together with the
// per-block table references.  s_bi[b] = {DC table, AC table, component}; table references are shared-memory
// word offsets or JB_K1B_GLOBAL | word offset into the global table array.
// ---------------------------------------------------------------------------------------------
#ifndef JB_K1B_SYMBOLS_PER_ROUND
#define JB_K1B_SYMBOLS_PER_ROUND 3
#endif
#define JB_K1B_GLOBAL 0x80000000u
#define JB_K1B_TABLE_WORDS (JB_LUT_SIZE + JB_LUT2_SUBTABLES * 64)

__device__ __forceinline__ void jb_k1b_setup(const JbDevImage &im, const JbHuffTable32 *__restrict__ tables,
                                             uint32_t *s_tab, uint4 *s_bi, int tid)
{
    __shared__ uint32_t s_ids[JB_K1B_TABLES];
    __shared__ uint32_t s_n;
    if (tid == 0) {
        uint32_t n = 0;
        for (int k = 0; k < im.ntables && n < JB_K1B_TABLES; k++) s_ids[n++] = im.table_index[k];
        s_n = n;
    }
    __syncthreads();
    const uint32_t ntab = s_n;
    for (uint32_t t = 0; t < ntab; t++) {
        const uint4 *src = reinterpret_cast<const uint4 *>(tables + s_ids[t]);
        uint4 *dst = reinterpret_cast<uint4 *>(s_tab + t * JB_K1B_TABLE_WORDS);
        for (int i = tid; i < JB_K1B_TABLE_WORDS / 4; i += JB_K1B_THREADS) dst[i] = __ldg(src + i);
    }
    if (tid < JB_MAX_BLOCKS_PER_MCU) {
        uint32_t ref[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint32_t id = im.table_index[c ? im.blk_ac[tid] : im.blk_dc[tid]];
            uint32_t j = 0;
            while (j < ntab && s_ids[j] != id) j++;
            ref[c] = j < ntab ? j * JB_K1B_TABLE_WORDS : (JB_K1B_GLOBAL | (uint32_t)(id * (sizeof(JbHuffTable32) / 4)));
        }
        s_bi[tid] = make_uint4(ref[0], ref[1], im.blk_comp[tid], 0);
    }
    __syncthreads();
}

// one Huffman symbol from the window's top bits: table entry (k_entropy_flat.cuh format) or JB_E32_BAD
__device__ __forceinline__ uint32_t jb_k1b_lookup(const uint32_t *s_tab, const uint32_t *tab_words, uint32_t toff, uint32_t hi)
{
    uint32_t e;
    if (toff & JB_K1B_GLOBAL) e = __ldg(tab_words + (toff & ~JB_K1B_GLOBAL) + (hi >> (32 - JB_LUT_BITS)));
    else e = s_tab[toff + (hi >> (32 - JB_LUT_BITS))];
    if ((e & 0xFFu) == 0) {
        uint32_t e2 = 0;
        if (e == JB_E32_BADLUT) return JB_E32_BAD;
        if (e != 0 && !(toff & JB_K1B_GLOBAL)) e2 = s_tab[toff + JB_LUT_SIZE + ((e >> 8) - 1) * 64 + ((hi >> 16) & 63)];
        if (e2 == 0) {
            // second-level miss or a table that is not cached: resolve against the table in global memory
            const uint32_t goff = (toff & JB_K1B_GLOBAL) ? (toff & ~JB_K1B_GLOBAL) : 0xFFFFFFFFu;
            if (goff != 0xFFFFFFFFu) e2 = jb_huff32_escape(reinterpret_cast<const JbHuffTable32 *>(tab_words + goff), e, hi >> 16);
            else e2 = 0; // cached table: the caller resolves it (it knows the table's global index)
        }
        e = e2;
    }
    return e;
}

// Bit reader over the clean stream: aligned big-endian words, one predicated 32-bit refill per symbol.
struct JbCleanReader {
    const uint32_t *words; // clean stream of the image (word pointer)
    uint32_t wpos;         // index of the next word to load
    uint32_t hi, lo;
    int n;
    __device__ __forceinline__ void seek(const uint8_t *d, uint32_t bitpos)
    {
        words = reinterpret_cast<const uint32_t *>(d);
        wpos = bitpos >> 5;
        hi = __byte_perm(__ldg(words + wpos), 0, 0x0123);
        lo = __byte_perm(__ldg(words + wpos + 1), 0, 0x0123);
        wpos += 2;
        n = 64;
        const int sk = (int)(bitpos & 31u);
        hi = __funnelshift_l(lo, hi, sk);
        lo <<= sk;
        n -= sk;
    }
    __device__ __forceinline__ uint32_t position() const { return wpos * 32 - (uint32_t)n; }
    __device__ __forceinline__ void refill()
    { // keeps n >= 32
        if (n < 32) {
            const uint32_t w = __byte_perm(__ldg(words + wpos), 0, 0x0123);
            wpos++;
            hi |= w >> n;
            lo |= __funnelshift_r(0u, w, n);
            n += 32;
        }
    }
    __device__ __forceinline__ void skip(uint32_t total)
    { // 0..32
        hi = __funnelshift_lc(lo, hi, total);
        lo = __funnelshift_lc(0u, lo, total);
        n -= (int)total;
    }
};

// ---------------------------------------------------------------------------------------------
// S: synchronisation rounds.  ROUND0: decode from the guess.  Otherwise: decode from the
// predecessor's exit state if it differs from the entry used last time.
// `exits`/`used` are indexed by sub_base + sub.  `changed` counts re-decodes in this round.
// ---------------------------------------------------------------------------------------------
// checkpoints [from, to) were written with prefix sums in this pass: turn them into suffix sums (total - prefix)
__device__ __forceinline__ void jb_k1b_close_checks(JbSubCheck *checks, uint32_t from, uint32_t to, const JbSubInfo &total)
{
    for (uint32_t i = from; i < to; i++) {
        JbSubCheck c = checks[i];
        c.nblk = total.nblk - c.nblk;
#pragma unroll
        for (int q = 0; q < 4; q++) c.dc[q] = total.dc[q] - c.dc[q];
        checks[i] = c;
    }
}

// MODE 0: guess round.  MODE 1: re-decode from the predecessor's exit state if it changed, stopping at the first
// checkpoint that reproduces the previous trajectory.  MODE 2: like 1 but always to the end (rebuilds `info`).
template <int MODE>
__global__ void __launch_bounds__(JB_K1B_THREADS)
jb_k1b_sync(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
            const JbHuffTable32 *__restrict__ tables, const uint8_t *__restrict__ clean,
            const uint32_t *__restrict__ clean_len, JbSubState *exits, JbSubState *__restrict__ used,
            JbSubInfo *__restrict__ info, JbSubCheck *__restrict__ checks, uint32_t *__restrict__ changed, int sub_shift)
{
    constexpr bool ROUND0 = MODE == 0;
    const uint32_t sub_bits = 1u << sub_shift;
    __shared__ __align__(16) uint32_t s_tab[JB_K1B_TABLES * JB_K1B_TABLE_WORDS];
    __shared__ uint4 s_bi[JB_MAX_BLOCKS_PER_MCU];
    __shared__ int s_dc[JB_K1B_THREADS][4];
    const uint32_t image = image_list[blockIdx.y];
    const JbDevImage &im = images[image];
    const int tid = threadIdx.x;
    const uint32_t total_bits = clean_len[image] * 8;
    if ((uint64_t)blockIdx.x * JB_K1B_THREADS * sub_bits >= total_bits && blockIdx.x > 0) return;
    const uint32_t sub = blockIdx.x * JB_K1B_THREADS + tid;
    const uint64_t start64 = (uint64_t)sub << sub_shift;
    const uint32_t start_bit = (uint32_t)start64;
    const uint32_t end_bit = min(start_bit + sub_bits, total_bits); // (reads stay inside the 64 bytes of padding)
    const uint32_t gi = im.sub_base + sub;
    bool work = start64 < total_bits;
    JbSubState entry;
    entry.p = 0; entry.bk = 0;
    if (work) {
        if (sub == 0) { entry.p = 0; entry.bk = 0; }
        else if (ROUND0) { entry.p = start_bit; entry.bk = 0; }
        else {
            // one 64-bit load: a neighbour may be rewriting its exit state in this very round
            const unsigned long long raw = *reinterpret_cast<const volatile unsigned long long *>(&exits[gi - 1]);
            entry.p = (uint32_t)raw; entry.bk = (uint32_t)(raw >> 32);
        }
        if (!ROUND0) {
            const JbSubState u = used[gi];
            if (u.p == entry.p && u.bk == entry.bk) work = false; // nothing new: my exit state stands
        }
    }
    if (!ROUND0 && !__syncthreads_or(work)) return; // the usual case after round 1: nobody in this CTA re-decodes
    jb_k1b_setup(im, tables, s_tab, s_bi, tid);
    if (!work) return;
    if (!ROUND0) atomicAdd(changed, 1u);
    *reinterpret_cast<uint2 *>(&used[gi]) = make_uint2(entry.p, entry.bk);

    const uint32_t bpm = im.bpm;
    const uint32_t *tab_words = reinterpret_cast<const uint32_t *>(tables);
    // bit window over the clean stream with two words of prefetch (words are big-endian in memory)
    const uint32_t *words = reinterpret_cast<const uint32_t *>(clean + im.data_off);
    uint32_t wpos = entry.p >> 5;
    uint32_t hi, lo;
    int n;
    {
        const uint32_t w0 = __byte_perm(__ldg(words + wpos), 0, 0x0123), w1 = __byte_perm(__ldg(words + wpos + 1), 0, 0x0123);
        const int sk = (int)(entry.p & 31u);
        hi = __funnelshift_l(w1, w0, sk);
        lo = w1 << sk;
        n = 64 - sk;
        wpos += 2;
    }
    uint32_t wnext = __ldg(words + wpos), wnext2 = __ldg(words + wpos + 1);
    uint32_t b = entry.bk >> 8, k = entry.bk & 0xFF;
    uint4 bi = s_bi[b];
    uint32_t nblk = 0;
    int *dcs = s_dc[tid];
    dcs[0] = dcs[1] = dcs[2] = dcs[3] = 0;
    int dcur = 0; // DC-difference sum of the current component
    uint32_t p = entry.p;
    const uint32_t cp_bits = sub_bits / JB_SUBSEQ_CHECKS;
    uint32_t next_cp = start_bit + cp_bits, cpj = 1;
    JbSubCheck *my_checks = checks + (size_t)gi * JB_SUBSEQ_CHECKS;
    // every symbol consumes at least one bit, so the loop ends after at most sub_bits symbols
    while (p < end_bit) {
        if (p >= next_cp) {
            while (p >= next_cp && cpj < JB_SUBSEQ_CHECKS) {
                JbSubCheck now;
                now.p = p; now.bk = (b << 8) | k; now.nblk = nblk; now.pad = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) now.dc[c] = dcs[c] + ((uint32_t)c == bi.z ? dcur : 0);
                if (MODE == 1) {
                    const JbSubCheck old = my_checks[cpj];
                    if (old.p == now.p && old.bk == now.bk) {
                        // same state at the same bit as last time: the rest of the sub-sequence decodes identically
                        JbSubInfo inf;
                        inf.nblk = now.nblk + old.nblk;
#pragma unroll
                        for (int c = 0; c < 4; c++) inf.dc[c] = now.dc[c] + old.dc[c];
                        info[gi] = inf;
                        jb_k1b_close_checks(my_checks, 1, cpj, inf);
                        return; // the exit state stands
                    }
                }
                my_checks[cpj] = now; // prefix sums for now; turned into suffix sums when the totals are known
                cpj++;
                next_cp += cp_bits;
            }
            if (cpj >= JB_SUBSEQ_CHECKS) next_cp = 0xFFFFFFFFu;
        }
        if (n < 32) {
            const uint32_t be = __byte_perm(wnext, 0, 0x0123);
            hi |= be >> n;
            lo |= __funnelshift_r(0u, be, n);
            n += 32;
            wpos++;
            wnext = wnext2;
            wnext2 = __ldg(words + wpos + 1);
        }
        // One symbol (the same step as K1's).  Up to three per round, like K1 (k_entropy_flat.cuh): the checkpoint test
        // and the refill are paid per round by the whole warp, so symbols per round is what the instruction count hangs
        // on.  A further symbol is taken while the window still holds 32 bits and neither the end of the sub-sequence
        // nor the next checkpoint has been reached (a checkpoint records the state at the first symbol boundary behind it).
        auto symbol = [&]() {
            const bool is_dc = k == 0;
            const uint32_t toff = is_dc ? bi.x : bi.y;
            uint32_t e = jb_k1b_lookup(s_tab, tab_words, toff, hi);
            if (e == 0) {
                const uint32_t id = im.table_index[is_dc ? im.blk_dc[b] : im.blk_ac[b]];
                const uint32_t e1 = s_tab[toff + (hi >> (32 - JB_LUT_BITS))];
                e = jb_huff32_escape(tables + id, e1, hi >> 16);
            }
            if (e == JB_E32_BAD) e = is_dc ? 0x01000101u : 0x40000101u; // invalid code while speculating: keep moving
            const uint32_t total = e & 0xFFu, len = (e >> 8) & 0xFFu, adv = e >> 24;
            const uint32_t s = total - len;
            const uint32_t x = __funnelshift_l(lo, hi, len);
            const uint32_t neg = ~(uint32_t)((int32_t)x >> 31);
            const uint32_t t = __funnelshift_l(x ^ neg, 0u, s); // the top s bits (0 for s = 0): one SHF, not two and a subtraction
            const int v = (int)((t ^ neg) - neg);
            hi = __funnelshift_lc(lo, hi, total);
            lo = __funnelshift_lc(0u, lo, total);
            n -= (int)total;
            p += total;
            if (is_dc) { dcur += v; nblk++; }
            k += adv;
        };
        symbol();
        // (further symbols stop at the end of a block: the block change below then runs once per round for the couple
        // of lanes that need it, not once per symbol step)
#if JB_K1B_SYMBOLS_PER_ROUND >= 2
        if (k < 64 && n >= 32 && p < end_bit && p < next_cp) symbol();
#endif
#if JB_K1B_SYMBOLS_PER_ROUND >= 3
        if (k < 64 && n >= 32 && p < end_bit && p < next_cp) symbol();
#endif
        if (k >= 64) {
            k = 0;
            b = b + 1 == bpm ? 0 : b + 1;
            const uint4 ni = s_bi[b];
            if (ni.z != bi.z) {
                dcs[bi.z] += dcur;
                dcur = 0;
            }
            bi = ni;
        }
    }
    dcs[bi.z] += dcur;
    *reinterpret_cast<volatile unsigned long long *>(&exits[gi]) =
        (unsigned long long)p | ((unsigned long long)((b << 8) | k) << 32);
    JbSubInfo inf;
    inf.nblk = nblk;
    inf.dc[0] = dcs[0]; inf.dc[1] = dcs[1]; inf.dc[2] = dcs[2]; inf.dc[3] = dcs[3];
    info[gi] = inf;
    jb_k1b_close_checks(my_checks, 1, cpj, inf);
}

// ---------------------------------------------------------------------------------------------
// P: per-image exclusive prefix sums of JbSubInfo over the sub-sequences (in place).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
jb_k1b_scan(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
            const uint32_t *__restrict__ clean_len, JbSubInfo *__restrict__ info, uint32_t *__restrict__ status, int sub_shift)
{
    const uint32_t image = image_list[blockIdx.x];
    const JbDevImage &im = images[image];
    const uint32_t total_bits = clean_len[image] * 8;
    const uint32_t nsub = (uint32_t)(((uint64_t)total_bits + (1u << sub_shift) - 1) >> sub_shift);
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
    if (tid == 0) { // slot nsub (every image reserves one more than it has sub-sequences): the totals
        JbSubInfo x;
        x.nblk = (uint32_t)s_carry[0]; x.dc[0] = s_carry[1]; x.dc[1] = s_carry[2]; x.dc[2] = s_carry[3]; x.dc[3] = s_carry[4];
        a[nsub] = x;
    }
    // (Fewer blocks in the stream than the frame needs is not an error by itself: the reference decodes on behind the
    // data, where PeekBits supplies 1-bits, and only fails on a symbol whose magnitude bits are not there
    // (JpegHuffmanScanDecoder.cs:81-110).  The last sub-sequence of the final pass does the same, see jb_k1b_descs.)
    (void)status;
}

// ---------------------------------------------------------------------------------------------
// W: final decode + coefficient output = the restart-segment decoder (jb_k1_huff_flat<true>, k_entropy_flat.cuh)
// run over the sub-sequences.  This kernel writes its descriptors from the converged entry states and the
// prefix sums: a sub-sequence first skips the tail of the block its predecessor owns, then decodes exactly the
// blocks whose DC symbol starts inside it.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
jb_k1b_descs(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
             const uint32_t *__restrict__ clean_len, const JbSubState *__restrict__ exits,
             const JbSubInfo *__restrict__ info, JbSegDesc *__restrict__ segs, int sub_shift)
{
    const uint32_t image = image_list[blockIdx.y];
    const JbDevImage &im = images[image];
    const uint32_t sub = blockIdx.x * 256 + threadIdx.x;
    if (sub >= im.sub_cap) return;
    const uint32_t total_bits = clean_len[image] * 8;
    // (a stream of fill bytes only leaves no bits at all: its one sub-sequence decodes the frame from the padding)
    const uint32_t nsub = max(1u, (uint32_t)(((uint64_t)total_bits + (1u << sub_shift) - 1) >> sub_shift));
    const uint32_t gi = im.sub_base + sub;
    const uint32_t total_blocks = im.total_mcus * im.bpm;
    JbSegDesc d;
    d.word0 = 0; d.lead = 0; d.nbytes = 0; d.nblocks = 0; d.coef_block = 0; d.image = image; d.flags = 0;
    d.pred[0] = d.pred[1] = d.pred[2] = d.pred[3] = 0;
    d.state = 0; d.endw = 0; d.pad[0] = d.pad[1] = 0;
    if (sub < nsub) {
        JbSubState entry;
        if (sub == 0) { entry.p = 0; entry.bk = 0; }
        else entry = exits[gi - 1];
        const JbSubInfo base = info[gi], next = info[gi + 1]; // exclusive prefixes; slot nsub holds the totals
        // the last sub-sequence decodes whatever the frame still needs: behind the data the window holds 1-bits like the
        // reference's PeekBits, and K1's per-symbol verdict decides whether that is an error
        const uint32_t first = min(base.nblk, total_blocks), last = sub + 1 == nsub ? total_blocks : min(next.nblk, total_blocks);
        const uint32_t k = entry.bk & 0xFFu;
        const bool skip = k != 0 && sub != 0;
        const uint64_t bit0 = im.data_off * 8 + entry.p;
        d.word0 = (uint32_t)(bit0 >> 5);
        d.lead = (uint32_t)(bit0 & 31u);
        d.nbytes = entry.p < total_bits ? total_bits - entry.p : 0;
        d.nblocks = (last - first) + (skip ? 1u : 0u);
        d.coef_block = im.coef_off + first;
        d.pred[0] = base.dc[0]; d.pred[1] = base.dc[1]; d.pred[2] = base.dc[2]; d.pred[3] = base.dc[3];
        d.state = (entry.bk >> 8) | (k << 8) | (skip ? 1u << 16 : 0u);
        d.endw = (uint32_t)(im.data_off >> 2) + ((clean_len[image] + 3) >> 2) + 8; // inside the 64 bytes of 0xFF padding
    }
    segs[gi] = d;
}
