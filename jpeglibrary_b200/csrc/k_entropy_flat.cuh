// k_entropy_flat.cuh -- K0b (per-segment un-stuffing) and K1 (flat restart-segment Huffman decode).
//
// Replaces, for a whole batch of images at once:
//   JpegBitReader.FillBuffer / PeekBits / TryReadBits / AdvanceAlignByte   (JpegBitReader.cs:29-204)
//   JpegHuffmanDecodingTable.Lookup / LookupSlow                            (JpegHuffmanDecodingTable.cs:73-113)
//   DecodeHuffmanCode / ReceiveAndExtend                                    (ScanDecoder/JpegHuffmanScanDecoder.cs:81-115)
//   ReadBlockBaseline + MCU loop + restart handling                         (ScanDecoder/JpegHuffmanBaselineScanDecoder.cs:99-222)
//
// Design (B200): the decoder is instruction-issue bound, so everything data dependent that is not the
// Huffman symbol itself is moved out of the per-symbol loop:
//   * K0b rewrites every restart segment as a "clean" stream: FF 00 -> FF, fill bytes dropped
//     (JpegBitReader.cs:108-128), bytes stored big-endian per 32-bit word and followed by 16 bytes of
//     1-bits (PeekBits pads with ones, :166).  The decoder's refill is then one predicated word append.
//   * Huffman tables carry 32-bit entries that already hold everything the symbol step needs
//     (bits to consume, code length, zig-zag run, zig-zag advance), built on the host by simulating the
//     reference's Lookup/LookupSlow, so DC and AC symbols share one branch-free step.
//   * Segments of the whole batch are numbered globally (image-major): warps are always full, whatever
//     the number of segments per image.
#pragma once
#include "jb_device.cuh"

// one restart segment of the batch, written by K0b
struct __align__(16) JbSegDesc {
    uint32_t word0;      // arena word (4-byte) index of the word that holds the segment's first byte
    uint32_t lead;       // bytes of that word in front of the segment (0..3)
    uint32_t nbytes;     // length of the segment in the stuffed stream
    uint32_t nblocks;    // blocks to decode (0: nothing to do)
    uint64_t coef_block; // first block of the segment in the coefficient store
    uint32_t image;
    uint32_t flags;      // bit 0: a restart marker / EOI must follow; bit 1: it does
    // Sub-sequences of the self-synchronising path (K1b) are decoded by the same kernel from the un-stuffed
    // stream: then word0/lead address the entry BIT (lead = 0..31 bits), nbytes is the number of real bits
    // from there to the end of the image's stream, and the decoder starts in the middle of the MCU sequence:
    int32_t pred[4];     // DC predictors at the entry
    uint32_t state;      // block-in-mcu | zig-zag index << 8 | (1 << 16: the first block belongs to the
                         // previous sub-sequence and is only skipped; it is counted in nblocks)
    uint32_t endw;       // last padded word of the image's clean stream
    uint32_t pad[2];
};

#define JB_K0B_THREADS 128

// K0b: restart-segment descriptors from K0's marker index (one thread per segment; a few microseconds).
__global__ void __launch_bounds__(JB_K0B_THREADS)
jb_k0b_segment_descs(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
                     const uint32_t *__restrict__ marks, const JbScanResult *__restrict__ scanres,
                     JbSegDesc *__restrict__ segs, uint32_t *__restrict__ status, uint32_t *__restrict__ mcu_limit, uint32_t *__restrict__ first_error)
{
    const uint32_t image = image_list[blockIdx.y];
    const JbDevImage &im = images[image];
    const uint32_t seg = blockIdx.x * JB_K0B_THREADS + threadIdx.x;
    if (seg >= im.nseg) return;
    const JbScanResult sr = scanres[image];
    const uint32_t *mk = marks + im.mark_base;
    const uint32_t dri = im.dri ? im.dri : im.total_mcus;
    const uint32_t my_nmcu = min(dri, im.total_mcus - seg * dri);
    uint32_t start = 0;
    bool reachable = true;
    if (seg > 0) {
        if (seg - 1 < sr.nmarkers && (mk[seg - 1] & 8u) == 0) start = (mk[seg - 1] >> 4) + 2;
        else reachable = false; // the previous interval is not followed by RSTn: "Expect restart marker."
    }
    uint32_t stop = seg < sr.nmarkers ? (mk[seg] >> 4) : sr.end_pos;
    if (!reachable || stop < start) stop = start;
    // the reference expects RSTn or EOI right after every *complete* interval (JpegHuffmanBaselineScanDecoder.cs:139-154)
    uint32_t flags = (im.dri != 0 && my_nmcu == dri) ? 1u : 0u;
    bool marker_ok = seg < sr.nmarkers;
    if (marker_ok && (mk[seg] & 8u) != 0) marker_ok = sr.end_marker == 0xD9u; // EOI ends the scan
    if (marker_ok) flags |= 2u;
    JbSegDesc d;
    const uint64_t a0 = im.data_off + start;
    d.word0 = (uint32_t)(a0 >> 2);
    d.lead = (uint32_t)(a0 & 3u);
    d.nbytes = stop - start;
    d.nblocks = reachable ? my_nmcu * im.bpm : 0;
    d.coef_block = im.coef_off + (uint64_t)seg * dri * im.bpm;
    d.image = image;
    d.flags = flags;
    d.pred[0] = d.pred[1] = d.pred[2] = d.pred[3] = 0;
    d.state = 0;
    d.endw = 0;
    d.pad[0] = d.pad[1] = 0;
    segs[im.seg_base + seg] = d;
    // An interval without RSTn in front of it is not decoded.  If the scan ended with EOI, that is how the reference
    // ends too, quietly (JpegHuffmanBaselineScanDecoder.cs:144-150: the EOI sits where a restart marker would); any other
    // marker there -- or none -- is "Expect restart marker.".  (If the EOI came in the middle of an interval, that
    // interval's own decode reports the premature end.)
    if (!reachable && sr.end_marker != 0xD9u) jb_report_error(status, first_error, image, JB_ST_EXPECT_RST, 0, seg - 1); // (behind interval seg - 1)
    // ... and the reference returns from the MCU loop there: WriteBlock is never called for the MCUs of the absent
    // intervals, so the renderer must leave their pixels alone.  Interval s is present iff markers 0..s-1 are RSTn.
    if (seg == 0 && im.dri != 0) {
        uint32_t rst = sr.nmarkers;
        if (rst > 0 && (mk[rst - 1] & 8u) != 0) rst--;
        if (rst + 1 < im.nseg) mcu_limit[image] = (rst + 1) * im.dri;
    }
}

// K0c: intervals that are not in the stream (see above) leave their blocks as the allocator made them: zero.  The
// coefficient store of sequential frames is not cleared per launch (K1 writes every block of every interval it
// decodes), so the absent ones are cleared here; one warp per interval, which returns at once in the normal case.
__global__ void __launch_bounds__(JB_K0B_THREADS)
jb_k0c_clear_absent(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
                    const uint32_t *__restrict__ marks, const JbScanResult *__restrict__ scanres, int16_t *__restrict__ coef)
{
    const uint32_t image = image_list[blockIdx.y];
    const JbDevImage &im = images[image];
    const uint32_t seg = blockIdx.x * (JB_K0B_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (seg == 0 || seg >= im.nseg) return;
    const JbScanResult sr = scanres[image];
    if (seg - 1 < sr.nmarkers && (marks[im.mark_base + seg - 1] & 8u) == 0) return; // present
    const uint32_t dri = im.dri ? im.dri : im.total_mcus;
    const uint64_t first = im.coef_off + (uint64_t)seg * dri * im.bpm;
    const uint64_t nblk = (uint64_t)min(dri, im.total_mcus - seg * dri) * im.bpm;
    uint4 *p = reinterpret_cast<uint4 *>(coef + first * 64);
    for (uint64_t i = lane; i < nblk * 8; i += 32) p[i] = make_uint4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------------------------
// Huffman table, 32-bit entries: byte 0 = bits to consume (code + magnitude), byte 1 = code length,
// byte 2 = zig-zag run before the coefficient, byte 3 = zig-zag advance after the symbol.
// byte 0 == 0: escape; byte 1 = 1 + second-level sub-table (indexed by the next 6 bits) or 0 = slow path.
// ---------------------------------------------------------------------------------------------
struct __align__(16) JbHuffTable32 {
    uint32_t lut[JB_LUT_SIZE];
    uint32_t lut2[JB_LUT2_SUBTABLES * 64];
    uint16_t maxcode[20];  // reference _maxCode[0..17] (left-aligned 16-bit), padded
    uint8_t valoffset[24]; // reference _valOffset[0..18], padded
    uint8_t values[256];
    uint32_t cls;          // 0 = DC, 1 = AC
    uint32_t pad[3];
};
static_assert(sizeof(JbHuffTable32) % 16 == 0, "table size");

#define JB_E32_BAD 0xFFFFFFFFu
// first-level look-up entry of a code (<= 10 bits) whose symbol is invalid (DC magnitude category > 16): it must not go
// to the slow path, which like the reference's LookupSlow starts at 9 bits and would take a short code for a long one
#define JB_E32_BADLUT 0x0000FF00u

// byte-permute selectors that left-align the kept bytes of a little-endian word in stream (big-endian) order:
// index = 4-bit mask of kept bytes (bit i = memory byte i), unused result bytes select the zero operand
__constant__ uint16_t jb_c_keepsel[16] = {0x4444, 0x0444, 0x1444, 0x0144, 0x2444, 0x0244, 0x1244, 0x0124,
                                          0x3444, 0x0344, 0x1344, 0x0134, 0x2344, 0x0234, 0x1234, 0x0123};

// ReadBlockBaseline's use of a decoded symbol (JpegHuffmanBaselineScanDecoder.cs:187-219)
__host__ __device__ inline uint32_t jb_entry32(uint32_t cls, uint32_t sym, uint32_t len)
{
    uint32_t s, run, adv;
    if (cls == 0) {
        if (sym > 16) return JB_E32_BAD;
        s = sym; run = 0; adv = 1;
    } else {
        s = sym & 15; run = sym >> 4;
        adv = s ? run + 1 : (run == 0 ? 64 : 16); // EOB; any other s == 0 symbol skips 16 (:213-219)
    }
    return (len + s) | (len << 8) | (run << 16) | (adv << 24);
}

// LookupSlow, JpegHuffmanDecodingTable.cs:88-113
__device__ __noinline__ uint32_t jb_huff32_slow(const JbHuffTable32 *t, uint32_t code16)
{
    int size = 9;
    while (code16 > __ldg(&t->maxcode[size])) size++;
    if (size > 16) return JB_E32_BAD;
    const uint32_t sym = __ldg(&t->values[(__ldg(&t->valoffset[size]) + (code16 >> (16 - size))) & 0xFF]);
    return jb_entry32(__ldg(&t->cls), sym, (uint32_t)size);
}

__device__ __noinline__ uint32_t jb_huff32_escape(const JbHuffTable32 *t, uint32_t e, uint32_t code16)
{
    if (e == JB_E32_BADLUT) return JB_E32_BAD;
    if (e != 0) {
        e = __ldg(&t->lut2[((e >> 8) - 1) * 64 + (code16 & 63)]);
        if (e != 0) return e;
    }
    return jb_huff32_slow(t, code16);
}

// ---------------------------------------------------------------------------------------------
// K1: one thread per restart segment of the batch.  Each lane assembles its current block in a private
// 128-byte shared-memory slot (zig-zag order); the per-symbol step is branch-free.  Completed blocks are
// moved to the coefficient store by the whole warp -- 8 lanes x 16 bytes per block, up to four blocks per
// round -- so that the load/store pipe sees full-width instructions (a per-lane copy issues 24 LSU
// instructions for the two lanes that finish in an average iteration).
// The Huffman look-up tables of the CTA's images live in shared memory (a look-up that misses L1 would
// stall all 32 lanes for an L2 round trip on nearly every symbol); a CTA whose images use more than
// JB_K1F_TABLES distinct tables reads the others through the read-only path.
// The CTA size is chosen at launch so that all segments of a batch are resident in one wave when possible.
// ---------------------------------------------------------------------------------------------
#ifndef JB_K1_OWNER_COPY
#define JB_K1_OWNER_COPY 0
#endif
#ifndef JB_K1_RANKED_COPY
#define JB_K1_RANKED_COPY 1
#endif
#ifndef JB_K1_SYMBOLS_PER_ROUND
#define JB_K1_SYMBOLS_PER_ROUND 3
#endif
#ifndef JB_K1F_MAX_THREADS
#define JB_K1F_MAX_THREADS (JB_K1_RANKED_COPY ? 992 : 1024) // (what fits 227 KB of shared memory next to four tables)
#endif
#define JB_K1F_TABLES 4
#define JB_K1F_TABLE_WORDS (JB_LUT_SIZE + JB_LUT2_SUBTABLES * 64)
// per-lane slot: 128 B coefficients | 16 B DC predictors | 10 x {DC table, AC table | component << 14} (uint16 pairs)
#ifndef JB_K1F_SLOT
#define JB_K1F_SLOT 192
#endif
#define JB_K1F_NOTAB 0x3FFFu // table reference: not cached in shared memory
__host__ __device__ inline size_t jb_k1f_smem_bytes(int threads)
{
    return (size_t)JB_K1F_TABLES * JB_K1F_TABLE_WORDS * 4 + (size_t)threads * (JB_K1F_SLOT + 4 + (JB_K1_RANKED_COPY ? 8 : 0));
}

// CLEAN = false: restart segments, read from the stuffed stream.  CLEAN = true: sub-sequences of K1b, read from
// the un-stuffed stream with an entry state (see JbSegDesc).
template <bool CLEAN>
__global__ void __launch_bounds__(JB_K1F_MAX_THREADS, 1)
jb_k1_huff_flat(const JbDevImage *__restrict__ images, const JbSegDesc *__restrict__ segs, uint32_t nsegs,
                const JbHuffTable32 *__restrict__ tables, const uint32_t *__restrict__ arena_words,
                int16_t *__restrict__ coef, uint32_t *__restrict__ status, uint32_t lanes_per_warp,
                uint32_t *__restrict__ first_error)
{
    extern __shared__ __align__(16) uint8_t jb_k1f_smem[];
    __shared__ uint32_t s_tab_id[JB_K1F_TABLES];
    __shared__ uint32_t s_ntab;
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31;
    uint32_t *s_tab = reinterpret_cast<uint32_t *>(jb_k1f_smem);
    uint8_t *s_slots = jb_k1f_smem + JB_K1F_TABLES * JB_K1F_TABLE_WORDS * 4;
    uint32_t *s_img = reinterpret_cast<uint32_t *>(s_slots + (size_t)nthreads * JB_K1F_SLOT);
    uint8_t *st = s_slots + tid * JB_K1F_SLOT;
    uint8_t *warp_slots = s_slots + (tid & ~31) * JB_K1F_SLOT;
    // Small batches spread their segments over more warps than they would fill (lanes_per_warp < 32): a warp steps
    // at the pace of its slowest lane and pays every lane's refills and block hand-offs, so a lane decodes two to three
    // times faster with a warp (nearly) to itself -- and a small batch has warps to spare.
    const uint32_t g = ((blockIdx.x * nthreads + tid) >> 5) * lanes_per_warp + lane;
#pragma unroll
    for (int i = 0; i < 9; i++) reinterpret_cast<uint4 *>(st)[i] = make_uint4(0, 0, 0, 0);
    JbSegDesc d;
    d.nblocks = 0; d.image = 0xFFFFFFFFu;
    if ((uint32_t)lane < lanes_per_warp && g < nsegs) d = segs[g];
    s_img[tid] = d.nblocks ? d.image : 0xFFFFFFFFu;
    __syncthreads();
    if (tid == 0) { // the CTA's table set: segments are numbered image-major, so images come in runs
        uint32_t n = 0, prev = 0xFFFFFFFFu;
        for (int i = 0; i < nthreads; i++) {
            const uint32_t im = s_img[i];
            if (im == prev || im == 0xFFFFFFFFu) continue;
            prev = im;
            const int nt = images[im].ntables;
            for (int k = 0; k < nt; k++) {
                const uint32_t id = images[im].table_index[k];
                uint32_t j = 0;
                while (j < n && s_tab_id[j] != id) j++;
                if (j == n && n < JB_K1F_TABLES) s_tab_id[n++] = id;
            }
        }
        s_ntab = n;
    }
    __syncthreads();
    const uint32_t ntab = s_ntab;
    for (uint32_t t = 0; t < ntab; t++) {
        const uint4 *src = reinterpret_cast<const uint4 *>(tables + s_tab_id[t]); // lut and lut2 are the first members
        uint4 *dst = reinterpret_cast<uint4 *>(s_tab + t * JB_K1F_TABLE_WORDS);
        for (int i = tid; i < JB_K1F_TABLE_WORDS / 4; i += nthreads) dst[i] = __ldg(src + i);
    }
    uint32_t left = d.nblocks;
    const JbDevImage *im = images + (left ? d.image : 0);
    const uint32_t bpm = left ? im->bpm : 0;
    uint32_t *s_bi = reinterpret_cast<uint32_t *>(st + 144);
    for (uint32_t k = 0; k < bpm; k++) {
        const uint4 bi = __ldg(&im->binfo[k]);
        const uint32_t ids[2] = {bi.x / (uint32_t)(sizeof(JbHuffTable32) / 4), bi.y / (uint32_t)(sizeof(JbHuffTable32) / 4)};
        uint32_t ref[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            uint32_t j = 0;
            while (j < ntab && s_tab_id[j] != ids[c]) j++;
            ref[c] = j < ntab ? j * JB_K1F_TABLE_WORDS : JB_K1F_NOTAB;
        }
        s_bi[k] = ref[0] | ((ref[1] | (bi.z << 14)) << 16);
    }
    __syncthreads();
    const uint32_t *tab_words = reinterpret_cast<const uint32_t *>(tables);

    // Bit window: hi:lo hold n + pad valid bits, left-aligned, fed 32 bits at a time from the STUFFED stream
    // (JpegBitReader.FillBuffer, JpegBitReader.cs:95-138).  The common refill -- a whole word inside the
    // segment without any 0xFF byte and no stuffed FF pending -- appends the byte-swapped word; everything else
    // (FF 00 -> FF, FF FF fill bytes, the partial first and last words, the 1-bit padding behind the segment
    // that PeekBits produces, :166) takes the byte-wise path below.
    // n counts the REAL bits of the window, pad the padding bits behind them (always the last bits of the window);
    // need = 32 - pad, so that "n < need" is "fewer than 32 bits in the window".  n goes negative when padding is
    // consumed; whether that is an error depends on what consumed it (see `under` below).
    const uint32_t a0r = d.lead, a1r = d.lead + d.nbytes;     // segment bytes, relative to word0 * 4
    // first word not entirely inside the segment (CLEAN: d.nbytes counts the real BITS behind the entry bit)
    const uint32_t endw = CLEAN ? d.endw : d.word0 + (a1r >> 2);
    const uint32_t realw = CLEAN ? d.word0 + ((d.lead + d.nbytes) >> 5) : 0u; // CLEAN: first word not entirely real
    uint32_t wabs = d.word0;
    uint32_t lim = CLEAN ? min(realw, endw) : (a0r ? 0u : endw); // fast refills while wabs < lim (0: byte-wise path)
    bool carry = false;                                       // the last byte appended was a stuffed 0xFF candidate
    uint32_t hi = 0, lo = 0, wnext = 0, wnext2 = 0; // wnext: word at wabs, wnext2: the one after (both prefetched)
    if (left) {
        wnext = __ldg(arena_words + wabs);
        wnext2 = __ldg(arena_words + wabs + 1);
    }
    int n = 0, pad = 0, need = 32;
    // DecodeHuffmanCode advances min(code size, bits available) and never fails; ReceiveAndExtend throws "The bit
    // stream ended prematurely." when its s bits are not all there (JpegHuffmanScanDecoder.cs:81-110).  Behind the
    // data the window holds 1-bits like PeekBits (JpegBitReader.cs:166), so the symbols come out the same either way
    // and only the verdict is at stake: a symbol is an error iff it has magnitude bits (s > 0) and they end behind the
    // data (n < 0 after it).  `under` collects n & -s: its sign bit is the verdict.
    int under = 0;
    uint32_t err = 0;

    uint32_t b = 0, k = 0; // block-in-mcu; next zig-zag index (0: the DC symbol comes next)
    int pred = 0;
    bool skip = false;     // CLEAN: the block under way belongs to the previous sub-sequence
    if (CLEAN && left) {
        b = d.state & 0xFFu;
        k = (d.state >> 8) & 0xFFu;
        skip = (d.state >> 16) & 1u;
        int *pp = reinterpret_cast<int *>(st + 128);
        pp[0] = d.pred[0]; pp[1] = d.pred[1]; pp[2] = d.pred[2]; pp[3] = d.pred[3];
        // window = the 64 bits at the entry bit
        const uint32_t w0 = __byte_perm(wnext, 0, 0x0123), w1 = __byte_perm(wnext2, 0, 0x0123);
        hi = __funnelshift_l(w1, w0, d.lead);
        lo = w1 << d.lead;
        n = (int)min(64u - d.lead, d.nbytes);
        pad = 64 - (int)d.lead - n;
        need = 32 - pad;
        wabs = min(wabs + 2, endw);
        wnext = __ldg(arena_words + wabs);
        wnext2 = __ldg(arena_words + min(wabs + 1, endw));
        pred = pp[(s_bi[b] >> 30)];
    }
    uint32_t bi = s_bi[b]; // low half: DC table, high half: AC table | component << 14 (shared-memory word offsets)
    uint32_t tdc = bi & 0x3FFFu, tac = (bi >> 16) & 0x3FFFu;
    uint64_t gptr = reinterpret_cast<uint64_t>(coef + d.coef_block * 64);
#if JB_K1_RANKED_COPY
    uint32_t gblk = (uint32_t)d.coef_block;  // the same as a block index (a store of 2^32 blocks would be 512 GB)
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint2 *s_fin = reinterpret_cast<uint2 *>(s_slots + (size_t)nthreads * (JB_K1F_SLOT + 4)); // one (lane, block) entry per lane
#endif

    while (__any_sync(0xFFFFFFFFu, left != 0)) {
        if (left != 0) {
            if (CLEAN) {
                if (n < need) { // un-stuffed stream: one predicated word append
                    const uint32_t be = __byte_perm(wnext, 0, 0x0123);
                    const int fill = n + pad;
                    hi |= be >> fill;
                    lo |= __funnelshift_r(0u, be, fill);
                    if (wabs < lim) n += 32;
                    else { // the word that holds the last real bits of the image's stream, or padding behind them
                        const int real = wabs == realw ? (int)((d.lead + d.nbytes) & 31u) : 0;
                        n += real;
                        pad += 32 - real;
                        need = 32 - pad;
                        lim = 0;
                    }
                    wabs = min(wabs + 1, endw);
                    wnext = wnext2;
                    wnext2 = __ldg(arena_words + min(wabs + 1, endw));
                }
            } else
            while (n < need) {
                const uint32_t w = wnext;
                const uint32_t inv = ~w;
                // 0x80 in every byte of w that is 0xFF (exact)
                const uint32_t F = ~(((inv & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | inv) & 0x80808080u;
                if (F == 0 && wabs < lim) {
                    const uint32_t be = __byte_perm(w, 0, 0x0123);
                    hi |= be >> n;
                    lo |= __funnelshift_r(0u, be, n);
                    n += 32;
                } else if (wabs < endw && (wabs != d.word0 || a0r == 0)) {
                    // a whole word inside the segment with 0xFF bytes or a pending stuffed zero: FF FF -> the first
                    // FF is a fill byte, FF 00 -> the zero is stuffing (JpegBitReader.cs:108-128), done on all
                    // four bytes at once
                    const uint32_t Z = ~(((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u;  // bytes that are 0x00
                    const uint32_t NF = (F >> 8) | ((wnext2 & 0xFFu) == 0xFFu ? 0x80000000u : 0u); // next byte is FF
                    const uint32_t PF = (F << 8) | (carry ? 0x80u : 0u);                       // previous byte is FF
                    const uint32_t K = ~((F & NF) | (Z & PF)) & 0x80808080u;
                    const uint32_t km = ((K >> 7) * 0x01020408u) >> 24;
                    const uint32_t outw = __byte_perm(w, 0, jb_c_keepsel[km]);
                    carry = (F >> 31) != 0;
                    lim = carry ? 0u : endw;
                    hi |= outw >> n;
                    lo |= __funnelshift_r(0u, outw, n);
                    n += 8 * __popc(km);
                } else {
                    // the partial first / last word of the segment and the 1-bit padding behind it
                    const uint32_t rel = (wabs - d.word0) * 4;
                    const uint32_t nb = wnext2 & 0xFFu;
                    uint32_t outw = 0, prv = carry ? 0xFFu : 0u;
                    int bits = 0, ones = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t p = rel + i, bv = (w >> (8 * i)) & 0xFFu;
                        const uint32_t nx = i < 3 ? (w >> (8 * i + 8)) & 0xFFu : nb;
                        if (p >= a1r) { // behind the segment: 1-bits
                            outw |= 0xFFu << (24 - bits - ones);
                            ones += 8;
                        } else if (p >= a0r) {
                            const uint32_t pv = p == a0r ? 0u : prv;
                            if (!((bv == 0xFFu && nx == 0xFFu) || (bv == 0u && pv == 0xFFu))) {
                                outw |= bv << (24 - bits);
                                bits += 8;
                            }
                        }
                        prv = bv;
                    }
                    carry = rel + 3 < a1r && (w >> 24) == 0xFFu;
                    lim = carry ? 0u : endw;
                    const int fill = n + pad;
                    hi |= outw >> fill;
                    lo |= __funnelshift_r(0u, outw, fill);
                    n += bits;
                    pad += ones;
                    need = 32 - pad;
                }
                wabs = min(wabs + 1, endw + 1);
                wnext = wnext2;
                wnext2 = __ldg(arena_words + wabs + 1);
            }
            // One symbol.  FIRST: the symbol behind the refill (DC when k == 0).  The second call of an iteration is an
            // AC symbol by construction (k > 0), so everything that concerns DC folds away there.
            auto symbol = [&](const bool FIRST) {
                const bool is_dc = FIRST && k == 0;
                const uint32_t toff = is_dc ? tdc : tac;
                uint32_t e = 0;
                if (toff != JB_K1F_NOTAB) e = s_tab[toff + (hi >> (32 - JB_LUT_BITS))];
                if ((e & 0xFFu) == 0) {
                    uint32_t e2 = 0;
                    if (e != 0 && e != JB_E32_BADLUT) e2 = s_tab[toff + JB_LUT_SIZE + ((e >> 8) - 1) * 64 + ((hi >> 16) & 63)];
                    if (e2 == 0) { // table not cached, a code longer than 16 bits / not in the second level, or a bad symbol
                        const uint4 gi = __ldg(&im->binfo[b]);
                        const uint32_t goff = is_dc ? gi.x : gi.y;
                        if (toff == JB_K1F_NOTAB) e = __ldg(tab_words + goff + (hi >> (32 - JB_LUT_BITS)));
                        e2 = (e & 0xFFu) ? e : jb_huff32_escape(reinterpret_cast<const JbHuffTable32 *>(tab_words + goff), e, hi >> 16);
                    }
                    e = e2;
                    if (e == JB_E32_BAD) { // invalid code or magnitude category: flag, then finish the block
                        err |= JB_ST_BAD_CODE;
                        e = is_dc ? 0x01000101u : 0x40000101u;
                    }
                }
                // (one PRMT per byte field: the shift-and-mask form costs two instructions for bytes 1 and 2)
                const uint32_t total = __byte_perm(e, 0, 0x4440), len = __byte_perm(e, 0, 0x4441), run = __byte_perm(e, 0, 0x4442),
                               adv = e >> 24;
                const uint32_t s = total - len;
                // ReceiveAndExtend (JpegHuffmanScanDecoder.cs:100-115): s magnitude bits follow the code
                const uint32_t x = __funnelshift_l(lo, hi, len);
                const uint32_t neg = ~(uint32_t)((int32_t)x >> 31); // all ones when the leading magnitude bit is 0
                const uint32_t t = __funnelshift_l(x ^ neg, 0u, s); // the top s bits (0 for s = 0): one SHF, not two and a subtraction
                int v = (int)((t ^ neg) - neg);
                hi = __funnelshift_lc(lo, hi, total);
                lo = __funnelshift_lc(0u, lo, total);
                n -= (int)total;
                under |= n & (int)(len - total); // sign bit: magnitude bits (s > 0) that end behind the data (n < 0)
                const uint32_t pos = min(k + run, 63u);
                if (is_dc) { v += pred; pred = v; }
                if ((s != 0 || is_dc) && !(CLEAN && skip)) *reinterpret_cast<int16_t *>(st + pos * 2) = (int16_t)v;
                k += adv;
            };
            symbol(true);
#if JB_K1_SYMBOLS_PER_ROUND >= 2
            // A second symbol in the same round when the block is not finished and the window still holds 32 bits (the
            // guarantee the refill gives the first one).  A warp pays for the refill and the block hand-off of a round
            // whatever the number of lanes that need them, so symbols per round is what the instruction count hangs on
            // (profiles/r2*_decode: 146 -> ... warp-instructions per 32 symbols).
            if (k < 64 && n >= need) symbol(false);
#endif
#if JB_K1_SYMBOLS_PER_ROUND >= 3
            if (k < 64 && n >= need) symbol(false);
#endif
#if JB_K1_SYMBOLS_PER_ROUND >= 4
            if (k < 64 && n >= need) symbol(false);
#endif
        }
        const bool finished = k >= 64;
#if JB_K1_OWNER_COPY
        // (A/B, kept for the record: every finished lane moves its own block -- eight 128-bit load / clear / store triples,
        // no ballot, no shuffles, no copy rounds: 24 warp-instructions per round against 58 below.  Measured on B200, 1024
        // frames: 19.1 ms against 11.2 ms (16.3 ms with a 208-byte slot stride that removes the bank conflicts of the
        // loads): a store instruction whose ~5 active lanes hit 5 different lines costs the load/store pipe 5 wavefronts, and
        // eight of them per round make the kernel LSU-bound.  The cooperative copy writes a full line per instruction.)
        if (finished && !(CLEAN && skip)) {
            uint4 *sp = reinterpret_cast<uint4 *>(st);
            uint4 *gp = reinterpret_cast<uint4 *>(gptr);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint4 q = sp[i];
                sp[i] = make_uint4(0, 0, 0, 0);
                gp[i] = q;
            }
        }
#elif JB_K1_RANKED_COPY
        // ---- completed blocks leave the SM as full 128-byte lines moved BY THE WARP: the finished lanes list themselves
        // (lane, destination block) in shared memory by rank, then every group of 8 lanes moves the (4r + g)-th listed
        // block in copy round r: ceil(finished / 4) rounds whatever groups the finished lanes sit in, no shuffles and no
        // vote per round.  (First generation: every group moved the blocks of its OWN 8 lanes -- 2.6 copy rounds of 19
        // instructions per decode round where 5 lanes finish; this one: 2 rounds of 11.)
        const uint32_t fin = __ballot_sync(0xFFFFFFFFu, finished && !(CLEAN && skip));
        if (fin) {
            uint2 *list = s_fin + (tid & ~31);
            if (finished && !(CLEAN && skip)) list[__popc(fin & lane_lt)] = make_uint2((uint32_t)lane, gblk);
            __syncwarp(); // the list, and the owners' coefficient stores above, are read by other lanes below
            const int nfin = __popc(fin);
            for (int idx = lane >> 3; idx < ((nfin + 3) & ~3); idx += 4) {
                if (idx < nfin) {
                    const uint2 ent = list[idx];
                    uint4 *sp = reinterpret_cast<uint4 *>(warp_slots + ent.x * JB_K1F_SLOT) + (lane & 7);
                    const uint4 q = *sp;
                    *sp = make_uint4(0, 0, 0, 0);
                    reinterpret_cast<uint4 *>(coef + (size_t)ent.y * 64)[lane & 7] = q;
                }
            }
            __syncwarp(); // the helpers' zeroing of the slots is ordered before the owners' next stores
        }
#else
        // ---- completed blocks leave the SM as full 128-byte lines: lanes 8g..8g+7 move the g-th finished lane's block
        const uint32_t fin = __ballot_sync(0xFFFFFFFFu, finished && !(CLEAN && skip));
        if (fin) {
            __syncwarp(); // the owners' coefficient stores above are read by other lanes below
            // every group of 8 lanes moves the finished blocks of its own 8 lanes, one block per round
            uint32_t mine = (fin >> (lane & 24)) & 0xFFu;
            do {
                const int t = 31 - __clz((int)mine); // highest finished lane of the group; -1: none left
                const int L = (lane & 24) + (t & 7);
                const uint32_t glo = __shfl_sync(0xFFFFFFFFu, (uint32_t)gptr, L);
                const uint32_t ghi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(gptr >> 32), L);
                if (t >= 0) {
                    uint4 *sp = reinterpret_cast<uint4 *>(warp_slots + L * JB_K1F_SLOT) + (lane & 7);
                    const uint4 q = *sp;
                    *sp = make_uint4(0, 0, 0, 0);
                    reinterpret_cast<uint4 *>(((uint64_t)ghi << 32) | glo)[lane & 7] = q;
                }
                mine &= ~(1u << (t & 31));
            } while (__any_sync(0xFFFFFFFFu, mine != 0));
            __syncwarp(); // the helpers' zeroing of the slots is ordered before the owners' next stores
        }
#endif
        if (finished) {
#if JB_K1_RANKED_COPY
            if (!(CLEAN && skip)) gblk++;
#endif
            if (!(CLEAN && skip)) gptr += 128;
            left--;
            k = 0;
            b = b + 1 == bpm ? 0 : b + 1;
            const uint32_t ni = s_bi[b];
            if ((ni ^ bi) >> 30) { // DC predictors are per component (:187-196)
                int *pp = reinterpret_cast<int *>(st + 128);
                if (!(CLEAN && skip)) pp[bi >> 30] = pred;
                pred = pp[ni >> 30];
            }
            skip = false;
            bi = ni;
            tdc = ni & 0x3FFFu;
            tac = (ni >> 16) & 0x3FFFu;
        }
    }
    if (d.nblocks == 0) return;
    if (under < 0) err |= JB_ST_PREMATURE_END; // "The bit stream ended prematurely." (ReceiveAndExtend)
    if (CLEAN) {
        if (err) atomicOr(status + d.image, err); // (no restart intervals: every failure is an InvalidDataException)
        return;
    }
    if (!(err & JB_ST_PREMATURE_END) && (d.flags & 1u)) {
        // AdvanceAlignByte + TryReadMarker: after dropping the partial byte no whole byte may remain before the
        // marker (fill bytes FF are skipped by FillBuffer)
        const uint8_t *bytes = reinterpret_cast<const uint8_t *>(arena_words + d.word0);
        uint32_t p = max((wabs - d.word0) * 4, a0r);
        while (p < a1r && bytes[p] == 0xFFu) p++;
        if (n >= 8 || p < a1r || !(d.flags & 2u)) err |= JB_ST_EXPECT_RST;
    }
    if (err) jb_report_error(status, first_error, d.image, err, 0, g - images[d.image].seg_base);
}
