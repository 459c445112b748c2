// k_idct_color_warp.cuh -- K2 (fast path, second generation): dequantise + un-zigzag + fp32 IDCT + level
// shift + clamp + chroma replication + YCbCr->RGB for 8-bit frames (same coverage and the same bit-exact
// arithmetic as k_idct_color_fast.cuh; see the notes there and in k_idct_color.cuh).
//
// What changed, and why (profiles/r1d_decode.txt): the first generation spent 2/3 of its issue slots around
// the fixed fp32 arithmetic -- gathers, two shared-memory transposes per block, uniform-datapath address
// arithmetic -- and its top stall was the CTA barrier between phases.  Here
//   * ONE THREAD OWNS ONE 8x8 BLOCK: its 64 coefficients arrive as eight 128-bit shared-memory loads, the
//     un-zigzag is a compile-time register permutation, both IDCT passes run in registers (16 independent
//     1-D transforms: plenty of ILP), no transposes;
//   * A WARP OWNS A UNIT of 32/BPM consecutive MCUs (4:2:0: 5 MCUs = 30 blocks = 80x16 pixels) end to end:
//     load -> IDCT -> colour -> store, synchronised with __syncwarp only, so warps drift freely and hide each
//     other's latencies;
//   * the unit's coefficient blocks are contiguous in the store (MCU scan order) and come in by 16-byte
//     cp.async copies that are issued one unit ahead; the unit's pixels leave as ONE 2-D TMA tensor store
//     (cp.async.bulk.tensor, SASS UTMASTG) through a per-image tensor map, which also clips at the image edges
//     (per-row bulk copies / plain stores remain for destinations that are not 16-byte aligned);
//   * dequantisation converts the 16-bit halves of the packed coefficient pairs directly (I2F.S16 on the XU
//     pipe, which nothing else in this kernel uses) and multiplies by the fp32 quantiser: two issue slots per
//     coefficient, none of them on the ALU pipe, which is the busiest one here.
#pragma once
#include "jb_device.cuh"
#include "k_idct_color.cuh"
#include "k_idct_color_fast.cuh"

#define JB_K2W_WARPS 4
#ifndef JB_K2W_MIN_CTAS
#define JB_K2W_MIN_CTAS 4 // CTAs per SM the register allocation aims at (128 registers; 5 -> 96 registers and spills)
#endif
#define JB_K2W_RAW_STRIDE 144 // bytes per block in the raw tile: 128 + 16 so that the eight 128-bit loads of the
                              // eight lanes of a quarter warp hit different banks

// natural (row-major) index -> zig-zag index (JpegZigZag.cs:15-25), usable as a compile-time constant
__host__ __device__ constexpr int jb_nat2zz_c(int n)
{
    constexpr int t[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                           41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                           46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
    return t[n];
}

// jb_chroma_terms with the -128 offsets folded into the rounding constants: (a*(c-128) + r) >> 16 is
// (a*c + (r - 128*a)) >> 16 in exact integer arithmetic
__device__ __forceinline__ JbChromaTerms jb_k2w_chroma_terms(uint32_t cb, uint32_t cr)
{
    const int rt = (91881 * (int)cr + (32768 - 128 * 91881)) >> 16;
    const int gt = (-22553 * (int)cb + (-46802 * (int)cr + (32768 + 128 * 22553 + 128 * 46802))) >> 16;
    const int bt = (116130 * (int)cb + (32768 - 128 * 116130)) >> 16;
    JbChromaTerms t;
    t.r2 = __byte_perm((uint32_t)rt, 0, 0x1010);
    t.g2 = __byte_perm((uint32_t)gt, 0, 0x1010);
    t.b2 = __byte_perm((uint32_t)bt, 0, 0x1010);
    return t;
}

template <int R>
__device__ __forceinline__ void jb_k2w_row(const uint32_t (&pk)[32], const float *__restrict__ qn, float (&d1)[64])
{
    // natural row R: dequantise (DequantizeBlockAndUnZigZag, JpegScanDecoder.cs:50-62) and transform along the row
    const float4 q0 = *reinterpret_cast<const float4 *>(qn + R * 8);
    const float4 q1 = *reinterpret_cast<const float4 *>(qn + R * 8 + 4);
    const float q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    float y[8], d[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int z = jb_nat2zz_c(R * 8 + e);
        // (float)(q*c) == fmul(float(q), float(c)): both round the same exact integer once.  The conversion reads
        // the 16-bit half of the packed pair directly (I2F.S16 on the otherwise idle XU pipe).
        float cf;
        if (z & 1) asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, hi; }" : "=f"(cf) : "r"(pk[z >> 1]));
        else asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, lo; }" : "=f"(cf) : "r"(pk[z >> 1]));
        y[e] = __fmul_rn(q[e], cf);
    }
    jb_idct8(y, d);
#pragma unroll
    for (int e = 0; e < 8; e++) d1[R * 8 + e] = d[e];
}

template <int C>
__device__ __forceinline__ void jb_k2w_col(const float (&d1)[64], uint32_t (&rows)[16])
{
    float y[8], d[8];
#pragma unroll
    for (int k = 0; k < 8; k++) y[k] = d1[k * 8 + C];
    jb_idct8(y, d); // along column C
#pragma unroll
    for (int k = 0; k < 8; k++) {
        // MultiplyInplace(0.125) + MathF.Round (half-to-even) + level shift in one fma, then clamp to 0..255
        const float t = __fmaf_rn(d[k], 0.125f, 12582912.0f + 128.0f);
        const uint32_t v = (uint32_t)__viaddmin_s32_relu(__float_as_int(t), -0x4B400000, 255);
        if ((C & 3) == 0) rows[k * 2 + (C >> 2)] = v;
        else rows[k * 2 + (C >> 2)] += v << (8 * (C & 3)); // disjoint bytes: one shift-add (LEA)
    }
}

#ifndef JB_K2_COMBINE_MAD
#define JB_K2_COMBINE_MAD 1 // the two halves of an RGB word are joined by a multiply-add (FMA pipe: the colour phase is ALU-heavy) instead of PRMT; A/B 10.49 -> 10.33 ms
#endif
#ifndef JB_K2_PACKED
#define JB_K2_PACKED 1 // both IDCT passes on packed fp32 pairs (FADD2 / FFMA2), see jb_idct8x2
#endif

// Packed variants of the two passes: rows R, R + 1 run in lockstep (one f32x2 lane each), then columns C, C + 1.
template <int R>
__device__ __forceinline__ void jb_k2w_row2(const uint32_t (&pk)[32], const float *__restrict__ qn2, jb_f2 (&d1)[32], const jb_f2 nz)
{
    // qn2: the quantisers of rows R, R + 1 interleaved: {q[R][e], q[R + 1][e]} per e
    jb_f2 y[8], d[8];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
        const ulonglong2 q = *reinterpret_cast<const ulonglong2 *>(qn2 + (R / 2) * 16 + e * 2);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int z0 = jb_nat2zz_c(R * 8 + e + u), z1 = jb_nat2zz_c((R + 1) * 8 + e + u);
            float c0, c1;
            if (z0 & 1) asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, hi; }" : "=f"(c0) : "r"(pk[z0 >> 1]));
            else asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, lo; }" : "=f"(c0) : "r"(pk[z0 >> 1]));
            if (z1 & 1) asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, hi; }" : "=f"(c1) : "r"(pk[z1 >> 1]));
            else asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, lo; }" : "=f"(c1) : "r"(pk[z1 >> 1]));
            // the product of two integers below 2^24 is exact: fma(q, c, -0.0) == fmul(q, c) == (float)(q * c)
            y[e + u] = jb_fma2(u ? q.y : q.x, jb_pack2(c0, c1), nz);
        }
    }
    jb_idct8x2(y, d, nz);
#pragma unroll
    for (int e = 0; e < 8; e++) d1[(R / 2) * 8 + e] = d[e]; // {d1[R][e], d1[R + 1][e]}
}

template <int C>
__device__ __forceinline__ void jb_k2w_col2(const jb_f2 (&d1)[32], uint32_t (&rows)[16], const jb_f2 nz)
{
    jb_f2 y[8], d[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { // {d1[k][C], d1[k][C + 1]}
        const jb_f2 a = d1[(k / 2) * 8 + C], b = d1[(k / 2) * 8 + C + 1];
        y[k] = (k & 1) ? jb_pack2(jb_hi2(a), jb_hi2(b)) : jb_pack2(jb_lo2(a), jb_lo2(b));
    }
    jb_idct8x2(y, d, nz); // along columns C, C + 1
    const jb_f2 eighth = jb_pack2(0.125f, 0.125f), bias = jb_pack2(12582912.0f + 128.0f, 12582912.0f + 128.0f);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        // MultiplyInplace(0.125) + MathF.Round (half-to-even) + level shift in one fma: the float's bits are
        // 0x4B400000 + sample, so their low halves are the int16 the reference stores in its block (the (short) cast of
        // ShiftDataLevel, JpegScanDecoder.cs:64-73, wrap-around included while |sample| < 2^22); two of them side by
        // side are clamped to 0..255 by one VIMNMX.S16x2.RELU, the clamp of the 8-bit writers
        const jb_f2 t = jb_fma2(d[k], eighth, bias);
        const uint32_t s2 = __byte_perm(__float_as_uint(jb_lo2(t)), __float_as_uint(jb_hi2(t)), 0x5410);
        const uint32_t v = __vimin_s16x2_relu(s2, 0x00FF00FFu); // bytes: v(C) 0 v(C + 1) 0
        if ((C & 3) == 0) rows[k * 2 + (C >> 2)] = v;
        else rows[k * 2 + (C >> 2)] = __byte_perm(rows[k * 2 + (C >> 2)], v, 0x6420);
    }
}

// FMT: 0 RGB24, 1 RGBA32, 2 YCBCR888.  HS,VS: chroma subsampling (1 or 2).  NC: 1 or 3 components.
template <int FMT, int NC, int HS, int VS>
__global__ void __launch_bounds__(JB_K2W_WARPS * 32, JB_K2W_MIN_CTAS)
jb_k2_idct_color_warp(const JbDevImage *__restrict__ images, const int16_t *__restrict__ coef,
                      const uint16_t *__restrict__ quant, const uint32_t *__restrict__ image_list,
                      int units_per_warp, const uint32_t *__restrict__ mcu_limit, const unsigned long long negzero2)
{
    constexpr int BPM = NC == 1 ? 1 : (HS * VS + 2);
    constexpr int UM = NC == 1 ? 16 : 32 / BPM; // MCUs per unit (grey: 16, which keeps the tiles inside 48 KB)
    constexpr int NB = UM * BPM;                // blocks (= busy lanes) per unit
    constexpr int TW = UM * 8 * HS;             // unit width in pixels (luma)
    constexpr int TH = 8 * VS;                  // unit height
    constexpr int CW = UM * 8;                  // chroma plane width
    constexpr int BPP = FMT == 1 ? 4 : 3;
    constexpr int ROW_BYTES = TW * BPP;
    constexpr int CHUNKS = NB * 8;              // 16-byte pieces of the unit's coefficient blocks
    static_assert(ROW_BYTES % 16 == 0, "rows must be bulk-copyable");

    // Shared-memory layout of a warp's sample planes.  The generic layout is dense (luma TW x TH, then Cb, then Cr,
    // CW x 8 each).  The 4:2:0 instances (the headline shape) use a layout in which every access of the three phases is
    // bank-conflict free (profiles/r2b_decode_batch1024.txt had 459 shared-memory wavefronts per unit against 198
    // ideal, which at one wavefront per clock was 10.5 of the kernel's 11.9 ms):
    //   * luma rows 8..15 start 64 bytes later, so the plane stores of the blocks of MCU row 0 and 1 (same columns,
    //     640 bytes = 0 banks apart in the dense plane) land 16 banks apart;
    //   * chroma rows are 80 bytes apart like luma rows (the relative position of luma and chroma stores is then the
    //     same for each of the eight row stores), the five chroma blocks of a row sit in the order 2 3 4 - 0 1, Cb
    //     starts at bank 4 and Cr at bank 20: the 16 lanes of either half warp store to 32 different banks;
    //   * phase B takes its items (4 pixels x 2 rows) in the order of s_item: 4 chroma rows x 8 neighbouring groups per
    //     round, whose luma loads (8 banks per chroma row), chroma loads and 12-byte staging stores (banks 3 g - 8 cy)
    //     are all distinct; the 20 groups of a row leave one round of 8 rows x 4 groups that is 2-way (the optimum).
    constexpr bool TUNED = NC == 3 && HS == 2 && VS == 2 && FMT != 1; // (RGBA32: the wider staging tile leaves no room)
    constexpr int YGAP = TUNED ? 64 : 0;          // bytes skipped after every band of 8 luma rows
    constexpr int CROW = TUNED ? 80 : CW;         // chroma row stride
    constexpr int CPLANE = NC == 1 ? 16 : CROW * 8;
    constexpr int CB0 = TUNED ? 1424 : TW * TH;   // byte offset of the Cb plane (tuned: word 356 = bank 4)
    constexpr int CR0 = CB0 + CPLANE + YGAP;      // (tuned: word 532 = bank 20)
    constexpr int PLANES = (CR0 + CPLANE + 127) & ~127;
    static_assert(!TUNED || (TW * TH + YGAP <= CB0 && (CB0 / 4) % 32 == 4 && (CR0 / 4) % 32 == 20), "bank plan");
    __shared__ __align__(16) uint8_t s_raw[JB_K2W_WARPS][NB * JB_K2W_RAW_STRIDE];
    __shared__ __align__(128) uint8_t s_pl[JB_K2W_WARPS][PLANES];
    __shared__ __align__(128) uint8_t s_stage[JB_K2W_WARPS][TH * ROW_BYTES];
    // quantisers in NATURAL order, one table per component (packed IDCT: rows 2r, 2r + 1 interleaved element-wise); the
    // tables are 68 floats apart: lanes of different components read the same 16-byte piece of their tables in one
    // LDS.128, and 64 floats apart those were three addresses on the same four banks (12 wavefronts instead of 4)
    constexpr int QSTRIDE = 68;
    __shared__ __align__(16) float s_qn[NC * QSTRIDE];
    constexpr int GROUPS = TW / 4;                // phase B works on items of 4 pixels x VS rows (one chroma row)
    constexpr int ITEMS = GROUPS * (TH / VS);
    static_assert(ITEMS % 32 == 0, "items must spread evenly over the lanes");
    static_assert(!TUNED || ITEMS == 160, "item order below");
    constexpr bool ITEM_TABLE = !(TUNED && FMT == 0); // (that instance has the order written out in phase B)
    __shared__ uint32_t s_item[ITEM_TABLE ? ITEMS : 1]; // per item: chroma | luma << 10 | (staging >> 2) << 21 byte offsets
    struct Im { // the fields of JbDevImage this kernel uses (the whole descriptor is 1.4 KB)
        uint64_t out_ptr, out_pitch, tmap_ptr, coef_off;
        uint32_t tmap_shift, quant_off, mcus_per_line, mcus_per_col, planar, width, height;
        uint32_t comp_plane_off[4], comp_plane_w[4];
    };
    __shared__ Im s_im;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t image = image_list[blockIdx.y];
    if (tid == 0) {
        const JbDevImage &g = images[image];
        s_im.out_ptr = g.out_ptr; s_im.out_pitch = g.out_pitch; s_im.tmap_ptr = g.tmap_ptr; s_im.coef_off = g.coef_off;
        s_im.tmap_shift = g.tmap_shift; s_im.quant_off = g.quant_off; s_im.mcus_per_line = g.mcus_per_line;
        s_im.mcus_per_col = g.mcus_per_col; s_im.planar = g.planar; s_im.width = g.width; s_im.height = g.height;
    } else if (tid < 5) {
        s_im.comp_plane_off[tid - 1] = images[image].comp_plane_off[tid - 1];
        s_im.comp_plane_w[tid - 1] = images[image].comp_plane_w[tid - 1];
    }
    __syncthreads();
    for (int i = tid; i < NC * 64; i += JB_K2W_WARPS * 32) {
#if JB_K2_PACKED
        const int n = i & 63, r = ((n >> 4) << 1) | (n & 1), e = (n >> 1) & 7; // slot (r/2)*16 + e*2 + (r&1)
        s_qn[(i >> 6) * QSTRIDE + (i & 63)] = (float)quant[s_im.quant_off + (i >> 6) * 64 + jb_c_nat2zz[r * 8 + e]];
#else
        s_qn[(i >> 6) * QSTRIDE + (i & 63)] = (float)quant[s_im.quant_off + (i >> 6) * 64 + jb_c_nat2zz[i & 63]];
#endif
    }
    for (int i = tid; i < (ITEM_TABLE ? ITEMS : 0); i += JB_K2W_WARPS * 32) {
        int cy, g;
        if (TUNED) {
            const int it = i >> 5, l = i & 31;
            if (it < 4) { cy = (it >> 1) * 4 + (l >> 3); g = (it & 1) * 8 + (l & 7); }
            else { cy = l >> 2; g = 16 + (l & 3); }
        } else {
            cy = i / GROUPS;
            g = i - cy * GROUPS;
        }
        const int gx = g * 4, cx = gx / HS; // first pixel of the group, luma and chroma
        const int coff = cy * CROW + (TUNED ? ((cx >> 3) + 4) % 6 * 8 + (cx & 7) : cx);
        const int yoff = cy * VS * TW + ((cy * VS) >> 3) * YGAP + gx;
        s_item[i] = (uint32_t)coff | ((uint32_t)yoff << 10) | ((uint32_t)((cy * VS * ROW_BYTES + gx * BPP) >> 2) << 21);
    }
    __syncthreads();

    // ---- per-lane constants: which block of the unit this lane owns
    const int j = lane;                         // block index inside the unit (scan order)
    const int m = j / BPM, b = j - m * BPM;
    int c = 0, bx, by;                          // component, block position inside the unit's component plane
    if (NC == 1 || b < HS * VS) {
        bx = m * HS + (b % HS);
        by = b / HS;
    } else {
        c = b - HS * VS + 1;
        bx = m;
        by = 0;
    }
    const float *qn = s_qn + c * QSTRIDE;
    const int W = s_im.width, H = s_im.height;
    const uint32_t mcus_per_line = s_im.mcus_per_line;
    const uint32_t upr = (mcus_per_line + UM - 1) / UM; // units per MCU row
    const uint32_t nunits = upr * s_im.mcus_per_col;
    uint8_t *const out = reinterpret_cast<uint8_t *>(s_im.out_ptr);
    const uint64_t pitch = s_im.out_pitch;
    const bool bulk_ok = ((s_im.out_ptr | pitch) & 15u) == 0;
    const uint64_t tmap = s_im.tmap_ptr;
    const int tmap_shift = (int)s_im.tmap_shift;
    const bool planar = s_im.planar != 0;
    // MCUs from here on were never decoded (the scan ended at an EOI on a restart boundary): the reference never calls
    // WriteBlock for them (JpegHuffmanBaselineScanDecoder.cs:144-150), their pixels are left as they are
    const uint32_t limit = mcu_limit ? mcu_limit[image] : 0xFFFFFFFFu;
    uint8_t *raw = s_raw[wid];
    uint8_t *yplane = s_pl[wid];
    uint8_t *stage = s_stage[wid];

    uint32_t unit = (blockIdx.x * JB_K2W_WARPS + wid) * (uint32_t)units_per_warp;
    const uint32_t unit_end = min(unit + (uint32_t)units_per_warp, nunits);
    if (unit >= unit_end) return;
    uint32_t mcu_row = unit / upr;
    uint32_t ucol = unit - mcu_row * upr;

    // 16-byte cp.async copies of one unit's coefficient blocks into the padded raw tile.  In the interleaved
    // store the unit's blocks are contiguous: chunk i = 32 r + lane comes from base + 16 i and goes to
    // 144 (i / 8) + 16 (i % 8), i.e. both sides advance by a constant per round.
    const uint32_t raw_lane = jb_smem_u32(raw) + lane * 16 + (lane >> 3) * 16;
    auto fetch = [&](uint32_t row, uint32_t uc) {
        const uint32_t col0 = uc * UM;
        const int nm = (int)min((uint32_t)UM, mcus_per_line - col0);
        if (!planar) {
            const uint8_t *src = reinterpret_cast<const uint8_t *>(coef + (s_im.coef_off + ((uint64_t)row * mcus_per_line + col0) * BPM) * 64) + lane * 16;
            const int limit = nm * BPM * 8 - lane;
#pragma unroll
            for (int r = 0; r < (CHUNKS + 31) / 32; r++)
                if (r * 32 < limit)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(raw_lane + r * (4 * JB_K2W_RAW_STRIDE)), "l"(src + r * 512) : "memory");
        } else {
#pragma unroll 1
            for (int i = lane; i < CHUNKS; i += 32) {
                const int jb = i >> 3, part = i & 7;
                const int mm = jb / BPM;
                if (mm < nm) {
                    const int bb = jb - mm * BPM;
                    int cc = 0, pbx, pby;
                    if (NC == 1 || bb < HS * VS) { pbx = mm * HS + (bb % HS); pby = bb / HS; }
                    else { cc = bb - HS * VS + 1; pbx = mm; pby = 0; }
                    const uint64_t blk = s_im.coef_off + s_im.comp_plane_off[cc] +
                                         (uint64_t)(row * (cc == 0 ? VS : 1) + pby) * s_im.comp_plane_w[cc] + (col0 * (cc == 0 ? HS : 1) + pbx);
                    const void *src = reinterpret_cast<const uint8_t *>(coef + blk * 64) + part * 16;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(jb_smem_u32(raw + jb * JB_K2W_RAW_STRIDE + part * 16)),
                                 "l"(src)
                                 : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(mcu_row, ucol);
    bool bulk_pending = false;

    for (; unit < unit_end; unit++) {
        const uint32_t mcu_col0 = ucol * UM;
        int nmcu = (int)min((uint32_t)UM, mcus_per_line - mcu_col0);
        const uint32_t mcu0 = mcu_row * mcus_per_line + mcu_col0;
        const bool clipped = mcu0 + (uint32_t)nmcu > limit; // (never on intact streams)
        if (clipped) nmcu = limit > mcu0 ? (int)(limit - mcu0) : 0;
        const bool valid = j < NB && m < nmcu;
        const uint32_t cur_row = mcu_row;
        if (++ucol == upr) { ucol = 0; mcu_row++; }

        // ------------------------------------------------ phase A: one block per lane, all in registers
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        uint32_t pk[32];
        if (valid) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(raw + j * JB_K2W_RAW_STRIDE);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint4 v = rp[i];
                pk[4 * i] = v.x; pk[4 * i + 1] = v.y; pk[4 * i + 2] = v.z; pk[4 * i + 3] = v.w;
            }
        }
        __syncwarp();
        if (unit + 1 < unit_end) fetch(mcu_row, ucol); // the raw tile is free again: prefetch the next unit
        if (valid) {
            uint32_t rows[16];
#if JB_K2_PACKED
            jb_f2 d1[32];
            jb_k2w_row2<0>(pk, qn, d1, negzero2); jb_k2w_row2<2>(pk, qn, d1, negzero2);
            jb_k2w_row2<4>(pk, qn, d1, negzero2); jb_k2w_row2<6>(pk, qn, d1, negzero2);
            jb_k2w_col2<0>(d1, rows, negzero2); jb_k2w_col2<2>(d1, rows, negzero2);
            jb_k2w_col2<4>(d1, rows, negzero2); jb_k2w_col2<6>(d1, rows, negzero2);
#else
            float d1[64];
            jb_k2w_row<0>(pk, qn, d1); jb_k2w_row<1>(pk, qn, d1); jb_k2w_row<2>(pk, qn, d1); jb_k2w_row<3>(pk, qn, d1);
            jb_k2w_row<4>(pk, qn, d1); jb_k2w_row<5>(pk, qn, d1); jb_k2w_row<6>(pk, qn, d1); jb_k2w_row<7>(pk, qn, d1);
            jb_k2w_col<0>(d1, rows); jb_k2w_col<1>(d1, rows); jb_k2w_col<2>(d1, rows); jb_k2w_col<3>(d1, rows);
            jb_k2w_col<4>(d1, rows); jb_k2w_col<5>(d1, rows); jb_k2w_col<6>(d1, rows); jb_k2w_col<7>(d1, rows);
#endif
            uint8_t *pl = c == 0 ? yplane + by * (8 * TW + YGAP) + bx * 8
                                 : yplane + (c == 1 ? CB0 : CR0) + (TUNED ? (bx + 4) % 6 : bx) * 8;
            const int pp = c == 0 ? TW : CROW;
#pragma unroll
            for (int k = 0; k < 8; k++) *reinterpret_cast<uint2 *>(pl + k * pp) = make_uint2(rows[2 * k], rows[2 * k + 1]);
        }
        // the staging tile must be free: the previous unit's bulk stores have to have read it
        if (bulk_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();

        // ------------------------------------------------ phase B: 4-pixel groups -> staging tile
        if constexpr (TUNED && FMT == 0) {
            // RGB24 from 4:2:0, the headline shape: the item order of s_item written out, so that every address is a
            // per-lane base plus a compile-time offset, and the twelve bytes of a group built as three words
            // [r0 g0 b0 r1] [g1 b1 r2 g2] [b2 r3 g3 b3] = even bytes | odd bytes << 8, where each half is ONE packed
            // add-and-clamp of a luma pair and a term pair: 5 + 6 + 3 instructions per row of four pixels instead of
            // 2 + 6 + 6, and the >> 16 of the chroma terms (jb_chroma_terms) is the byte selection that pairs them.
            const int q = lane >> 3, r = lane & 7, q2 = lane >> 2, r2 = lane & 3;
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const uint8_t *cp, *yp;
                uint8_t *sp;
                if (it < 4) { // chroma rows 4 (it / 2) + q, groups 8 (it % 2) + r
                    cp = yplane + CB0 + q * CROW + 2 * r + (it >> 1) * (4 * CROW) + ((it & 1) ? 0 : 32);
                    yp = yplane + q * (2 * TW) + 4 * r + (it >> 1) * (8 * TW + YGAP) + (it & 1) * 32;
                    sp = stage + q * (2 * ROW_BYTES) + 12 * r + (it >> 1) * (8 * ROW_BYTES) + (it & 1) * 96;
                } else {      // chroma rows q2, groups 16 + r2
                    cp = yplane + CB0 + q2 * CROW + 2 * r2 + 16;
                    yp = yplane + q2 * (2 * TW) + (q2 >> 2) * YGAP + 4 * r2 + 64;
                    sp = stage + q2 * (2 * ROW_BYTES) + 12 * r2 + 192;
                }
                const uint32_t cbw = *reinterpret_cast<const uint16_t *>(cp);
                const uint32_t crw = *reinterpret_cast<const uint16_t *>(cp + (CR0 - CB0));
                const int cb0 = (int)(cbw & 0xFF), cb1 = (int)(cbw >> 8), cr0 = (int)(crw & 0xFF), cr1 = (int)(crw >> 8);
                // the products of jb_k2w_chroma_terms before their >> 16
                const uint32_t pr0 = (uint32_t)(91881 * cr0 + (32768 - 128 * 91881));
                const uint32_t pr1 = (uint32_t)(91881 * cr1 + (32768 - 128 * 91881));
                const uint32_t pg0 = (uint32_t)(-22553 * cb0 + (-46802 * cr0 + (32768 + 128 * 22553 + 128 * 46802)));
                const uint32_t pg1 = (uint32_t)(-22553 * cb1 + (-46802 * cr1 + (32768 + 128 * 22553 + 128 * 46802)));
                const uint32_t pb0 = (uint32_t)(116130 * cb0 + (32768 - 128 * 116130));
                const uint32_t pb1 = (uint32_t)(116130 * cb1 + (32768 - 128 * 116130));
                // term pairs: the upper halves of two products side by side
                const uint32_t t_rb0 = __byte_perm(pr0, pb0, 0x7632), t_gr0 = __byte_perm(pg0, pr0, 0x7632);
                const uint32_t t_g0r1 = __byte_perm(pg0, pr1, 0x7632), t_b0g1 = __byte_perm(pb0, pg1, 0x7632);
                const uint32_t t_bg1 = __byte_perm(pb1, pg1, 0x7632), t_rb1 = __byte_perm(pr1, pb1, 0x7632);
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const uint32_t y4 = *reinterpret_cast<const uint32_t *>(yp + rr * TW);
                    const uint32_t y00 = __byte_perm(y4, 0, 0x4040), y01 = __byte_perm(y4, 0, 0x4140);
                    const uint32_t y12 = __byte_perm(y4, 0, 0x4241), y23 = __byte_perm(y4, 0, 0x4342);
                    const uint32_t y33 = __byte_perm(y4, 0, 0x4343);
                    uint32_t *dst = reinterpret_cast<uint32_t *>(sp + rr * ROW_BYTES);
#if JB_K2_COMBINE_MAD
                    dst[0] = jb_addclamp2(y00, t_rb0) + jb_addclamp2(y01, t_gr0) * 256u;   // r0 g0 b0 r1
                    dst[1] = jb_addclamp2(y12, t_g0r1) + jb_addclamp2(y12, t_b0g1) * 256u; // g1 b1 r2 g2
                    dst[2] = jb_addclamp2(y23, t_bg1) + jb_addclamp2(y33, t_rb1) * 256u;   // b2 r3 g3 b3
#else
                    dst[0] = __byte_perm(jb_addclamp2(y00, t_rb0), jb_addclamp2(y01, t_gr0), 0x6240);   // r0 g0 b0 r1
                    dst[1] = __byte_perm(jb_addclamp2(y12, t_g0r1), jb_addclamp2(y12, t_b0g1), 0x6240); // g1 b1 r2 g2
                    dst[2] = __byte_perm(jb_addclamp2(y23, t_bg1), jb_addclamp2(y33, t_rb1), 0x6240);   // b2 r3 g3 b3
#endif
                }
            }
        } else
#pragma unroll
        for (int it = 0; it < ITEMS / 32; it++) {
            const uint32_t io = s_item[it * 32 + lane];
            const uint8_t *cp = yplane + CB0 + (io & 1023u);
            const uint8_t *yp = yplane + ((io >> 10) & 2047u);
            uint8_t *sp = stage + ((io >> 21) << 2);
            JbChromaTerms t0, t1, t2, t3;
            uint32_t cb4 = 0x80808080u, cr4 = 0x80808080u; // grey: Cb = Cr = 128 (DecodeAction.cs:58-66)
            if (NC == 3) {
                if (HS == 2) {
                    cb4 = *reinterpret_cast<const uint16_t *>(cp);
                    cr4 = *reinterpret_cast<const uint16_t *>(cp + (CR0 - CB0));
                } else {
                    cb4 = *reinterpret_cast<const uint32_t *>(cp);
                    cr4 = *reinterpret_cast<const uint32_t *>(cp + (CR0 - CB0));
                }
            }
            if (FMT != 2) {
                t0 = jb_k2w_chroma_terms(cb4 & 0xFF, cr4 & 0xFF);
                t1 = jb_k2w_chroma_terms(__byte_perm(cb4, 0, 0x4441), __byte_perm(cr4, 0, 0x4441));
                if (HS == 1) {
                    t2 = jb_k2w_chroma_terms(__byte_perm(cb4, 0, 0x4442), __byte_perm(cr4, 0, 0x4442));
                    t3 = jb_k2w_chroma_terms(cb4 >> 24, cr4 >> 24);
                }
            }
#pragma unroll
            for (int rr = 0; rr < VS; rr++) {
                const uint32_t y4 = *reinterpret_cast<const uint32_t *>(yp + rr * TW);
                uint32_t o0, o1, o2, o3 = 0;
                if (FMT == 2) {
                    uint32_t cbx, crx; // 4 chroma bytes for the 4 pixels
                    if (HS == 2) {
                        cbx = __byte_perm(cb4, 0, 0x1100);
                        crx = __byte_perm(cr4, 0, 0x1100);
                    } else {
                        cbx = cb4;
                        crx = cr4;
                    }
                    // bytes: y0 cb0 cr0 y1 | cb1 cr1 y2 cb2 | cr2 y3 cb3 cr3
                    const uint32_t a = __byte_perm(y4, cbx, 0x1040);
                    o0 = __byte_perm(a, crx, 0x3410);
                    const uint32_t bq = __byte_perm(cbx, crx, 0x2051);
                    o1 = __byte_perm(bq, y4, 0x3610);
                    const uint32_t cq = __byte_perm(crx, cbx, 0x3702);
                    o2 = __byte_perm(cq, y4, 0x3270);
                } else {
                    const uint32_t y01 = __byte_perm(y4, 0, 0x4140); // two 16-bit lanes: y0, y1
                    const uint32_t y23 = __byte_perm(y4, 0, 0x4342);
                    uint32_t r01, g01, b01, r23, g23, b23;
                    if (HS == 2) {
                        r01 = jb_addclamp2(y01, t0.r2); g01 = jb_addclamp2(y01, t0.g2); b01 = jb_addclamp2(y01, t0.b2);
                        r23 = jb_addclamp2(y23, t1.r2); g23 = jb_addclamp2(y23, t1.g2); b23 = jb_addclamp2(y23, t1.b2);
                    } else {
                        const uint32_t ra = __byte_perm(t0.r2, t1.r2, 0x5410), ga = __byte_perm(t0.g2, t1.g2, 0x5410),
                                       ba = __byte_perm(t0.b2, t1.b2, 0x5410);
                        const uint32_t rb = __byte_perm(t2.r2, t3.r2, 0x5410), gb = __byte_perm(t2.g2, t3.g2, 0x5410),
                                       bb = __byte_perm(t2.b2, t3.b2, 0x5410);
                        r01 = jb_addclamp2(y01, ra); g01 = jb_addclamp2(y01, ga); b01 = jb_addclamp2(y01, ba);
                        r23 = jb_addclamp2(y23, rb); g23 = jb_addclamp2(y23, gb); b23 = jb_addclamp2(y23, bb);
                    }
                    if (FMT == 0) {
                        // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
                        const uint32_t rg01 = __byte_perm(r01, g01, 0x6240);
                        const uint32_t rg23 = __byte_perm(r23, g23, 0x6240);
                        o0 = __byte_perm(rg01, b01, 0x2410);
                        const uint32_t t = __byte_perm(rg01, b01, 0x0063);
                        o1 = __byte_perm(t, rg23, 0x5410);
                        o2 = __byte_perm(b23, rg23, 0x2760);
                    } else {
                        o0 = __byte_perm(__byte_perm(r01, g01, 0x0040), b01, 0x0410) | 0xFF000000u;
                        o1 = __byte_perm(__byte_perm(r01, g01, 0x0062), b01, 0x0610) | 0xFF000000u;
                        o2 = __byte_perm(__byte_perm(r23, g23, 0x0040), b23, 0x0410) | 0xFF000000u;
                        o3 = __byte_perm(__byte_perm(r23, g23, 0x0062), b23, 0x0610) | 0xFF000000u;
                    }
                }
                uint32_t *dst = reinterpret_cast<uint32_t *>(sp + rr * ROW_BYTES);
                if (BPP == 4) {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(o0, o1, o2, o3);
                } else {
                    dst[0] = o0; dst[1] = o1; dst[2] = o2;
                }
            }
        }
        // ------------------------------------------------ phase C: staging tile -> global
        const int x0 = mcu_col0 * 8 * HS, y0 = cur_row * TH;
        const int rows_out = min(TH, H - y0);
        const int row_bytes = clipped ? min(nmcu * 8 * HS, W - x0) * BPP : min(ROW_BYTES, (W - x0) * BPP);
        if (tmap && !clipped) {
            // one 2-D TMA tensor store for the whole unit; the hardware clips at the right and bottom image edges
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> async proxy
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap),
                             "r"((x0 * BPP) >> tmap_shift), "r"(y0), "r"(jb_smem_u32(stage))
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            bulk_pending = true;
        } else if (bulk_ok && row_bytes == ROW_BYTES && !clipped) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> async proxy
            __syncwarp();
            if (lane < rows_out) {
                uint8_t *g = out + (uint64_t)(y0 + lane) * pitch + (uint64_t)x0 * BPP;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g),
                             "r"(jb_smem_u32(stage + lane * ROW_BYTES)), "n"(ROW_BYTES)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            bulk_pending = true;
        } else {
            __syncwarp();
            const int words = row_bytes >> 2;
            for (int i = lane; i < rows_out * words; i += 32) {
                const int rr = i / words, wv = i - rr * words;
                uint8_t *g = out + (uint64_t)(y0 + rr) * pitch + (uint64_t)x0 * BPP;
                const uint32_t v = *reinterpret_cast<const uint32_t *>(stage + rr * ROW_BYTES + wv * 4);
                if ((reinterpret_cast<uint64_t>(g) & 3u) == 0) reinterpret_cast<uint32_t *>(g)[wv] = v;
                else {
                    g[wv * 4] = (uint8_t)v; g[wv * 4 + 1] = (uint8_t)(v >> 8);
                    g[wv * 4 + 2] = (uint8_t)(v >> 16); g[wv * 4 + 3] = (uint8_t)(v >> 24);
                }
            }
            const int tail = row_bytes & 3;
            if (tail)
                for (int i = lane; i < rows_out * tail; i += 32) {
                    const int rr = i / tail, tb = (row_bytes & ~3) + i % tail;
                    out[(uint64_t)(y0 + rr) * pitch + (uint64_t)x0 * BPP + tb] = stage[rr * ROW_BYTES + tb];
                }
            __syncwarp(); // the staging tile is reused by the next unit
        }
    }
    if (bulk_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
