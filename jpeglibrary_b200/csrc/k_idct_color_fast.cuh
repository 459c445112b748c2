// k_idct_color_fast.cuh -- K2 (fast path): dequantise + un-zigzag + fp32 IDCT + level shift + clamp +
// chroma replication + YCbCr->RGB for 8-bit frames with 1 or 3 components whose luma is sampled at
// (Hmax,Vmax) and whose chroma planes share one sampling ratio of 1 or 2 per axis (4:4:4, 4:2:2,
// 4:4:0, 4:2:0, grey).  Everything else goes through the generic kernel (k_idct_color.cuh).
//
// Same reference functions as k_idct_color.cuh; same bit-exact arithmetic:
//   * dequantisation  (float)(q*c) == fmul(float(q), float(c)): both round the same exact integer once;
//   * IDCT passes use only __fadd_rn/__fsub_rn/__fmul_rn in the reference's order (jb_idct8);
//   * MultiplyInplace(0.125) + MathF.Round (half-to-even) + levelShift is one fma with the magic
//     constant 1.5*2^23 + 128: the product by 0.125 is exact, so the fma rounds the exact value
//     d/8 + 128 + 1.5*2^23 once, to an integer, ties-to-even -- identical to rint(d/8) + 128 because
//     the constant is an even integer.
//
// Structure per CTA (384 threads = 8 threads per 8x8 block, 48 blocks = one strip of MCUs):
//   A. coalesced 128-bit loads of the strip's coefficient blocks -> raw int16 tile in smem;
//      each thread gathers one natural-order row (8 ld.shared.s16 with per-thread constant offsets),
//      dequantises with its quant row held in registers, runs pass 1, transposes through a padded
//      fp32 tile, runs pass 2, and writes clamped 8-bit samples into per-component planes in smem;
//   B. colour: each thread converts 4-pixel groups (chroma terms computed once per chroma sample,
//      clamps done two pixels at a time with the DPX add-min-max instruction) into an RGB staging
//      tile in smem;
//   C. store: full rows of the staging tile leave the SM as bulk asynchronous copies
//      (cp.async.bulk shared->global, one per pixel row) when the destination is 16-byte aligned,
//      else as plain vector/byte stores with edge clipping.
#pragma once
#include "jb_device.cuh"
#include "k_idct_color.cuh"

#define JB_K2F_THREADS 384
#define JB_K2F_BLOCKS 48
#define JB_K2F_RAW_STRIDE 80   // int16 per block in the raw tile (64 + 16 padding: spreads banks)
#define JB_K2F_F_STRIDE 72     // floats per block in the fp32 transpose tile
#define JB_K2F_MAX_ROW_BYTES 1536 // widest strip row: 384 px * 4 B (grey RGBA)
#define JB_K2F_STAGE_BYTES 12288  // 8 rows * 1536 B, or 16 rows * 512 B ... all shapes fit

__constant__ uint8_t jb_c_nat2zz[64] = { // JpegZigZag.cs:15-25: natural index -> zig-zag index
    0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
    41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
    46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

__device__ __forceinline__ uint32_t jb_smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// clamp(a + b, 0, 255) for two packed int16 lanes (DPX VIADDMNMX)
__device__ __forceinline__ uint32_t jb_addclamp2(uint32_t a, uint32_t b)
{
    return __viaddmin_s16x2_relu(a, b, 0x00FF00FFu);
}

struct JbChromaTerms {
    uint32_t r2, g2, b2; // each term duplicated in both 16-bit lanes
};

// apps/JpegDecode/JpegYCbCrToRgbConverter.cs:93-121,171-205 (see jb_ycc_to_rgb)
__device__ __forceinline__ JbChromaTerms jb_chroma_terms(int cb, int cr)
{
    const int cbv = cb - 128, crv = cr - 128;
    const int rt = (91881 * crv + 32768) >> 16;
    const int gt = (-22553 * cbv + 32768 + -46802 * crv) >> 16;
    const int bt = (116130 * cbv + 32768) >> 16;
    JbChromaTerms t;
    t.r2 = __byte_perm((uint32_t)rt, 0, 0x1010);
    t.g2 = __byte_perm((uint32_t)gt, 0, 0x1010);
    t.b2 = __byte_perm((uint32_t)bt, 0, 0x1010);
    return t;
}

// FMT: 0 RGB24, 1 RGBA32, 2 YCBCR888.  HS,VS: chroma subsampling (1 or 2).  NC: 1 or 3 components.
template <int FMT, int NC, int HS, int VS>
__global__ void __launch_bounds__(JB_K2F_THREADS)
jb_k2_idct_color_fast(const JbDevImage *__restrict__ images, const int16_t *__restrict__ coef,
                      const uint16_t *__restrict__ quant, const uint32_t *__restrict__ image_list,
                      int tiles_per_cta)
{
    constexpr int BPM = NC == 1 ? 1 : (HS * VS + 2);
    constexpr int TILE_MCUS = JB_K2F_BLOCKS / BPM;
    constexpr int TW = TILE_MCUS * 8 * HS;      // strip width in pixels (luma)
    constexpr int TH = 8 * VS;                  // strip height
    constexpr int CW = TILE_MCUS * 8;           // chroma plane width
    constexpr int BPP = FMT == 1 ? 4 : 3;
    constexpr int ROW_BYTES = TW * BPP;
    static_assert(ROW_BYTES * TH <= JB_K2F_STAGE_BYTES, "staging tile too small");

    __shared__ __align__(16) int16_t s_raw[JB_K2F_BLOCKS * JB_K2F_RAW_STRIDE];
    __shared__ __align__(16) float s_f[JB_K2F_BLOCKS * JB_K2F_F_STRIDE];
    __shared__ __align__(16) uint8_t s_y[TW * TH];
    __shared__ __align__(16) uint8_t s_c[2][NC == 1 ? 16 : CW * 8];
    __shared__ __align__(128) uint8_t s_stage[JB_K2F_STAGE_BYTES];
    __shared__ JbDevImage s_im;

    const int tid = threadIdx.x;
    const uint32_t image = image_list[blockIdx.y];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(images + image);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_im);
        for (int i = tid; i < (int)(sizeof(JbDevImage) / 4); i += JB_K2F_THREADS) dst[i] = src[i];
    }
    __syncthreads();

    // ---- per-thread constants: which block, which row, its quant row and gather offsets
    const int j = tid >> 3, r = tid & 7;
    const int m = j / BPM, b = j - m * BPM;
    int c = 0, bx, by; // component, block position inside the strip's component plane
    if (NC == 1 || b < HS * VS) {
        bx = m * HS + (b % HS);
        by = b / HS;
    } else {
        c = b - HS * VS + 1;
        bx = m;
        by = 0;
    }
    // the image's own block order (scan order) may differ from Y,Cb,Cr: map through blk_comp
    // (fast path is only selected when blk_comp follows the frame order, checked on the host)
    float qrow[8];
    int goff[8];
    {
        const uint16_t *q = quant + s_im.quant_off + c * 64;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int z = jb_c_nat2zz[r * 8 + e];
            goff[e] = z;
            qrow[e] = (float)q[z];
        }
    }
    const int W = s_im.width, H = s_im.height;
    const uint32_t mcus_per_line = s_im.mcus_per_line;
    const uint32_t strips = (mcus_per_line + TILE_MCUS - 1) / TILE_MCUS;
    const uint32_t ntiles = strips * s_im.mcus_per_col;
    uint8_t *const out = reinterpret_cast<uint8_t *>(s_im.out_ptr);
    const uint64_t pitch = s_im.out_pitch;
    const bool bulk_ok = ((s_im.out_ptr | pitch) & 15u) == 0 && (ROW_BYTES % 16) == 0;
    const unsigned wm = 0xFFFFFFFFu;
    int16_t *rawb = s_raw + j * JB_K2F_RAW_STRIDE;
    float *fb = s_f + j * JB_K2F_F_STRIDE;
    bool bulk_pending = false;

    // strips are walked with incremental (row, column) counters and the next strip's coefficients are
    // prefetched into registers while the current strip is transformed (hides the HBM latency)
    uint32_t tile = blockIdx.x * tiles_per_cta;
    uint32_t mcu_row = tile / strips;
    uint32_t strip = tile - mcu_row * strips;
    const uint32_t tile_end = min(tile + (uint32_t)tiles_per_cta, ntiles);
    // progressive frames keep their coefficients in per-component planes (bx, by are the block's
    // position inside the strip's component plane, which starts at MCU column col0)
    const bool planar = s_im.planar != 0;
    const uint64_t plane_base = s_im.coef_off + s_im.comp_plane_off[c];
    const uint32_t plane_w = s_im.comp_plane_w[c];
    auto load_raw = [&](uint32_t row, uint32_t st) -> uint4 {
        const uint32_t col0 = st * TILE_MCUS;
        const int nm = (int)min((uint32_t)TILE_MCUS, mcus_per_line - col0);
        if (m >= nm) return make_uint4(0, 0, 0, 0);
        uint64_t blk;
        if (!planar) blk = s_im.coef_off + ((uint64_t)row * mcus_per_line + col0) * BPM + j;
        else         blk = plane_base + (uint64_t)(row * (c == 0 ? VS : 1) + by) * plane_w + (col0 * (c == 0 ? HS : 1) + bx);
        return __ldg(reinterpret_cast<const uint4 *>(coef + blk * 64) + r);
    };
    uint4 raw_next = make_uint4(0, 0, 0, 0);
    if (tile < tile_end) raw_next = load_raw(mcu_row, strip);

    for (; tile < tile_end; tile++) {
        const uint32_t mcu_col0 = strip * TILE_MCUS;
        const int nmcu = (int)min((uint32_t)TILE_MCUS, mcus_per_line - mcu_col0);
        const bool valid = m < nmcu;
        const uint32_t cur_row = mcu_row;
        // advance the counters and issue the next strip's loads
        if (++strip == strips) { strip = 0; mcu_row++; }
        const uint4 raw = raw_next;
        if (tile + 1 < tile_end) raw_next = load_raw(mcu_row, strip);

        // ------------------------------------------------ phase A
        if (valid) *reinterpret_cast<uint4 *>(rawb + r * 8) = raw;
        __syncwarp(wm);
        float y[8], d[8];
        if (valid) {
#pragma unroll
            for (int e = 0; e < 8; e++) y[e] = __fmul_rn(qrow[e], (float)rawb[goff[e]]);
            jb_idct8(y, d); // pass 1: along row r
#pragma unroll
            for (int k = 0; k < 8; k++) fb[k * 8 + r] = d[k];
        }
        __syncwarp(wm);
        if (valid) {
            const float4 lo = *reinterpret_cast<const float4 *>(fb + r * 8);
            const float4 hi = *reinterpret_cast<const float4 *>(fb + r * 8 + 4);
            y[0] = lo.x; y[1] = lo.y; y[2] = lo.z; y[3] = lo.w;
            y[4] = hi.x; y[5] = hi.y; y[6] = hi.z; y[7] = hi.w;
            jb_idct8(y, d); // pass 2: along column r
            uint8_t *pl = (c == 0 ? s_y + (by * 8) * TW : s_c[c - 1]) + bx * 8 + r;
            const int pp = c == 0 ? TW : CW;
#pragma unroll
            for (int mr = 0; mr < 8; mr++) {
                // (x * 0.125) rounded half-to-even, + 128, clamped to 0..255
                const float t = __fmaf_rn(d[mr], 0.125f, 12582912.0f + 128.0f);
                const int v = __viaddmin_s32_relu(__float_as_int(t), -0x4B400000, 255);
                pl[mr * pp] = (uint8_t)v;
            }
        }
        // staging tile must be free: the previous iteration's bulk stores have to have read it
        if (bulk_pending && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();

        // ------------------------------------------------ phase B: 4-pixel groups -> staging tile
        constexpr int GROUPS = TW / 4;
        constexpr int ITEMS = GROUPS * (TH / VS);
        for (int item = tid; item < ITEMS; item += JB_K2F_THREADS) {
            const int cy = item / GROUPS, gx = (item - cy * GROUPS) * 4;
            JbChromaTerms t0, t1, t2, t3;
            uint32_t cb4 = 0x80808080u, cr4 = 0x80808080u; // grey: Cb = Cr = 128 (DecodeAction.cs:58-66)
            if (NC == 3) {
                if (HS == 2) {
                    cb4 = *reinterpret_cast<const uint16_t *>(s_c[0] + cy * CW + (gx >> 1));
                    cr4 = *reinterpret_cast<const uint16_t *>(s_c[1] + cy * CW + (gx >> 1));
                } else {
                    cb4 = *reinterpret_cast<const uint32_t *>(s_c[0] + cy * CW + gx);
                    cr4 = *reinterpret_cast<const uint32_t *>(s_c[1] + cy * CW + gx);
                }
            }
            if (FMT != 2) {
                t0 = jb_chroma_terms(cb4 & 0xFF, cr4 & 0xFF);
                t1 = jb_chroma_terms((cb4 >> 8) & 0xFF, (cr4 >> 8) & 0xFF);
                if (HS == 1) {
                    t2 = jb_chroma_terms((cb4 >> 16) & 0xFF, (cr4 >> 16) & 0xFF);
                    t3 = jb_chroma_terms(cb4 >> 24, cr4 >> 24);
                }
            }
#pragma unroll
            for (int rr = 0; rr < VS; rr++) {
                const int py = cy * VS + rr;
                const uint32_t y4 = *reinterpret_cast<const uint32_t *>(s_y + py * TW + gx);
                uint32_t o0, o1, o2, o3 = 0;
                if (FMT == 2) {
                    // interleaved Y Cb Cr with replicated chroma
                    uint32_t cbx, crx; // 4 chroma bytes for the 4 pixels
                    if (HS == 2) {
                        cbx = __byte_perm(cb4, 0, 0x1100);
                        crx = __byte_perm(cr4, 0, 0x1100);
                    } else {
                        cbx = cb4;
                        crx = cr4;
                    }
                    // bytes: y0 cb0 cr0 y1 | cb1 cr1 y2 cb2 | cr2 y3 cb3 cr3
                    const uint32_t a = __byte_perm(y4, cbx, 0x1040);       // y0 cb0 . y1 -> fix below
                    o0 = __byte_perm(a, crx, 0x3410);                      // y0 cb0 cr0 y1
                    const uint32_t bq = __byte_perm(cbx, crx, 0x2051);     // cb1 cr1 . cb2
                    o1 = __byte_perm(bq, y4, 0x3610);                      // cb1 cr1 y2 cb2
                    const uint32_t cq = __byte_perm(crx, cbx, 0x3702);     // cr2 . cb3 cr3
                    o2 = __byte_perm(cq, y4, 0x3270);                      // cr2 y3 cb3 cr3
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
                    // lanes hold 0..255 in bytes 0 and 2
                    if (FMT == 0) {
                        // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
                        const uint32_t rg01 = __byte_perm(r01, g01, 0x6240); // r0 g0 r1 g1
                        const uint32_t rg23 = __byte_perm(r23, g23, 0x6240); // r2 g2 r3 g3
                        o0 = __byte_perm(rg01, b01, 0x2410);                 // r0 g0 b0 r1
                        const uint32_t t = __byte_perm(rg01, b01, 0x0063);   // g1 b1 . .
                        o1 = __byte_perm(t, rg23, 0x5410);                   // g1 b1 r2 g2
                        const uint32_t u = __byte_perm(b23, rg23, 0x2760);   // b2 r3 g3 b3
                        o2 = u;
                    } else {
                        o0 = __byte_perm(__byte_perm(r01, g01, 0x0040), b01, 0x0410) | 0xFF000000u;
                        o1 = __byte_perm(__byte_perm(r01, g01, 0x0062), b01, 0x0610) | 0xFF000000u;
                        o2 = __byte_perm(__byte_perm(r23, g23, 0x0040), b23, 0x0410) | 0xFF000000u;
                        o3 = __byte_perm(__byte_perm(r23, g23, 0x0062), b23, 0x0610) | 0xFF000000u;
                    }
                }
                uint32_t *dst = reinterpret_cast<uint32_t *>(s_stage + py * ROW_BYTES + gx * BPP);
                if (BPP == 4) {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(o0, o1, o2, o3);
                } else {
                    dst[0] = o0; dst[1] = o1; dst[2] = o2;
                }
            }
        }
        // ------------------------------------------------ phase C: staging tile -> global
        const int x0 = mcu_col0 * 8 * HS, y0 = cur_row * TH;
        const int rows = min(TH, H - y0);
        const int row_bytes = min(ROW_BYTES, (W - x0) * BPP);
        if (bulk_ok && row_bytes == ROW_BYTES) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> async proxy
            __syncthreads();
            if (tid == 0) {
                for (int rr = 0; rr < rows; rr++) {
                    uint8_t *g = out + (uint64_t)(y0 + rr) * pitch + (uint64_t)x0 * BPP;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g),
                                 "r"(jb_smem_u32(s_stage + rr * ROW_BYTES)), "n"(ROW_BYTES)
                                 : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            bulk_pending = true;
        } else {
            __syncthreads();
            const int words = row_bytes >> 2;
            for (int i = tid; i < rows * words; i += JB_K2F_THREADS) {
                const int rr = i / words, wv = i - rr * words;
                uint8_t *g = out + (uint64_t)(y0 + rr) * pitch + (uint64_t)x0 * BPP;
                const uint32_t v = *reinterpret_cast<const uint32_t *>(s_stage + rr * ROW_BYTES + wv * 4);
                if ((reinterpret_cast<uint64_t>(g) & 3u) == 0) reinterpret_cast<uint32_t *>(g)[wv] = v;
                else {
                    g[wv * 4] = (uint8_t)v; g[wv * 4 + 1] = (uint8_t)(v >> 8);
                    g[wv * 4 + 2] = (uint8_t)(v >> 16); g[wv * 4 + 3] = (uint8_t)(v >> 24);
                }
            }
            const int tail = row_bytes & 3;
            if (tail && tid < rows * tail) {
                const int rr = tid / tail, tb = (row_bytes & ~3) + tid % tail;
                out[(uint64_t)(y0 + rr) * pitch + (uint64_t)x0 * BPP + tb] = s_stage[rr * ROW_BYTES + tb];
            }
            __syncthreads(); // staging tile is reused by the next iteration
        }
    }
    if (bulk_pending && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
