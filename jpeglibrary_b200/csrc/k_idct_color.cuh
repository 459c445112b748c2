// k_idct_color.cuh -- K2: dequantise + un-zigzag + fp32 IDCT + level shift + pixel replication +
// clamp + YCbCr->RGB, fused, one pass from the coefficient store to pitch-linear pixels.
//
// Replaces:
//   DequantizeBlockAndUnZigZag / ShiftDataLevel   (ScanDecoder/JpegScanDecoder.cs:50-73)
//   FastFloatingPointDCT.TransformIDCT            (FastFloatingPointDCT.cs:54-185)
//   WriteBlock / WriteBlockSlow (replication)     (ScanDecoder/JpegHuffmanBaselineScanDecoder.cs:225-268,
//                                                  JpegBlockAllocator.cs:151-190)
//   JpegBufferOutputWriter8Bit.WriteBlock         (apps/JpegDecode/JpegBufferOutputWriter8Bit.cs:28-60)
//   JpegBufferOutputWriterGreaterThan8Bit         (apps/JpegDecode/JpegBufferOutputWriterGreaterThan8Bit.cs:34-68)
//   JpegYCbCrToRgbConverter.ConvertYCbCr8ToRgb24  (apps/JpegDecode/JpegYCbCrToRgbConverter.cs:134-205)
//
// Bit-exactness: the reference IDCT is fp32 with separately rounded multiplies and adds in a fixed
// order (RyuJIT never contracts).  Every operation below is an explicit __fmul_rn/__fadd_rn/__fsub_rn,
// which nvcc never fuses into FMA, so the int16 samples are bit-identical to the reference's.
#pragma once
#include "jb_device.cuh"

#define JB_K2_MAX_BLOCKS 48
#define JB_K2_THREADS (JB_K2_MAX_BLOCKS * 8)
#define JB_K2_BLOCK_STRIDE 72 // floats per block in shared memory (64 + 8: conflict-free transposes)

__constant__ uint8_t jb_c_zigzag[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// One 1-D pass of FastFloatingPointDCT.IDCT8x4_{Left,Right}Part (FastFloatingPointDCT.cs:79-185),
// same operations in the same order.  y[] in, d[] out.
__device__ __forceinline__ void jb_idct8(const float y[8], float d[8])
{
    const float C_1_175876 = 1.175875602f, C_1_961571 = -1.961570560f, C_0_390181 = -0.390180644f,
                C_0_899976 = -0.899976223f, C_2_562915 = -2.562915447f, C_0_298631 = 0.298631336f,
                C_2_053120 = 2.053119869f, C_3_072711 = 3.072711026f, C_1_501321 = 1.501321110f,
                C_0_541196 = 0.541196100f, C_1_847759 = -1.847759065f, C_0_765367 = 0.765366865f;
    const float my1 = y[1], my7 = y[7], my3 = y[3], my5 = y[5];
    float mz0 = __fadd_rn(my1, my7);
    float mz2 = __fadd_rn(my3, my7);
    float mz1 = __fadd_rn(my3, my5);
    float mz3 = __fadd_rn(my1, my5);
    float mz4 = __fmul_rn(__fadd_rn(mz0, mz1), C_1_175876);
    mz2 = __fadd_rn(__fmul_rn(mz2, C_1_961571), mz4);
    mz3 = __fadd_rn(__fmul_rn(mz3, C_0_390181), mz4);
    mz0 = __fmul_rn(mz0, C_0_899976);
    mz1 = __fmul_rn(mz1, C_2_562915);
    const float mb3 = __fadd_rn(__fadd_rn(__fmul_rn(my7, C_0_298631), mz0), mz2);
    const float mb2 = __fadd_rn(__fadd_rn(__fmul_rn(my5, C_2_053120), mz1), mz3);
    const float mb1 = __fadd_rn(__fadd_rn(__fmul_rn(my3, C_3_072711), mz1), mz2);
    const float mb0 = __fadd_rn(__fadd_rn(__fmul_rn(my1, C_1_501321), mz0), mz3);
    const float my2 = y[2], my6 = y[6], my0 = y[0], my4 = y[4];
    mz4 = __fmul_rn(__fadd_rn(my2, my6), C_0_541196);
    mz0 = __fadd_rn(my0, my4);
    mz1 = __fsub_rn(my0, my4);
    mz2 = __fadd_rn(mz4, __fmul_rn(my6, C_1_847759));
    mz3 = __fadd_rn(mz4, __fmul_rn(my2, C_0_765367));
    const float a0 = __fadd_rn(mz0, mz3);
    const float a3 = __fsub_rn(mz0, mz3);
    const float a1 = __fadd_rn(mz1, mz2);
    const float a2 = __fsub_rn(mz1, mz2);
    d[0] = __fadd_rn(a0, mb0);
    d[7] = __fsub_rn(a0, mb0);
    d[1] = __fadd_rn(a1, mb1);
    d[6] = __fsub_rn(a1, mb1);
    d[2] = __fadd_rn(a2, mb2);
    d[5] = __fsub_rn(a2, mb2);
    d[3] = __fadd_rn(a3, mb3);
    d[4] = __fsub_rn(a3, mb3);
}

// The same pass on TWO independent 1-D transforms at once, with packed fp32 (Blackwell add/sub/fma.rn.f32x2: SASS FADD2 /
// FFMA2).  Every lane of a pair sees exactly the scalar operations above in the same order, each rounded once:
//   * add.rn.f32x2 / sub.rn.f32x2 are the two scalar operations;
//   * a product is written fma.rn.f32x2(a, C, -0.0): x + (-0.0) == x for every x including both zeros, so the result is
//     RN(a * C), the scalar multiply.  It is NOT written mul.rn.f32x2: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2
//     into one FFMA2 (single rounding) even though both carry .rn and even with --fmad=false, and it folds a literal
//     -0.0 addend back into a multiply first.  The -0.0 pair therefore comes in as a kernel argument (`nz`), which
//     ptxas cannot see through; an fma result feeding an add cannot be contracted any further.
typedef unsigned long long jb_f2;
__device__ __forceinline__ jb_f2 jb_pack2(float lo, float hi) { jb_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float jb_lo2(jb_f2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ float jb_hi2(jb_f2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return hi; }
__device__ __forceinline__ jb_f2 jb_add2(jb_f2 a, jb_f2 b) { jb_f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ jb_f2 jb_sub2(jb_f2 a, jb_f2 b) { jb_f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ jb_f2 jb_fma2(jb_f2 a, jb_f2 b, jb_f2 c) { jb_f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ jb_f2 jb_mulc2(jb_f2 a, float c, jb_f2 nz) { return jb_fma2(a, jb_pack2(c, c), nz); }

__device__ __forceinline__ void jb_idct8x2(const jb_f2 y[8], jb_f2 d[8], const jb_f2 nz)
{
    const float C_1_175876 = 1.175875602f, C_1_961571 = -1.961570560f, C_0_390181 = -0.390180644f,
                C_0_899976 = -0.899976223f, C_2_562915 = -2.562915447f, C_0_298631 = 0.298631336f,
                C_2_053120 = 2.053119869f, C_3_072711 = 3.072711026f, C_1_501321 = 1.501321110f,
                C_0_541196 = 0.541196100f, C_1_847759 = -1.847759065f, C_0_765367 = 0.765366865f;
    const jb_f2 my1 = y[1], my7 = y[7], my3 = y[3], my5 = y[5];
    jb_f2 mz0 = jb_add2(my1, my7);
    jb_f2 mz2 = jb_add2(my3, my7);
    jb_f2 mz1 = jb_add2(my3, my5);
    jb_f2 mz3 = jb_add2(my1, my5);
    jb_f2 mz4 = jb_mulc2(jb_add2(mz0, mz1), C_1_175876, nz);
    mz2 = jb_add2(jb_mulc2(mz2, C_1_961571, nz), mz4);
    mz3 = jb_add2(jb_mulc2(mz3, C_0_390181, nz), mz4);
    mz0 = jb_mulc2(mz0, C_0_899976, nz);
    mz1 = jb_mulc2(mz1, C_2_562915, nz);
    const jb_f2 mb3 = jb_add2(jb_add2(jb_mulc2(my7, C_0_298631, nz), mz0), mz2);
    const jb_f2 mb2 = jb_add2(jb_add2(jb_mulc2(my5, C_2_053120, nz), mz1), mz3);
    const jb_f2 mb1 = jb_add2(jb_add2(jb_mulc2(my3, C_3_072711, nz), mz1), mz2);
    const jb_f2 mb0 = jb_add2(jb_add2(jb_mulc2(my1, C_1_501321, nz), mz0), mz3);
    const jb_f2 my2 = y[2], my6 = y[6], my0 = y[0], my4 = y[4];
    mz4 = jb_mulc2(jb_add2(my2, my6), C_0_541196, nz);
    mz0 = jb_add2(my0, my4);
    mz1 = jb_sub2(my0, my4);
    mz2 = jb_add2(mz4, jb_mulc2(my6, C_1_847759, nz));
    mz3 = jb_add2(mz4, jb_mulc2(my2, C_0_765367, nz));
    const jb_f2 a0 = jb_add2(mz0, mz3);
    const jb_f2 a3 = jb_sub2(mz0, mz3);
    const jb_f2 a1 = jb_add2(mz1, mz2);
    const jb_f2 a2 = jb_sub2(mz1, mz2);
    d[0] = jb_add2(a0, mb0);
    d[7] = jb_sub2(a0, mb0);
    d[1] = jb_add2(a1, mb1);
    d[6] = jb_sub2(a1, mb1);
    d[2] = jb_add2(a2, mb2);
    d[5] = jb_sub2(a2, mb2);
    d[3] = jb_add2(a3, mb3);
    d[4] = jb_sub2(a3, mb3);
}

__device__ __forceinline__ int jb_clamp255(int v) { return min(max(v, 0), 255); }

// One sample of a P-bit frame as the 8-bit value the reference's application writers store:
//   P == 8: clamp                                   (apps/JpegDecode/JpegBufferOutputWriter8Bit.cs:28-60)
//   P  > 8: clamp(sample >> (P - 8))                (JpegBufferOutputWriterGreaterThan8Bit.cs:34-68)
//   P  < 8: clamp to [0, 2^P - 1], then repeat the P-bit pattern up to 8 bits; when P does not divide 8 the
//           last partial copy takes the pattern's LOW bits (JpegBufferOutputWriterLessThan8Bit.cs:35-92)
__device__ __forceinline__ int jb_sample_to_u8(int v, int precision)
{
    if (precision >= 8) return jb_clamp255(v >> (precision - 8));
    uint32_t bits = (uint32_t)min(max(v, 0), (1 << precision) - 1);
    int have = precision;
    while (have < 8) { bits |= bits << precision; have += precision; }
    if (have > 8) {
        bits >>= precision; have -= precision;
        const int rem = 8 - have;
        bits = (bits << rem) | (bits & ((1u << rem) - 1u));
    }
    return (int)(bits & 255u);
}

// apps/JpegDecode/JpegYCbCrToRgbConverter.cs:93-121 evaluated in fp32 as the C# does:
// d1 = Fix(2-2*0.299) = 91881, d2 = -Fix(0.299*f1/0.587) = -46802, d3 = Fix(2-2*0.114) = 116130,
// d4 = -Fix(0.114*f3/0.587) = -22553; Code2V is the identity for the default reference black/white.
__device__ __forceinline__ void jb_ycc_to_rgb(int y, int cb, int cr, int &r, int &g, int &b)
{
    const int cbv = cb - 128, crv = cr - 128;
    r = jb_clamp255(y + ((91881 * crv + 32768) >> 16));
    g = jb_clamp255(y + ((-22553 * cbv + 32768 + -46802 * crv) >> 16));
    b = jb_clamp255(y + ((116130 * cbv + 32768) >> 16));
}

struct JbK2Geom {
    int plane_off[4]; // sample offset of each component plane in s_plane
    int plane_pitch[4];
    int hshift[4], vshift[4];
};

__global__ void __launch_bounds__(JB_K2_THREADS)
jb_k2_idct_color(const JbDevImage *__restrict__ images,
                 const int16_t *__restrict__ coef, const uint16_t *__restrict__ quant,
                 const uint32_t *__restrict__ image_list, const uint32_t *__restrict__ mcu_limit,
                 const uint32_t *__restrict__ comp_limit)
{
    __shared__ __align__(16) float s_f[JB_K2_MAX_BLOCKS * JB_K2_BLOCK_STRIDE];
    __shared__ __align__(16) int16_t s_plane[JB_K2_MAX_BLOCKS * 64];
    __shared__ uint16_t s_q[4 * 64];
    __shared__ JbDevImage s_im;
    __shared__ JbK2Geom s_g;

    // grid = (tiles per image, images); a tile is a strip of consecutive MCUs of one MCU row
    JbTileWork tw;
    tw.image = image_list ? image_list[blockIdx.y] : blockIdx.y;
    const int tid = threadIdx.x;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(images + tw.image);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_im);
        for (int i = tid; i < (int)(sizeof(JbDevImage) / 4); i += JB_K2_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const int ncomp = s_im.ncomp, bpm = s_im.bpm;
    {
        const uint32_t tile_mcus = JB_K2_MAX_BLOCKS / bpm;
        const uint32_t strips = (s_im.mcus_per_line + tile_mcus - 1) / tile_mcus;
        tw.mcu_row = blockIdx.x / strips;
        tw.mcu_col0 = (blockIdx.x - tw.mcu_row * strips) * tile_mcus;
        tw.nmcu = min(tile_mcus, s_im.mcus_per_line - tw.mcu_col0);
        if (tw.mcu_row >= s_im.mcus_per_col) return;
        // MCUs that were never decoded (scan ended at an EOI on a restart boundary) are not written (see K0b)
        const uint32_t limit = mcu_limit ? mcu_limit[tw.image] : 0xFFFFFFFFu, mcu0 = tw.mcu_row * s_im.mcus_per_line + tw.mcu_col0;
        if (mcu0 >= limit) return;
        tw.nmcu = min(tw.nmcu, limit - mcu0);
    }
    const int nmcu = tw.nmcu;
    if (tid < ncomp * 64) s_q[tid] = quant[s_im.quant_off + tid];
    if (tid == 0) {
        int off = 0;
        for (int c = 0; c < ncomp; c++) {
            s_g.plane_off[c] = off;
            s_g.plane_pitch[c] = nmcu * s_im.comp_h[c] * 8;
            off += nmcu * s_im.comp_h[c] * s_im.comp_v[c] * 64;
            int hs = s_im.hmax / s_im.comp_h[c], vs = s_im.vmax / s_im.comp_v[c];
            s_g.hshift[c] = hs == 4 ? 2 : hs == 2 ? 1 : 0;
            s_g.vshift[c] = vs == 4 ? 2 : vs == 2 ? 1 : 0;
        }
    }
    __syncthreads();

    // ------------------------------------------------------------ phase A: dequant + IDCT
    const int nblk = nmcu * bpm;
    const int j = tid >> 3, r = tid & 7;
    const unsigned wm = __ballot_sync(0xFFFFFFFFu, j < nblk); // the 8 threads of a block share a warp
    if (j < nblk) {
        const int m = j / bpm, b = j - m * bpm;
        const int c = s_im.blk_comp[b];
        // scan-list frames: a component no scan names is never written by the reference (its samples stay 0), and
        // neither are the MCUs behind the point where the component's scans ended at an EOI on a restart boundary
        const bool unwritten = s_im.planar && (!((s_im.covered >> c) & 1u) ||
                                               (comp_limit && s_im.seq_dri &&
                                                tw.mcu_row * s_im.mcus_per_line + tw.mcu_col0 + (uint32_t)m >= comp_limit[tw.image * 4 + c]));
        uint64_t blk;
        if (!s_im.planar) {
            blk = s_im.coef_off + ((uint64_t)tw.mcu_row * s_im.mcus_per_line + tw.mcu_col0) * bpm + j;
        } else { // progressive: per-component planes of MCU-padded block grids
            const int bi0 = b - s_im.comp_blk_off[c], hc0 = s_im.comp_h[c];
            blk = s_im.coef_off + s_im.comp_plane_off[c] +
                  (uint64_t)(tw.mcu_row * s_im.comp_v[c] + bi0 / hc0) * s_im.comp_plane_w[c] +
                  (tw.mcu_col0 + m) * hc0 + bi0 % hc0;
        }
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(coef + blk * 64) + r);
        const uint32_t wv[4] = {raw.x, raw.y, raw.z, raw.w};
        float *fb = s_f + j * JB_K2_BLOCK_STRIDE;
        const uint16_t *q = s_q + c * 64;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int z = r * 8 + e;
            const int cv = (int)(int16_t)((wv[e >> 1] >> ((e & 1) * 16)) & 0xFFFF);
            // product in int32, then int -> float (JpegScanDecoder.cs:60)
            fb[jb_c_zigzag[z]] = __int2float_rn((int)q[z] * cv);
        }
        __syncwarp(wm);
        // pass 1: 1-D IDCT along row r of the block (src.TransposeInto(temp); IDCT8x4 on temp)
        float y[8], d[8];
        {
            const float4 lo = *reinterpret_cast<const float4 *>(fb + r * 8);
            const float4 hi = *reinterpret_cast<const float4 *>(fb + r * 8 + 4);
            y[0] = lo.x; y[1] = lo.y; y[2] = lo.z; y[3] = lo.w;
            y[4] = hi.x; y[5] = hi.y; y[6] = hi.z; y[7] = hi.w;
        }
        jb_idct8(y, d);
        __syncwarp(wm);
#pragma unroll
        for (int k = 0; k < 8; k++) fb[k * 8 + r] = d[k]; // O1[k][r]
        __syncwarp(wm);
        // pass 2: for column k = r: 1-D IDCT over O1[k][0..7]
        {
            const float4 lo = *reinterpret_cast<const float4 *>(fb + r * 8);
            const float4 hi = *reinterpret_cast<const float4 *>(fb + r * 8 + 4);
            y[0] = lo.x; y[1] = lo.y; y[2] = lo.z; y[3] = lo.w;
            y[4] = hi.x; y[5] = hi.y; y[6] = hi.z; y[7] = hi.w;
        }
        jb_idct8(y, d);
        // scale, round half-to-even, level shift (MultiplyInplace(0.125) + ShiftDataLevel)
        const int shift = 1 << (s_im.precision - 1);
        const int bi = b - s_im.comp_blk_off[c];
        const int hc = s_im.comp_h[c];
        const int bx = m * hc + bi % hc, by = bi / hc;
        int16_t *pl = s_plane + s_g.plane_off[c] + (by * 8) * s_g.plane_pitch[c] + bx * 8 + r;
#pragma unroll
        for (int mrow = 0; mrow < 8; mrow++) {
            const int v = __float2int_rn(__fmul_rn(d[mrow], 0.125f)) + shift;
            pl[mrow * s_g.plane_pitch[c]] = unwritten ? (int16_t)0 : (int16_t)v;
        }
    }
    __syncthreads();

    // ------------------------------------------------------------ phase B: replicate + colour + store
    const int W = s_im.width, H = s_im.height;
    const int TW = nmcu * 8 * s_im.hmax, TH = 8 * s_im.vmax;
    const int x0 = tw.mcu_col0 * 8 * s_im.hmax, y0 = tw.mcu_row * 8 * s_im.vmax;
    const int fmt = s_im.out_format;
    const int pshift = s_im.precision > 8 ? s_im.precision - 8 : 0;
    uint8_t *out = reinterpret_cast<uint8_t *>(s_im.out_ptr);
    const uint64_t pitch = s_im.out_pitch;
    const int groups_per_row = TW >> 2;
    const int ngroups = groups_per_row * TH;

    for (int g = tid; g < ngroups; g += JB_K2_THREADS) {
        const int gy = g / groups_per_row, gx = (g - gy * groups_per_row) << 2;
        const int x = x0 + gx, yy = y0 + gy;
        if (yy >= H || x >= W) continue;
        const int npx = min(4, W - x);
        int s[3][4];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (c < ncomp) {
                const int16_t *row = s_plane + s_g.plane_off[c] + (gy >> s_g.vshift[c]) * s_g.plane_pitch[c];
#pragma unroll
                for (int p = 0; p < 4; p++) s[c][p] = row[(gx + p) >> s_g.hshift[c]];
            } else {
#pragma unroll
                for (int p = 0; p < 4; p++) s[c][p] = 128 << pshift;
            }
        }
        if (fmt == 3 /* JB_OUT_PLANAR_I16 */) {
            for (int c = 0; c < ncomp; c++) {
                int16_t *dst = reinterpret_cast<int16_t *>(out + ((uint64_t)c * H + yy) * pitch) + x;
                if (c < 3) {
                    for (int p = 0; p < npx; p++) dst[p] = (int16_t)s[c][p];
                } else {
                    const int16_t *row = s_plane + s_g.plane_off[c] + (gy >> s_g.vshift[c]) * s_g.plane_pitch[c];
                    for (int p = 0; p < npx; p++) dst[p] = row[(gx + p) >> s_g.hshift[c]];
                }
            }
            continue;
        }
        uint8_t px[16];
        const int bpp = fmt == 1 ? 4 : 3;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int yv = jb_sample_to_u8(s[0][p], s_im.precision);
            const int cb = ncomp > 1 ? jb_sample_to_u8(s[1][p], s_im.precision) : 128;
            const int cr = ncomp > 2 ? jb_sample_to_u8(s[2][p], s_im.precision) : 128;
            if (fmt == 2 /* JB_OUT_YCBCR888 */) {
                px[p * 3] = (uint8_t)yv; px[p * 3 + 1] = (uint8_t)cb; px[p * 3 + 2] = (uint8_t)cr;
            } else {
                int rr, gg, bb;
                jb_ycc_to_rgb(yv, cb, cr, rr, gg, bb);
                px[p * bpp] = (uint8_t)rr; px[p * bpp + 1] = (uint8_t)gg; px[p * bpp + 2] = (uint8_t)bb;
                if (bpp == 4) px[p * 4 + 3] = 255;
            }
        }
        uint8_t *dst = out + (uint64_t)yy * pitch + (uint64_t)x * bpp;
        const int nbytes = npx * bpp;
        if (npx == 4 && ((reinterpret_cast<uint64_t>(dst) & 3u) == 0)) {
            uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (i * 4 < nbytes)
                    d32[i] = px[i * 4] | (px[i * 4 + 1] << 8) | (px[i * 4 + 2] << 16) | ((uint32_t)px[i * 4 + 3] << 24);
            }
        } else {
            for (int i = 0; i < nbytes; i++) dst[i] = px[i];
        }
    }
}
