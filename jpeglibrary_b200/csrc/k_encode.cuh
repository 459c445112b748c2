// k_encode.cuh -- encoder kernels (north_star item (4)), batched over images.
//
//   K3  jb_k3_fdct_quant      RGB->YCbCr (E1), zero-padded block read + box downsample (E2, E3), level shift
//                             + fp32 FDCT (E4), divide by the quantiser + round half-to-even + zig-zag (E5)
//                             -> coefficient store in MCU scan order
//   K3b jb_k3b_histogram      DC-difference / run-size symbol histogram per (table class, id) (E6)
//   K3c jb_k3c_build_tables   optimised table construction, one thread per table: Annex K.1-K.3 with the
//                             reference's tie rules + the runtime's introsort (E7); also callable on the host
//   K4a jb_k4a_block_bits     bits each block will occupy
//   K4b jb_k4b_scan           per-image exclusive prefix sum -> bit offset of every block
//   K4c jb_k4c_pack           parallel bit packing into an un-stuffed stream (E8)
//   K4d jb_k4d_stuff          FF -> FF 00 byte stuffing + 1-bit padding of the last byte (E9)
//
// Reference: JpegEncoder.TransformBlocks/BuildHuffmanTables/WritePreparedScanData (JpegEncoder.cs:414-656),
// EncodeBlock/EncodeRunLength (:828-918), FastFloatingPointDCT.TransformFDCT (FastFloatingPointDCT.cs:195-362),
// JpegHuffmanEncodingTableBuilder (:69-282), JpegWriter bit mode (JpegWriter.cs:104-227),
// apps/JpegEncode/JpegRgbToYCbCrConverter.cs:26-93, apps/JpegEncode/JpegBufferInputReader.cs:26-50.
#pragma once
#include "jb_device.cuh"
#include "k_idct_color_fast.cuh" // jb_c_nat2zz
#include "k_idct_color_warp.cuh" // jb_nat2zz_c

struct JbEncImage {
    uint64_t pix_ptr;   // device address of the input pixels
    uint64_t pix_pitch;
    uint64_t coef_off;  // first block in the coefficient store
    uint64_t bits_off;  // first entry in the per-block bit arrays
    uint64_t raw_off;   // byte offset of the un-stuffed stream
    uint64_t raw_cap;   // bytes reserved for it
    uint64_t out_off;   // byte offset of the stuffed scan bytes
    uint64_t out_cap;
    uint32_t quant_off; // ncomp natural-order... (zig-zag uint16[64] per component)
    uint32_t total_mcus, mcus_per_line, mcus_per_col;
    uint32_t table_base; // first of this image's 8 encoder tables: [class * 4 + id]
    uint16_t width, height;
    uint8_t ncomp, bpm, hs, vs, in_format; // in_format: 0 RGB24, 1 YCbCr888, 2 grey8
    uint8_t comp_td[4], comp_ta[4];
    uint8_t blk_comp[JB_MAX_BLOCKS_PER_MCU];
    uint8_t pad[5];
    uint32_t dri;       // transcoding only: MCUs per restart interval (0: none)
    uint32_t nint;      // number of restart intervals (1 without DRI)
};

struct JbEncTable {     // one optimised Huffman table
    uint16_t code[256];
    uint8_t len[256];
    uint8_t bits[16];   // DHT counts
    uint8_t vals[256];  // DHT symbols
    uint32_t nvals;
    uint32_t pad[3];
};

// One 1-D pass of FastFloatingPointDCT.FDCT8x4_{Left,Right}Part (FastFloatingPointDCT.cs:195-314)
__device__ __forceinline__ void jb_fdct8(const float s[8], float d[8])
{
    const float t0 = __fadd_rn(s[0], s[7]), t7 = __fsub_rn(s[0], s[7]);
    const float t1 = __fadd_rn(s[1], s[6]), t6 = __fsub_rn(s[1], s[6]);
    const float t2 = __fadd_rn(s[2], s[5]), t5 = __fsub_rn(s[2], s[5]);
    const float t3 = __fadd_rn(s[3], s[4]), t4 = __fsub_rn(s[3], s[4]);
    float c0 = __fadd_rn(t0, t3), c3 = __fsub_rn(t0, t3);
    float c1 = __fadd_rn(t1, t2), c2 = __fsub_rn(t1, t2);
    d[0] = __fadd_rn(c0, c1);
    d[4] = __fsub_rn(c0, c1);
    d[2] = __fadd_rn(__fmul_rn(0.541196f, c2), __fmul_rn(1.306563f, c3));
    d[6] = __fsub_rn(__fmul_rn(0.541196f, c3), __fmul_rn(1.306563f, c2));
    c3 = __fadd_rn(__fmul_rn(1.175876f, t4), __fmul_rn(0.785695f, t7));
    c0 = __fsub_rn(__fmul_rn(1.175876f, t7), __fmul_rn(0.785695f, t4));
    c2 = __fadd_rn(__fmul_rn(1.387040f, t5), __fmul_rn(0.275899f, t6));
    c1 = __fsub_rn(__fmul_rn(1.387040f, t6), __fmul_rn(0.275899f, t5));
    d[3] = __fsub_rn(c0, c2);
    d[5] = __fsub_rn(c3, c1);
    c0 = __fmul_rn(__fadd_rn(c0, c2), 0.707107f);
    c3 = __fmul_rn(__fadd_rn(c3, c1), 0.707107f);
    d[1] = __fadd_rn(c0, c3);
    d[7] = __fsub_rn(c0, c3);
}


// apps/JpegEncode/JpegRgbToYCbCrConverter.cs:40-56 evaluated in fp32 like the C#:
// Fix(0.299)=19595 Fix(0.587)=38470 Fix(0.114)=7471 Fix(0.168735892)=11058 Fix(0.331264108)=21710
// Fix(0.5)=32768 Fix(0.418687589)=27439 Fix(0.081312411)=5329
__device__ __forceinline__ void jb_rgb_to_ycc(int r, int g, int b, int &y, int &cb, int &cr)
{
    y = (19595 * r + 38470 * g + (7471 * b + 32768)) >> 16;
    cb = (-11058 * r + -21710 * g + (32768 * b + (128 << 16) + 32767)) >> 16;
    cr = ((32768 * r + (128 << 16) + 32767) + -27439 * g + -5329 * b) >> 16;
    y &= 0xFF; cb &= 0xFF; cr &= 0xFF; // (byte) casts
}

// ---------------------------------------------------------------------------------------------
// K3 (second generation; the first one used 8 threads per block with CTA-wide phases): warp-autonomous, ONE THREAD PER 8x8
// BLOCK like the decoder's K2 (k_idct_color_warp.cuh).  A warp owns a unit of 32/BPM MCUs: it converts the unit's
// pixels into component planes in shared memory (4 pixels = three 32-bit loads per lane), then every lane gathers
// its block (box filter for sub-sampled chroma), runs both FDCT passes in registers, quantises with true fp32
// division and writes its 128-byte zig-zag block with eight 128-bit stores.  No transposes, no CTA barriers.
// ---------------------------------------------------------------------------------------------
#define JB_K3W_WARPS 4

template <int K>
__device__ __forceinline__ void jb_k3w_col(const float (&p1)[64], float (&F)[64])
{
    float y[8], d[8];
#pragma unroll
    for (int e = 0; e < 8; e++) y[e] = p1[e * 8 + K];
    jb_fdct8(y, d); // pass 2: d[mr] = F[vertical mr][horizontal K]
#pragma unroll
    for (int mr = 0; mr < 8; mr++) F[mr * 8 + K] = d[mr];
}

template <int NC, int HS, int VS>
#ifndef JB_K3_MIN_CTAS
#define JB_K3_MIN_CTAS 6 // 80 registers: 24 warps per SM hide the shared-memory and load latency better than 16 (A/B: 42.6 vs 43.4 ms per 512 frames)
#endif
__global__ void __launch_bounds__(JB_K3W_WARPS * 32, JB_K3_MIN_CTAS)
jb_k3_fdct_quant_warp(const JbEncImage *__restrict__ images, const uint32_t *__restrict__ image_list,
                      const uint16_t *__restrict__ quant, int16_t *__restrict__ coef, int units_per_warp)
{
    constexpr int BPM = NC == 1 ? 1 : HS * VS + 2;
    constexpr int UM = NC == 1 ? 16 : 32 / BPM; // MCUs per unit
    constexpr int NB = UM * BPM;
    constexpr int TW = UM * 8 * HS, TH = 8 * VS;
    __shared__ __align__(16) uint8_t s_c[JB_K3W_WARPS][NC][TH][TW]; // component planes at full resolution, 0 outside the image
    // 1 / (8 q) in fp64, NATURAL order.  ZigZagAndQuantizeBlock divides the fp32 coefficient (after the exact
    // MultiplyInplace(0.125)) by the quantiser in fp32 (JpegEncoder.cs:812-826).  RN32(F / (8q)) is obtained here as
    // RN32(RN64((double)F * RN64(1 / (8q)))): for a 24-bit F and an integer q < 2^16 the exact quotient is never an
    // fp32 rounding tie (q's odd part would need a 25-bit multiple to fit 24 bits) and lies at least 2^-41 (relative)
    // from the nearest fp32 rounding boundary unless it is exactly representable, while the fp64 product is within
    // 2^-52 of it -- so the double rounding is innocuous.  Checked on 9.2e8 random and adversarial (F, q) pairs against
    // IEEE fp32 division (0 mismatches); three instructions instead of the ~12-instruction div.rn sequence with its
    // slow-path branch per coefficient.
    __shared__ __align__(16) double s_rq[NC * 64];
    __shared__ JbEncImage s_im;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t image = image_list[blockIdx.y];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(images + image);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_im);
        for (int i = tid; i < (int)(sizeof(JbEncImage) / 4); i += JB_K3W_WARPS * 32) dst[i] = src[i];
    }
    __syncthreads();
    for (int i = tid; i < NC * 64; i += JB_K3W_WARPS * 32)
        s_rq[i] = 0.125 / (double)quant[s_im.quant_off + (i >> 6) * 64 + jb_c_nat2zz[i & 63]];
    __syncthreads();

    const int j = lane;
    const int m = j / BPM, b = j - m * BPM;
    int c = 0, bx, by;
    if (NC == 1 || b < HS * VS) { bx = m * HS + (b % HS); by = b / HS; }
    else { c = b - HS * VS + 1; bx = m; by = 0; }
    const double *rq = s_rq + c * 64;
    const int W = s_im.width, H = s_im.height;
    const uint32_t mpl = s_im.mcus_per_line;
    const uint32_t upr = (mpl + UM - 1) / UM, nunits = upr * s_im.mcus_per_col;
    const uint8_t *pix = reinterpret_cast<const uint8_t *>(s_im.pix_ptr);
    const uint64_t pitch = s_im.pix_pitch;
    const int fmt = s_im.in_format;
    const bool vec_ok = fmt != 2 && ((s_im.pix_ptr | pitch) & 3u) == 0; // 4 pixels = 12 bytes as three aligned words
    uint8_t (*pl)[TH][TW] = s_c[wid];

    uint32_t unit = (blockIdx.x * JB_K3W_WARPS + wid) * (uint32_t)units_per_warp;
    const uint32_t unit_end = min(unit + (uint32_t)units_per_warp, nunits);
    for (; unit < unit_end; unit++) {
        const uint32_t mcu_row = unit / upr, ucol = unit - mcu_row * upr;
        const uint32_t mcu_col0 = ucol * UM;
        const int nmcu = (int)min((uint32_t)UM, mpl - mcu_col0);
        const int x0 = mcu_col0 * 8 * HS, y0 = mcu_row * TH;
        const int wlim = min(W - x0, nmcu * 8 * HS); // pixels of this unit's rows that exist

        // ---- E1 + E2: pixels -> component planes (JpegBufferInputReader zero-fills outside the image).
        // All loads of the unit are issued before the first conversion so that their latencies overlap.
        __syncwarp();
        constexpr int NG = (TW * TH / 4 + 31) / 32; // 4-pixel groups per lane
        uint32_t lw[NG][3];
#pragma unroll
        for (int t = 0; t < NG; t++) {
            const int g = lane + 32 * t;
            const int py = g / (TW / 4), px = (g - py * (TW / 4)) * 4;
            lw[t][0] = lw[t][1] = lw[t][2] = 0;
            if (g < TW * TH / 4 && y0 + py < H && vec_ok && px + 4 <= wlim) {
                const uint32_t *p = reinterpret_cast<const uint32_t *>(pix + (uint64_t)(y0 + py) * pitch + (uint64_t)(x0 + px) * 3);
                lw[t][0] = __ldg(p); lw[t][1] = __ldg(p + 1); lw[t][2] = __ldg(p + 2);
            }
        }
#pragma unroll
        for (int t = 0; t < NG; t++) {
            const int g = lane + 32 * t;
            if (g >= TW * TH / 4) break;
            const int py = g / (TW / 4), px = (g - py * (TW / 4)) * 4;
            const int y = y0 + py;
            uint32_t o0 = 0, o1 = 0, o2 = 0; // four samples of each component
            if (y < H && px < wlim) {
                if (vec_ok && px + 4 <= wlim) {
                    const uint32_t w0 = lw[t][0], w1 = lw[t][1], w2 = lw[t][2];
                    // bytes: a0 b0 c0 a1 | b1 c1 a2 b2 | c2 a3 b3 c3
                    const uint32_t a4 = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);  // a0 a1 a2 a3
                    const uint32_t b4 = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);  // b0 b1 b2 b3
                    const uint32_t c4 = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);  // c0 c1 c2 c3
                    if (fmt == 1) { o0 = a4; o1 = b4; o2 = c4; }
                    else {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            int yy, cb, cr;
                            jb_rgb_to_ycc((a4 >> (8 * i)) & 0xFF, (b4 >> (8 * i)) & 0xFF, (c4 >> (8 * i)) & 0xFF, yy, cb, cr);
                            o0 |= (uint32_t)yy << (8 * i); o1 |= (uint32_t)cb << (8 * i); o2 |= (uint32_t)cr << (8 * i);
                        }
                    }
                } else {
                    const uint8_t *p = pix + (uint64_t)y * pitch + (uint64_t)(x0 + px) * (fmt == 2 ? 1 : 3);
#pragma unroll 1
                    for (int i = 0; i < 4; i++) {
                        if (px + i >= wlim) break;
                        const uint8_t *q = p + i * (fmt == 2 ? 1 : 3);
                        int yy = 0, cb = 0, cr = 0;
                        if (fmt == 0) jb_rgb_to_ycc(q[0], q[1], q[2], yy, cb, cr);
                        else if (fmt == 1) { yy = q[0]; cb = q[1]; cr = q[2]; }
                        else yy = q[0];
                        o0 |= (uint32_t)yy << (8 * i); o1 |= (uint32_t)cb << (8 * i); o2 |= (uint32_t)cr << (8 * i);
                    }
                }
            }
            *reinterpret_cast<uint32_t *>(&pl[0][py][px]) = o0;
            if (NC == 3) {
                *reinterpret_cast<uint32_t *>(&pl[1][py][px]) = o1;
                *reinterpret_cast<uint32_t *>(&pl[2][py][px]) = o2;
            }
        }
        __syncwarp();

        // ---- E3 + E4 + E5: one block per lane
        if (j < NB && m < nmcu) {
            float p1[64];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                float s8[8], d8[8];
                if (c == 0) {
                    const uint2 v = *reinterpret_cast<const uint2 *>(&pl[0][by * 8 + r][bx * 8]);
#pragma unroll
                    for (int e = 0; e < 8; e++) // (float)(sample - 128) through the fp32 magic number (no XU conversion)
                        s8[e] = __fsub_rn(__int_as_float(0x4B400000 + (int)(((e < 4 ? v.x : v.y) >> (8 * (e & 3))) & 0xFF)), 12582912.0f + 128.0f);
                } else {
                    // box filter: sum of HS x VS samples, (sum + delta) >> shift (JpegEncoder.cs:777-785), on 16-bit
                    // SIMD lanes: bytes of HS*8 consecutive samples of VS rows
                    constexpr int SH = (HS == 2 ? 1 : 0) + (VS == 2 ? 1 : 0);
                    uint32_t acc[4] = {0, 0, 0, 0}; // eight 16-bit sums: acc[i] = sum[2i] | sum[2i+1] << 16
#pragma unroll
                    for (int dy = 0; dy < VS; dy++) {
                        const uint8_t *row = &pl[NC == 3 ? c : 0][r * VS + dy][bx * 8 * HS];
                        if (HS == 2) {
                            const uint4 v = *reinterpret_cast<const uint4 *>(row);
                            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int i = 0; i < 4; i++) acc[i] += (w[i] & 0x00FF00FFu) + ((w[i] >> 8) & 0x00FF00FFu); // pairs
                        } else {
                            const uint2 v = *reinterpret_cast<const uint2 *>(row);
                            acc[0] += __byte_perm(v.x, 0, 0x4140); acc[1] += __byte_perm(v.x, 0, 0x4342);
                            acc[2] += __byte_perm(v.y, 0, 0x4140); acc[3] += __byte_perm(v.y, 0, 0x4342);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        int sum = (int)((acc[e >> 1] >> (16 * (e & 1))) & 0xFFFFu);
                        if constexpr (SH > 0) sum = (sum + (1 << (SH - 1))) >> SH;
                        s8[e] = __fsub_rn(__int_as_float(0x4B400000 + sum), 12582912.0f + 128.0f);
                    }
                }
                jb_fdct8(s8, d8); // pass 1: along row r
#pragma unroll
                for (int k = 0; k < 8; k++) p1[r * 8 + k] = d8[k];
            }
            float F[64];
            jb_k3w_col<0>(p1, F); jb_k3w_col<1>(p1, F); jb_k3w_col<2>(p1, F); jb_k3w_col<3>(p1, F);
            jb_k3w_col<4>(p1, F); jb_k3w_col<5>(p1, F); jb_k3w_col<6>(p1, F); jb_k3w_col<7>(p1, F);
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; i++) pk[i] = 0;
#pragma unroll
            for (int n = 0; n < 64; n++) {
                // MultiplyInplace(0.125), coefficient / element, MathF.Round -> (short)  (JpegEncoder.cs:812-826)
                const float v = (float)((double)F[n] * rq[n]);
                // MathF.Round (half to even) through the fp32 magic number: exact for |v| < 2^22
                const uint32_t q16 = (uint32_t)(__float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000) & 0xFFFFu;
                const int z = jb_nat2zz_c(n);
                pk[z >> 1] |= q16 << (16 * (z & 1));
            }
            const uint64_t blk = s_im.coef_off + ((uint64_t)mcu_row * mpl + mcu_col0) * BPM + j;
            uint4 *dst = reinterpret_cast<uint4 *>(coef + blk * 64);
#pragma unroll
            for (int i = 0; i < 8; i++) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
    }
}

__device__ __forceinline__ int jb_bit_count(int a) { return 32 - __clz(a); } // BitCountTable (:938-953)

// index of the previous block of the same component in scan order, or -1
__device__ __forceinline__ int64_t jb_prev_block(const JbEncImage &im, uint32_t blk)
{
    const uint32_t mcu = blk / im.bpm, b = blk - mcu * im.bpm;
    const int c = im.blk_comp[b];
    if (b > 0 && im.blk_comp[b - 1] == c) return (int64_t)blk - 1;
    if (mcu == 0) return -1;
    if (im.dri != 0 && mcu % im.dri == 0) return -1; // DC prediction restarts with every restart interval
    int last = b; // last block of component c inside an MCU
    while (last + 1 < im.bpm && im.blk_comp[last + 1] == c) last++;
    return (int64_t)(mcu - 1) * im.bpm + last;
}

// ---- E6: symbol histogram (GatherBlockStatistics :551-597)
__global__ void __launch_bounds__(256)
jb_k3b_histogram(const JbEncImage *__restrict__ images, const int16_t *__restrict__ coef, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_h[8 * 256];
    const JbEncImage &im = images[blockIdx.y];
    const uint32_t total = im.total_mcus * im.bpm;
    if (blockIdx.x * 256 >= total) return;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) s_h[i] = 0;
    __syncthreads();
    const uint32_t blk = blockIdx.x * 256 + threadIdx.x;
    if (blk < total) {
        const int16_t *p = coef + (im.coef_off + blk) * 64;
        const int c = im.blk_comp[blk % im.bpm];
        const int64_t pb = jb_prev_block(im, blk);
        const int pred = pb < 0 ? 0 : coef[(im.coef_off + pb) * 64];
        uint32_t *hd = s_h + im.comp_td[c] * 256, *ha = s_h + (4 + im.comp_ta[c]) * 256;
        int t = p[0] - pred;
        atomicAdd(hd + jb_bit_count(abs(t)), 1u);
        int run = 0;
#pragma unroll 1
        for (int i8 = 0; i8 < 8; i8++) {
            const uint4 raw = reinterpret_cast<const uint4 *>(p)[i8];
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int e = 0; e < 8; e++) {
                if (i8 == 0 && e == 0) continue;
                t = (int)(int16_t)((w[e >> 1] >> ((e & 1) * 16)) & 0xFFFF);
                if (t == 0) { run++; continue; }
                while (run > 15) { atomicAdd(ha + 0xF0, 1u); run -= 16; }
                atomicAdd(ha + ((run << 4) | jb_bit_count(abs(t))), 1u);
                run = 0;
            }
        }
        if (run > 0) atomicAdd(ha, 1u);
    }
    __syncthreads();
    uint32_t *g = hist + (uint64_t)blockIdx.y * 8 * 256;
    for (int i = threadIdx.x; i < 8 * 256; i += 256)
        if (s_h[i]) atomicAdd(g + i, s_h[i]);
}

// ---- E7: JpegHuffmanEncodingTableBuilder.BuildUsingStandardMethod (:69-176) + BuildCanonicalCode (:240-282).
// Runs as one GPU thread per table, and on the host for callers that bring their own histograms.
struct JbHSym { long long freq; short value; unsigned short code_size; short others; };

__host__ __device__ inline int jb_hs_cmp(const JbHSym &a, const JbHSym &b) { return (a.code_size > b.code_size) - (a.code_size < b.code_size); }
__host__ __device__ inline void jb_hs_swap(JbHSym *k, int i, int j) { if (i != j) { JbHSym t = k[i]; k[i] = k[j]; k[j] = t; } }
__host__ __device__ inline void jb_hs_swap_if_greater(JbHSym *k, int i, int j) { if (i != j && jb_hs_cmp(k[i], k[j]) > 0) jb_hs_swap(k, i, j); }

// .NET Core ArraySortHelper<T>.IntroSort(Comparison<T>): unstable, decides the symbol order inside a code length
__host__ __device__ inline void jb_hs_introsort(JbHSym *keys, int n0)
{
    int depth0 = 0;
    for (int t = n0; t > 0; t >>= 1) depth0++;
    depth0 *= 2;
    // explicit stack instead of recursion (right part pushed, loop on the left part)
    int stack_lo[64], stack_n[64], stack_d[64], sp = 0;
    stack_lo[0] = 0; stack_n[0] = n0; stack_d[0] = depth0; sp = 1;
    while (sp > 0) {
        sp--;
        JbHSym *k = keys + stack_lo[sp];
        int n = stack_n[sp], depth = stack_d[sp];
        const int lo0 = stack_lo[sp];
        while (n > 1) {
            if (n <= 16) {
                if (n == 2) { jb_hs_swap_if_greater(k, 0, 1); break; }
                if (n == 3) { jb_hs_swap_if_greater(k, 0, 1); jb_hs_swap_if_greater(k, 0, 2); jb_hs_swap_if_greater(k, 1, 2); break; }
                for (int i = 0; i < n - 1; i++) { // insertion sort
                    JbHSym t = k[i + 1];
                    int jj = i;
                    while (jj >= 0 && jb_hs_cmp(t, k[jj]) < 0) { k[jj + 1] = k[jj]; jj--; }
                    k[jj + 1] = t;
                }
                break;
            }
            if (depth == 0) { // heap sort
                for (int i = n / 2; i >= 1; i--) {
                    int ii = i; JbHSym d = k[ii - 1];
                    while (ii <= n / 2) { int ch = 2 * ii; if (ch < n && jb_hs_cmp(k[ch - 1], k[ch]) < 0) ch++; if (!(jb_hs_cmp(d, k[ch - 1]) < 0)) break; k[ii - 1] = k[ch - 1]; ii = ch; }
                    k[ii - 1] = d;
                }
                for (int i = n; i > 1; i--) {
                    jb_hs_swap(k, 0, i - 1);
                    int ii = 1, nn = i - 1; JbHSym d = k[0];
                    while (ii <= nn / 2) { int ch = 2 * ii; if (ch < nn && jb_hs_cmp(k[ch - 1], k[ch]) < 0) ch++; if (!(jb_hs_cmp(d, k[ch - 1]) < 0)) break; k[ii - 1] = k[ch - 1]; ii = ch; }
                    k[ii - 1] = d;
                }
                break;
            }
            depth--;
            const int hi = n - 1, mid = hi >> 1;
            jb_hs_swap_if_greater(k, 0, mid);
            jb_hs_swap_if_greater(k, 0, hi);
            jb_hs_swap_if_greater(k, mid, hi);
            const JbHSym pivot = k[mid];
            jb_hs_swap(k, mid, hi - 1);
            int left = 0, right = hi - 1;
            while (left < right) {
                while (jb_hs_cmp(k[++left], pivot) < 0) ;
                while (jb_hs_cmp(pivot, k[--right]) < 0) ;
                if (left >= right) break;
                jb_hs_swap(k, left, right);
            }
            if (left != hi - 1) jb_hs_swap(k, left, hi - 1);
            // recurse on the right part (deferred), continue with the left part
            stack_lo[sp] = lo0 + (int)(k - (keys + lo0)) + left + 1;
            stack_n[sp] = n - (left + 1);
            stack_d[sp] = depth;
            sp++;
            n = left;
        }
    }
}

// NOTE on order: the reference recurses into the RIGHT part first and then loops on the left part; both
// parts are disjoint sub-arrays, so the deferred execution order does not change the result.
__host__ __device__ inline int jb_build_encoder_table(const uint32_t *freq, JbEncTable *out, JbHSym *sy /* [257] scratch */)
{
    int count = 0;
    for (int i = 0; i < 256; i++)
        if (freq[i]) { sy[count].value = (short)i; sy[count].freq = freq[i]; sy[count].code_size = 0; sy[count].others = -1; count++; }
    for (int i = 0; i < 256; i++) { out->code[i] = 0; out->len[i] = 0; out->vals[i] = 0; }
    for (int i = 0; i < 16; i++) out->bits[i] = 0;
    out->nvals = 0;
    if (count == 0) return 0;
    const int n = count + 1;
    sy[count].value = -1; sy[count].freq = 1; sy[count].code_size = 0; sy[count].others = -1;
    for (;;) { // FindHuffmanCodeSize :178-238
        int v1 = -1, v2 = -1;
        long long f1 = -1, f2 = -1;
        for (int i = 0; i < n; i++) { const long long f = sy[i].freq; if (f >= 0 && (v1 == -1 || f < f1)) { v1 = i; f1 = f; } }
        for (int i = 0; i < n; i++) { const long long f = sy[i].freq; if (f >= 0 && i != v1 && (v2 == -1 || f < f2)) { v2 = i; f2 = f; } }
        if (v2 == -1) break;
        sy[v1].freq += sy[v2].freq;
        sy[v2].freq = -1;
        sy[v1].code_size++;
        while (sy[v1].others != -1) { v1 = sy[v1].others; sy[v1].code_size++; }
        sy[v1].others = (short)v2;
        sy[v2].code_size++;
        while (sy[v2].others != -1) { v2 = sy[v2].others; sy[v2].code_size++; }
    }
    uint8_t bits[264];
    for (int i = 0; i < 264; i++) bits[i] = 0;
    int index = 32;
    for (int i = 0; i < n; i++) { const int cs = sy[i].code_size; if (cs > 0) { if (cs > index) index = cs; bits[cs - 1]++; } }
    // The reference counts the codes of a size in BYTES (`Span<byte> bits`, :117): 256 codes of one size -- 255 symbols of
    // equal weight and the sentinel, a perfectly balanced tree -- wrap to 0, the searches below run off the front of the
    // span and the reference dies of an IndexOutOfRangeException.  No coefficient histogram gets there (at most 162 AC
    // symbols at 8 bits, 242 at 12), a caller of jb_build_huffman_table can: the host build stops (-1) where the reference
    // throws; the device code, which only ever sees K3b's histograms, is left as it was validated.
#ifndef __CUDA_ARCH__
#define JB_BUILDER_BOUND(i) if ((i) < 0) return -1
#else
#define JB_BUILDER_BOUND(i)
#endif
    for (;;) { // K.3 :129-160
        while (bits[index] > 0) {
            int jj = index - 1;
            do { jj -= 1; JB_BUILDER_BOUND(jj); } while (bits[jj] == 0);
            bits[index] -= 2; bits[index - 1] += 1; bits[jj + 1] += 2; bits[jj] -= 1;
        }
        index -= 1;
        if (index != 15) continue;
        while (bits[index] == 0) { index--; JB_BUILDER_BOUND(index); }
        bits[index]--;
        break;
    }
#undef JB_BUILDER_BOUND
    for (int i = 0; i < n; i++) if (sy[i].value == -1) sy[i].code_size = 0xFFFF;
    jb_hs_introsort(sy, n);
    int code = 0, kk = 0;
    for (int l = 1; l <= 16; l++) {
        out->bits[l - 1] = bits[l - 1];
        for (int i = 0; i < bits[l - 1] && kk < count; i++, kk++) {
            const int v = (uint8_t)sy[kk].value;
            out->vals[kk] = (uint8_t)v;
            out->code[v] = (uint16_t)code;
            out->len[v] = (uint8_t)l;
            code++;
        }
        code <<= 1;
    }
    out->nvals = (uint32_t)count;
    return count;
}

__global__ void jb_k3c_build_tables(const uint32_t *__restrict__ hist, JbEncTable *__restrict__ tables, JbHSym *__restrict__ scratch, int ntables)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntables) return;
    jb_build_encoder_table(hist + (uint64_t)t * 256, tables + t, scratch + (uint64_t)t * 257);
}

// ---- E8: bits per block / packing (EncodeBlock :828-870, EncodeRunLength :893-918)
// the block's 64 coefficients as eight 128-bit loads, and the DC predictor (previous block of the same component)
__device__ __forceinline__ void jb_load_block(const JbEncImage &im, const int16_t *__restrict__ coef, uint32_t blk,
                                              uint32_t (&w)[32], int &pred)
{
    const uint4 *p4 = reinterpret_cast<const uint4 *>(coef + (im.coef_off + blk) * 64);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint4 q = __ldg(p4 + j);
        w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w;
    }
    const int64_t pb = jb_prev_block(im, blk);
    pred = pb < 0 ? 0 : coef[(im.coef_off + pb) * 64];
}

template <bool PACK>
__device__ __forceinline__ uint32_t jb_encode_block_regs(const JbEncImage &im, const uint32_t (&w)[32], int pred, uint32_t blk,
                                                         const JbEncTable *__restrict__ tables,
                                                         uint32_t *__restrict__ raw_words, uint64_t bitpos)
{
    const int c = im.blk_comp[blk % im.bpm];
    const JbEncTable *dct = tables + im.table_base + im.comp_td[c];
    const JbEncTable *act = tables + im.table_base + 4 + im.comp_ta[c];
    uint32_t nbits = 0;
    // local accumulator: bits are emitted MSB-first into 32-bit big-endian words
    uint64_t acc = 0;
    int nacc = 0;
    uint64_t word = bitpos >> 5;
    const int lead = (int)(bitpos & 31);
    bool first = true;
    auto emit = [&](uint32_t bits, int len) {
        nbits += len;
        if (!PACK || len == 0) return;
        acc = (acc << len) | (bits & ((1u << len) - 1u));
        nacc += len;
        const int room = first ? 32 - lead : 32;
        if (nacc >= room) {
            const uint32_t out = (uint32_t)(acc >> (nacc - room));
            if (first) atomicOr(raw_words + word, out); // shares its word with the previous block
            else raw_words[word] = out;                 // whole word is mine
            word++;
            nacc -= room;
            acc &= (1ull << nacc) - 1ull;
            first = false;
        }
    };
    auto runlen = [&](const JbEncTable *t, int run, int value) {
        const int a = abs(value), b2 = value < 0 ? value - 1 : value;
        const int nb = jb_bit_count(a);
        const int sym = (run << 4) | nb;
        emit(t->code[sym], t->len[sym]);
        if (nb > 0) emit((uint32_t)b2, nb);
    };
    // (the block's 64 coefficients stay in registers: the unrolled loop below indexes them at compile time; 64 scalar
    // loads per thread left this kernel latency-bound at 16-24 % issue rate)
    runlen(dct, 0, (int)(int16_t)(w[0] & 0xFFFFu) - pred);
    int run = 0;
#pragma unroll
    for (int i = 1; i < 64; i++) {
        const int t = (i & 1) ? ((int)w[i >> 1] >> 16) : (int)(int16_t)(w[i >> 1] & 0xFFFFu);
        if (t == 0) { run++; continue; }
        while (run > 15) { emit(act->code[0xF0], act->len[0xF0]); run -= 16; }
        runlen(act, run, t);
        run = 0;
    }
    if (run > 0) emit(act->code[0], act->len[0]);
    if (PACK && nacc > 0) {
        // trailing partial word: shared with the next block (or with the first block's own lead bits)
        const int room = first ? 32 - lead : 32;
        const uint32_t out = (uint32_t)(acc << (room - nacc));
        atomicOr(raw_words + word, out);
    }
    return nbits;
}

template <bool PACK>
__device__ __forceinline__ uint32_t jb_encode_block(const JbEncImage &im, const int16_t *__restrict__ coef, uint32_t blk,
                                                    const JbEncTable *__restrict__ tables, uint32_t *__restrict__ raw_words,
                                                    uint64_t bitpos)
{
    uint32_t w[32];
    int pred;
    jb_load_block(im, coef, blk, w, pred);
    return jb_encode_block_regs<PACK>(im, w, pred, blk, tables, raw_words, bitpos);
}

// Restart intervals (transcoding): every interval starts on a byte boundary behind the previous interval's 1-bit
// padding and the two marker bytes (JpegOptimizer.cs:794-812).  One warp per interval adds that gap to the bit count
// of the interval's last block, so that the plain prefix sum of K4b yields the right offsets.
__global__ void __launch_bounds__(256)
jb_k4a_interval_gaps(const JbEncImage *__restrict__ images, uint32_t *__restrict__ block_bits)
{
    const JbEncImage &im = images[blockIdx.y];
    const uint32_t k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (im.dri == 0 || k + 1 >= im.nint) return; // (the last interval is padded by K4d like any scan end)
    const uint32_t per = im.dri * im.bpm;
    uint32_t *a = block_bits + im.bits_off + (uint64_t)k * per;
    uint32_t sum = 0;
    for (uint32_t i = lane; i < per; i += 32) sum += a[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if (lane == 0) a[per - 1] += ((8u - (sum & 7u)) & 7u) + 16u;
}

__global__ void __launch_bounds__(256)
jb_k4a_block_bits(const JbEncImage *__restrict__ images, const int16_t *__restrict__ coef,
                  const JbEncTable *__restrict__ tables, uint32_t *__restrict__ block_bits)
{
    const JbEncImage &im = images[blockIdx.y];
    const uint32_t total = im.total_mcus * im.bpm;
    const uint32_t blk = blockIdx.x * 256 + threadIdx.x;
    if (blk >= total) return;
    block_bits[im.bits_off + blk] = jb_encode_block<false>(im, coef, blk, tables, nullptr, 0);
}

// per-image exclusive scan of block_bits (in place -> bit offsets); total bits per image in totals[]
__global__ void __launch_bounds__(1024)
jb_k4b_scan(const JbEncImage *__restrict__ images, uint32_t *__restrict__ block_bits, unsigned long long *__restrict__ totals,
            uint32_t *__restrict__ status)
{
    const JbEncImage &im = images[blockIdx.x];
    const uint32_t total = im.total_mcus * im.bpm;
    uint32_t *a = block_bits + im.bits_off;
    __shared__ uint32_t s_w[32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < total; base += 1024) {
        const uint32_t i = base + tid;
        const uint32_t v = i < total ? a[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += t; }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        uint32_t off = 0, tot = 0;
        for (int w = 0; w < 32; w++) { if (w < wid) off += s_w[w]; tot += s_w[w]; }
        const unsigned long long pos = s_carry + off + x - v;
        if (i < total) a[i] = (uint32_t)pos; // (a scan of 2^32 bits = 512 MiB or more is refused below)
        __syncthreads();
        if (tid == 0) s_carry += tot;
        __syncthreads();
    }
    if (tid == 0) {
        totals[blockIdx.x] = s_carry;
        if (s_carry >> 32) atomicOr(status + blockIdx.x, 16u); // block offsets are 32-bit: nothing is packed
    }
}

__global__ void __launch_bounds__(256)
jb_k4c_pack(const JbEncImage *__restrict__ images, const int16_t *__restrict__ coef, const JbEncTable *__restrict__ tables,
            const uint32_t *__restrict__ block_bits, const unsigned long long *__restrict__ totals,
            uint8_t *__restrict__ raw, uint32_t *__restrict__ status)
{
    const JbEncImage &im = images[blockIdx.y];
    const uint32_t total = im.total_mcus * im.bpm;
    const uint32_t blk = blockIdx.x * 256 + threadIdx.x;
    if (blk >= total) return;
    if ((totals[blockIdx.y] >> 32) != 0) return;    // refused by K4b
    if (totals[blockIdx.y] + 64 > im.raw_cap * 8) { // reserved space too small: reported, nothing written; the host
        if (blk == 0) atomicOr(status + blockIdx.y, 8u); // reserves what the totals ask for and packs again
        return;
    }
    uint32_t *words = reinterpret_cast<uint32_t *>(raw + im.raw_off);
    const uint32_t at = block_bits[im.bits_off + blk];
    const uint32_t nbits = jb_encode_block<true>(im, coef, blk, tables, words, at);
    if (im.dri != 0) {
        // the last block of a restart interval also writes the 1-bit padding and the RSTn marker behind it
        const uint32_t per = im.dri * im.bpm, k = blk / per;
        if (blk % per == per - 1 && k + 1 < im.nint) {
            const uint32_t end = at + nbits, padbits = (8u - (end & 7u)) & 7u;
            const uint32_t v = (((1u << padbits) - 1u) << 16) | 0xFFD0u | (k & 7u); // pad, FF, D0 + (k mod 8)
            const uint32_t len = padbits + 16;                                      // <= 23 bits, MSB first at `end`
            const uint32_t w = end >> 5, sh = end & 31u;
            const uint64_t placed = ((uint64_t)v << (64 - len)) >> sh;
            atomicOr(words + w, (uint32_t)(placed >> 32));
            if ((uint32_t)placed) atomicOr(words + w + 1, (uint32_t)placed);
        }
    }
}

// ---- E9: byte stuffing + padding (JpegWriter.FlushRegister :104-128, ExitBitMode :141-167).
// The un-stuffed stream holds big-endian 32-bit words; one CTA per image.
__global__ void __launch_bounds__(256)
jb_k4d_stuff(const JbEncImage *__restrict__ images, const unsigned long long *__restrict__ totals,
             const uint32_t *__restrict__ block_bits, const uint8_t *__restrict__ raw, uint8_t *__restrict__ out,
             uint32_t *__restrict__ out_len, uint32_t *__restrict__ status)
{
    const JbEncImage &im = images[blockIdx.x];
    const unsigned long long bits = totals[blockIdx.x];
    const uint32_t nbytes = (uint32_t)((bits + 7) >> 3);
    const uint32_t *words = reinterpret_cast<const uint32_t *>(raw + im.raw_off);
    uint8_t *dst = out + im.out_off;
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    if ((bits >> 32) != 0 || bits + 64 > im.raw_cap * 8) { if (tid == 0) out_len[blockIdx.x] = 0; return; }
    const int padbits = (int)((8 - (bits & 7)) & 7);
    // restart markers sit in the un-stuffed stream already (K4c); their FF must not be stuffed.  Marker k occupies the
    // two bytes in front of interval k + 1, whose first block's bit offset is in block_bits.
    const uint32_t per = im.dri * im.bpm;
    const uint32_t nmark = im.dri ? im.nint - 1 : 0;
    const uint32_t *first_bits = block_bits + im.bits_off;
    auto marker_at = [&](uint32_t k) { return (first_bits[(uint64_t)(k + 1) * per] >> 3) - 2u; }; // byte offset of its FF
    for (uint32_t tile = 0; tile < nbytes; tile += 256 * 16) {
        const uint32_t pos0 = tile + tid * 16;
        uint8_t b[16];
        uint32_t cnt = 0;
        uint32_t plain = 0; // bit e: byte pos0 + e is the FF of a marker (copied as it is)
        if (nmark && pos0 < nbytes) {
            uint32_t lo = 0, hi = nmark; // first marker at or behind pos0
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (marker_at(mid) < pos0) lo = mid + 1; else hi = mid;
            }
            for (; lo < nmark; lo++) {
                const uint32_t m = marker_at(lo);
                if (m >= pos0 + 16) break;
                plain |= 1u << (m - pos0);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t w = pos0 + q * 4 < nbytes ? words[(pos0 >> 2) + q] : 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                uint32_t v = (w >> (24 - 8 * e)) & 0xFF;
                const uint32_t p = pos0 + q * 4 + e;
                if (p == nbytes - 1 && padbits) v |= (1u << padbits) - 1u; // final partial byte: 1-bits
                b[q * 4 + e] = (uint8_t)v;
                if (p < nbytes) cnt += (v == 0xFF && !((plain >> (q * 4 + e)) & 1u)) ? 2 : 1;
            }
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_w[wid] = incl;
        __syncthreads();
        uint32_t off = s_base + incl - cnt, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { if (w < wid) off += s_w[w]; tot += s_w[w]; }
        if ((unsigned long long)off + cnt <= im.out_cap) { // (cnt <= 32; a tile that does not fit makes s_base > out_cap below)
#pragma unroll
            for (int e = 0; e < 16; e++)
                if (pos0 + e < nbytes) { dst[off++] = b[e]; if (b[e] == 0xFF && !((plain >> e) & 1u)) dst[off++] = 0; }
        }
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    if (tid == 0) {
        if (s_base > im.out_cap) { atomicOr(status + blockIdx.x, 8u); out_len[blockIdx.x] = 0; }
        else out_len[blockIdx.x] = s_base;
    }
}
