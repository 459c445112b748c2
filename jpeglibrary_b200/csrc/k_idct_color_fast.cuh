// k_idct_color_fast.cuh -- helpers of the fast K2 path (k_idct_color_warp.cuh): zig-zag table, DPX clamp, chroma terms.
// (The first-generation kernel that lived here -- 8 threads per block, CTA-wide phases -- was replaced by the
// warp-autonomous kernel; profiles/r1d_decode.txt vs profiles/r1i_decode_batch1024.txt.)
#pragma once
#include "jb_device.cuh"
#include "k_idct_color.cuh"

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
