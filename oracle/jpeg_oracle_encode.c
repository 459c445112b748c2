/*
 * jpeg_oracle_encode.c -- CPU restatement of the reference's baseline encoder with optimised Huffman
 * coding.  TEST INFRASTRUCTURE ONLY (see jpeg_oracle.h).
 *
 * PARITY UNPINNED: the reference has no encoder tests or golden vectors (tests/JpegLibrary.Tests has
 * only Decoder/, Optimizer/, Utils/).  This file follows the source line by line and is anchored on
 * decode(encode(x)) round trips through the pinned decoder and through libjpeg-turbo.  The symbol
 * order inside one code length depends on .NET's unstable introsort (JpegHuffmanEncodingTableBuilder.cs
 * :170); it is emulated from the runtime's published algorithm and is flagged
 * "pending confirmation on a .NET machine".  Code LENGTHS per symbol do not depend on it.
 */
#include "jpeg_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const uint8_t kZigzagToNatural[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* JpegStandardQuantizationTable.cs:9-31 (zig-zag order) */
static const uint16_t kStdLuma[64] = {16, 11, 12, 14, 12, 10, 16, 14, 13, 14, 18, 17, 16, 19, 24, 40,
                                      26, 24, 22, 22, 24, 49, 35, 37, 29, 40, 58, 51, 61, 60, 57, 51,
                                      56, 55, 64, 72, 92, 78, 64, 68, 87, 69, 55, 56, 80, 109, 81, 87,
                                      95, 98, 103, 104, 103, 62, 77, 113, 121, 112, 100, 120, 92, 101, 103, 99};
static const uint16_t kStdChroma[64] = {17, 18, 18, 24, 21, 24, 47, 26, 26, 47, 99, 66, 56, 66, 99, 99,
                                        99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                        99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                        99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* ScaleByQuality JpegStandardQuantizationTable.cs:64-89 */
void jo_std_quant_table(int chroma, int quality, uint16_t out_zz[64])
{
    const uint16_t *src = chroma ? kStdChroma : kStdLuma;
    int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 64; i++) {
        int x = (src[i] * scale + 50) / 100;
        out_zz[i] = (uint16_t)(x < 1 ? 1 : (x > 255 ? 255 : x));
    }
}

/* apps/JpegEncode/JpegRgbToYCbCrConverter.cs:26-93 */
static int fix16f(float x) { return (int)((x * 65536.0f) + 0.5f); }
void jo_rgb_to_ycbcr(const uint8_t *rgb, uint8_t *ycbcr, size_t n)
{
    const int yr = fix16f(0.299f), yg = fix16f(0.587f), yb = fix16f(0.114f);
    const int cbr = -fix16f(0.168735892f), cbg = -fix16f(0.331264108f), half = fix16f(0.5f);
    const int crg = -fix16f(0.418687589f), crb = -fix16f(0.081312411f);
    const int off = (128 << 16) + 32768 - 1;
    for (size_t i = 0; i < n; i++) {
        int r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        ycbcr[3 * i] = (uint8_t)((yr * r + yg * g + (yb * b + 32768)) >> 16);
        ycbcr[3 * i + 1] = (uint8_t)((cbr * r + cbg * g + (half * b + off)) >> 16);
        ycbcr[3 * i + 2] = (uint8_t)(((half * r + off) + crg * g + crb * b) >> 16);
    }
}

/* FastFloatingPointDCT.FDCT8x4_{Left,Right}Part :195-314: 1-D pass over rows V0..V7, per column */
static void fdct_pass(const float *s, float *d)
{
    for (int c = 0; c < 8; c++) {
        float c0 = s[c], c1 = s[56 + c];
        float t0 = c0 + c1, t7 = c0 - c1;
        c1 = s[48 + c]; c0 = s[8 + c];
        float t1 = c0 + c1, t6 = c0 - c1;
        c1 = s[40 + c]; c0 = s[16 + c];
        float t2 = c0 + c1, t5 = c0 - c1;
        c0 = s[24 + c]; c1 = s[32 + c];
        float t3 = c0 + c1, t4 = c0 - c1;
        c0 = t0 + t3;
        float c3 = t0 - t3;
        c1 = t1 + t2;
        float c2 = t1 - t2;
        d[c] = c0 + c1;
        d[32 + c] = c0 - c1;
        float w0 = 0.541196f, w1 = 1.306563f;
        d[16 + c] = (w0 * c2) + (w1 * c3);
        d[48 + c] = (w0 * c3) - (w1 * c2);
        w0 = 1.175876f; w1 = 0.785695f;
        c3 = (w0 * t4) + (w1 * t7);
        c0 = (w0 * t7) - (w1 * t4);
        w0 = 1.387040f; w1 = 0.275899f;
        c2 = (w0 * t5) + (w1 * t6);
        c1 = (w0 * t6) - (w1 * t5);
        d[24 + c] = c0 - c2;
        d[40 + c] = c3 - c1;
        const float invsqrt2 = 0.707107f;
        c0 = (c0 + c2) * invsqrt2;
        c3 = (c3 + c1) * invsqrt2;
        d[8 + c] = c0 + c3;
        d[56 + c] = c0 - c3;
    }
}
static void transpose8(const float *s, float *d)
{
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) d[j * 8 + i] = s[i * 8 + j];
}

/* ShiftDataLevel :801-810, TransformFDCT FastFloatingPointDCT.cs:346-362, ZigZagAndQuantizeBlock :812-826 */
static void fdct_quant_block(const int16_t samples[64], const uint16_t q_zz[64], int16_t out_zz[64])
{
    float a[64], b[64];
    for (int i = 0; i < 64; i++) a[i] = (float)(samples[i] - 128);
    transpose8(a, b);
    fdct_pass(b, a);
    transpose8(a, b);
    fdct_pass(b, a);
    for (int i = 0; i < 64; i++) a[i] = a[i] * 0.125f;
    for (int i = 0; i < 64; i++) out_zz[i] = (int16_t)rintf(a[kZigzagToNatural[i]] / (float)q_zz[i]);
}

/* apps/JpegEncode/JpegBufferInputReader.ReadBlock :26-50 */
static void read_block(const uint8_t *ycbcr, int W, int H, int ncomp, int ci, int x, int y, int16_t blk[64], int clear_partial)
{
    int bw = W - x < 8 ? W - x : 8, bh = H - y < 8 ? H - y : 8;
    if (bw < 0) bw = 0;
    if (bh < 0) bh = 0;
    if ((bw != 8 || bh != 8) && clear_partial) memset(blk, 0, 128);
    for (int oy = 0; oy < bh; oy++)
        for (int ox = 0; ox < bw; ox++) blk[oy * 8 + ox] = ycbcr[((size_t)(y + oy) * W + x + ox) * ncomp + ci];
}

/* ReadBlock / ReadBlockWithSubsample / CopySubsampleBlock JpegEncoder.cs:743-799; dst is the
   zero-initialised allocator block (:465-468) */
static void read_block_subsampled(const uint8_t *ycbcr, int W, int H, int ncomp, int ci, int x, int y, int hs, int vs,
                                  int16_t dst[64])
{
    if (hs == 1 && vs == 1) {
        read_block(ycbcr, W, H, ncomp, ci, x, y, dst, 1);
        return;
    }
    int hshift = hs == 4 ? 2 : hs == 2 ? 1 : 0, vshift = vs == 4 ? 2 : vs == 2 ? 1 : 0;
    int16_t tmp[64];
    memset(tmp, 0, sizeof tmp); /* SkipInit'd in the reference; only matters for partial edge blocks,
                                   where ReadBlock clears it first */
    for (int v = 0; v < vs; v++)
        for (int h = 0; h < hs; h++) {
            read_block(ycbcr, W, H, ncomp, ci, x + 8 * h, y + 8 * v, tmp, 1);
            int box = h << (3 - hshift), boy = v << (3 - vshift);
            for (int yy = 0; yy < 8; yy++)
                for (int xx = 0; xx < 8; xx++) dst[(boy + (yy >> vshift)) * 8 + box + (xx >> hshift)] += tmp[yy * 8 + xx];
        }
    int total = hshift + vshift;
    if (total > 0) {
        int delta = 1 << (total - 1);
        for (int i = 0; i < 64; i++) dst[i] = (int16_t)((dst[i] + delta) >> total);
    }
}

/* BitCountTable :938-953 */
static int bit_count(int a)
{
    int n = 0;
    while (a) { n++; a >>= 1; }
    return n;
}

/* ------------------------------------------------------------------------- */
/* JpegHuffmanEncodingTableBuilder.BuildUsingStandardMethod :69-176            */
typedef struct { long long freq; short value; unsigned short code_size; short others; } hsym;

/* .NET Core ArraySortHelper<T>.IntroSort with a Comparison<T> */
static int cmp_size(const hsym *a, const hsym *b) { return (a->code_size > b->code_size) - (a->code_size < b->code_size); }
static void swap_if_greater(hsym *k, int i, int j)
{
    if (i != j && cmp_size(&k[i], &k[j]) > 0) { hsym t = k[i]; k[i] = k[j]; k[j] = t; }
}
static void swap_sym(hsym *k, int i, int j) { if (i != j) { hsym t = k[i]; k[i] = k[j]; k[j] = t; } }
static void insertion_sort(hsym *k, int n)
{
    for (int i = 0; i < n - 1; i++) {
        hsym t = k[i + 1];
        int j = i;
        while (j >= 0 && cmp_size(&t, &k[j]) < 0) { k[j + 1] = k[j]; j--; }
        k[j + 1] = t;
    }
}
static void down_heap(hsym *k, int i, int n)
{
    hsym d = k[i - 1];
    while (i <= n / 2) {
        int child = 2 * i;
        if (child < n && cmp_size(&k[child - 1], &k[child]) < 0) child++;
        if (!(cmp_size(&d, &k[child - 1]) < 0)) break;
        k[i - 1] = k[child - 1];
        i = child;
    }
    k[i - 1] = d;
}
static void heap_sort(hsym *k, int n)
{
    for (int i = n / 2; i >= 1; i--) down_heap(k, i, n);
    for (int i = n; i > 1; i--) { swap_sym(k, 0, i - 1); down_heap(k, 1, i - 1); }
}
static int pick_pivot_and_partition(hsym *k, int n)
{
    int hi = n - 1, mid = hi >> 1;
    swap_if_greater(k, 0, mid);
    swap_if_greater(k, 0, hi);
    swap_if_greater(k, mid, hi);
    hsym pivot = k[mid];
    swap_sym(k, mid, hi - 1);
    int left = 0, right = hi - 1;
    while (left < right) {
        while (cmp_size(&k[++left], &pivot) < 0) ;
        while (cmp_size(&pivot, &k[--right]) < 0) ;
        if (left >= right) break;
        swap_sym(k, left, right);
    }
    if (left != hi - 1) swap_sym(k, left, hi - 1);
    return left;
}
static void intro_sort(hsym *k, int n, int depth)
{
    while (n > 1) {
        if (n <= 16) {
            if (n == 2) { swap_if_greater(k, 0, 1); return; }
            if (n == 3) { swap_if_greater(k, 0, 1); swap_if_greater(k, 0, 2); swap_if_greater(k, 1, 2); return; }
            insertion_sort(k, n);
            return;
        }
        if (depth == 0) { heap_sort(k, n); return; }
        depth--;
        int p = pick_pivot_and_partition(k, n);
        intro_sort(k + p + 1, n - (p + 1), depth);
        n = p;
    }
}

int jo_build_huffman_table(const uint32_t freq[256], uint8_t bits_out[16], uint8_t vals[256])
{
    hsym sy[257];
    int count = 0;
    for (int i = 0; i < 256; i++)
        if (freq[i]) { sy[count].value = (short)i; sy[count].freq = freq[i]; sy[count].code_size = 0; sy[count].others = -1; count++; }
    memset(bits_out, 0, 16);
    if (count == 0) return 0;
    int n = count + 1;
    sy[count].value = -1; sy[count].freq = 1; sy[count].code_size = 0; sy[count].others = -1;
    /* FindHuffmanCodeSize :178-238 (ties -> lowest index) */
    for (;;) {
        int v1 = -1, v2 = -1;
        long long f1 = -1, f2 = -1;
        for (int i = 0; i < n; i++) { long long f = sy[i].freq; if (f >= 0 && (v1 == -1 || f < f1)) { v1 = i; f1 = f; } }
        for (int i = 0; i < n; i++) { long long f = sy[i].freq; if (f >= 0 && i != v1 && (v2 == -1 || f < f2)) { v2 = i; f2 = f; } }
        if (v2 == -1) break;
        sy[v1].freq += sy[v2].freq;
        sy[v2].freq = -1;
        sy[v1].code_size++;
        while (sy[v1].others != -1) { v1 = sy[v1].others; sy[v1].code_size++; }
        sy[v1].others = (short)v2;
        sy[v2].code_size++;
        while (sy[v2].others != -1) { v2 = sy[v2].others; sy[v2].code_size++; }
    }
    /* K.2 / K.3 :111-160 */
    uint8_t bits[300];
    memset(bits, 0, sizeof bits);
    int index = 32;
    for (int i = 0; i < n; i++) {
        int cs = sy[i].code_size;
        if (cs > 0) { if (cs > index) index = cs; bits[cs - 1]++; }
    }
    for (;;) {
        while (bits[index] > 0) {
            int j = index - 1;
            do { j -= 1; if (j < 0) return -1; } while (bits[j] == 0); /* IndexOutOfRangeException in the reference */
            bits[index] -= 2;
            bits[index - 1] += 1;
            bits[j + 1] += 2;
            bits[j] -= 1;
        }
        index -= 1;
        if (index != 15) continue;
        /* `Span<byte> bits` (:117): 256 codes of one size wrap to 0 and the reference runs off the front of the span */
        while (bits[index] == 0) { index--; if (index < 0) return -1; }
        bits[index]--;
        break;
    }
    /* sort :162-170 */
    for (int i = 0; i < n; i++) if (sy[i].value == -1) sy[i].code_size = 0xFFFF;
    int depth = 0;
    for (int t = n; t > 0; t >>= 1) depth++; /* floor(log2(n)) + 1 */
    intro_sort(sy, n, 2 * depth);
    for (int i = 0; i < 16; i++) bits_out[i] = bits[i];
    for (int i = 0; i < count; i++) vals[i] = (uint8_t)sy[i].value;
    return count;
}

/* ------------------------------------------------------------------------- */
/* JpegHuffmanEncodingTableBuilder.BuildUsingPackageMerge :287-413 (MostOptimalCoding = true).
   Four of its five sorts are Array.Sort / List<T>.Sort with a Comparison: the runtime's unstable introsort again, here
   over arbitrary elements (gsort_*: the same algorithm as intro_sort above with the comparison as a parameter). */
typedef int (*gcmp)(const void *, const void *);
#define GS_MAX 24
static void gs_swap(char *k, size_t sz, int i, int j)
{
    if (i == j) return;
    char t[GS_MAX];
    memcpy(t, k + i * sz, sz); memcpy(k + i * sz, k + j * sz, sz); memcpy(k + j * sz, t, sz);
}
static void gs_swap_if_greater(char *k, size_t sz, gcmp c, int i, int j)
{
    if (i != j && c(k + i * sz, k + j * sz) > 0) gs_swap(k, sz, i, j);
}
static void gs_insertion(char *k, size_t sz, gcmp c, int n)
{
    for (int i = 0; i < n - 1; i++) {
        char t[GS_MAX];
        memcpy(t, k + (i + 1) * sz, sz);
        int j = i;
        while (j >= 0 && c(t, k + j * sz) < 0) { memcpy(k + (j + 1) * sz, k + j * sz, sz); j--; }
        memcpy(k + (j + 1) * sz, t, sz);
    }
}
static void gs_down_heap(char *k, size_t sz, gcmp c, int i, int n)
{
    char d[GS_MAX];
    memcpy(d, k + (i - 1) * sz, sz);
    while (i <= n / 2) {
        int child = 2 * i;
        if (child < n && c(k + (child - 1) * sz, k + child * sz) < 0) child++;
        if (!(c(d, k + (child - 1) * sz) < 0)) break;
        memcpy(k + (i - 1) * sz, k + (child - 1) * sz, sz);
        i = child;
    }
    memcpy(k + (i - 1) * sz, d, sz);
}
static void gs_heap_sort(char *k, size_t sz, gcmp c, int n)
{
    for (int i = n / 2; i >= 1; i--) gs_down_heap(k, sz, c, i, n);
    for (int i = n; i > 1; i--) { gs_swap(k, sz, 0, i - 1); gs_down_heap(k, sz, c, 1, i - 1); }
}
static int gs_partition(char *k, size_t sz, gcmp c, int n)
{
    int hi = n - 1, mid = hi >> 1;
    gs_swap_if_greater(k, sz, c, 0, mid);
    gs_swap_if_greater(k, sz, c, 0, hi);
    gs_swap_if_greater(k, sz, c, mid, hi);
    char pivot[GS_MAX];
    memcpy(pivot, k + mid * sz, sz);
    gs_swap(k, sz, mid, hi - 1);
    int left = 0, right = hi - 1;
    while (left < right) {
        while (c(k + (++left) * sz, pivot) < 0) ;
        while (c(pivot, k + (--right) * sz) < 0) ;
        if (left >= right) break;
        gs_swap(k, sz, left, right);
    }
    if (left != hi - 1) gs_swap(k, sz, left, hi - 1);
    return left;
}
static void gs_intro(char *k, size_t sz, gcmp c, int n, int depth)
{
    while (n > 1) {
        if (n <= 16) {
            if (n == 2) { gs_swap_if_greater(k, sz, c, 0, 1); return; }
            if (n == 3) { gs_swap_if_greater(k, sz, c, 0, 1); gs_swap_if_greater(k, sz, c, 0, 2); gs_swap_if_greater(k, sz, c, 1, 2); return; }
            gs_insertion(k, sz, c, n);
            return;
        }
        if (depth == 0) { gs_heap_sort(k, sz, c, n); return; }
        depth--;
        int p = gs_partition(k, sz, c, n);
        gs_intro(k + (size_t)(p + 1) * sz, sz, c, n - (p + 1), depth);
        n = p;
    }
}
static void dotnet_sort(void *base, int n, size_t sz, gcmp c)
{
    if (n < 2) return;
    int depth = 0;
    for (int t = n; t > 0; t >>= 1) depth++; /* floor(log2(n)) + 1 */
    gs_intro((char *)base, sz, c, n, 2 * depth);
}

typedef struct { long long freq; int index; int left, right; } pmnode; /* Node :456-476; left == -1: a leaf */
static const pmnode *pm_pool; /* the comparisons below sort node HANDLES (ints) by the frequency of the node they name */
static int cmp_sym_freq_desc(const void *a, const void *b)
{ /* (x, y) => y.Frequency.CompareTo(x.Frequency) :346 */
    long long x = ((const hsym *)a)->freq, y = ((const hsym *)b)->freq;
    return (y > x) - (y < x);
}
static int cmp_node_desc(const void *a, const void *b)
{
    long long x = pm_pool[*(const int *)a].freq, y = pm_pool[*(const int *)b].freq;
    return (y > x) - (y < x);
}
static int cmp_node_asc(const void *a, const void *b) { return cmp_node_desc(b, a); }
static int cmp_symbol_comparer(const void *a, const void *b)
{ /* SymbolComparer :428-453: code size ascending, then frequency descending */
    const hsym *x = a, *y = b;
    if (x->code_size > y->code_size) return 1;
    if (x->code_size < y->code_size) return -1;
    if (x->freq > y->freq) return -1;
    if (x->freq < y->freq) return 1;
    return 0;
}
static void pm_traverse(const pmnode *pool, int node, hsym *sy)
{ /* TraverseNode :394-409 */
    if (pool[node].left < 0) { sy[pool[node].index].code_size++; return; }
    pm_traverse(pool, pool[node].left, sy);
    pm_traverse(pool, pool[node].right, sy);
}

int jo_build_huffman_table_optimal(const uint32_t freq[256], uint8_t bits_out[16], uint8_t vals[256])
{
    hsym sy[257];
    int count = 0;
    for (int i = 0; i < 256; i++)
        if (freq[i]) { sy[count].value = (short)i; sy[count].freq = freq[i]; sy[count].code_size = 0; sy[count].others = 0; count++; }
    memset(bits_out, 0, 16);
    if (count == 0) return 0; /* (the reference would go on with the sentinel alone; the encoder never asks) */
    sy[count].value = -1; sy[count].freq = 0; sy[count].code_size = 0; sy[count].others = 0; /* :316-321 */
    const int n = count + 1;
    /* RunPackageMerge :344-410 */
    dotnet_sort(sy, n, sizeof(hsym), cmp_sym_freq_desc);
    /* 16 levels of n leaves each; level l - 1 receives at most half of level l's nodes as packages: < 2n per level */
    const int cap = 2 * n + 2;
    pmnode *pool = malloc(sizeof(pmnode) * (size_t)(16 * cap));
    int *lists = malloc(sizeof(int) * (size_t)(16 * cap));
    int npool = 0, len[16];
    for (int l = 15; l >= 0; l--) {
        for (int i = 0; i < n; i++) {
            pool[npool].freq = sy[i].freq; pool[npool].index = i; pool[npool].left = pool[npool].right = -1;
            lists[l * cap + i] = npool++;
        }
        len[l] = n;
    }
    pm_pool = pool;
    for (int l = 15; l > 0; l--) {
        int *nodes = lists + l * cap, *next = lists + (l - 1) * cap;
        dotnet_sort(nodes, len[l], sizeof(int), cmp_node_desc);
        while (len[l] >= 2) { /* package the two smallest (the last two) and merge the package into the next level */
            const int n1 = nodes[len[l] - 1], n2 = nodes[len[l] - 2];
            len[l] -= 2;
            pool[npool].freq = pool[n1].freq + pool[n2].freq; pool[npool].index = 0;
            pool[npool].left = n1; pool[npool].right = n2;
            next[len[l - 1]++] = npool++;
        }
    }
    dotnet_sort(lists, len[0], sizeof(int), cmp_node_asc);
    int select = 2 * (n - 1);
    if (select < 1) select = 1;
    for (int i = 0; i < select; i++) pm_traverse(pool, lists[i], sy);
    free(pool);
    free(lists);
    /* :331-343: order by (code size, frequency), drop the sentinel */
    dotnet_sort(sy, n, sizeof(hsym), cmp_symbol_comparer);
    int at = 0;
    for (int i = n - 1; i >= 0; i--) if (sy[i].value == -1) { at = i; break; }
    for (int i = at; i < n - 1; i++) sy[i] = sy[i + 1];
    /* BuildCanonicalCode(symbols) :478-509 assigns consecutive codes in this order; as a DHT that is the counts per code
       size and the symbols in this order (JpegHuffmanEncodingTable.TryWrite :50-86) */
    for (int i = 0; i < count; i++) {
        if (sy[i].code_size >= 1 && sy[i].code_size <= 16) bits_out[sy[i].code_size - 1]++;
        vals[i] = (uint8_t)sy[i].value;
    }
    return count;
}

/* canonical codes from (bits, vals): BuildCanonicalCode :240-282 */
typedef struct { uint16_t code[256]; uint8_t len[256]; } enc_table;
static void make_enc_table(const uint8_t bits[16], const uint8_t *vals, int count, enc_table *t)
{
    memset(t, 0, sizeof *t);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < bits[l - 1] && k < count; i++, k++) { t->code[vals[k]] = (uint16_t)code; t->len[vals[k]] = (uint8_t)l; code++; }
        code <<= 1;
    }
}

/* JpegWriter bit mode: WriteBits :207-227, FlushRegister :104-128, ExitBitMode :141-167 */
typedef struct { uint8_t *p; size_t n, cap; uint64_t acc; int nbits; } bitw;
static void bw_byte(bitw *w, uint8_t b)
{
    if (w->n + 2 > w->cap) { w->cap = w->cap * 2 + 65536; w->p = realloc(w->p, w->cap); }
    w->p[w->n++] = b;
    if (b == 0xFF) w->p[w->n++] = 0;
}
static void bw_bits(bitw *w, uint32_t bits, int len)
{
    if (len == 0) return;
    w->acc = (w->acc << len) | (bits & ((1u << len) - 1u));
    w->nbits += len;
    while (w->nbits >= 8) { bw_byte(w, (uint8_t)(w->acc >> (w->nbits - 8))); w->nbits -= 8; }
}
static void bw_finish(bitw *w)
{
    if (w->nbits > 0) { int pad = 8 - w->nbits; bw_bits(w, (1u << pad) - 1u, pad); }
}

static void put_marker_bytes(bitw *w, const uint8_t *b, size_t n)
{ /* raw bytes, no stuffing */
    if (w->n + n > w->cap) { w->cap = (w->cap + n) * 2 + 65536; w->p = realloc(w->p, w->cap); }
    memcpy(w->p + w->n, b, n);
    w->n += n;
}

int jo_encode_ycbcr(const uint8_t *ycbcr, const jo_encode_params *p, jo_encoded *out)
{
    memset(out, 0, sizeof *out);
    const int W = p->width, H = p->height, nc = p->ncomp;
    int hmax = 1, vmax = 1;
    for (int c = 0; c < nc; c++) { if (p->h[c] > hmax) hmax = p->h[c]; if (p->v[c] > vmax) vmax = p->v[c]; }
    const int mpl = (W + 8 * hmax - 1) / (8 * hmax), mpc = (H + 8 * vmax - 1) / (8 * vmax);
    const int wblk = (W + 7) / 8, hblk = (H + 7) / 8;
    int16_t *dummy = out->dummy; /* (zero: memset above) */
    for (int c = 0; c < nc; c++) {
        int hs = hmax / p->h[c], vs = vmax / p->v[c];
        out->alloc_w[c] = (wblk + hs - 1) / hs;
        out->alloc_h[c] = (hblk + vs - 1) / vs;
        out->coef[c] = calloc((size_t)out->alloc_w[c] * out->alloc_h[c] * 64, sizeof(int16_t));
        if (!p->qt_present[p->tq[c]]) { snprintf(out->error, sizeof out->error, "Quantization table is not defined."); return JO_ERR_INVALID_OP; }
    }
#define BLK(c, bx, by) (((bx) >= out->alloc_w[c] || (by) >= out->alloc_h[c]) ? dummy : out->coef[c] + ((size_t)(by) * out->alloc_w[c] + (bx)) * 64)
    /* TransformBlocks :414-483 (MCU-padding blocks alias the allocator's dummy block, quirk Q4) */
    for (int my = 0; my < mpc; my++)
        for (int mx = 0; mx < mpl; mx++)
            for (int c = 0; c < nc; c++) {
                int hs = hmax / p->h[c], vs = vmax / p->v[c];
                for (int y = 0; y < p->v[c]; y++)
                    for (int x = 0; x < p->h[c]; x++) {
                        int bx = mx * p->h[c] + x, by = my * p->v[c] + y;
                        int16_t *blk = BLK(c, bx, by);
                        int16_t tmp[64];
                        read_block_subsampled(ycbcr, W, H, nc, c, bx * 8 * hs, by * 8 * vs, hs, vs, blk);
                        memcpy(tmp, blk, 128);
                        fdct_quant_block(tmp, p->qt[p->tq[c]], blk);
                    }
            }
    /* BuildHuffmanTables / GatherBlockStatistics :491-597 */
    int pred[JO_MAX_COMP] = {0, 0, 0, 0};
    for (int my = 0; my < mpc; my++)
        for (int mx = 0; mx < mpl; mx++)
            for (int c = 0; c < nc; c++)
                for (int y = 0; y < p->v[c]; y++)
                    for (int x = 0; x < p->h[c]; x++) {
                        const int16_t *blk = BLK(c, mx * p->h[c] + x, my * p->v[c] + y);
                        int t = blk[0] - pred[c];
                        pred[c] = blk[0];
                        out->hist[0][p->td[c]][bit_count(t < 0 ? -t : t)]++;
                        int run = 0;
                        for (int i = 1; i < 64; i++) {
                            t = blk[i];
                            if (t == 0) { run++; continue; }
                            while (run > 15) { out->hist[1][p->ta[c]][0xF0]++; run -= 16; }
                            out->hist[1][p->ta[c]][(run << 4) | bit_count(t < 0 ? -t : t)]++;
                            run = 0;
                        }
                        if (run > 0) out->hist[1][p->ta[c]][0]++;
                    }
    /* tables: one per (class, id) that some component uses, in EncodeAction's insertion order
       (dc0, ac0, dc1, ac1): apps/JpegEncode/EncodeAction.cs:42-45 */
    enc_table enc[2][4];
    int used[2][4];
    memset(used, 0, sizeof used);
    for (int c = 0; c < nc; c++) { used[0][p->td[c]] = 1; used[1][p->ta[c]] = 1; }
    for (int id = 0; id < 4; id++)
        for (int cls = 0; cls < 2; cls++)
            if (used[cls][id]) {
                out->dht_nvals[cls][id] = (p->optimize == 2 ? jo_build_huffman_table_optimal : jo_build_huffman_table)(
                    out->hist[cls][id], out->dht_bits[cls][id], out->dht_vals[cls][id]);
                if (out->dht_nvals[cls][id] == 0) { snprintf(out->error, sizeof out->error, "No symbol is recorded."); return JO_ERR_INVALID_OP; }
                make_enc_table(out->dht_bits[cls][id], out->dht_vals[cls][id], out->dht_nvals[cls][id], &enc[cls][id]);
            }
    /* stream: SOI, DQT, SOF0, DHT, SOS, data, EOI (JpegEncoder.Encode :255-290) */
    bitw w = {0};
    uint8_t hdr[1024];
    size_t k = 0;
    hdr[k++] = 0xFF; hdr[k++] = 0xD8;
    /* DQT: all tables in one segment, SetQuantizationTable order = identifier order here */
    int nq = 0;
    for (int i = 0; i < 4; i++) nq += p->qt_present[i] ? 1 : 0;
    hdr[k++] = 0xFF; hdr[k++] = 0xDB;
    hdr[k++] = (uint8_t)((nq * 65 + 2) >> 8); hdr[k++] = (uint8_t)(nq * 65 + 2);
    for (int i = 0; i < 4; i++)
        if (p->qt_present[i]) { hdr[k++] = (uint8_t)i; for (int j = 0; j < 64; j++) hdr[k++] = (uint8_t)p->qt[i][j]; }
    /* SOF0 */
    hdr[k++] = 0xFF; hdr[k++] = 0xC0;
    int sof_len = 6 + 3 * nc + 2;
    hdr[k++] = (uint8_t)(sof_len >> 8); hdr[k++] = (uint8_t)sof_len;
    hdr[k++] = 8; hdr[k++] = (uint8_t)(H >> 8); hdr[k++] = (uint8_t)H; hdr[k++] = (uint8_t)(W >> 8); hdr[k++] = (uint8_t)W; hdr[k++] = (uint8_t)nc;
    for (int c = 0; c < nc; c++) { hdr[k++] = (uint8_t)(c + 1); hdr[k++] = (uint8_t)((p->h[c] << 4) | p->v[c]); hdr[k++] = (uint8_t)p->tq[c]; }
    put_marker_bytes(&w, hdr, k);
    /* DHT: one segment */
    k = 0;
    int dht_len = 2;
    for (int id = 0; id < 4; id++) for (int cls = 0; cls < 2; cls++) if (used[cls][id]) dht_len += 17 + out->dht_nvals[cls][id];
    hdr[k++] = 0xFF; hdr[k++] = 0xC4; hdr[k++] = (uint8_t)(dht_len >> 8); hdr[k++] = (uint8_t)dht_len;
    put_marker_bytes(&w, hdr, k);
    for (int id = 0; id < 4; id++)
        for (int cls = 0; cls < 2; cls++)
            if (used[cls][id]) {
                uint8_t tc = (uint8_t)((cls << 4) | id);
                put_marker_bytes(&w, &tc, 1);
                put_marker_bytes(&w, out->dht_bits[cls][id], 16);
                put_marker_bytes(&w, out->dht_vals[cls][id], (size_t)out->dht_nvals[cls][id]);
            }
    /* SOS */
    k = 0;
    int sos_len = 6 + 2 * nc;
    hdr[k++] = 0xFF; hdr[k++] = 0xDA; hdr[k++] = (uint8_t)(sos_len >> 8); hdr[k++] = (uint8_t)sos_len; hdr[k++] = (uint8_t)nc;
    for (int c = 0; c < nc; c++) { hdr[k++] = (uint8_t)(c + 1); hdr[k++] = (uint8_t)((p->td[c] << 4) | p->ta[c]); }
    hdr[k++] = 0; hdr[k++] = 63; hdr[k++] = 0;
    put_marker_bytes(&w, hdr, k);
    out->scan_offset = w.n;
    /* WritePreparedScanData :605-656 / EncodeBlock :828-870 */
    memset(pred, 0, sizeof pred);
    for (int my = 0; my < mpc; my++)
        for (int mx = 0; mx < mpl; mx++)
            for (int c = 0; c < nc; c++)
                for (int y = 0; y < p->v[c]; y++)
                    for (int x = 0; x < p->h[c]; x++) {
                        const int16_t *blk = BLK(c, mx * p->h[c] + x, my * p->v[c] + y);
                        const enc_table *dct = &enc[0][p->td[c]], *act = &enc[1][p->ta[c]];
                        int t = blk[0] - pred[c];
                        pred[c] = blk[0];
                        int a = t < 0 ? -t : t, b = t < 0 ? t - 1 : t, nb = bit_count(a);
                        bw_bits(&w, dct->code[nb], dct->len[nb]);
                        if (nb) bw_bits(&w, (uint32_t)b & ((1u << nb) - 1u), nb);
                        int run = 0;
                        for (int i = 1; i < 64; i++) {
                            t = blk[i];
                            if (t == 0) { run++; continue; }
                            while (run > 15) { bw_bits(&w, act->code[0xF0], act->len[0xF0]); run -= 16; }
                            a = t < 0 ? -t : t; b = t < 0 ? t - 1 : t; nb = bit_count(a);
                            int sym = (run << 4) | nb;
                            bw_bits(&w, act->code[sym], act->len[sym]);
                            bw_bits(&w, (uint32_t)b & ((1u << nb) - 1u), nb);
                            run = 0;
                        }
                        if (run > 0) bw_bits(&w, act->code[0], act->len[0]);
                    }
    bw_finish(&w);
    out->scan_len = w.n - out->scan_offset;
    uint8_t eoi[2] = {0xFF, 0xD9};
    put_marker_bytes(&w, eoi, 2);
    out->bytes = w.p;
    out->len = w.n;
    return JO_OK;
#undef BLK
}

void jo_encoded_free(jo_encoded *e)
{
    free(e->bytes);
    e->bytes = NULL;
    for (int i = 0; i < JO_MAX_COMP; i++) { free(e->coef[i]); e->coef[i] = NULL; }
}
