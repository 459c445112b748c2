/*
 * jpeg_oracle.h -- CPU restatement of yigolden/JpegLibrary's Huffman JPEG hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks the decode side
 * bit-exactly against every Huffman golden vector the reference's own tests hold
 * (tests/Assets/{baseline,huffman_sequential,huffman_progressive}/ *.jpg with their
 * .high.png/.low-diff.png pairs; reference loader tests/JpegLibrary.Tests/Utils/
 * ImageHelper.cs:12-91).  The encoder side has no golden vectors in the reference
 * ("parity unpinned" for E1-E9); it is anchored on the source and on
 * decode(encode(x)) round trips.
 *
 * All file:line citations are relative to the reference tree (src/JpegLibrary/...).
 */
#ifndef JPEG_ORACLE_H
#define JPEG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JO_OK 0
#define JO_ERR_INVALID_DATA (-1)   /* InvalidDataException in the reference          */
#define JO_ERR_INVALID_OP (-2)     /* InvalidOperationException (e.g. missing RST)   */
#define JO_ERR_UNSUPPORTED (-3)    /* NotSupportedException / arithmetic / lossless  */
#define JO_ERR_NOMEM (-4)

#define JO_MAX_COMP 4
#define JO_MAX_SCANS 64

typedef struct {
    int ncomp;               /* Ns */
    int comp_index[JO_MAX_COMP]; /* index into frame components */
    int td[JO_MAX_COMP], ta[JO_MAX_COMP];
    int ss, se, ah, al;
    size_t entropy_offset;   /* byte offset of the first entropy-coded byte */
    int restart_interval;    /* DRI in force for this scan */
} jo_scan_info;

typedef struct {
    /* frame */
    int sof;                 /* 0,1,2 */
    int precision, width, height, ncomp;
    int comp_id[JO_MAX_COMP], comp_h[JO_MAX_COMP], comp_v[JO_MAX_COMP], comp_tq[JO_MAX_COMP];
    int hmax, vmax, mcus_per_line, mcus_per_col;
    /* quantisation tables actually used by each component at IDCT time (zig-zag order) */
    uint16_t qt[JO_MAX_COMP][64];
    /* scans */
    int nscans;
    jo_scan_info scans[JO_MAX_SCANS];
    /* coefficient store: per component, MCU-padded block grid, row-major blocks,
       each block 64 int16 in ZIG-ZAG order with absolute (DC-predicted) DC.
       For progressive frames blocks outside the reference allocator's grid
       (bx >= alloc_w || by >= alloc_h) are "don't care" (the reference routes them
       to a shared dummy block, JpegBlockAllocator.cs:108-111). */
    int coef_w[JO_MAX_COMP], coef_h[JO_MAX_COMP];   /* padded grid (blocks) */
    int alloc_w[JO_MAX_COMP], alloc_h[JO_MAX_COMP]; /* reference allocator grid */
    int16_t *coef[JO_MAX_COMP];
    /* component planes after IDCT + level shift + pixel replication, cropped to
       width x height, UNCLAMPED int16 (what WriteBlock receives). plane-major. */
    int16_t *planes;         /* [ncomp][height][width] */
    /* app-level outputs (apps/JpegDecode): interleaved YCbCr888 and RGB24.
       For ncomp==1 Cb=Cr=128 (DecodeAction.cs:58-66). NULL if ncomp not in {1,3}. */
    uint8_t *ycbcr;          /* [height][width][3] */
    uint8_t *rgb;            /* [height][width][3] */
    size_t consumed;         /* bytes consumed up to and including EOI (or len) */
    char error[160];
    /* sequential frames: one byte per block of the padded grid, 1 = some scan read the block and handed it to
       WriteBlock (...BaselineScanDecoder.cs:119-134).  Blocks behind an EOI that sits on a restart boundary
       (:144-150) and components no scan names are never written: their samples in `planes` keep the 0 of a fresh
       buffer.  NULL for progressive / lossless frames. */
    uint8_t *written[JO_MAX_COMP];
} jo_image;

/* flags for jo_decode */
#define JO_WANT_COEF 1
#define JO_WANT_PLANES 2
#define JO_WANT_RGB 4
#define JO_WANT_ALL 7

/* Decode one JPEG (SOF0/SOF1/SOF2 Huffman).  Returns JO_OK or an error code;
   img->error holds the message.  Always call jo_free(img) afterwards. */
int jo_decode(const uint8_t *data, size_t len, int flags, jo_image *img);
/* The same behind JpegDecoder.LoadTables(tables) (JpegDecoder.cs:313-360): an abbreviated stream whose DHT / DQT / DRI
   segments live in a separate tables stream. */
int jo_decode_with_tables(const uint8_t *tables, size_t tables_len, const uint8_t *data, size_t len, int flags,
                          jo_image *img);
void jo_free(jo_image *img);

/* Stand-alone block math (used by unit tests of the kernels). */
/* D6+D7+D8: dequantise zig-zag block, fp32 IDCT, round-half-even, + level shift. */
void jo_dequant_idct_block(const int16_t coef_zz[64], const uint16_t q_zz[64], int level_shift,
                           int16_t out[64]);
/* D11: apps/JpegDecode/JpegYCbCrToRgbConverter.cs:171-205 for n pixels. */
void jo_ycbcr_to_rgb(const uint8_t *ycbcr, uint8_t *rgb, size_t n);
/* E1: apps/JpegEncode/JpegRgbToYCbCrConverter.cs:64-93 for n pixels. */
void jo_rgb_to_ycbcr(const uint8_t *rgb, uint8_t *ycbcr, size_t n);

/* Decode a batch with `threads` POSIX threads (one image per task) producing only
   RGB24 into caller buffers -- the CPU baseline of bench.py.  out[i] must hold
   3*W*H bytes (may be NULL to discard).  Returns number of failed images. */
int jo_decode_batch_rgb(const uint8_t *const *data, const size_t *len, int n, int threads,
                        uint8_t *const *out);

/* ------------------------------------------------------------------ encoder */
typedef struct {
    int width, height;
    int ncomp;
    int h[JO_MAX_COMP], v[JO_MAX_COMP];      /* sampling factors */
    int tq[JO_MAX_COMP], td[JO_MAX_COMP], ta[JO_MAX_COMP];
    uint16_t qt[4][64];                      /* zig-zag order, as SetQuantizationTable */
    int qt_present[4];
    int optimize;                            /* 1: optimised Huffman (config 5); 2: MostOptimalCoding (package merge) */
} jo_encode_params;

typedef struct {
    uint8_t *bytes; size_t len;              /* full JPEG stream */
    /* coefficient store as the reference's allocator lays it out (no dummy):
       per component alloc_w x alloc_h blocks, zig-zag */
    int alloc_w[JO_MAX_COMP], alloc_h[JO_MAX_COMP];
    int16_t *coef[JO_MAX_COMP];
    uint32_t hist[2][4][256];                /* [class dc=0/ac=1][table id][symbol] */
    uint8_t dht_bits[2][4][16];
    uint8_t dht_vals[2][4][256];
    int dht_nvals[2][4];
    size_t scan_offset, scan_len;            /* entropy-coded bytes inside `bytes` */
    char error[160];
    /* the allocator's dummy block (JpegBlockAllocator.cs:73-78,108-111) as the scan write finds it: every
       MCU-padding block aliases it, so it holds what the LAST padding block of TransformBlocks left there */
    int16_t dummy[64];
} jo_encoded;

/* Annex-K tables scaled like JpegStandardQuantizationTable.ScaleByQuality (:64-89). */
void jo_std_quant_table(int chroma, int quality, uint16_t out_zz[64]);
/* Encode interleaved YCbCr888 (or gray when ncomp==1) exactly like JpegEncoder.Encode()
   fed by apps/JpegEncode/JpegBufferInputReader.  */
int jo_encode_ycbcr(const uint8_t *ycbcr, const jo_encode_params *p, jo_encoded *out);
void jo_encoded_free(jo_encoded *e);
/* Optimised table construction from a histogram: BuildUsingStandardMethod
   (JpegHuffmanEncodingTableBuilder.cs:69-176). Returns number of symbols. */
int jo_build_huffman_table(const uint32_t freq[256], uint8_t bits[16], uint8_t vals[256]);
/* The same with MostOptimalCoding = true: BuildUsingPackageMerge (JpegHuffmanEncodingTableBuilder.cs:287-413). */
int jo_build_huffman_table_optimal(const uint32_t freq[256], uint8_t bits[16], uint8_t vals[256]);

#ifdef __cplusplus
}
#endif
#endif
