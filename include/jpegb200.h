/*
 * jpegb200.h -- C-ABI of the B200-native JPEG hot path (libjpegb200.so).
 *
 * This is the drop-in boundary for yigolden/JpegLibrary's per-block hot path
 * (SURVEY.md section 8b).  The reference has no FFI of its own (it is 100 % managed C#);
 * the entry points below are what a `JpegLibrary.Cuda` P/Invoke layer binds so that
 *
 *   JpegScanDecoder.ProcessScan            (src/JpegLibrary/ScanDecoder/JpegScanDecoder.cs:14)
 *   JpegHuffmanBaselineScanDecoder         (ScanDecoder/JpegHuffmanBaselineScanDecoder.cs:51-268)
 *   JpegHuffmanProgressiveScanDecoder      (ScanDecoder/JpegHuffmanProgressiveScanDecoder.cs:57-470)
 *   JpegHuffmanLosslessScanDecoder         (ScanDecoder/JpegHuffmanLosslessScanDecoder.cs:52-223)
 *   JpegBlockOutputWriter.WriteBlock sinks (JpegBlockOutputWriter.cs:17; app sinks
 *                                           apps/JpegDecode/JpegBufferOutputWriter8Bit.cs:28-60,
 *                                           apps/JpegDecode/JpegYCbCrToRgbConverter.cs:171-205)
 *   JpegEncoder.TransformBlocks / BuildHuffmanTables / WritePreparedScanData
 *                                          (JpegEncoder.cs:414-483, 491-597, 605-656)
 *
 * run on the GPU.  Marker and header parsing stays on the host (JpegDecoder.Identify /
 * the marker loop JpegDecoder.cs:509-617); the host hands parsed headers down in
 * jb_image_desc.  The library never parses markers except restart-marker discovery.
 *
 * Conventions: plain C, blittable structs, status-code returns (0 = ok, negative
 * codes map 1:1 to the reference's exception classes), no exceptions or aborts
 * across the ABI, no global state besides an explicit context.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with
 * JB_ERR_NO_DEVICE.
 */
#ifndef JPEGB200_H
#define JPEGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_API __attribute__((visibility("default")))

/* ---- status codes (reference exception each one maps to) ------------------ */
#define JB_OK 0
#define JB_ERR_INVALID_DATA (-1)      /* InvalidDataException  (JpegDecoder.cs:365-375)          */
#define JB_ERR_INVALID_OPERATION (-2) /* InvalidOperationException, incl. "Expect restart marker."
                                         (ScanDecoder/JpegHuffmanBaselineScanDecoder.cs:151-154) */
#define JB_ERR_NOT_SUPPORTED (-3)     /* NotSupportedException (JpegDecoder.cs:627-630)          */
#define JB_ERR_ARGUMENT (-4)          /* ArgumentException / ArgumentOutOfRange / ArgumentNull   */
#define JB_ERR_NO_DEVICE (-5)         /* no CUDA device / driver: there is no CPU fallback       */
#define JB_ERR_CUDA (-6)              /* CUDA runtime failure; see jb_last_error                 */
#define JB_ERR_NOMEM (-7)

#define JB_MAX_COMPONENTS 4

/* ---- output formats of the decode path ------------------------------------ */
#define JB_OUT_RGB24 0      /* D9+D10+D11: replicate, clamp, YCbCr->RGB; 3 B/pixel             */
#define JB_OUT_RGBA32 1     /* same, A = 255 (JpegYCbCrToRgbConverter.cs:134-169)              */
#define JB_OUT_YCBCR888 2   /* D9+D10 only: what JpegBufferOutputWriter8Bit leaves in memory   */
#define JB_OUT_PLANAR_I16 3 /* D9 only: unclamped int16 component planes [c][H][W] -- exactly the
                               samples WriteBlock receives; the compatibility path replays
                               WriteBlock calls from it on the host                            */
#define JB_OUT_COEFFICIENTS 4 /* D4/P3-P5 only: zig-zag int16 blocks with absolute DC, in the device
                                 store layout described by jb_coef_layout                       */

/* DHT as it appears in the stream (JpegHuffmanDecodingTable.TryParse :249-291) */
typedef struct jb_huff_spec {
    uint8_t table_class; /* 0 = DC, 1 = AC */
    uint8_t identifier;  /* 0..3 */
    uint8_t bits[16];
    uint8_t values[256];
    uint16_t value_count;
} jb_huff_spec;

/* One SOS (JpegScanHeader.cs) with the tables in force when it was met. */
typedef struct jb_scan_desc {
    uint8_t component_count;                       /* Ns */
    uint8_t component_index[JB_MAX_COMPONENTS];    /* index into the frame's components */
    int16_t dc_table[JB_MAX_COMPONENTS];           /* index into jb_image_desc.tables, -1 = undefined */
    int16_t ac_table[JB_MAX_COMPONENTS];
    uint8_t ss, se, ah, al;
    uint32_t restart_interval;                     /* DRI in force (JpegDecoder.cs:635-650) */
    uint64_t entropy_offset;                       /* first entropy-coded byte, relative to data */
    uint64_t entropy_length;                       /* bytes up to (not including) the next non-RST marker */
} jb_scan_desc;

/* One image: everything JpegDecoder.Identify + the marker loop know before the hot path. */
typedef struct jb_image_desc {
    const uint8_t *data; /* host pointer (pinned memory from jb_pinned_alloc avoids a staging copy) */
    uint64_t length;
    uint8_t sof;         /* 0 baseline, 1 extended, 2 progressive, 3 lossless (JpegMarker.StartOfFrame0..3) */
    uint8_t precision;   /* P */
    uint8_t component_count;
    uint8_t reserved0;
    uint16_t width, height;
    uint8_t h[JB_MAX_COMPONENTS], v[JB_MAX_COMPONENTS];
    uint16_t quant[JB_MAX_COMPONENTS][64]; /* zig-zag order; the table each component is rendered with */
    uint32_t scan_count;
    const jb_scan_desc *scans;
    uint32_t table_count;
    const jb_huff_spec *tables;
} jb_image_desc;

/* Where one image's result goes. */
typedef struct jb_output_desc {
    void *dst;          /* device pointer if on_device, else host pointer (pinned preferred) */
    uint64_t pitch;     /* bytes per pixel row (plane row for PLANAR_I16); 0 = tightly packed */
    uint64_t capacity;  /* bytes available at dst; mandatory (0 is refused): "Destination buffer is too small." */
    int32_t format;     /* JB_OUT_* */
    int32_t on_device;
} jb_output_desc;

/* Layout of the device coefficient store of one image (JB_OUT_COEFFICIENTS). */
typedef struct jb_coef_layout {
    int32_t interleaved;            /* 1: [mcu][block-in-mcu][64] scan order (baseline);
                                       0: per-component planes of MCU-padded block grids (progressive) */
    int32_t mcus_per_line, mcus_per_column, blocks_per_mcu;
    int32_t comp_block_offset[JB_MAX_COMPONENTS]; /* interleaved: first block-in-mcu index of comp;
                                                     planar: first block of the component's plane */
    int32_t comp_blocks_w[JB_MAX_COMPONENTS], comp_blocks_h[JB_MAX_COMPONENTS];
    uint64_t total_blocks;
} jb_coef_layout;

typedef struct jb_ctx jb_ctx;
typedef struct jb_batch jb_batch;

/* ---- context ---------------------------------------------------------------- */
JB_API int jb_device_count(void);
JB_API int jb_ctx_create(int device, jb_ctx **out);
JB_API void jb_ctx_destroy(jb_ctx *ctx);
/* Message of the last failure on this context (image index + detail). */
JB_API const char *jb_last_error(jb_ctx *ctx);
/* cudaStream_t the context launches its kernels on (for CUDA-event timing by the caller). */
JB_API void *jb_ctx_stream(jb_ctx *ctx);
JB_API int jb_ctx_synchronize(jb_ctx *ctx);
/* Returns the device memory the context's allocation pool has cached from destroyed batches to the driver
   (batches allocate stream-ordered from a per-context pool that otherwise keeps everything for the next batch). */
JB_API int jb_ctx_trim(jb_ctx *ctx);

/* ---- memory ------------------------------------------------------------------- */
JB_API int jb_pinned_alloc(jb_ctx *ctx, size_t bytes, void **out);
JB_API int jb_pinned_free(jb_ctx *ctx, void *p);
JB_API int jb_device_alloc(jb_ctx *ctx, size_t bytes, void **out);
JB_API int jb_device_free(jb_ctx *ctx, void *p);
JB_API int jb_memcpy_d2h(jb_ctx *ctx, void *dst_host, const void *src_device, size_t bytes);
JB_API int jb_memcpy_h2d(jb_ctx *ctx, void *dst_device, const void *src_host, size_t bytes);

/* ---- decode path: replaces JpegScanDecoder.ProcessScan + Dispose for a batch ---- */
/* Validates the descriptors, builds device Huffman/quant tables, allocates the device
   input + coefficient store.  Nothing is copied or launched yet. */
JB_API int jb_decode_batch_create(jb_ctx *ctx, const jb_image_desc *images, const jb_output_desc *outputs,
                                  int count, jb_batch **out);
/* H2D of the compressed bytes (cudaMemcpyAsync on the context stream). */
JB_API int jb_decode_batch_upload(jb_batch *b);
/* Launches the decode kernels (restart scan, Huffman, IDCT+colour) on the context stream;
   results land in device memory (the outputs' dst when on_device, else a device staging area). */
JB_API int jb_decode_batch_launch(jb_batch *b);
/* D2H of results for outputs with on_device == 0, then waits and collects per-image status. */
JB_API int jb_decode_batch_finish(jb_batch *b);
/* upload + launch + finish. Returns the first non-zero per-image status (or 0). */
JB_API int jb_decode_batch_run(jb_batch *b);
/* Per-image status codes after finish (JB_OK / JB_ERR_*). */
JB_API int jb_decode_batch_status(jb_batch *b, int32_t *status, int count);
JB_API int jb_decode_batch_coef_layout(jb_batch *b, int image, jb_coef_layout *out);
/* Number of kernels launched by the last jb_decode_batch_launch. */
JB_API int jb_decode_batch_launch_count(jb_batch *b);
/* Per-kernel timing with CUDA events on the context stream.  While profiling is on, every
   jb_decode_batch_launch brackets each kernel with event records (no host synchronisation);
   jb_decode_batch_profile then waits for the stream and returns, per kernel, the AVERAGE duration
   in ms over all launches since profiling was switched on.  Returns the number of kernels. */
JB_API int jb_decode_batch_set_profiling(jb_batch *b, int on);
JB_API int jb_decode_batch_profile(jb_batch *b, char (*names)[48], float *ms, int cap);
/* Progressive frames, jb_decode_batch_set_profiling(b, 2): the schedule of the last launch's scan decoder warps (an
   instrumented instance of the kernel runs instead of the production one), four words per job:
   (image << 32 | scan << 16 | first segment), start ns, end ns, ns spent waiting for producer scans.
   Returns the number of jobs written (all jobs when cap <= 0 and out == NULL is a size query). */
JB_API int jb_decode_batch_scan_trace(jb_batch *b, uint64_t *out, int cap);
JB_API void jb_decode_batch_destroy(jb_batch *b);

/* One-call convenience: create + run + destroy. */
JB_API int jb_decode(jb_ctx *ctx, const jb_image_desc *images, const jb_output_desc *outputs, int count,
                     int32_t *status);

/* ---- stand-alone stages (unit-parity entry points) ------------------------------ */
/* D6-D11 on caller-provided coefficient blocks (device pointers): the IDCT+colour kernel alone.
   `coef` uses the layout jb_coef_layout describes for `image`. */
JB_API int jb_render_from_coefficients(jb_ctx *ctx, const jb_image_desc *image, const int16_t *coef_device,
                                       const jb_output_desc *output);

/* ---- encode path: replaces JpegEncoder.TransformBlocks / BuildHuffmanTables / WritePreparedScanData
   (JpegEncoder.cs:414-483, 491-597, 605-656) for a batch.  The host keeps writing SOI/DQT/SOF0/DHT/SOS/EOI
   (JpegEncoder.Encode :255-290) around the scan bytes this path produces. ------------------------------- */
#define JB_IN_RGB24 0     /* interleaved RGB; converted like apps/JpegEncode/JpegRgbToYCbCrConverter.cs:64-93 */
#define JB_IN_YCBCR888 1  /* interleaved YCbCr as JpegBufferInputReader reads it (apps/JpegEncode/JpegBufferInputReader.cs) */
#define JB_IN_GRAY8 2
#define JB_IN_COEFFICIENTS 3 /* `pixels` is a DEVICE pointer to quantised zig-zag blocks in MCU scan order (the layout
                                JB_OUT_COEFFICIENTS produces): transcoding, JpegOptimizer.Scan/Optimize
                                (JpegOptimizer.cs:72-154, 546-879) without any DCT */

typedef struct jb_encode_desc {
    const void *pixels;   /* host (pinned preferred) or device pointer */
    uint64_t pitch;       /* bytes per row; 0 = tightly packed */
    int32_t on_device;
    int32_t format;       /* JB_IN_* */
    uint16_t width, height;
    uint8_t component_count;                 /* 1 or 3 */
    uint8_t h[JB_MAX_COMPONENTS], v[JB_MAX_COMPONENTS];   /* AddComponent sampling factors (JpegEncoder.cs:175) */
    uint8_t tq[JB_MAX_COMPONENTS], td[JB_MAX_COMPONENTS], ta[JB_MAX_COMPONENTS];
    uint8_t reserved;
    uint16_t restart_interval;               /* JB_IN_COEFFICIENTS only: MCUs per restart interval of the scan that is
                                                re-packed (JpegOptimizer.CopyScanBaseline, JpegOptimizer.cs:772-812:
                                                DC prediction restarts, 1-bit padding, RSTn written between intervals) */
    uint16_t quant[4][64];                   /* SetQuantizationTable, zig-zag order, by identifier */
    uint8_t quant_present[4];
} jb_encode_desc;

typedef struct jb_encode_batch jb_encode_batch;

JB_API int jb_encode_batch_create(jb_ctx *ctx, const jb_encode_desc *images, int count, jb_encode_batch **out);
/* H2D of host pixels, then colour conversion + downsample + FDCT + quantisation (E1-E5) and the symbol
   histograms (E6), all on the context stream. */
JB_API int jb_encode_batch_transform(jb_encode_batch *b);
/* Histograms for a host-side table builder (the reference's JpegHuffmanEncodingTableBuilder):
   out[image][class*4 + id][256].  Synchronises. */
JB_API int jb_encode_batch_histograms(jb_encode_batch *b, uint32_t *out, int count);
/* Either build the optimised tables on the GPU (same algorithm, one thread per table) ... */
JB_API int jb_encode_batch_build_tables(jb_encode_batch *b);
/* ... or install tables built by the host (DHT form). */
JB_API int jb_encode_batch_set_table(jb_encode_batch *b, int image, const jb_huff_spec *table);
/* Bit lengths, prefix sum, parallel bit packing, byte stuffing (E8, E9). */
JB_API int jb_encode_batch_pack(jb_encode_batch *b);
/* Waits; returns the first per-image error. */
JB_API int jb_encode_batch_finish(jb_encode_batch *b);
JB_API int jb_encode_batch_get_table(jb_encode_batch *b, int image, int table_class, int identifier, jb_huff_spec *out);
JB_API int jb_encode_batch_scan_length(jb_encode_batch *b, int image, uint64_t *length);
JB_API int jb_encode_batch_read_scan(jb_encode_batch *b, int image, uint8_t *dst_host, uint64_t capacity);
/* Quantised zig-zag coefficient blocks in MCU scan order (parity checks, JpegOptimizer-style reuse). */
JB_API int jb_encode_batch_read_coefficients(jb_encode_batch *b, int image, int16_t *dst_host, uint64_t capacity_blocks);
JB_API int jb_encode_batch_launch_count(jb_encode_batch *b);
JB_API void jb_encode_batch_destroy(jb_encode_batch *b);
/* The table builder on the host (JpegHuffmanEncodingTableBuilder.Build(optimal: false), :62-176). */
JB_API int jb_build_huffman_table(const uint32_t frequencies[256], int table_class, int identifier, jb_huff_spec *out);
/* The same with MostOptimalCoding = true: JpegHuffmanEncodingTableBuilder.Build(optimal: true), package merge (:287-413).
   Host only: histograms from jb_encode_batch_histograms, tables back in through jb_encode_batch_set_table. */
JB_API int jb_build_huffman_table_optimal(const uint32_t frequencies[256], int table_class, int identifier, jb_huff_spec *out);

/* Host only (no device needed): how a frame that is decoded through the scan list (progressive, or sequential with
   several scans) is planned.  Ten words per scan: its place in the job order, number of producer scans (255 = every
   earlier scan), bit mask of producers that are waited for as a whole (the others are followed block by block), six
   producer scan indices (-1 = unused), 1 if a later scan consumes it.  Returns the number of scans (0 for frames on
   the single-scan fast paths) or JB_ERR_*. */
JB_API int jb_plan_scans(const jb_image_desc *image, int32_t *out, int cap);

JB_API const char *jb_version(void);

#ifdef __cplusplus
}
#endif
#endif
