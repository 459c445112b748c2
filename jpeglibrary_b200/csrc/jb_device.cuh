// jb_device.cuh -- device-side data structures shared by the decode kernels.
//
// Data layout in HBM (see DESIGN.md "Data layout"):
//   * compressed bytes: one arena, every image's entropy-coded segment starts 256-B aligned;
//   * restart-marker index: uint32 per marker, (byte position << 4) | (RSTn index, or 8 for a terminator);
//   * coefficient store: int16 blocks of 64 in ZIG-ZAG order with absolute DC.  Baseline frames use
//     MCU scan order [mcu][block-in-mcu][64]; progressive frames use per-component planes of
//     MCU-padded block grids (non-interleaved scans address blocks by component coordinates);
//   * pixels: caller-provided pitch-linear buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define JB_LUT_BITS 10
#define JB_LUT_SIZE (1 << JB_LUT_BITS)
#define JB_MAX_BLOCKS_PER_MCU 10
#define JB_MAX_TABLE_SLOTS 8

// Huffman decoding table, device form.  Built on the host by *simulating* the reference's
// JpegHuffmanDecodingTable.Lookup/LookupSlow (JpegHuffmanDecodingTable.cs:73-113) for every
// 10-bit prefix, so every code -- valid or not -- resolves exactly as in the reference.
#define JB_LUT2_SUBTABLES 8
struct __align__(16) JbHuffTable {
    uint16_t lut[JB_LUT_SIZE]; // (symbol << 8) | code size; size 0 => escape: high byte = 1 + second-level
                               // sub-table (indexed by the remaining 6 bits), or 0 => slow path
    uint16_t lut2[JB_LUT2_SUBTABLES * 64]; // (symbol << 8) | code size; 0 => slow path
    uint16_t maxcode[20];      // reference _maxCode[0..17] (left-aligned 16-bit), padded
    uint8_t valoffset[24];     // reference _valOffset[0..18], padded
    uint8_t values[256];
    uint32_t cls;              // 0 = DC, 1 = AC
    uint32_t pad[3];
};
static_assert(sizeof(JbHuffTable) % 16 == 0, "table must be copyable as uint4");

struct __align__(16) JbDevImage {
    // flat restart-segment decode (K0b/K1): per block-in-mcu x = DC table, y = AC table (32-bit word offsets into
    // the JbHuffTable32 array), z = component
    uint4 binfo[JB_MAX_BLOCKS_PER_MCU];
    uint32_t seg_base;   // global index of the image's first restart segment in the batch
    uint32_t tmap_shift; // log2 of the element size of the output tensor map (x coordinates are in elements)
    uint64_t tmap_ptr;   // device address of the CUtensorMap of the pixel output (0: no TMA tensor stores)
    // compressed input
    uint64_t data_off;   // offset of the entropy-coded bytes in the device arena (256-B aligned)
    uint32_t data_len;   // upper bound of entropy-coded length (bytes)
    // restart structure
    uint32_t dri;        // MCUs per restart interval (0: none)
    uint32_t nseg;       // number of restart segments = ceil(total_mcus / dri) (1 if dri == 0)
    uint32_t mark_base;  // first entry of this image in the marker index
    uint32_t mark_cap;   // entries reserved (nseg + 1)
    // geometry
    uint32_t total_mcus, mcus_per_line, mcus_per_col;
    uint16_t width, height;
    uint8_t ncomp, precision, bpm, sof;
    uint8_t hmax, vmax;
    uint8_t comp_h[4], comp_v[4];
    uint8_t comp_blk_off[4];            // first block-in-mcu of each component (interleaved layout)
    uint8_t blk_comp[JB_MAX_BLOCKS_PER_MCU]; // component of block-in-mcu b
    uint8_t blk_dc[JB_MAX_BLOCKS_PER_MCU];   // table slot (0..7) of block b
    uint8_t blk_ac[JB_MAX_BLOCKS_PER_MCU];
    uint16_t table_index[JB_MAX_TABLE_SLOTS]; // slot -> index into the device table array
    uint8_t ntables;
    uint8_t seq_dri;     // scan-list frame of a SEQUENTIAL process whose scans have restart intervals: a scan may end at an EOI
                         // on a restart boundary and leave the intervals behind it unwritten (per-component MCU limits)
    uint8_t pad0[2];
    // self-synchronising decode (scans without restart markers)
    uint32_t use_selfsync; // 1: K1b path
    uint32_t sub_base;     // first sub-sequence slot of this image in the sub-sequence arrays
    uint32_t sub_cap;      // slots reserved
    uint32_t chunk_base;   // first 64 KB un-stuff chunk of this image in the per-chunk counters
    // progressive frames: per-scan descriptors and a planar coefficient store
    uint32_t scan_base, nscans;     // JbDevScan entries of this image
    uint32_t planar;                // 1: per-component planes of MCU-padded block grids
    uint32_t comp_plane_off[4];     // first block of each component plane (relative to coef_off)
    uint32_t comp_plane_w[4];       // blocks per row of each plane
    uint32_t covered;               // scan-list frames: bit c = some scan names component c.  The reference's sequential
                                    // decoder never calls WriteBlock for the others: their samples stay 0
    // lossless (SOF3): planes reuse comp_plane_off (x64 samples) / comp_plane_w (samples per row); every SOS is a JbDevScan
    // (scan_base, nscans; ss = predictor selection, al = point transform); ll_comp_scan[c] = the LAST scan that names
    // component c (0xFF: none) -- the reference decodes scan by scan into the same scanline store, later scans win
    uint8_t ll_comp_scan[4];
    int32_t ll_reserved;
    // coefficient store
    uint64_t coef_off;   // first block of this image in the coefficient store (in blocks)
    uint32_t quant_off;  // first of ncomp quant tables (64 x uint16 each) in the quant array
    // output
    uint64_t out_ptr;    // device address of the image's pixels
    uint64_t out_pitch;
    int32_t out_format;
    int32_t pad1;
};

// byte range K0 indexes: one per sequential image, one per scan of a progressive image
struct JbScanRange {
    uint64_t data_off;
    uint32_t data_len;
    uint32_t mark_base, mark_cap;
    uint32_t pad;
};

// one SOS of a progressive frame (JpegHuffmanProgressiveScanDecoder.ProcessScan, :57-90)
#define JB_PROG_MAX_DEPS 6
struct JbDevScan {
    uint64_t data_off;   // arena offset of the scan's entropy-coded bytes
    uint32_t data_len;
    uint32_t range;      // index into the K0 ranges / results
    uint32_t mark_base;
    uint32_t dri, nseg;
    uint32_t nunits;     // MCUs (interleaved scan) or blocks (single-component scan)
    uint32_t wb, hb;     // single-component scans: block grid of the component (:146-147)
    uint8_t ncomp, ss, se, ah, al;
    uint8_t comp[4];
    uint8_t level;       // dependency level: scans of one level touch disjoint (component, band) sets
    uint8_t ndep;        // producers: earlier scans of the image that share a component and overlap in band
    uint8_t dep_all;     // bit i: dep[i] must be COMPLETE before this scan starts (its unit order differs from mine);
                         // otherwise unit u only needs the producer's units 0..u (same component(s), same order)
    uint8_t has_consumer; // a later scan reads what this one writes: publish progress
    uint8_t seq;         // scan of a SEQUENTIAL frame decoded through the scan list (several scans, or a scan that does
                         // not name every component once): whole blocks, walked MCU by MCU with the component's own
                         // h x v whatever the number of components (reference quirk Q2)
    uint16_t dep[JB_PROG_MAX_DEPS]; // scan indices inside the image; ndep == 0xFF: wait for every earlier scan
    uint16_t dc_tab[4], ac_tab[4]; // device table indices, 0xFFFF = not defined
};

// K1c work.  A lane entry is one restart segment of one scan; a job is one warp: either ONE segment of an AC
// refinement scan, decoded by the whole warp, or up to 32 entries of other scans (of different images, same place
// in their scan scripts), one per lane.  Warps take jobs in list order through a ticket counter, and the list holds
// producers in front of their consumers, so a waiting warp's producers are always running or done.
struct JbProgLane {
    uint32_t image;  // batch index
    uint32_t scan;   // scan index inside the image
    uint32_t seg;    // restart segment
    uint32_t pad;
};
struct JbProgJob {
    uint32_t first;  // first lane entry
    uint32_t lanes;  // entries (1..32); 1 for a whole-warp job
    uint32_t coop;   // whole-warp AC refinement
    uint32_t pad;
};

// per-range result of K0 (restart scan)
struct JbScanResult {
    uint32_t nmarkers;   // entries written to the marker index (RSTn + at most one terminator)
    uint32_t end_pos;    // position of the terminator marker (or data_len)
    uint32_t end_marker; // terminator marker byte (0 if none found)
    uint32_t pad;
};

// work item of the IDCT+colour kernel: a CTA renders a strip of consecutive MCUs of one MCU row
struct JbTileWork {
    uint32_t image;
    uint32_t mcu_row;
    uint32_t mcu_col0;
    uint32_t nmcu;
};

// status bits written by kernels (per image)
#define JB_ST_BAD_CODE 1u        // invalid Huffman code / magnitude category  -> InvalidDataException
#define JB_ST_PREMATURE_END 2u   // ran out of bits inside a segment            -> InvalidDataException
#define JB_ST_EXPECT_RST 4u      // restart marker missing / misplaced          -> InvalidOperationException
#define JB_ST_STALLED 8u         // K1c gave up waiting for a producer scan (never expected) -> JB_ERR_CUDA

// Which exception does the reference throw for a stream with several defects?  The one it meets FIRST: it decodes scan by
// scan, interval by interval, and stops at the first failure.  The kernels decode everything at once and OR their status
// bits per image, so next to the bits every failure also takes part in an atomicMin on a key that orders it in the
// stream: (scan << 20 | 2 * interval + after) << 2 | class, `after` = 1 for what is checked behind an interval's data
// (the restart marker), class 1 = InvalidDataException (bad code, premature end), 2 = InvalidOperationException
// ("Expect restart marker.").  The host reports the class of the smallest key.
__device__ __forceinline__ void jb_report_error(uint32_t *status, uint32_t *first_error, uint32_t image, uint32_t err,
                                                uint32_t scan, uint32_t interval)
{
    if (!err) return;
    atomicOr(status + image, err);
    if (!first_error) return;
    const bool data = (err & (JB_ST_BAD_CODE | JB_ST_PREMATURE_END)) != 0; // (met inside the interval: before its end check)
    const uint32_t unit = min(2u * interval + (data ? 0u : 1u), (1u << 20) - 1u);
    atomicMin(first_error + image, ((min(scan, 1023u) << 20 | unit) << 2) | (data ? 1u : 2u));
}

