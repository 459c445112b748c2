// k_lossless.cuh -- K1d/K5: lossless (SOF3) Huffman frames.
//
// Replaces JpegHuffmanLosslessScanDecoder.ProcessScan (ScanDecoder/JpegHuffmanLosslessScanDecoder.cs:52-205),
// ReadSampleLossless (:207-223) and the JpegPartialScanlineAllocator flush with pixel replication
// (JpegPartialScanlineAllocator.cs:103-220).
//
//   jb_k1d_lossless_entropy  one thread per (image, scan, restart segment): Huffman-decode the difference of every
//                            sample of the segment's MCUs into the component planes (mod 2^16, as the reference's
//                            (short) cast makes the reconstruction arithmetic modular).  A frame may hold several
//                            scans (one per component is the usual non-interleaved form): launched scan index by
//                            scan index, so that a later scan over the same component replaces the earlier one
//   jb_k1d_lossless_predict  one warp per (image, component): raster-order reconstruction with the reference's
//                            predictor selection rules, which are a pure function of the sample position:
//                            first MCU row or first MCU after a restart -> 1-D/initial rules (:109-139),
//                            first MCU column -> Rb (:141-144), else predictor Ss (:145-159)
//   jb_k5_lossless_output    replicate to full resolution + the same sinks as K2 (planar int16 / YCbCr888 / RGB)
#pragma once
#include "jb_device.cuh"
#include "k_entropy_decode.cuh"
#include "k_idct_color.cuh"

#define JB_MAX_COMPONENTS_DEV 4
#ifndef JB_K1D_PREFETCH
#define JB_K1D_PREFETCH 8 // steps of the reconstruction wavefront whose inputs are loaded at once (1: a load per step)
#endif

__global__ void __launch_bounds__(32)
jb_k1d_lossless_entropy(const JbDevImage *__restrict__ images, const JbDevScan *__restrict__ scans,
                        const uint32_t *__restrict__ image_list, uint32_t scan_index,
                        const JbHuffTable *__restrict__ tables, const uint8_t *__restrict__ arena,
                        const uint32_t *__restrict__ marks, const JbScanResult *__restrict__ scanres,
                        int16_t *__restrict__ store, uint32_t *__restrict__ status, int lanes_per_warp,
                        uint32_t *__restrict__ first_error)
{
    const uint32_t image = image_list[blockIdx.y];
    const JbDevImage &im = images[image];
    const int lane = threadIdx.x;
    if (lane >= lanes_per_warp || scan_index >= im.nscans) return;
    const JbDevScan &sc = scans[im.scan_base + scan_index];
    const uint32_t seg = blockIdx.x * lanes_per_warp + lane;
    if (seg >= sc.nseg) return;
    const JbScanResult sr = scanres[sc.range];
    const uint32_t *mk = marks + sc.mark_base;
    // the bit reader wants a 4-byte aligned base: the image's (256-byte aligned) arena slot, scan-relative positions shifted
    const uint8_t *data = arena + im.data_off;
    const uint32_t rel = (uint32_t)(sc.data_off - im.data_off);
    const uint32_t per_seg = sc.dri ? sc.dri : im.total_mcus;
    const uint32_t first = seg * per_seg, count = min(per_seg, im.total_mcus - first);
    uint32_t start = 0, err = 0;
    if (seg > 0) {
        if (seg - 1 < sr.nmarkers && (mk[seg - 1] & 8u) == 0) start = (mk[seg - 1] >> 4) + 2;
        else {
            // no RSTn in front of this interval.  EOI at a restart boundary ends the scan quietly (:172-176): the
            // intervals behind it are simply not there; any other marker is the previous interval's error
            if (sr.end_marker != 0xD9u) jb_report_error(status, first_error, image, JB_ST_EXPECT_RST, scan_index, seg - 1);
            return;
        }
    }
    const uint32_t stop = (seg < sr.nmarkers ? (mk[seg] >> 4) : sr.end_pos) + rel;
    start += rel;
    JbBitReader br;
    br.init(data, start, stop);
    int16_t *base = store + im.coef_off * 64;
    const uint8_t *tab_base = reinterpret_cast<const uint8_t *>(tables);
    for (uint32_t u = first; u < first + count && !err; u++) {
        const uint32_t row = u / im.mcus_per_line, col = u - row * im.mcus_per_line;
        for (int i = 0; i < sc.ncomp && !err; i++) { // the scan's components in scan order, h x v samples each (:86-100)
            const int c = sc.comp[i];
            const int h = im.comp_h[c], v = im.comp_v[c];
            const JbHuffTable *tab = reinterpret_cast<const JbHuffTable *>(tab_base + (size_t)sc.dc_tab[i] * sizeof(JbHuffTable));
            int16_t *plane = base + (size_t)im.comp_plane_off[c] * 64;
            for (int s = 0; s < h * v; s++) {
                const int x = s % h, y = s / h;
                br.ensure32();
                uint32_t e = jb_huff_lookup(tab, br.peek16());
                if (e == 0xFFFFFFFFu) { err |= JB_ST_BAD_CODE; break; }
                br.skip_code(e & 0xFF);
                const int t = (int)(e >> 8);
                int d = 0;
                if (t == 16) d = 32768;
                else if (t > 16) { err |= JB_ST_BAD_CODE; break; }
                else if (t != 0) d = jb_extend((int)br.take(t), t);
                plane[(size_t)(row * v + y) * im.comp_plane_w[c] + col * h + x] = (int16_t)d;
            }
        }
    }
    if (br.n < br.pad) err |= JB_ST_PREMATURE_END;
    if (!err && sc.dri != 0 && count == per_seg) {
        const int real = br.n - br.pad;
        uint32_t p = br.pos;
        while (p < stop && data[p] == 0xFF) p++;
        bool marker_ok = seg < sr.nmarkers;
        if (marker_ok && (mk[seg] & 8u) != 0) marker_ok = sr.end_marker == 0xD9u;
        if (real >= 8 || p < stop || !marker_ok) err |= JB_ST_EXPECT_RST;
    }
    if (err) jb_report_error(status, first_error, image, err, scan_index, seg);
}

__device__ __forceinline__ int jb_lossless_px(int predictor, int ra, int rb, int rc)
{
    switch (predictor) {
    case 1: return ra;
    case 2: return rb;
    case 3: return rc;
    case 4: return ra + rb - rc;
    case 5: return ra + ((rb - rc) >> 1);
    case 6: return rb + ((ra - rc) >> 1);
    case 7: return (ra + rb) >> 1;
    default: return 0;
    }
}

// One warp per (image, component).  The predictors only look at the left, upper and upper-left neighbours, so
// 32 consecutive rows are reconstructed as a wavefront: lane l handles row r0 + l and runs l columns behind lane
// l - 1, whose last two outputs arrive by shuffle (Rb, Rc); Ra is the lane's own previous output.
__global__ void __launch_bounds__(32 * JB_MAX_COMPONENTS_DEV)
jb_k1d_lossless_predict(const JbDevImage *__restrict__ images, const JbDevScan *__restrict__ scans,
                        const uint32_t *__restrict__ image_list,
                        const uint32_t *__restrict__ marks, const JbScanResult *__restrict__ scanres,
                        int16_t *__restrict__ store)
{
    const uint32_t image = image_list[blockIdx.x];
    const JbDevImage &im = images[image];
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (c >= im.ncomp) return;
    const int h = im.comp_h[c], v = im.comp_v[c];
    const int w = (int)im.comp_plane_w[c], rows = (int)im.mcus_per_col * v, cols = (int)im.mcus_per_line * h;
    int16_t *plane = store + im.coef_off * 64 + (size_t)im.comp_plane_off[c] * 64;
    if (im.ll_comp_scan[c] == 0xFF) { // in no scan: JpegPartialScanlineAllocator's zeros (the store is not cleared)
        const int vs = im.vmax / v, hc = ((int)im.height + vs - 1) / vs;
        for (int i = lane; i < w * hc; i += 32) plane[i] = 0;
        return;
    }
    // the parameters of the scan that decoded this component (the last one that names it)
    const JbDevScan &sc = scans[im.scan_base + im.ll_comp_scan[c]];
    const int predictor = sc.ss;
    // 1 << (P - Pt - 1) as C# evaluates it (JpegHuffmanLosslessScanDecoder.cs:81): the shift count is taken modulo 32, so a
    // damaged Pt >= P yields a value whose low 16 bits are 0 instead of an error
    const int initial = (int)(1u << ((im.precision - sc.al - 1) & 31));
    const uint32_t dri = sc.dri, mpl = im.mcus_per_line;
    // MCUs of intervals that are not in the stream (EOI at a restart boundary) keep the allocator's zeros
    uint32_t nrst = scanres[sc.range].nmarkers;
    if (nrst && (marks[sc.mark_base + nrst - 1] & 8u)) nrst--;
    const uint32_t valid = dri ? min(im.total_mcus, (nrst + 1u) * dri) : im.total_mcus;
    for (int r0 = 0; r0 < rows; r0 += 32) {
        const int cy = r0 + lane;
        const bool rowok = cy < rows;
        int16_t *line = plane + (size_t)cy * w;
        const int row = cy / v, y = cy - row * v;
        int ra = 0, rb = 0, rc = 0, out = 0;
        // position of the lane's next sample, kept incrementally (a division per step was most of a step's instructions):
        // MCU column, sample inside the MCU, MCU number and its place inside the restart interval
        int col = 0, x = 0;
        uint32_t mcu = (uint32_t)row * mpl, in_interval = dri ? mcu % dri : 1u;
        // The differences of the lane's row and (lane 0) the row above the band do not depend on the reconstruction: they
        // are fetched JB_K1D_PREFETCH steps at a time, so that a step of the serial chain no longer waits for two L2 round
        // trips (measured on 64 frames of 1024 x 1024 x 3: one load per step 23.6 ms, eight at a time 11.6 ms, and 7.2 ms with
        // the sample position kept incrementally; profiles/r2f_lossless_launches_*.csv).
        for (int t0 = 0; t0 < cols + 31; t0 += JB_K1D_PREFETCH) {
            int dv[JB_K1D_PREFETCH], av[JB_K1D_PREFETCH];
#pragma unroll
            for (int j = 0; j < JB_K1D_PREFETCH; j++) {
                const int cx = t0 + j - lane;
                const bool act = rowok && cx >= 0 && cx < cols;
                dv[j] = act ? (int)line[cx] : 0;
                av[j] = (lane == 0 && act && cy > 0) ? (int)__ldcg(line - w + cx) : 0; // last row of the previous band
            }
#pragma unroll
            for (int j = 0; j < JB_K1D_PREFETCH; j++) {
                int above = __shfl_up_sync(0xFFFFFFFFu, out, 1);
                const int cx = t0 + j - lane;
                const bool act = rowok && cx >= 0 && cx < cols;
                if (lane == 0) above = av[j];
                rc = rb;
                rb = above;
                if (act) {
                    const bool after_restart = in_interval == 0; // (dri == 0: never)
                    int pred;
                    if (row == 0 || after_restart) { // :109-139
                        if (col == 0 && x == 0) pred = initial;
                        else pred = jb_lossless_px(predictor, ra, y == 0 ? initial : rb, y == 0 ? initial : rc);
                    } else if (col == 0) {           // :141-144
                        pred = rb;
                    } else {                         // :145-159
                        pred = jb_lossless_px(predictor, ra, rb, rc);
                    }
                    out = mcu < valid ? (int16_t)(dv[j] + pred) : 0;
                    line[cx] = (int16_t)out;
                    ra = out;
                    if (++x == h) {
                        x = 0;
                        col++;
                        mcu++;
                        if (++in_interval == dri) in_interval = 0;
                    }
                }
            }
        }
        __syncwarp();
    }
}

// one thread per output pixel
__global__ void __launch_bounds__(256)
jb_k5_lossless_output(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
                      const int16_t *__restrict__ store)
{
    const JbDevImage &im = images[image_list[blockIdx.y]];
    const int W = im.width, H = im.height;
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= (uint32_t)W * H) return;
    const int yy = i / W, x = i - yy * W;
    int s[4] = {0, 0, 0, 0};
    for (int c = 0; c < im.ncomp; c++) {
        const int hs = im.hmax / im.comp_h[c], vs = im.vmax / im.comp_v[c];
        s[c] = store[im.coef_off * 64 + (size_t)im.comp_plane_off[c] * 64 + (size_t)(yy / vs) * im.comp_plane_w[c] + x / hs];
    }
    uint8_t *out = reinterpret_cast<uint8_t *>(im.out_ptr);
    const uint64_t pitch = im.out_pitch;
    const int fmt = im.out_format;
    if (fmt == 3) {
        for (int c = 0; c < im.ncomp; c++) reinterpret_cast<int16_t *>(out + ((uint64_t)c * H + yy) * pitch)[x] = (int16_t)s[c];
        return;
    }
    const int yv = jb_sample_to_u8(s[0], im.precision);
    const int cb = im.ncomp == 3 ? jb_sample_to_u8(s[1], im.precision) : 128;
    const int cr = im.ncomp == 3 ? jb_sample_to_u8(s[2], im.precision) : 128;
    const int bpp = fmt == 1 ? 4 : 3;
    uint8_t *dst = out + (uint64_t)yy * pitch + (uint64_t)x * bpp;
    if (fmt == 2) { dst[0] = (uint8_t)yv; dst[1] = (uint8_t)cb; dst[2] = (uint8_t)cr; return; }
    int r, g, b;
    jb_ycc_to_rgb(yv, cb, cr, r, g, b);
    dst[0] = (uint8_t)r; dst[1] = (uint8_t)g; dst[2] = (uint8_t)b;
    if (bpp == 4) dst[3] = 255;
}
