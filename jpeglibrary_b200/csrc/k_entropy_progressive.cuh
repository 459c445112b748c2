// k_entropy_progressive.cuh -- K1c: every scan of the progressive (SOF2) Huffman frames of a batch, ONE launch.
//
// Replaces JpegHuffmanProgressiveScanDecoder.ProcessScan and its block readers
// (ScanDecoder/JpegHuffmanProgressiveScanDecoder.cs:57-419):
//   interleaved DC scans (:92-138), single-component DC/AC scans (:140-194), HandleRestart (:196-224),
//   ReadBlockProgressiveDC (:227-253), ReadBlockProgressiveAC (:255-311),
//   ReadBlockProgressiveACRefined (:313-419).
// The coefficient store (JpegBlockAllocator, JpegBlockAllocator.cs:35-114) is a planar, MCU-padded
// grid in HBM, zero-initialised per batch; blocks the reference routes to its shared dummy block
// (out-of-range MCU padding, :108-111) simply land in the padding here and are never rendered.
//
// Parallelism.  A scan is a serial bit stream (restart intervals apart) and a refinement scan reads the
// coefficient history earlier scans left, so it cannot be decoded speculatively.  What the format does allow:
//   * scans commute unless they share a component and overlap in band (the host records those producers per scan);
//   * a consumer needs block u of its producer only when it gets to block u itself: dependent scans run
//     CONCURRENTLY, the consumer a few blocks behind the producer.  Producers publish their block count with a
//     release store every JB_K1C_PUBLISH units, consumers poll it (relaxed) and read coefficients past L1 (ld.cg).
// libjpeg's 10-scan script has the chain {Y 1-5, Y 6-63} -> Y refine -> Y refine: instead of three passes in a row
// (43 + 61 + 111 ms for a batch of 1080p frames) the frame takes about as long as its slowest scan.
// One warp per job (JbProgJob): AC refinement scans are decoded by the whole warp (the serial symbol chain of one
// stream is the bound, so its latency is what counts); the other scans are cheap per stream and are packed one per
// lane, 32 images to a warp, so that they do not take warp slots and issue cycles from the refinement warps.  Jobs
// are handed out through a ticket counter in list order -- the host puts producers in front of consumers (and long
// dependency chains first) -- so a waiting warp's producers are always running or done.
#pragma once
#include "jb_device.cuh"
#include "k_entropy_decode.cuh"

__device__ __forceinline__ uint32_t jb_prog_bits(JbBitReader &br, int k)
{ // k in 1..16 raw bits (TryReadBits)
    br.ensure32();
    return br.take(k);
}

#define JB_K1C_PUBLISH 32u        // a single-segment scan publishes its progress every so many units
#define JB_K1C_SPIN_LIMIT (1u << 19) // x 8 us: a producer that does not move for seconds -> JB_ST_STALLED, never a hang

// A consumer polls its producer's progress with a RELAXED load.  An acquire load is a load plus CCTL.IVALL -- it throws the
// SM's whole L1 away, and with it the Huffman tables and stream lines of every other warp of the SM (150 -> 144 ms per
// 1024 frames).  Nothing a producer publishes is ever read through L1 here: coefficients are read with ld.cg, and those
// loads are issued behind the branch that tests the progress value.
__device__ __forceinline__ uint32_t jb_ld_progress(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void jb_st_release(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool TRACE> // TRACE: record the schedule (jb_decode_batch_scan_trace); a separate instance keeps its registers
__global__ void __launch_bounds__(32, 32)   // out of the production kernel
jb_k1c_progressive_scans(const JbDevImage *__restrict__ images, const JbDevScan *__restrict__ scans,
                         const JbProgJob *__restrict__ jobs, const JbProgLane *__restrict__ entries, uint32_t njobs,
                         const JbHuffTable *__restrict__ tables,
                         const uint8_t *__restrict__ arena, const uint32_t *__restrict__ marks,
                         const JbScanResult *__restrict__ scanres, int16_t *coef, uint32_t *__restrict__ status,
                         uint32_t *progress, uint32_t *ticket, unsigned long long *trace, uint32_t *scan_limit,
                         uint32_t *first_error)
{
    const int lane = threadIdx.x;
    uint32_t turn = 0;
    if (lane == 0) turn = atomicAdd(ticket, 1u);
    turn = __shfl_sync(0xFFFFFFFFu, turn, 0);
    if (turn >= njobs) return;
    unsigned long long t_start = 0, t_wait = 0; // profiling only (jb_decode_batch_scan_trace)
    if (TRACE) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    const JbProgJob job = jobs[turn];
    const bool coop = job.coop != 0;
    const bool active = coop || (uint32_t)lane < job.lanes;
    const JbProgLane ent = entries[job.first + (active && !coop ? (uint32_t)lane : 0u)];
    const uint32_t image = ent.image;
    const JbDevImage &im = images[image];
    const JbDevScan sc = scans[im.scan_base + ent.scan]; // by value: its fields are used in every inner loop
    uint32_t *const my_progress = progress + im.scan_base + ent.scan;
    // AC refinement scans are decoded by the whole warp on a shared-memory copy of the current block (prefetched one
    // block ahead): the refinement loop reads and rewrites the block's coefficients one by one, which from global
    // memory costs an L2 round trip per coefficient
    // (host: coop == (sc.ncomp == 1 && sc.ss != 0 && sc.ah != 0))
    __shared__ uint32_t s_blk[32];
    const uint32_t seg = ent.seg;

    const JbScanResult sr = scanres[sc.range];
    const uint32_t *mk = marks + sc.mark_base;
    // the bit reader wants a 4-byte aligned base: use the image's (256-byte aligned) arena slot and
    // shift the scan-relative marker positions
    const uint8_t *data = arena + im.data_off;
    const uint32_t rel = (uint32_t)(sc.data_off - im.data_off);
    const uint32_t per_seg = sc.dri ? sc.dri : sc.nunits;
    const uint32_t first = min(seg * per_seg, sc.nunits);
    uint32_t count = active ? min(per_seg, sc.nunits - first) : 0u;
    uint32_t start = 0, stop, err = 0;
    if (active && seg > 0) {
        if (seg - 1 < sr.nmarkers && (mk[seg - 1] & 8u) == 0) start = (mk[seg - 1] >> 4) + 2;
        else {
            // no RSTn in front of this interval.  EOI at a restart boundary ends the scan quietly (HandleRestart
            // :203-207): the intervals behind it are simply not there; any other marker is an error
            // ("no RSTn in front of this interval" is met behind the previous one: reported with that place, here and now --
            // nothing else can fail in a job that decodes nothing)
            if (sr.end_marker != 0xD9u) jb_report_error(status, first_error, image, JB_ST_EXPECT_RST, ent.scan, seg - 1);
            // a scan of a sequential frame that ends here never calls WriteBlock for the MCUs behind this point
            // (JpegHuffmanBaselineScanDecoder.cs:144-150): the renderer takes the first such point of every scan
            else if (sc.seq) atomicMin(scan_limit + im.scan_base + ent.scan, first);
            count = 0;
        }
    }
    stop = seg < sr.nmarkers ? (mk[seg] >> 4) : sr.end_pos;
    start += rel;
    stop += rel;
    if (active && count == 0) start = stop;
    const bool needs_marker = sc.dri != 0 && count == per_seg;

    // ---- producers.  `avail` = units of this scan whose history is final (every producer is past them)
    uint32_t avail = sc.ndep ? 0u : 0xFFFFFFFFu;
    auto wait_for = [&](uint32_t u) { // returns once unit u may be decoded
        if (u < avail) return;
        unsigned long long t0 = 0, t1 = 0;
        if (TRACE) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        uint32_t a = 0xFFFFFFFFu;
        const int nd = sc.ndep == 0xFF ? (int)ent.scan : (int)sc.ndep;
        for (int i = 0; i < nd; i++) {
            const uint32_t ps = sc.ndep == 0xFF ? (uint32_t)i : (uint32_t)sc.dep[i];
            const JbDevScan &pd = scans[im.scan_base + ps];
            const bool whole = sc.ndep == 0xFF || ((sc.dep_all >> i) & 1u) || pd.nseg > 1;
            const uint32_t total = pd.nseg > 1 ? pd.nseg : pd.nunits; // the count that means "complete"
            const uint32_t *pp = progress + im.scan_base + ps;
            uint32_t v, spins = 0;
            for (;;) {
                v = jb_ld_progress(pp);
                if (v >= total) { v = 0xFFFFFFFFu; break; }
                if (!whole && v > u) break;
                // back off quickly: a poll costs issue slots that the decoding warps of the SM need, and the producer
                // only publishes every JB_K1C_PUBLISH units (tens of microseconds)
                __nanosleep(spins < 4 ? 500 : spins < 16 ? 2000 : 8000);
                if (++spins > JB_K1C_SPIN_LIMIT) { err |= JB_ST_STALLED; v = 0xFFFFFFFFu; break; }
            }
            a = min(a, v);
        }
        avail = a;
        if (TRACE) {
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            t_wait += t1 - t0;
        }
    };
    uint32_t next_pub = first + JB_K1C_PUBLISH;
    const bool publishes = sc.nseg == 1 && sc.has_consumer != 0;
    auto publish = [&](uint32_t done) { // single-segment scans: units 0..done-1 are final
        if (!publishes || done < next_pub) return;
        next_pub = done + JB_K1C_PUBLISH;
        if (coop) { // every lane wrote part of the blocks
            __threadfence();
            __syncwarp();
        } // (a lane with a stream of its own: the release store orders its own writes, and costs no L1 invalidation)
        if (!coop || lane == 0) jb_st_release(my_progress, done);
    };

    JbBitReader br;
    br.init(data, start, stop);
    int16_t *store = coef + im.coef_off * 64;
    const int al = sc.al;
    int eobrun = 0;
    int pred[4] = {0, 0, 0, 0};
    const uint8_t *tab_base = reinterpret_cast<const uint8_t *>(tables);

    auto huff = [&](uint32_t table_index) -> int {
        br.ensure32();
        uint32_t e = jb_huff_lookup(reinterpret_cast<const JbHuffTable *>(tab_base + (size_t)table_index * sizeof(JbHuffTable)),
                                    br.peek16());
        if (e == 0xFFFFFFFFu) { err |= JB_ST_BAD_CODE; e = 0x0001u; }
        br.skip_code(e & 0xFF);
        return (int)(e >> 8);
    };
    auto dc_block = [&](int16_t *blk, int slot) {
        if (sc.ah == 0) { // first scan (:231-242)
            int s = huff(sc.dc_tab[slot]);
            if (s > 16) { err |= JB_ST_BAD_CODE; s = 0; }
            int v = 0;
            if (s != 0) v = jb_extend((int)jb_prog_bits(br, s), s);
            const int p = slot == 0 ? pred[0] : slot == 1 ? pred[1] : slot == 2 ? pred[2] : pred[3];
            v += p;
            if (slot == 0) pred[0] = v; else if (slot == 1) pred[1] = v; else if (slot == 2) pred[2] = v; else pred[3] = v;
            blk[0] = (int16_t)(v << al);
        } else { // refinement (:244-252)
            // blk[0] |= bit << al without a load in the decoder's dependency chain: an OR reduction on the block's
            // first word, resolved in L2 (blk[1], the other half of the word, may be written by an AC scan meanwhile)
            if (jb_prog_bits(br, 1) != 0) atomicOr(reinterpret_cast<unsigned int *>(blk), 1u << al);
        }
    };

    // ---- one whole block of a sequential frame: ReadBlockBaseline (JpegHuffmanBaselineScanDecoder.cs:187-219)
    auto seq_block = [&](int16_t *blk, int slot) {
        uint4 *b4 = reinterpret_cast<uint4 *>(blk); // "outputBuffer = default" (:121): a later scan over the same
#pragma unroll                                      // component replaces the block
        for (int j = 0; j < 8; j++) b4[j] = make_uint4(0, 0, 0, 0);
        int s = huff(sc.dc_tab[slot]);
        if (s > 16) { err |= JB_ST_BAD_CODE; s = 0; }
        int v = s != 0 ? jb_extend((int)jb_prog_bits(br, s), s) : 0;
        const int p = slot == 0 ? pred[0] : slot == 1 ? pred[1] : slot == 2 ? pred[2] : pred[3];
        v += p;
        if (slot == 0) pred[0] = v; else if (slot == 1) pred[1] = v; else if (slot == 2) pred[2] = v; else pred[3] = v;
        blk[0] = (int16_t)v;
        for (int k = 1; k < 64 && !err;) {
            const int sym = huff(sc.ac_tab[slot]);
            const int r = sym >> 4, sz = sym & 15;
            if (sz != 0) {
                blk[min(k + r, 63)] = (int16_t)jb_extend((int)jb_prog_bits(br, sz), sz);
                k += r + 1;
            } else if (r == 0) {
                break;      // end of block
            } else {
                k += 16;    // ZRL -- and, like the reference (:213-219), any other symbol with s == 0
            }
        }
    };

    const int p1 = 1 << al, m1 = -(1 << al);
    if (coop) {
        // ---- AC refinement, one stream per warp, decoded by the WHOLE warp.  Every lane runs the same bit reader and
        // Huffman decode (identical state, broadcast loads), lane l owns coefficients l and l + 32 of the block.
        // ReadBlockProgressiveACRefined (:313-419) walks the band position by position: a nonzero coefficient costs one
        // correction bit, a zero one counts down the symbol's run.  The history of a block does not change while the
        // scan is at it (positions only move forward), so per block every lane ranks its two positions among the
        // zero-history and among the nonzero-history positions of the band and the zero positions are listed in shared
        // memory.  Per symbol, "the (r+1)-th zero from k on" is then one table look-up, the number of correction bits
        // in front of it is arithmetic (k_end - k - r), they are read in one go and every lane picks its own by rank.
        const int c = sc.comp[0];
        const uint32_t *plane = reinterpret_cast<const uint32_t *>(store + (size_t)im.comp_plane_off[c] * 64);
        const uint32_t pw = im.comp_plane_w[c];
        const int ss = sc.ss, se = sc.se;
        const uint64_t band = (se >= 63 ? ~0ull : ((1ull << (se + 1)) - 1ull)) & ~((1ull << ss) - 1ull);
        const uint32_t band_lo = (uint32_t)band, band_hi = (uint32_t)(band >> 32);
        const uint32_t below = (1u << lane) - 1u; // lanes in front of mine
        const bool in_lo = (band_lo >> lane) & 1u, in_hi = (band_hi >> lane) & 1u;
        __shared__ uint8_t s_zpos[64];  // position of the t-th zero-history coefficient of the band
        __shared__ uint16_t s_lut[1 << JB_LUT_BITS]; // first-level Huffman look-up of the scan's AC table
        const JbHuffTable *actab = reinterpret_cast<const JbHuffTable *>(tab_base + (size_t)sc.ac_tab[0] * sizeof(JbHuffTable));
        for (int i = lane; i < (1 << JB_LUT_BITS) / 2; i += 32)
            reinterpret_cast<uint32_t *>(s_lut)[i] = __ldg(reinterpret_cast<const uint32_t *>(actab->lut) + i);
        const uint32_t wb = sc.wb;
        uint32_t by = first / wb, bx = first - by * wb;
        if (count) wait_for(first);
        uint32_t nxt = count ? __ldcg(plane + ((size_t)by * pw + bx) * 32 + lane) : 0u;
        for (uint32_t u = first; u < first + count; u++) {
            uint32_t *gblk = const_cast<uint32_t *>(plane) + ((size_t)by * pw + bx) * 32;
            s_blk[lane] = nxt;
            if (++bx == wb) { bx = 0; by++; }
            if (u + 1 < first + count) { // prefetch
                wait_for(u + 1);
                nxt = __ldcg(plane + ((size_t)by * pw + bx) * 32 + lane);
            }
            __syncwarp();
            int16_t *b16 = reinterpret_cast<int16_t *>(s_blk);
            int c_lo = b16[lane], c_hi = b16[lane + 32];
            bool beyond = false; // a coefficient was placed just behind the band (see below)
            __syncwarp();
            if (!err) {
                const uint32_t z_lo = __ballot_sync(0xFFFFFFFFu, c_lo == 0) & band_lo;  // zero history, in band
                const uint32_t z_hi = __ballot_sync(0xFFFFFFFFu, c_hi == 0) & band_hi;
                const uint32_t h_lo = band_lo & ~z_lo, h_hi = band_hi & ~z_hi;          // nonzero history, in band
                const int nz_lo = __popc(z_lo), nzeros = nz_lo + __popc(z_hi);
                const int nh_lo = __popc(h_lo), nhist = nh_lo + __popc(h_hi);
                const bool mz_lo = (z_lo >> lane) & 1u, mz_hi = (z_hi >> lane) & 1u;
                // my ranks among the nonzero-history positions (-1000: not one of them)
                const int hr_lo = in_lo && !mz_lo ? __popc(h_lo & below) : -1000;
                const int hr_hi = in_hi && !mz_hi ? nh_lo + __popc(h_hi & below) : -1000;
                if (mz_lo) s_zpos[__popc(z_lo & below)] = (uint8_t)lane;
                if (mz_hi) s_zpos[nz_lo + __popc(z_hi & below)] = (uint8_t)(lane + 32);
                __syncwarp();
                int zbase = 0, hbase = 0; // zero- / nonzero-history positions of the band in front of k
                // n correction bits for the nonzero-history positions of rank hbase .. hbase + n - 1 (first bit = lowest)
                auto correct = [&](int n) {
                    for (int done = 0; done < n; done += 16) {
                        const int cnt = min(16, n - done);
                        if (br.n < cnt) br.ensure32();
                        const uint32_t v = br.take(cnt);
                        const int i_lo = hr_lo - hbase - done, i_hi = hr_hi - hbase - done;
                        if ((unsigned)i_lo < (unsigned)cnt && ((v >> (cnt - 1 - i_lo)) & 1u) && (c_lo & p1) == 0) c_lo += c_lo >= 0 ? p1 : m1;
                        if ((unsigned)i_hi < (unsigned)cnt && ((v >> (cnt - 1 - i_hi)) & 1u) && (c_hi & p1) == 0) c_hi += c_hi >= 0 ? p1 : m1;
                    }
                    hbase += n;
                };
                int k = ss;
                if (eobrun == 0) {
                    while (k <= se) {
                        br.ensure32(); // >= 33 bits: the code (<= 16), a sign bit or <= 14 run bits, and 16 more
                        uint32_t e = s_lut[br.peek16() >> (16 - JB_LUT_BITS)];
                        if ((e & 0xFFu) == 0) {
                            e = jb_huff_lookup(actab, br.peek16());
                            if (e == 0xFFFFFFFFu) { err |= JB_ST_BAD_CODE; e = 0x0001u; }
                        }
                        br.skip_code(e & 0xFFu);
                        const int r = (int)(e >> 12), sz = (int)(e >> 8) & 15;
                        int s = 0;
                        if (sz != 0) {
                            s = br.take(1) != 0 ? p1 : m1;
                        } else if (r != 15) {
                            eobrun = 1 << r;
                            if (r != 0) eobrun += (int)br.take(r);
                            break;
                        }
                        // the (r+1)-th zero-history position at or behind k; the band end if there are fewer
                        const int target = zbase + r;
                        int k_end, n;
                        if (target < nzeros) {
                            k_end = s_zpos[target];
                            n = k_end - k - r;
                            zbase = target + 1;
                        } else {
                            k_end = se + 1;
                            n = nhist - hbase;
                            beyond |= s != 0 && se < 63; // (the reference writes at se + 1 when the run overshoots the band)
                        }
                        correct(n);
                        if (s != 0 && k_end < 64) {
                            if (k_end < 32) { if (lane == k_end) c_lo = s; }
                            else if (lane == k_end - 32) c_hi = s;
                        }
                        k = k_end + 1;
                    }
                }
                if (eobrun > 0) {
                    if (k <= se) correct(nhist - hbase);
                    --eobrun;
                }
            }
            // write back the scan's band only: scans of the same level refine other bands of the same block concurrently
            int16_t *g16 = reinterpret_cast<int16_t *>(gblk);
            if (lane >= ss && lane <= se) g16[lane] = (int16_t)c_lo;
            if (lane + 32 >= ss && lane + 32 <= se) g16[lane + 32] = (int16_t)c_hi;
            if (beyond && lane == ((se + 1) & 31)) // the reference's write just behind the band (corrupt streams only)
                g16[se + 1] = (int16_t)(se + 1 < 32 ? c_lo : c_hi);
            publish(u + 1);
        }
    } else if (sc.ncomp > 1 || sc.seq) {
        // ---- interleaved DC scan (:92-138), or any scan of a sequential frame (whole blocks, same MCU walk).  Component geometry is hoisted (compile-time indexed, so it stays in
        // registers) and the MCU position is stepped instead of divided out per MCU.
        int16_t *cbase[4];
        uint32_t cpitch[4];
        int ch[4], cv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = sc.comp[i < sc.ncomp ? i : 0];
            cbase[i] = store + (size_t)im.comp_plane_off[c] * 64;
            cpitch[i] = im.comp_plane_w[c];
            ch[i] = im.comp_h[c]; cv[i] = im.comp_v[c];
        }
        uint32_t my = first / im.mcus_per_line, mx = first - my * im.mcus_per_line;
        const uint32_t mpl = im.mcus_per_line;
        for (uint32_t u = first; u < first + count && !err; u++) {
            wait_for(u);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (i >= sc.ncomp) break;
                for (int y = 0; y < cv[i]; y++)
                    for (int x = 0; x < ch[i]; x++) {
                        int16_t *blk = cbase[i] + ((size_t)(my * cv[i] + y) * cpitch[i] + mx * ch[i] + x) * 64;
                        if (sc.seq) seq_block(blk, i); else dc_block(blk, i);
                    }
            }
            if (++mx == mpl) { mx = 0; my++; }
            publish(u + 1);
        }
    } else if (sc.ss == 0) {
        // ---- single-component DC scan (:140-194 with ReadBlockProgressiveDC)
        const int c = sc.comp[0];
        const uint32_t wb = sc.wb, pitch = im.comp_plane_w[c];
        int16_t *base = store + (size_t)im.comp_plane_off[c] * 64;
        uint32_t by = first / wb, bx = first - by * wb;
        for (uint32_t u = first; u < first + count && !err; u++) {
            wait_for(u);
            dc_block(base + ((size_t)by * pitch + bx) * 64, 0);
            if (++bx == wb) { bx = 0; by++; }
            publish(u + 1);
        }
    } else {
        // ---- AC first scan (:259-305), ONE loop over symbols: the lanes of a packed warp are in different blocks of
        // different images, and a loop nest per block would make every lane wait for the longest block of the warp
        // at each block end.  (AC refinement scans never get here: they are whole-warp jobs.)
        const int c = sc.comp[0];
        const int ss = sc.ss, se = sc.se;
        const uint32_t wb = sc.wb, pitch = im.comp_plane_w[c];
        int16_t *base = store + (size_t)im.comp_plane_off[c] * 64;
        const JbHuffTable *actab = reinterpret_cast<const JbHuffTable *>(tab_base + (size_t)sc.ac_tab[0] * sizeof(JbHuffTable));
        const uint32_t end = first + count;
        uint32_t u = first;
        uint32_t by = first / wb, bx = first - by * wb;
        int16_t *blk = base + ((size_t)by * pitch + bx) * 64;
        int i = ss;
        if (count) wait_for(first);
        while (u < end && !err) {
            bool block_done = false;
            if (eobrun != 0) { // the whole run of end-of-band blocks at once
                const uint32_t skip = min((uint32_t)eobrun, end - u);
                eobrun -= (int)skip;
                u += skip;
                by = u / wb; bx = u - by * wb;
                blk = base + ((size_t)by * pitch + bx) * 64;
                i = ss;
                // the skipped blocks count as done by this scan only when its own producers are past them: a consumer
                // that follows this scan relies on "block u done here => done in every scan this one follows"
                wait_for(u - 1);
                publish(u);
                if (u < end) wait_for(u);
                continue;
            }
            br.ensure32(); // >= 33 bits: the code (<= 16) and up to 16 magnitude / run bits
            uint32_t e = jb_huff_lookup(actab, br.peek16());
            if (e == 0xFFFFFFFFu) { err |= JB_ST_BAD_CODE; e = 0x0001u; }
            br.skip_code(e & 0xFFu);
            const int r = (int)(e >> 12), sz = (int)(e >> 8) & 15;
            if (sz != 0) {
                i += r;
                const int v = jb_extend((int)br.take(sz), sz);
                blk[min(i, 63)] = (int16_t)(v << al);
                i++;
            } else if (r == 15) {
                i += 16;
            } else {
                eobrun = 1 << r;
                if (r != 0) eobrun += (int)br.take(r);
                --eobrun;
                block_done = true;
            }
            if (block_done || i > se) {
                u++;
                i = ss;
                if (++bx == wb) { bx = 0; by++; }
                blk = base + ((size_t)by * pitch + bx) * 64;
                publish(u);
                if (u < end) wait_for(u);
            }
        }
    }

    if (active && (!coop || lane == 0)) {
        if (br.n < br.pad) err |= JB_ST_PREMATURE_END;
        if (needs_marker && !err) {
            const int real = br.n - br.pad;
            uint32_t p = br.pos;
            while (p < stop && data[p] == 0xFF) p++;
            bool marker_ok = seg < sr.nmarkers;
            if (marker_ok && (mk[seg] & 8u) != 0) marker_ok = sr.end_marker == 0xD9u;
            if (real >= 8 || p < stop || !marker_ok) err |= JB_ST_EXPECT_RST;
        }
        if (err) jb_report_error(status, first_error, image, err, ent.scan, seg);
    }
    if (TRACE && lane == 0) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        trace[4 * turn] = ((unsigned long long)image << 32) | (ent.scan << 16) | min(ent.seg, 0xFFFFu);
        trace[4 * turn + 1] = t_start; trace[4 * turn + 2] = t_end; trace[4 * turn + 3] = t_wait;
    }
    // ---- this job's part of the scan is final, whatever happened: consumers must never wait for an error path
    if (coop ? sc.has_consumer != 0 : (active && sc.has_consumer)) {
        __threadfence();
        if (coop) __syncwarp();
        if (!coop || lane == 0) {
            if (sc.nseg == 1) jb_st_release(my_progress, sc.nunits);
            else atomicAdd(my_progress, 1u);
        }
    }
}

// Scan-list frames of a sequential process with restart intervals: MCUs a component was written for.  A block reaches
// WriteBlock iff some scan that names its component got to its MCU; a scan that met EOI on a restart boundary stopped at
// scan_limit (0xFFFFFFFF: it ran to the end).  comp_limit[image * 4 + c] = the furthest any scan naming c got,
// mcu_limit[image] = the furthest any component got (the renderer and the partial D2H copy stop there).
__global__ void jb_k1c_sequential_limits(const JbDevImage *__restrict__ images, const JbDevScan *__restrict__ scans,
                                         const uint32_t *__restrict__ image_list, int nimg,
                                         const uint32_t *__restrict__ scan_limit, uint32_t *__restrict__ comp_limit,
                                         uint32_t *__restrict__ mcu_limit)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nimg) return;
    const uint32_t image = image_list[i];
    const JbDevImage &im = images[image];
    if (!im.seq_dri) return;
    uint32_t lim[4] = {0, 0, 0, 0}, most = 0;
    for (uint32_t k = 0; k < im.nscans; k++) {
        const JbDevScan &sc = scans[im.scan_base + k];
        const uint32_t l = min(scan_limit[im.scan_base + k], im.total_mcus);
        for (int a = 0; a < sc.ncomp; a++) lim[sc.comp[a] & 3] = max(lim[sc.comp[a] & 3], l);
    }
    for (int c = 0; c < 4; c++) {
        comp_limit[image * 4 + c] = lim[c] >= im.total_mcus ? 0xFFFFFFFFu : lim[c];
        if (c < im.ncomp) most = max(most, lim[c]);
    }
    mcu_limit[image] = most >= im.total_mcus ? 0xFFFFFFFFu : most;
}
