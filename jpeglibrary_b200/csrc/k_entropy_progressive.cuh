// k_entropy_progressive.cuh -- K1c: one scan of a progressive (SOF2) Huffman frame, for a whole batch.
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
// Parallelism: one thread per (image, scan, restart segment).  A scan only depends on earlier scans
// that touch the same component with an overlapping spectral band (refinement of what they wrote), so
// the host sorts scans into dependency levels and launches one kernel per level: libjpeg's 10-scan
// script needs 4 launches ({DC}, {Y 1-5, Cr, Cb, Y 6-63}, {Y refine, DC refine, Cr refine, Cb refine},
// {Y refine}).  Refinement scans read coefficient history, so they cannot be decoded speculatively;
// streams without restart markers expose one thread per image and scan (batch-level parallelism only).
// One lane per warp is used in that case so that every serial stream gets its own scheduler slot.
#pragma once
#include "jb_device.cuh"
#include "k_entropy_decode.cuh"

__device__ __forceinline__ uint32_t jb_prog_bits(JbBitReader &br, int k)
{ // k in 1..16 raw bits (TryReadBits)
    br.ensure32();
    return br.take(k);
}

__global__ void __launch_bounds__(32)
jb_k1c_progressive_scan(const JbDevImage *__restrict__ images, const uint32_t *__restrict__ image_list,
                        const JbDevScan *__restrict__ scans, int level, const JbHuffTable *__restrict__ tables,
                        const uint8_t *__restrict__ arena, const uint32_t *__restrict__ marks,
                        const JbScanResult *__restrict__ scanres, int16_t *__restrict__ coef,
                        uint32_t *__restrict__ status, int lanes_per_warp)
{
    const uint32_t image = image_list[blockIdx.y];
    const JbDevImage &im = images[image];
    const uint32_t scan_index = blockIdx.z; // every scan of the requested dependency level runs concurrently
    if (scan_index >= im.nscans) return;
    const JbDevScan sc = scans[im.scan_base + scan_index]; // by value: its fields are used in every inner loop
    if (sc.level != level) return;
    const int lane = threadIdx.x;
    // AC refinement scans of serial streams (one stream per warp) are decoded by lane 0 on a shared-memory copy of
    // the current block that the whole warp loads (prefetched one block ahead) and stores back: the refinement
    // loop reads and rewrites the block's coefficients one by one, which from global memory costs an L2 round
    // trip per coefficient (a store evicts the line from L1)
    const bool coop = lanes_per_warp == 1 && sc.ncomp == 1 && sc.ss != 0 && sc.ah != 0;
    __shared__ uint32_t s_blk[32];
    if (!coop && lane >= lanes_per_warp) return;
    const uint32_t seg = coop ? blockIdx.x : blockIdx.x * lanes_per_warp + lane;
    if (seg >= sc.nseg) return;

    const JbScanResult sr = scanres[sc.range];
    const uint32_t *mk = marks + sc.mark_base;
    // the bit reader wants a 4-byte aligned base: use the image's (256-byte aligned) arena slot and
    // shift the scan-relative marker positions
    const uint8_t *data = arena + im.data_off;
    const uint32_t rel = (uint32_t)(sc.data_off - im.data_off);
    const uint32_t per_seg = sc.dri ? sc.dri : sc.nunits;
    const uint32_t first = seg * per_seg;
    const uint32_t count = min(per_seg, sc.nunits - first);
    uint32_t start = 0, stop, err = 0;
    if (seg > 0) {
        if (seg - 1 < sr.nmarkers && (mk[seg - 1] & 8u) == 0) start = (mk[seg - 1] >> 4) + 2;
        else { atomicOr(status + image, JB_ST_EXPECT_RST); return; }
    }
    stop = seg < sr.nmarkers ? (mk[seg] >> 4) : sr.end_pos;
    start += rel;
    stop += rel;
    const bool needs_marker = sc.dri != 0 && count == per_seg;

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
        br.skip(e & 0xFF);
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
            const int bit = (int)jb_prog_bits(br, 1);
            blk[0] = (int16_t)(blk[0] | (int16_t)(bit << al));
        }
    };

    const int p1 = 1 << al, m1 = -(1 << al);
    // ---- AC refinement of one block (:313-419)
    auto refine_block = [&](int16_t *blk, int ss, int se) {
        int k = ss;
        if (eobrun == 0) {
            for (; k <= se; k++) {
                const int sym = huff(sc.ac_tab[0]);
                int r = sym >> 4, s = sym & 15;
                if (s != 0) {
                    s = jb_prog_bits(br, 1) != 0 ? p1 : m1;
                } else if (r != 15) {
                    eobrun = 1 << r;
                    if (r != 0) eobrun += (int)jb_prog_bits(br, r);
                    break;
                }
                do {
                    int cv = blk[k];
                    if (cv != 0) {
                        if (jb_prog_bits(br, 1) != 0) {
                            if ((cv & p1) == 0) blk[k] = (int16_t)(cv + (cv >= 0 ? p1 : m1));
                        }
                    } else {
                        if (--r < 0) break;
                    }
                    k++;
                } while (k <= se);
                if (s != 0 && k < 64) blk[k] = (int16_t)s;
            }
        }
        if (eobrun > 0) {
            for (; k <= se; k++) {
                int cv = blk[k];
                if (cv != 0) {
                    if (jb_prog_bits(br, 1) != 0) {
                        if ((cv & p1) == 0) blk[k] = (int16_t)(cv + (cv > 0 ? p1 : m1));
                    }
                }
            }
            --eobrun;
        }
    };

    if (coop) {
        // ---- AC refinement, one stream per warp, decoded by the WHOLE warp.  Every lane runs the same bit reader and
        // Huffman decode (identical state, broadcast loads), lane l owns coefficients l and l + 32 of the block.
        // ReadBlockProgressiveACRefined (:313-419) walks the band position by position: a nonzero coefficient costs one
        // correction bit, a zero one counts down the symbol's run.  Here the nonzero history of the block is a 64-bit
        // ballot, so "the (r+1)-th zero from k on" is a prefix-popcount + ballot, the correction bits in front of it
        // are read in one go and applied by all lanes at once.
        const int c = sc.comp[0];
        const uint32_t *plane = reinterpret_cast<const uint32_t *>(store + (size_t)im.comp_plane_off[c] * 64);
        const uint32_t pw = im.comp_plane_w[c];
        const int ss = sc.ss, se = sc.se;
        const uint64_t band = (se >= 63 ? ~0ull : ((1ull << (se + 1)) - 1ull)) & ~((1ull << ss) - 1ull);
        const uint64_t below_lo = (1ull << lane) - 1ull, below_hi = (1ull << (lane + 32)) - 1ull; // positions in front of mine
        uint32_t by = first / sc.wb, bx = first - by * sc.wb;
        uint32_t nxt = count ? __ldg(plane + ((size_t)by * pw + bx) * 32 + lane) : 0u;
        for (uint32_t u = first; u < first + count; u++) {
            uint32_t *gblk = const_cast<uint32_t *>(plane) + ((size_t)by * pw + bx) * 32;
            s_blk[lane] = nxt;
            if (++bx == sc.wb) { bx = 0; by++; }
            if (u + 1 < first + count) nxt = __ldg(plane + ((size_t)by * pw + bx) * 32 + lane); // prefetch
            __syncwarp();
            int16_t *b16 = reinterpret_cast<int16_t *>(s_blk);
            int c_lo = b16[lane], c_hi = b16[lane + 32];
            bool beyond = false; // a coefficient was placed just behind the band (see below)
            __syncwarp();
            if (!err) {
                const uint64_t Z = (((uint64_t)__ballot_sync(0xFFFFFFFFu, c_hi != 0) << 32) | __ballot_sync(0xFFFFFFFFu, c_lo != 0)) & band;
                // correction bits for the nonzero-history positions in `nz` (ascending), read 16 at a time
                auto correct = [&](uint64_t nz) {
                    const int n = __popcll(nz);
                    const int my_lo = (nz >> lane) & 1ull ? __popcll(nz & below_lo) : -1;
                    const int my_hi = (nz >> (lane + 32)) & 1ull ? __popcll(nz & below_hi) : -1;
                    for (int done = 0; done < n; done += 16) {
                        const int cnt = min(16, n - done);
                        const uint32_t v = jb_prog_bits(br, cnt); // first bit = lowest position
                        if (my_lo >= done && my_lo < done + cnt && ((v >> (cnt - 1 - (my_lo - done))) & 1u) && (c_lo & p1) == 0)
                            c_lo += c_lo >= 0 ? p1 : m1;
                        if (my_hi >= done && my_hi < done + cnt && ((v >> (cnt - 1 - (my_hi - done))) & 1u) && (c_hi & p1) == 0)
                            c_hi += c_hi >= 0 ? p1 : m1;
                    }
                };
                int k = ss;
                if (eobrun == 0) {
                    while (k <= se) {
                        const int sym = huff(sc.ac_tab[0]);
                        int r = sym >> 4, s = sym & 15;
                        if (s != 0) {
                            s = jb_prog_bits(br, 1) != 0 ? p1 : m1;
                        } else if (r != 15) {
                            eobrun = 1 << r;
                            if (r != 0) eobrun += (int)jb_prog_bits(br, r);
                            break;
                        }
                        // the (r+1)-th zero-history position at or behind k; the band end if there are fewer
                        const uint64_t from_k = ~((1ull << k) - 1ull);
                        const uint64_t zeros = ~Z & band & from_k;
                        const bool z_lo = (zeros >> lane) & 1ull, z_hi = (zeros >> (lane + 32)) & 1ull;
                        const uint32_t hit_lo = __ballot_sync(0xFFFFFFFFu, z_lo && __popcll(zeros & below_lo) == r);
                        const uint32_t hit_hi = __ballot_sync(0xFFFFFFFFu, z_hi && __popcll(zeros & below_hi) == r);
                        const int k_end = hit_lo ? __ffs(hit_lo) - 1 : hit_hi ? 32 + __ffs(hit_hi) - 1 : se + 1;
                        const uint64_t upto = k_end >= 64 ? ~0ull : ((1ull << k_end) - 1ull);
                        correct(Z & from_k & upto);
                        if (s != 0 && k_end < 64) { // (the reference writes at se + 1 when the run overshoots the band)
                            if (k_end < 32) { if (lane == k_end) c_lo = s; }
                            else if (lane == k_end - 32) c_hi = s;
                            beyond |= k_end > se;
                        }
                        k = k_end + 1;
                    }
                }
                if (eobrun > 0) {
                    if (k <= se) correct(Z & ~((1ull << k) - 1ull));
                    --eobrun;
                }
            }
            // write back the scan's band only: scans of the same level refine other bands of the same block concurrently
            int16_t *g16 = reinterpret_cast<int16_t *>(gblk);
            if (lane >= ss && lane <= se) g16[lane] = (int16_t)c_lo;
            if (lane + 32 >= ss && lane + 32 <= se) g16[lane + 32] = (int16_t)c_hi;
            if (beyond && lane == ((se + 1) & 31)) // the reference's write just behind the band (corrupt streams only)
                g16[se + 1] = (int16_t)(se + 1 < 32 ? c_lo : c_hi);
        }
        if (lane != 0) return;
    } else if (sc.ncomp > 1) {
        // ---- interleaved DC scan (:92-138)
        for (uint32_t u = first; u < first + count && !err; u++) {
            const uint32_t my = u / im.mcus_per_line, mx = u - my * im.mcus_per_line;
            for (int i = 0; i < sc.ncomp; i++) {
                const int c = sc.comp[i];
                const int h = im.comp_h[c], v = im.comp_v[c];
                for (int y = 0; y < v; y++)
                    for (int x = 0; x < h; x++) {
                        int16_t *blk = store + ((size_t)im.comp_plane_off[c] + (size_t)(my * v + y) * im.comp_plane_w[c] + mx * h + x) * 64;
                        dc_block(blk, i);
                    }
            }
        }
    } else {
        const int c = sc.comp[0];
        const int ss = sc.ss, se = sc.se;
        for (uint32_t u = first; u < first + count && !err; u++) {
            const uint32_t by = u / sc.wb, bx = u - by * sc.wb;
            int16_t *blk = store + ((size_t)im.comp_plane_off[c] + (size_t)by * im.comp_plane_w[c] + bx) * 64;
            if (ss == 0) {
                dc_block(blk, 0);
            } else if (sc.ah == 0) {
                // ---- AC first scan (:259-305)
                if (eobrun != 0) { // the whole run of end-of-band blocks at once
                    const uint32_t skip = min((uint32_t)eobrun, first + count - u);
                    eobrun -= (int)skip;
                    u += skip - 1;
                    continue;
                }
                for (int i = ss; i <= se; i++) {
                    const int sym = huff(sc.ac_tab[0]);
                    const int r = sym >> 4, s = sym & 15;
                    i += r;
                    if (s != 0) {
                        const int v = jb_extend((int)jb_prog_bits(br, s), s);
                        blk[min(i, 63)] = (int16_t)(v << al);
                    } else if (r != 15) {
                        eobrun = 1 << r;
                        if (r != 0) eobrun += (int)jb_prog_bits(br, r);
                        --eobrun;
                        break;
                    }
                }
            } else {
                // ---- AC refinement (:313-419)
                refine_block(blk, ss, se);
            }
        }
    }

    if (br.n < br.pad) err |= JB_ST_PREMATURE_END;
    if (needs_marker && !err) {
        const int real = br.n - br.pad;
        uint32_t p = br.pos;
        while (p < stop && data[p] == 0xFF) p++;
        bool marker_ok = seg < sr.nmarkers;
        if (marker_ok && (mk[seg] & 8u) != 0) marker_ok = sr.end_marker == 0xD9u;
        if (real >= 8 || p < stop || !marker_ok) err |= JB_ST_EXPECT_RST;
    }
    if (err) atomicOr(status + image, err);
}
