// jpeg_host.cpp -- host-side marker/header walk producing jb_image_desc.
//
// Mirrors the control flow of JpegDecoder.Decode (src/JpegLibrary/JpegDecoder.cs:509-617):
// SOI, then a marker loop keeping a registry of the Huffman tables (DHT), quantisation tables (DQT)
// and the restart interval (DRI) in force; at every SOS the tables the scan refers to are
// snapshotted into the descriptor and the entropy-coded segment is delimited by searching for the
// next marker that is neither a restart marker nor FF00/FFFF (JpegReader.TryReadMarker,
// JpegReader.cs:120-158).  No decode arithmetic happens here.
#include "../../include/jpegb200_host.h"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <atomic>

struct jbh_parsed {
    jb_image_desc desc{};
    std::vector<jb_scan_desc> scans;
    std::vector<jb_huff_spec> tables;
    uint64_t consumed = 0;
    int sof_marker = 0;
};

static thread_local std::string g_parse_error;

namespace {

struct Walker {
    const uint8_t *d;
    uint64_t n, pos = 0;
    // JpegReader.TryReadMarker
    int next_marker()
    {
        while (pos + 1 < n) {
            if (d[pos] == 0xFF) {
                uint8_t b2 = d[pos + 1];
                if (b2 == 0xFF) { pos += 1; continue; }
                if (b2 == 0x00) { pos += 2; continue; }
                pos += 2;
                return b2;
            }
            const void *q = memchr(d + pos, 0xFF, n - pos);
            if (!q) { pos = n; return -1; }
            pos = (uint64_t)((const uint8_t *)q - d);
        }
        pos = n;
        return -1;
    }
};

int perr(int code, uint64_t offset, const char *msg)
{
    char buf[256];
    snprintf(buf, sizeof buf, "Failed to decode JPEG data at offset %llu. %s", (unsigned long long)offset, msg);
    g_parse_error = buf;
    return code;
}

// position of the next marker that ends an entropy-coded segment (not RSTn, not FF00, not FFFF)
uint64_t find_scan_end(const uint8_t *d, uint64_t n, uint64_t from)
{
    uint64_t p = from;
    while (p + 1 < n) {
        const void *q = memchr(d + p, 0xFF, n - p);
        if (!q) return n;
        p = (uint64_t)((const uint8_t *)q - d);
        if (p + 1 >= n) return n;
        uint8_t b2 = d[p + 1];
        if (b2 == 0x00) { p += 2; continue; }
        if (b2 == 0xFF) { p += 1; continue; }
        if (b2 >= 0xD0 && b2 <= 0xD7) { p += 2; continue; }
        return p;
    }
    return n;
}

} // namespace

extern "C" {

const char *jbh_last_parse_error(void) { return g_parse_error.c_str(); }

// `tables`: what JpegDecoder.LoadTables was given before the stream (abbreviated streams, e.g. the JPEGTables of a
// TIFF file): its DHT / DQT / DRI segments are the state the marker loop of the stream starts from.
// tables_used != nullptr: only the tables stream is walked (what LoadTables itself would raise), and the caller learns
// how many of its bytes the walk took (an EOI, or bytes without a further marker, are left out).
static int parse_impl(const uint8_t *tables, uint64_t tables_len, const uint8_t *data, uint64_t length, jbh_parsed **out,
                      uint64_t *tables_used = nullptr)
{
    jbh_parsed *dummy = nullptr;
    if (tables_used) { out = &dummy; *tables_used = 0; }
    if (!out) return JB_ERR_ARGUMENT;
    *out = nullptr;
    if (!tables_used && (!data || length == 0)) {
        g_parse_error = "Input buffer is not specified.";
        return JB_ERR_INVALID_OPERATION;
    }
    jbh_parsed *p = new jbh_parsed;
    struct Guard {
        jbh_parsed *p;
        ~Guard() { delete p; }
    } guard{p};
    jb_image_desc &im = p->desc;
    im.data = data;
    im.length = length;

    int huff_latest[2][4];
    for (auto &a : huff_latest)
        for (auto &b : a) b = -1;
    uint16_t qt[4][64] = {}; // (lossless frames carry no DQT: their descriptor holds zeros, not stack contents)
    bool qt_present[4] = {false, false, false, false};
    uint8_t comp_id[4] = {0, 0, 0, 0}, comp_tq[4] = {0, 0, 0, 0};
    // The reference's sequential and lossless scan decoders are created when the frame header is read and take the
    // restart interval ONCE, in their constructors (JpegDecoder.cs:569, JpegHuffmanBaselineScanDecoder.cs:38,
    // JpegHuffmanLosslessScanDecoder.cs:32): every scan of such a frame uses the value JpegDecoder holds at the SOF.  This
    // walk stands for Identify() followed by Decode() (apps/JpegDecode, the reference's tests): Identify() has walked the
    // whole stream before, so unless a DRI segment precedes the SOF that value is the LAST DRI of the stream.
    // Progressive scans read it per scan (JpegHuffmanProgressiveScanDecoder.cs:78): the DRI in force at the SOS.
    uint32_t restart_interval = 0, restart_at_sof = 0;
    bool dri_before_sof = false, dri_seen = false;
    bool have_frame = false;
    // progressive `_components` slot emulation (JpegHuffmanProgressiveScanDecoder.cs:21,69,431-462)
    int slot_comp[4] = {-1, -1, -1, -1};
    uint16_t slot_qt[4][64] = {};

    // DHT: possibly several tables per segment (JpegDecoder.ProcessDefineHuffmanTable)
    auto on_dht = [&](const uint8_t *b, uint64_t n, uint64_t seg_at) -> int {
        while (n > 0) {
            if (n < 17) return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse Huffman table.");
            jb_huff_spec s{};
            s.table_class = b[0] >> 4;
            s.identifier = b[0] & 15;
            int cnt = 0;
            for (int i = 0; i < 16; i++) { s.bits[i] = b[1 + i]; cnt += b[1 + i]; }
            if (cnt > 256 || n < 17ull + cnt || s.table_class > 1 || s.identifier > 3)
                return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse Huffman table.");
            memcpy(s.values, b + 17, (size_t)cnt);
            s.value_count = (uint16_t)cnt;
            huff_latest[s.table_class][s.identifier] = (int)p->tables.size();
            p->tables.push_back(s);
            b += 17 + cnt;
            n -= 17 + cnt;
        }
        return JB_OK;
    };
    auto on_dqt = [&](const uint8_t *b, uint64_t n, uint64_t seg_at) -> int {
        while (n > 0) {
            int pq = b[0] >> 4, tq = b[0] & 15;
            uint64_t need = pq ? 129 : 65;
            if (pq > 1 || tq > 3 || n < need) return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse quantization table.");
            for (int i = 0; i < 64; i++) qt[tq][i] = pq ? (uint16_t)((b[1 + 2 * i] << 8) | b[2 + 2 * i]) : b[1 + i];
            qt_present[tq] = true;
            b += need;
            n -= need;
        }
        return JB_OK;
    };
    auto on_dri = [&](const uint8_t *b, uint64_t n, uint64_t seg_at, bool in_stream) -> int {
        if (n < 2) return perr(JB_ERR_INVALID_DATA, seg_at, "Unexpected end of input data when reading segment content.");
        restart_interval = (uint32_t)((b[0] << 8) | b[1]);
        if (in_stream) dri_seen = true;
        return JB_OK;
    };

    // JpegDecoder.LoadTables (JpegDecoder.cs:319-360): SOI and RSTn are passed over, DHT / DQT / DRI are taken, anything
    // else with a length is skipped, EOI or the end of the data ends the walk without complaint.  Its DRI is the value the
    // decoder object holds when Identify() starts: it stays in force unless the stream brings its own.
    if (tables && tables_len) {
        Walker t{tables, tables_len, 0};
        while (t.pos < tables_len) {
            if (tables_used) *tables_used = t.pos;
            int m = t.next_marker();
            if (m < 0 || m == 0xD9) break;
            if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7)) continue;
            if (t.pos + 2 > tables_len) return perr(JB_ERR_INVALID_DATA, t.pos, "Unexpected end of input data when reading segment length.");
            uint64_t seglen = ((uint64_t)tables[t.pos] << 8) | tables[t.pos + 1];
            if (seglen < 2 || t.pos + seglen > tables_len) return perr(JB_ERR_INVALID_DATA, t.pos, "Unexpected end of input data reached.");
            const uint8_t *b = tables + t.pos + 2;
            const uint64_t n = seglen - 2, seg_at = t.pos;
            t.pos += seglen;
            int rc = JB_OK;
            if (m == 0xC4) rc = on_dht(b, n, seg_at);
            else if (m == 0xDB) rc = on_dqt(b, n, seg_at);
            else if (m == 0xDD) rc = on_dri(b, n, seg_at, false);
            if (rc) return rc;
            if (tables_used) *tables_used = t.pos;
        }
    }
    if (tables_used) return JB_OK; // (the guard frees p)

    if (length < 2 || data[0] != 0xFF || data[1] != 0xD8) return perr(JB_ERR_INVALID_DATA, 0, "Marker StartOfImage not found.");
    Walker w{data, length, 2};
    bool eoi = false;
    while (!eoi && w.pos < length) {
        int m = w.next_marker();
        if (m < 0) return perr(JB_ERR_INVALID_DATA, w.pos, "No marker found.");
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) { eoi = true; break; }
        if (w.pos + 2 > length) return perr(JB_ERR_INVALID_DATA, w.pos, "Unexpected end of input data when reading segment length.");
        uint64_t seglen = ((uint64_t)data[w.pos] << 8) | data[w.pos + 1];
        if (seglen < 2 || w.pos + seglen > length) return perr(JB_ERR_INVALID_DATA, w.pos, "Unexpected end of input data reached.");
        const uint8_t *b = data + w.pos + 2;
        uint64_t n = seglen - 2;
        uint64_t seg_at = w.pos;
        w.pos += seglen;
        switch (m) {
        case 0xC0: case 0xC1: case 0xC2: case 0xC3: case 0xC9: case 0xCA: {
            if (n < 6) return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse frame header.");
            p->sof_marker = m;
            im.sof = (uint8_t)(m - 0xC0);
            im.precision = b[0];
            im.height = (uint16_t)((b[1] << 8) | b[2]);
            im.width = (uint16_t)((b[3] << 8) | b[4]);
            im.component_count = b[5];
            if (im.component_count < 1 || im.component_count > 4 || n < 6 + 3ull * im.component_count)
                return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse frame header.");
            // what no scan decoder can be built for is refused here, where the frame header is read (the device-side
            // planning repeats these checks for descriptors that do not come from this walker)
            if (im.width == 0 || im.height == 0 || im.precision < 2 || im.precision > 16)
                return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse frame header.");
            for (int i = 0; i < im.component_count; i++) {
                comp_id[i] = b[6 + 3 * i];
                im.h[i] = b[7 + 3 * i] >> 4;
                im.v[i] = b[7 + 3 * i] & 15;
                comp_tq[i] = b[8 + 3 * i];
                if (comp_tq[i] > 3 || im.h[i] < 1 || im.h[i] > 4 || im.v[i] < 1 || im.v[i] > 4)
                    return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse frame header.");
            }
            have_frame = true;
            restart_at_sof = restart_interval;
            dri_before_sof = dri_seen;
            break;
        }
        case 0xC5: case 0xC6: case 0xC7: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
            return perr(JB_ERR_INVALID_DATA, seg_at, "This type of JPEG stream is not supported.");
        case 0xC4:
            if (int rc = on_dht(b, n, seg_at)) return rc;
            break;
        case 0xDB:
            if (int rc = on_dqt(b, n, seg_at)) return rc;
            break;
        case 0xDD:
            if (int rc = on_dri(b, n, seg_at, true)) return rc;
            break;
        case 0xDA: {
            if (!have_frame) return perr(JB_ERR_INVALID_DATA, seg_at, "Scan header appears before frame header.");
            if (n < 1) return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse scan header.");
            jb_scan_desc s{};
            s.component_count = b[0];
            if (s.component_count < 1 || s.component_count > 4 || n < 1 + 2ull * s.component_count + 3)
                return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse scan header.");
            for (int i = 0; i < s.component_count; i++) {
                int sel = b[1 + 2 * i], found = -1;
                for (int j = 0; j < im.component_count; j++)
                    if (comp_id[j] == sel) found = j;
                if (found < 0) return perr(JB_ERR_INVALID_DATA, seg_at, "The specified component is missing.");
                s.component_index[i] = (uint8_t)found;
                int td = b[2 + 2 * i] >> 4, ta = b[2 + 2 * i] & 15;
                if (td > 3 || ta > 3) return perr(JB_ERR_INVALID_DATA, seg_at, "Failed to parse scan header.");
                s.dc_table[i] = (int16_t)huff_latest[0][td];
                s.ac_table[i] = (int16_t)huff_latest[1][ta];
                if (im.sof != 3 && !qt_present[comp_tq[found]]) // lossless frames carry no DQT
                    return perr(JB_ERR_INVALID_DATA, seg_at, "Quantization table of component is not defined.");
                // table the component is rendered with: baseline = at its scan; progressive = slot state
                slot_comp[i] = found;
                memcpy(slot_qt[i], qt[comp_tq[found]], 128);
                if (im.sof != 2) memcpy(im.quant[found], qt[comp_tq[found]], 128);
            }
            const uint8_t *t = b + 1 + 2 * s.component_count;
            s.ss = t[0];
            s.se = t[1];
            s.ah = t[2] >> 4;
            s.al = t[2] & 15;
            s.restart_interval = restart_interval;
            s.entropy_offset = w.pos;
            uint64_t end = find_scan_end(data, length, w.pos);
            s.entropy_length = end - w.pos;
            w.pos = end;
            p->scans.push_back(s);
            break;
        }
        default:
            break; // APPn, COM, DAC, ...: skipped (ProcessOtherMarker)
        }
    }
    if (!have_frame) {
        g_parse_error = "Frame header was not found.";
        return JB_ERR_INVALID_OPERATION;
    }
    if (im.sof == 2) {
        int seen[4] = {0, 0, 0, 0};
        for (int i = 0; i < im.component_count; i++) {
            if (slot_comp[i] < 0) return perr(JB_ERR_INVALID_DATA, w.pos, "progressive frame leaves a component slot without scans");
            memcpy(im.quant[slot_comp[i]], slot_qt[i], 128);
            seen[slot_comp[i]]++;
        }
        for (int i = 0; i < im.component_count; i++)
            if (seen[i] != 1)
                return perr(JB_ERR_NOT_SUPPORTED, w.pos, "progressive scan order leaves component slots inconsistent (reference quirk P6)");
    }
    if (im.sof != 2)
        for (jb_scan_desc &s : p->scans) s.restart_interval = dri_before_sof ? restart_at_sof : restart_interval;
    p->consumed = w.pos;
    im.scan_count = (uint32_t)p->scans.size();
    im.scans = p->scans.data();
    im.table_count = (uint32_t)p->tables.size();
    im.tables = p->tables.data();
    guard.p = nullptr;
    *out = p;
    return JB_OK;
}

int jbh_parse(const uint8_t *data, uint64_t length, jbh_parsed **out) { return parse_impl(nullptr, 0, data, length, out); }

int jbh_parse_with_tables(const uint8_t *tables, uint64_t tables_length, const uint8_t *data, uint64_t length, jbh_parsed **out)
{
    return parse_impl(tables, tables_length, data, length, out);
}

int jbh_check_tables(const uint8_t *tables, uint64_t tables_length, uint64_t *used)
{
    uint64_t u = 0;
    int rc = parse_impl(tables, tables_length, nullptr, 0, nullptr, &u);
    if (used) *used = u;
    return rc;
}

const jb_image_desc *jbh_desc(const jbh_parsed *p) { return p ? &p->desc : nullptr; }
uint64_t jbh_consumed(const jbh_parsed *p) { return p ? p->consumed : 0; }
int jbh_sof_marker(const jbh_parsed *p) { return p ? p->sof_marker : 0; }
void jbh_free(jbh_parsed *p) { delete p; }

int jbh_parse_batch(const uint8_t *const *data, const uint64_t *length, int count, int threads, jbh_parsed **out)
{
    return jbh_parse_batch_with_tables(nullptr, 0, data, length, count, threads, out);
}

int jbh_parse_batch_with_tables(const uint8_t *tables, uint64_t tables_length, const uint8_t *const *data,
                                const uint64_t *length, int count, int threads, jbh_parsed **out)
{
    if (!data || !length || !out || count < 0) return -1;
    if (threads < 1) threads = 1;
    std::atomic<int> next{0}, failed{0};
    auto worker = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= count) break;
            if (parse_impl(tables, tables_length, data[i], length[i], &out[i]) != JB_OK) failed.fetch_add(1);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads && t < count; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    return failed.load();
}

int jbh_collect_descs(jbh_parsed *const *parsed, int count, jb_image_desc *descs)
{
    if (!parsed || !descs) return JB_ERR_ARGUMENT;
    for (int i = 0; i < count; i++) {
        if (!parsed[i]) return JB_ERR_ARGUMENT;
        descs[i] = parsed[i]->desc;
    }
    return JB_OK;
}

} // extern "C"
