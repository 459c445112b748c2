"""Host marker walk (JpegDecoder.Identify / marker loop mirror) -- CPU only.
Known answers: tests/JpegLibrary.Tests/Decoder/MetadataIdentifyTests.cs:19-130."""
import pytest

import jpeglibrary_b200 as J
import oracle_ffi as O
import synth
from conftest import golden_bytes


@pytest.mark.parametrize("name", ["cramps.jpg", "testorig12.jpg", "progress.jpg", "yellowcat_progressive_restart.jpg"])
def test_identify_known_answers(name, golden):
    k = golden["identify"][name]
    dec = J.JpegDecoder()
    dec.SetInput(golden_bytes(name))
    consumed = dec.Identify()
    assert (dec.Width, dec.Height, dec.NumberOfComponents, dec.Precision) == (k["Width"], k["Height"], k["NumberOfComponents"], k["Precision"])
    assert consumed == k["JpegStreamSize"]


def test_identify_requires_input():
    dec = J.JpegDecoder()
    with pytest.raises(J.InvalidOperationException):
        dec.Identify()
    with pytest.raises(J.InvalidOperationException):
        dec.Width
    dec.SetInput(b"\x00\x01\x02\x03")
    with pytest.raises(J.InvalidDataException):
        dec.Identify()


@pytest.mark.parametrize("name", ["lake.jpg", "cramps.jpg", "testorig12.jpg", "progress.jpg", "yellowcat_progressive_restart.jpg"])
def test_descriptor_matches_oracle_headers(name):
    blob = golden_bytes(name)
    p = J.Parsed(blob)
    d = p.desc
    o = O.decode(blob, want_rgb=False)
    assert (d.width, d.height, d.component_count, d.precision, d.sof) == (o.width, o.height, o.ncomp, o.precision, o.sof)
    assert d.scan_count == o.nscans
    for c in range(o.ncomp):
        assert (d.h[c], d.v[c]) == (o.comp_h[c], o.comp_v[c])
        assert list(d.quant[c]) == list(o.qt[c])
    for i in range(o.nscans):
        s, t = d.scans[i], o.scans[i]
        assert (s.ss, s.se, s.ah, s.al, s.restart_interval, s.entropy_offset) == (t.ss, t.se, t.ah, t.al, t.restart_interval, t.entropy_offset)
        assert blob[s.entropy_offset + s.entropy_length] == 0xFF  # the scan ends at a marker


def test_restart_interval_and_sampling_getters():
    blob = synth.synth_jpeg(2, 80, 48, restart_rows=1)
    dec = J.JpegDecoder()
    dec.SetInput(blob)
    dec.Identify()
    assert dec.GetMaximumHorizontalSampling() == 2 and dec.GetMaximumVerticalSampling() == 2
    assert dec.GetHorizontalSampling(1) == 1 and dec.GetVerticalSampling(0) == 2
    with pytest.raises(J.ArgumentException):
        dec.GetHorizontalSampling(3)
    assert J.Parsed(blob).desc.scans[0].restart_interval == 5  # one MCU row of an 80-px-wide 4:2:0 image


def test_decode_call_order_errors():
    dec = J.JpegDecoder()
    with pytest.raises(J.InvalidOperationException):
        dec.Decode()
    dec.SetInput(synth.synth_jpeg(2, 32, 32))
    with pytest.raises(J.InvalidOperationException):
        dec.Decode()  # "The output buffer is not specified."
    with pytest.raises(J.ArgumentException):
        dec.SetOutputWriter(None)


# ----------------------------------------------------------------------------- scan-list planning (host only)
def _plan(blob):
    import ctypes as C
    p = J.Parsed(blob)
    n = p.desc.scan_count
    out = (C.c_int32 * (10 * max(n, 1)))()
    k = J._native.cuda.jb_plan_scans(C.byref(p.desc), out, n)
    assert k >= 0, k
    rows = [list(out[10 * i:10 * i + 10]) for i in range(k)]
    return [dict(rank=r[0], ndep=r[1], whole=r[2], deps=[d for d in r[3:9] if d >= 0], consumed=bool(r[9])) for r in rows]


def test_progressive_scan_dependencies_follow_component_and_band_overlap():
    """libjpeg's script: 0 DC first (all), 1 Y 1-5, 2 Cr, 3 Cb, 4 Y 6-63, 5 Y refine, 6 DC refine, 7 Cr refine, 8 Cb refine,
    9 Y refine.  Producers are transitively reduced and followed block by block (same component list, one segment)."""
    import synth
    plan = _plan(synth.synth_jpeg(3, 160, 112, progressive=True, subsampling="4:2:0"))
    assert len(plan) == 10
    # (Y 6-63 follows Y 1-5 although their bands are disjoint: in a damaged stream the first may write up to 15 places
    # behind its band, and the reference's scan order decides who wins)
    assert [p["deps"] for p in plan] == [[], [], [], [], [1], [4], [0], [2], [3], [5]]
    assert all(p["whole"] == 0 for p in plan)                                  # every producer has the consumer's unit order
    assert [p["consumed"] for p in plan] == [True, True, True, True, True, True, False, False, False, False]
    for k, p in enumerate(plan):                                               # job order: producers rank in front of consumers
        assert all(plan[d]["rank"] < p["rank"] for d in p["deps"]), k
    assert sorted(p["rank"] for p in plan) == list(range(10))
    assert plan[9]["rank"] < plan[6]["rank"] and plan[5]["rank"] < plan[7]["rank"]   # the long luma chain starts first


def test_progressive_scans_with_restart_intervals_are_waited_for_as_a_whole():
    import synth
    plan = _plan(synth.synth_jpeg(3, 160, 112, progressive=True, subsampling="4:4:4", restart_blocks=6))
    assert plan[4]["deps"] == [1] and plan[4]["whole"] == 1                    # producers in several segments: no block-wise following
    assert plan[5]["deps"] == [4] and plan[5]["whole"] == 1
    assert plan[9]["deps"] == [5] and plan[9]["whole"] == 1


def test_sequential_scan_lists():
    import oracle_ffi as O
    import synth
    src = synth.synth_jpeg(3, 96, 64, subsampling="4:2:0")
    assert _plan(src) == []                                                    # the fast path: no scan list
    d = O.decode(src, want_rgb=False)
    plan = _plan(synth.resequence_scans(src, d, [[0], [1], [2]]))
    assert [p["deps"] for p in plan] == [[], [], []]                           # different components: the scans commute
    plan = _plan(synth.resequence_scans(src, d, [[0, 1, 2], [0]]))
    assert plan[1]["deps"] == [0] and plan[1]["whole"] == 1                    # same blocks, another walk: wait for all of it
    plan = _plan(synth.resequence_scans(src, d, [[0], [0]]))
    assert plan[1]["deps"] == [0] and plan[1]["whole"] == 0                    # same walk: follow block by block


def _dri(n):
    return b"\xff\xdd\x00\x04" + n.to_bytes(2, "big")


def test_sequential_frames_take_the_restart_interval_at_the_frame_header():
    """The reference builds its sequential / lossless scan decoder when it reads the SOF and the decoder takes the restart
    interval once, in its constructor (JpegDecoder.cs:569, JpegHuffmanBaselineScanDecoder.cs:38, ...Lossless...:32): every
    scan of the frame uses the value JpegDecoder holds at that moment -- after Identify() that is the LAST DRI of the
    stream, unless a DRI segment precedes the SOF.  Progressive scans read it per scan (...Progressive...:78).
    Walker and oracle must agree on it."""
    blob, _ = synth.synth_lossless_scans(3, 32, 24, scans=[dict(components=[0]), dict(components=[1, 2])], restart=5)
    sos = [i for i in range(len(blob) - 1) if blob[i] == 0xFF and blob[i + 1] == 0xDA]
    sof = blob.index(b"\xff\xc3")
    assert len(sos) == 2 and blob.count(b"\xff\xdd") == 1
    two = blob[:sos[1]] + _dri(9) + blob[sos[1]:]            # a second DRI in front of the second scan
    early = two[:sof] + _dri(3) + two[sof:]                  # and one in front of the frame header
    for stream, want in ((blob, 5), (two, 9), (early, 3)):
        d = J.Parsed(stream).desc
        assert [d.scans[i].restart_interval for i in range(d.scan_count)] == [want, want]
        try:
            o = O.decode(stream, want_rgb=False)
            assert [s.restart_interval for s in o.scans] == [want, want]
        except O.OracleError as e:                            # (coded with 5: the other intervals do not fit the data)
            assert want != 5 and e.code in (-1, -2)
    # progressive: the interval in force at each SOS
    prog = synth.encode_jpeg(synth.synth_rgb(1, 48, 32), subsampling="4:4:4", progressive=True, restart_blocks=4)
    sos = [i for i in range(len(prog) - 1) if prog[i] == 0xFF and prog[i + 1] == 0xDA]
    d = J.Parsed(prog[:sos[2]] + _dri(0) + prog[sos[2]:]).desc
    assert [d.scans[i].restart_interval for i in range(3)] == [4, 4, 0]


def test_scan_header_checks_follow_the_reference():
    """Two streams the corrupted-stream campaign found (profiles/r2_fuzz_campaign.txt).  The reference never validates the
    spectral selection: a single-component scan with Ss = 0 is a DC scan whatever Se says (JpegHuffmanProgressiveScanDecoder
    .cs:149-166) -- and it asks for the DC table of EVERY such scan, DC refinement scans included (:149-152)."""
    import ctypes as C
    import os
    here = os.path.join(os.path.dirname(__file__), "fixtures")
    accepted = open(os.path.join(here, "fuzz_gray_progressive_dc_scan_with_se.jpg"), "rb").read()      # first scan: Ss 0, Se 119
    refused = open(os.path.join(here, "fuzz_gray_progressive_dc_refine_without_table.jpg"), "rb").read()  # DHT of the DC table swallowed
    O.decode(accepted, want_rgb=False)
    with pytest.raises(O.OracleError) as e:
        O.decode(refused, want_rgb=False)
    assert e.value.code == -1 and "Huffman table" in str(e.value)
    for blob, ok in ((accepted, True), (refused, False)):
        p = J.Parsed(blob)
        out = (C.c_int32 * (10 * p.desc.scan_count))()
        k = J._native.cuda.jb_plan_scans(C.byref(p.desc), out, p.desc.scan_count)
        assert (k == p.desc.scan_count) if ok else (k == J._native.JB_ERR_INVALID_DATA), k


# ---- abbreviated streams behind JpegDecoder.LoadTables (JpegDecoder.cs:313-360) -------------------------------------
def _same_descriptor(a, b, offset_shift):
    assert (a.width, a.height, a.component_count, a.precision, a.sof, a.scan_count) == (b.width, b.height, b.component_count, b.precision, b.sof, b.scan_count)
    for c in range(a.component_count):
        assert (a.h[c], a.v[c]) == (b.h[c], b.v[c]) and list(a.quant[c]) == list(b.quant[c])
    for i in range(a.scan_count):
        s, t = a.scans[i], b.scans[i]
        assert (s.ss, s.se, s.ah, s.al, s.restart_interval, s.entropy_length, s.component_count) == (t.ss, t.se, t.ah, t.al, t.restart_interval, t.entropy_length, t.component_count)
        assert s.entropy_offset + offset_shift == t.entropy_offset
        for k in range(s.component_count):
            for x, y in ((s.dc_table[k], t.dc_table[k]), (s.ac_table[k], t.ac_table[k])):
                assert (x < 0) == (y < 0)
                if x >= 0:
                    u, v = a.tables[x], b.tables[y]
                    assert (u.table_class, u.identifier, list(u.bits), u.value_count, list(u.values)[:u.value_count]) == \
                           (v.table_class, v.identifier, list(v.bits), v.value_count, list(v.values)[:v.value_count])


@pytest.mark.parametrize("kw", [dict(), dict(restart_rows=1), dict(subsampling="4:4:4"), dict(progressive=True)], ids=str)
def test_abbreviated_stream_behind_load_tables_gives_the_descriptor_of_the_whole_stream(kw):
    blob = synth.synth_jpeg(5, 72, 40, **kw)
    progressive = kw.get("progressive", False)
    # (a progressive file of libjpeg carries further DHT segments between its scans: only the ones in front move)
    tables, rest = synth.split_tables(blob, move=(0xC4, 0xDB, 0xDD))
    assert len(tables) > 100 and len(rest) < len(blob)
    with pytest.raises(J.InvalidDataException):  # the abbreviated stream alone refers to tables nobody defined
        if progressive:
            raise J.InvalidDataException("n/a")
        J.Parsed(rest)
    whole, abbr = J.Parsed(blob), J.Parsed(rest, tables)
    _same_descriptor(abbr.desc, whole.desc, len(blob) - len(rest))
    # the oracle walks the two streams the same way
    o = O.decode(rest, want_rgb=False, tables=tables)
    w = O.decode(blob, want_rgb=False)
    assert all((a == b).all() for a, b in zip(o.coef, w.coef)) and (o.planes == w.planes).all()
    assert o.restart_interval == w.restart_interval


def test_load_tables_call_sequence_and_errors():
    blob = synth.synth_jpeg(6, 48, 32, restart_rows=1)
    dht_dqt, rest = synth.split_tables(blob, move=(0xC4, 0xDB, 0xDD))
    dqt_only, _ = synth.split_tables(blob, move=(0xDB,))
    dht_only, _ = synth.split_tables(blob, move=(0xC4,))
    dri_only = b"\xff\xd8\xff\xdd\x00\x04" + (3).to_bytes(2, "big")  # no EOI: the walk ends with the data
    dec = J.JpegDecoder()
    dec.SetInput(rest)
    with pytest.raises(J.InvalidDataException):
        dec.Identify()  # "Quantization table of component is not defined."
    # two calls add up (the first stream's EOI ends ITS walk only), a DRI loaded this way stays in force when the
    # stream brings none -- and the value Identify() leaves is the stream's own when it has one
    dec.LoadTables(dqt_only)
    dec.LoadTables(dht_only)
    dec.Identify()
    assert dec._parsed.desc.scans[0].restart_interval == 0  # the DRI segment went away with the tables
    dec.LoadTables(dri_only)
    dec.Identify()
    assert dec._parsed.desc.scans[0].restart_interval == 3 == J.Parsed(blob).desc.scans[0].restart_interval
    o = O.decode(rest, want_rgb=False, tables=dqt_only[:-2] + dht_only[:-2] + dri_only)
    assert o.restart_interval == 3
    assert (dec.Width, dec.Height) == (48, 32)
    dec.ResetTables()
    with pytest.raises(J.InvalidDataException):
        dec.Identify()
    # what LoadTables itself raises: a DHT segment cut short, a segment running past the end of the data
    bad = bytearray(dht_only)
    bad[4:6] = (5).to_bytes(2, "big")
    with pytest.raises(J.InvalidDataException, match="Failed to parse Huffman table"):
        dec.LoadTables(bytes(bad))
    with pytest.raises(J.InvalidDataException, match="Unexpected end of input data"):
        dec.LoadTables(dht_only[:40])
    with pytest.raises(O.OracleError, match="Failed to parse Huffman table"):
        O.decode(rest, want_rgb=False, tables=bytes(bad))
    dec.LoadTables(b"")            # nothing to walk
    dec.LoadTables(b"\x00\x01\x02")  # no marker at all: ends silently (JpegDecoder.cs:326-329)


def test_damaged_tables_streams_get_the_oracles_verdict():
    """Differential: 400 damaged tables streams (byte flips, cuts, duplicated and dropped segments) in front of one
    abbreviated image -- the host walk (LoadTables, then Identify + the walk of Decode) and the oracle accept and refuse
    the same ones, with the same message, and agree on the quantisation tables of what they accept."""
    import ctypes as C
    import numpy as np
    from jpeglibrary_b200 import _native as N
    rng = np.random.default_rng(77)
    blob = synth.synth_jpeg(8, 64, 48, restart_rows=1)
    tables, rest = synth.split_tables(blob, move=(0xC4, 0xDB, 0xDD))
    verdicts = {"ok": 0, "load": 0, "walk": 0, "plan": 0}
    for trial in range(400):
        t = bytearray(tables)
        kind = trial % 4
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                t[int(rng.integers(2, len(t)))] = int(rng.integers(0, 256))
        elif kind == 1:
            t = t[:int(rng.integers(2, len(t)))]
        elif kind == 2:
            a = int(rng.integers(2, len(t) - 4))
            t = t[:a] + t[a + int(rng.integers(1, 40)):]
        else:
            a = int(rng.integers(2, len(t) - 4))
            t = t[:a] + bytes([0xFF, int(rng.integers(0xC0, 0x100))]) + t[a:]
        t = bytes(t)
        try:
            want, werr = O.decode(rest, want_rgb=False, tables=t), None
        except O.OracleError as e:
            want, werr = None, str(e)
        dec = J.JpegDecoder()
        got, gerr, stage = None, None, "load"
        try:
            dec.LoadTables(t)
            stage = "walk"
            dec.SetInput(rest)
            dec.Identify()
            got = dec._parsed.desc
            stage = "plan"  # jb_decode_batch_create's host-side planning (never touches the device): missing tables, ...
            rc = N.cuda.jb_plan_scans(C.byref(got), None, 0)
            if rc < 0:
                gerr = f"planning refused the descriptor ({rc})"
        except (J.InvalidDataException, J.InvalidOperationException) as e:
            gerr = str(e)
        if werr is not None and any(k in werr for k in ("Invalid Huffman code", "magnitude category", "bit stream ended", "restart marker", "end of JPEG data stream")):
            # the tables parse but the image does not decode with them: the scan's failure is the GPU path's to report
            assert gerr is None, (trial, werr, gerr)
            continue
        assert (werr is None) == (gerr is None), (trial, werr, gerr)
        if werr is None:
            verdicts["ok"] += 1
            for c in range(3):
                assert list(got.quant[c]) == list(want.qt[c]), trial
        else:
            if stage != "plan":
                assert gerr.split(". ", 1)[-1].rstrip(".") in werr, (trial, werr, gerr)
            verdicts[stage] += 1
    assert verdicts["ok"] > 20 and verdicts["load"] > 50, verdicts


@pytest.mark.parametrize("name", ["lake.jpg", "cramps.jpg", "testorig12.jpg", "progress.jpg", "lossless1_s22.jpg", "lossless5_s22.jpg"])
def test_golden_assets_split_into_tables_and_abbreviated_streams(name):
    """The reference's own assets as TIFF-style pairs (tables stream + abbreviated stream): same descriptor from the host
    walk, same coefficients and planes from the oracle as the whole file -- whose planes are pinned on the reference's
    golden PNGs (tests/test_oracle_golden.py)."""
    blob = golden_bytes(name)
    tables, rest = synth.split_tables(blob, move=(0xC4, 0xDB, 0xDD))
    whole, abbr = J.Parsed(blob), J.Parsed(rest, tables)
    _same_descriptor(abbr.desc, whole.desc, len(blob) - len(rest))
    o, w = O.decode(rest, want_rgb=False, tables=tables), O.decode(blob, want_rgb=False)
    assert (o.planes == w.planes).all() and all((a == b).all() for a, b in zip(o.coef, w.coef))
