"""CPU differential fuzz of the HOST side of the path: the marker walk (jpeglibrary_b200/host/jpeg_host.cpp, standing
for JpegDecoder.Identify + the marker loop of Decode, JpegDecoder.cs:75-162, :509-617) followed by the planning that
jb_decode_batch_create runs on the host (jb_plan_scans: never touches the device) must accept and refuse damaged
headers like the oracle's walk does.  What only the scan decoders can find (bad codes, premature end, missing restart
markers) is the kernels' to report and is covered by the -m gpu fuzz tests."""
import ctypes as C
import os

import numpy as np
import pytest

import jpeglibrary_b200 as J
import oracle_ffi as O
import synth
from conftest import golden_bytes
from jpeglibrary_b200 import _native as N
from test_gpu_fuzz import mutate_header

SCAN_FAILURES = ("Invalid Huffman code", "magnitude category", "bit stream ended", "restart marker", "end of JPEG data stream")


def _scan_list(scans, restart):
    src = synth.synth_jpeg(33, 120, 88, subsampling="4:2:0", quality=88)
    return synth.resequence_scans(src, O.decode(src, want_rgb=False), scans, restart)


def _bases():
    rgb = synth.synth_rgb(40, 160, 112)
    return {
        "restart": synth.encode_jpeg(rgb, quality=85, subsampling="4:2:0", restart_blocks=4),
        "plain_444": synth.encode_jpeg(rgb, quality=90, subsampling="4:4:4", optimize=True),
        "progressive": synth.encode_jpeg(rgb, quality=85, subsampling="4:2:0", progressive=True),
        "lossless": synth.synth_lossless(42, 96, 64, predictor=4, restart=24)[0],
        "gray_rows": synth.encode_jpeg(rgb, quality=80, gray=True, restart_rows=1),
        "extended_12bit": golden_bytes("testorig12.jpg"),
        "progressive_restart": golden_bytes("yellowcat_progressive_restart.jpg"),
        "lossless_golden": golden_bytes("lossless3_s22.jpg"),
        "scan_list": _scan_list([[0, 1], [2]], 5),
        "scan_list_second_pass": _scan_list([[0, 1, 2], [0]], 0),
        "lossless_scans": synth.synth_lossless_scans(44, 48, 32, [dict(components=[0], predictor=1), dict(components=[1, 2], predictor=4)])[0],
    }


BASES = _bases()


def _segments(blob):
    """(position, length) of every marker segment that carries a length: the headers in front of and BETWEEN the scans"""
    out, i = [], 2
    while i + 4 <= len(blob):
        if blob[i] != 0xFF or blob[i + 1] in (0x00, 0xFF) or 0xD0 <= blob[i + 1] <= 0xD9:
            i += 1
            continue
        ln = int.from_bytes(blob[i + 2:i + 4], "big")
        out.append((i, 2 + ln))
        i += 2 + ln
    return out


def mutate_any_segment(blob, rng, kind, segments):
    """like mutate_header, anywhere a segment lies: the tables, DRI and scan headers between the scans of a progressive
    frame or a scan list as well"""
    at, ln = segments[int(rng.integers(len(segments)))]
    b = bytearray(blob)
    p = at + int(rng.integers(0, min(ln, 24)))
    if kind == 0:
        b[p] ^= 1 << int(rng.integers(8))
    elif kind == 1:
        b[p] = int(rng.integers(256))
    elif kind == 2:
        del b[p:p + int(rng.integers(1, 4))]
    else:
        b[p:p] = bytes(rng.integers(0, 256, int(rng.integers(1, 4))).astype(np.uint8))
    return bytes(b)


@pytest.mark.parametrize("name", list(BASES))
def test_damaged_headers_get_the_oracles_verdict_from_walk_and_planning(name):
    blob = BASES[name]
    rng = np.random.default_rng(11 + sum(map(ord, name)))
    counts = {"ok": 0, "refused": 0, "scan": 0, "out of scope": 0, "deviation 6": 0, "deviation 3": 0}
    segments = _segments(blob)
    problems = []
    for trial in range(int(os.environ.get("JB_HOST_FUZZ_TRIALS", "600"))):
        bad = mutate_header(blob, rng, trial % 4) if trial % 2 else mutate_any_segment(blob, rng, (trial // 2) % 4, segments)
        try:
            want, werr, wcode = O.decode(bad, want_rgb=False), None, 0
        except O.OracleError as e:
            want, werr, wcode = None, str(e), e.code  # JO_ERR_* = JB_ERR_* for the three exception classes
        if werr is not None and ("outside the oracle's scope" in werr or "out of memory" in werr or "overhangs the component plane" in werr):
            continue  # (the last: lossless sampling factors with which the reference indexes out of its scanline store and
            #            dies of an ArgumentOutOfRangeException; refused as NOT_SUPPORTED here)
        desc, gerr, gcode = None, None, 0
        try:
            p = J.Parsed(bad)
            desc = p.desc
            if desc.scan_count:  # (a frame without scans never reaches the C-ABI)
                gcode = min(0, N.cuda.jb_plan_scans(C.byref(desc), None, 0))
                if gcode:
                    gerr = f"planning refused the descriptor ({gcode})"
        except J.InvalidDataException as e:
            gerr, gcode = str(e), N.JB_ERR_INVALID_DATA
        except J.InvalidOperationException as e:
            gerr, gcode = str(e), N.JB_ERR_INVALID_OPERATION
        except J.NotSupportedException as e:
            gerr, gcode = str(e), N.JB_ERR_NOT_SUPPORTED
        if gcode == N.JB_ERR_INVALID_DATA and desc is not None and (werr is None or any(k in werr for k in SCAN_FAILURES)) and any(
                desc.scans[i].component_count == 1 and desc.scans[i].ss > 0 and (desc.scans[i].se > 63 or desc.scans[i].ss > desc.scans[i].se)
                for i in range(desc.scan_count)) and desc.sof == 2:
            # documented deviation (deviation 3): single-component AC scans with Ss > Se or Se > 63 are
            # refused ("Failed to parse scan header."); the reference validates neither (the first decodes nothing, the
            # second clamps or -- refinement scans -- runs past the block, where the oracle stops as well)
            counts["deviation 3"] += 1
            continue
        if gcode == N.JB_ERR_NOT_SUPPORTED and (werr is None or any(k in werr for k in SCAN_FAILURES)):
            counts["out of scope"] += 1  # a layout the GPU path documents as unsupported: no CPU path, no verdict
            continue
        if werr is not None and any(k in werr for k in SCAN_FAILURES):
            # the oracle walks into the scan and fails there.  The host either leaves that to the kernels or has refused
            # the headers already (a later scan without its table, ...): then with the SAME exception class
            counts["scan"] += 1
            if gerr is not None and wcode == N.JB_ERR_INVALID_OPERATION and gcode == N.JB_ERR_INVALID_DATA:
                # documented deviation (deviation 6): a stream with SEVERAL defects whose earlier scan
                # fails with "Expect restart marker." (InvalidOperationException) while the host-side checks, made for the
                # whole file before anything is decoded, have met a later defect (a scan without its table, a component
                # slot without scans) and raise InvalidDataException
                counts["deviation 6"] += 1
                continue
            if gerr is not None and gcode != wcode:
                problems.append(f"trial {trial}: the scan fails in the oracle [{werr}] but the host refused the headers [{gerr}]")
            continue
        if (werr is None) != (gerr is None):
            problems.append(f"trial {trial}: oracle [{werr}] host [{gerr}]")
            continue
        if werr is None:
            counts["ok"] += 1
            same = (desc.width, desc.height, desc.component_count, desc.precision, desc.sof, desc.scan_count) == \
                   (want.width, want.height, want.ncomp, want.precision, want.sof, want.nscans)
            same = same and all(list(desc.quant[c]) == list(want.qt[c]) for c in range(want.ncomp) if want.sof != 3)
            same = same and all((desc.scans[i].ss, desc.scans[i].se, desc.scans[i].ah, desc.scans[i].al, desc.scans[i].restart_interval, desc.scans[i].entropy_offset) ==
                                (want.scans[i].ss, want.scans[i].se, want.scans[i].ah, want.scans[i].al, want.scans[i].restart_interval, want.scans[i].entropy_offset)
                                for i in range(want.nscans))
            if not same:
                problems.append(f"trial {trial}: both accept, descriptors differ")
        else:
            counts["refused"] += 1
            if gcode != wcode:
                problems.append(f"trial {trial}: exception classes differ: oracle [{werr}] host [{gerr}]")
    print(name, counts, len(problems))
    assert not problems, "\n".join(problems[:20])
    assert counts["ok"] > 10 and counts["refused"] > 10 and counts["deviation 6"] <= max(1, counts["scan"] // 50)


# ---- descriptors that do not come from the walker: the C-ABI's own validation ------------------------------------------
def _clone(d, keep):
    """deep copy of a descriptor into arrays the test owns (appended to `keep` so that they stay alive)"""
    scans = (N.ScanDesc * max(1, d.scan_count))()
    for i in range(d.scan_count):
        C.memmove(C.byref(scans[i]), C.byref(d.scans[i]), C.sizeof(N.ScanDesc))
    tables = (N.HuffSpec * max(1, d.table_count))()
    for i in range(d.table_count):
        C.memmove(C.byref(tables[i]), C.byref(d.tables[i]), C.sizeof(N.HuffSpec))
    out = N.ImageDesc()
    C.memmove(C.byref(out), C.byref(d), C.sizeof(N.ImageDesc))
    out.scans = C.cast(scans, type(out.scans))
    out.tables = C.cast(tables, type(out.tables))
    keep.extend([scans, tables])
    return out, scans, tables


@pytest.mark.parametrize("name", list(BASES))
def test_corrupted_descriptors_are_refused_or_planned_never_trusted(name):
    """jb_image_desc is caller-provided memory (the C# binding fills it from JpegDecoder's state): whatever the fields say,
    the planning of jb_decode_batch_create (jb_plan_scans: host only) answers with a status code.  Random field damage
    (counts, indices, table references, offsets and lengths up to 2^64, sampling factors, precision, frame type); the
    same loop ran 33 000 descriptors under AddressSanitizer without a report (profiles/r2g_sanitizer_host_asan.txt)."""
    rng = np.random.default_rng(99 + sum(map(ord, name)))
    p = J.Parsed(BASES[name])
    values = [0, 1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 63, 64, 65, 127, 128, 255, 256, 1023, 65535, 2 ** 31 - 1, 2 ** 32 - 1, 2 ** 64 - 8]
    refused = 0
    for trial in range(int(os.environ.get("JB_HOST_FUZZ_TRIALS", "300"))):
        keep = []
        d, scans, tables = _clone(p.desc, keep)
        for _ in range(int(rng.integers(1, 4))):
            k, v = int(rng.integers(0, 16)), int(values[int(rng.integers(len(values)))])
            si, ci, ti = int(rng.integers(0, max(1, d.scan_count))), int(rng.integers(0, 4)), int(rng.integers(0, max(1, d.table_count)))
            if k == 0: d.width = v & 0xFFFF
            elif k == 1: d.height = v & 0xFFFF
            elif k == 2: d.component_count = v & 0xFF
            elif k == 3: d.precision = v & 0xFF
            elif k == 4: d.sof = v & 0xFF
            elif k == 5: d.h[ci] = v & 0xFF
            elif k == 6: d.v[ci] = v & 0xFF
            elif k == 7: d.scan_count = min(v & 0xFFFFFFFF, len(scans))    # (a count beyond the array is the caller lying about
            elif k == 8: d.table_count = min(v & 0xFFFFFFFF, len(tables))  #  its own memory: outside any contract)
            elif k == 9: scans[si].component_count = v & 0xFF
            elif k == 10: scans[si].component_index[ci] = v & 0xFF
            elif k == 11: scans[si].dc_table[ci] = ((v + 2 ** 15) % 2 ** 16) - 2 ** 15
            elif k == 12: scans[si].ac_table[ci] = ((v + 2 ** 15) % 2 ** 16) - 2 ** 15
            elif k == 13:
                f = int(rng.integers(0, 5))
                setattr(scans[si], ["ss", "se", "ah", "al", "restart_interval"][f], v & (0xFF if f < 4 else 0xFFFFFFFF))
            elif k == 14:
                setattr(scans[si], "entropy_offset" if rng.integers(2) else "entropy_length", v)
            elif rng.integers(2): tables[ti].value_count = v & 0xFFFF
            else: tables[ti].bits[int(rng.integers(16))] = v & 0xFF
        rc = N.cuda.jb_plan_scans(C.byref(d), None, 0)
        assert rc >= 0 or rc in (N.JB_ERR_INVALID_DATA, N.JB_ERR_NOT_SUPPORTED, N.JB_ERR_ARGUMENT, N.JB_ERR_INVALID_OPERATION), (trial, rc)
        refused += rc < 0
    assert refused > 20
    # what must be refused whatever else the descriptor says
    L = p.desc.length
    for damage in (lambda s: setattr(s, "entropy_offset", L + 5),
                   lambda s: (setattr(s, "entropy_offset", 2 ** 64 - 8), setattr(s, "entropy_length", 16)),
                   lambda s: s.component_index.__setitem__(0, 7)):
        keep = []
        d, scans, tables = _clone(p.desc, keep)
        damage(scans[d.scan_count - 1])
        assert N.cuda.jb_plan_scans(C.byref(d), None, 0) < 0
