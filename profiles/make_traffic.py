"""Extract per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, per launch) from an ncu --set full
report and write profiles/traffic.json, which bench.py reads to fill roofline.traffic.

usage: python profiles/make_traffic.py gpurun_out/prof.ncu-rep <images in the profiled batch> [note]"""
import csv
import json
import os
import subprocess
import sys

NAMES = {  # kernel function -> the interval name jb_decode_batch_profile reports
    "jb_k0_restart_scan": "jb_k0_restart_scan",
    "jb_k0b_segment_descs": "jb_k0b_segment_descs",
    "jb_k1_huff_flat": "jb_k1_huff_segments",
    "jb_k2_idct_color_warp": "jb_k2_idct_color",
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(rep, images, note=""):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    res = {}
    for r in rows[2:]:
        for fn, name in NAMES.items():
            if r[ik].startswith(fn) or (" " + fn) in r[ik]:
                b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
                res[name] = {"dram_bytes_per_image": b / images, "dram_bytes_per_launch": b, "kernel": r[ik][:80]}
    doc = {"source": os.path.basename(rep), "images_in_profiled_batch": images, "note": note, "kernels": res}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    json.dump(doc, open(path, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "")
