import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jpeglibrary_b200 as J, oracle_ffi as O, test_gpu_fuzz as F
for f in ("fuzz_progressive_192", "fuzz_progressive_906", "fuzz_progressive_499", "fuzz_progressive_restart_949"):
    blob = open(f"scratch/{f}.jpg", "rb").read()
    want, werr = F.run_oracle(blob)
    got, gerr = F.run_gpu(blob)
    print(f, "oracle", werr, "gpu", gerr)
    if want is None or got is None: continue
    print("  planes equal", np.array_equal(got, want.planes))
    lay, coef = J.decode_coefficients(blob)
    o = want
    for c in range(o.ncomp):
        w, h = lay.comp_blocks_w[c], lay.comp_blocks_h[c]
        plane = coef[lay.comp_block_offset[c]:lay.comp_block_offset[c] + w * h].reshape(h, w, 64)
        aw, ah = o.alloc_w[c], o.alloc_h[c]
        d = plane[:ah, :aw] != o.coef[c][:ah, :aw]
        if d.any():
            by, bx, k = np.nonzero(d)
            print(f"  comp {c}: {d.sum()} coefficient(s) differ in {len(set(zip(by.tolist(), bx.tolist())))} blocks; first at block ({by[0]},{bx[0]}) k={k[0]} gpu={plane[by[0],bx[0],k[0]]} oracle={o.coef[c][by[0],bx[0],k[0]]}; zigzag positions {sorted(set(k.tolist()))[:12]}")
    p = J.Parsed(blob); d = p.desc
    for i in range(d.scan_count):
        s = d.scans[i]; print('   scan', i, list(s.component_index)[:s.component_count], 'ss', s.ss, 'se', s.se, 'ah', s.ah, 'al', s.al, 'len', s.entropy_length)
