import os, sys
os.environ["JB_DEBUG_STATUS"] = "1"
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, synth, oracle_ffi as O, jpeglibrary_b200 as J
for cut in (0, 2, 3, 4):
    for env in ({}, {"JB_SS_SHIFT": "15"}):
        os.environ.pop("JB_SS_SHIFT", None); os.environ.update(env)
        blob, coef = synth.handmade_grey(64, 48, seed=0, cut=cut)
        print("cut", cut, env, flush=True)
        try:
            lay, c = J.decode_coefficients(blob)
            want = O.scan_order_coefficients(O.decode(blob, want_rgb=False)).reshape(-1, 64)
            bad = np.nonzero((c != want).any(axis=1))[0]
            print("  decoded; differing blocks:", bad[:10], len(bad), flush=True)
        except Exception as e:
            print("  raised", type(e).__name__, e, flush=True)
