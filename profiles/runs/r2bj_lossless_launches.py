"""Launch list of the lossless chain (K0, K1d entropy, K1d predict, K5 output): 64 frames of 1024 x 1024 x 3, 8 bits,
predictor 1, one restart interval per row.  Run under ncu --metrics gpu__time_duration.sum (profiles/runs/r2bj.sh)."""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import multiprocessing as mp
import synth
import jpeglibrary_b200 as J


def gen(i):
    return synth.synth_lossless(i, 1024, 1024, precision=8, predictor=1, restart=1024)[0]


if __name__ == "__main__":
    with mp.get_context("fork").Pool(4) as pool:
        blobs = pool.map(gen, range(4))
    batch = [np.frombuffer(blobs[i % 4], dtype=np.uint8) for i in range(64)]
    with J.JpegBatchDecoder(batch, J.JB_OUT_PLANAR_I16, device_output=True) as dec:
        dec.upload()
        for _ in range(3):
            dec.launch()
        dec.finish()
        assert dec.status() == [0] * 64
