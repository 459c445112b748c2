"""Throughput of the two SURVEY 8(f) rows that bench.py has no leg for, with their parity gates (test infrastructure:
uses the oracle as the checker; run on a GPU box from the repo root):

  python tests/campaigns/next_rows_bench.py [--frames 256] [--distinct 16] [--steps 5]

(f)1 JpegOptimizer: a batch of 4K 4:2:0 baseline frames (restart interval = one MCU row, libjpeg's standard tables) through
     JpegBatchOptimizer -- K0/K1 decode to coefficient stores in HBM, K3b histograms, K3c optimised tables, K4 re-pack --
     inputs resident in HBM, a step = launch() + finish(); reported in streams/s, MP/s and bytes saved.  Gate: every
     distinct output decodes (oracle) to the coefficients of its source and is smaller.
(f)3 lossless SOF3: a batch of 1024x1024 three-component 8-bit frames (predictor 1, one restart interval per row) through
     JpegBatchDecoder to unclamped int16 planes in HBM.  Gate: planes equal the coded samples and the oracle's planes.
     CPU figure next to it: the oracle on all host threads, one frame per task.
One JSON line per row.  Times are wall-clock around K steps between two context synchronisations (both legs wait on the
host once per step for the decoder's verdicts)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402

W4K, H4K = 3840, 2160
WL = HL = 1024


def _gen_4k(i):
    return synth.synth_jpeg(1000 + i, W4K, H4K, quality=85, subsampling="4:2:0", restart_rows=1)


def _gen_lossless(i):
    return synth.synth_lossless(i, WL, HL, precision=8, predictor=1, restart=WL)


def pool_map(fn, n):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(max(1, min(n, os.cpu_count() or 2, 32))) as pool:
        return pool.map(fn, range(n))


def optimizer_row(J, O, ctx, args):
    blobs = pool_map(_gen_4k, args.distinct)
    pinned = []
    for b in blobs:
        a = ctx.pinned_array(len(b))
        a[:] = np.frombuffer(b, dtype=np.uint8)
        pinned.append(a)
    batch = [pinned[i % len(pinned)] for i in range(args.frames)]
    with J.JpegBatchOptimizer(batch, context=ctx, parse_threads=min(32, os.cpu_count() or 1)) as opt:
        opt.upload()
        for _ in range(3):
            opt.launch()
            opt.finish()
        saved = []
        for i in range(len(blobs)):  # gate on every distinct stream
            out = opt.stream(i)
            a, c = O.decode(blobs[i], want_rgb=False), O.decode(out, want_rgb=False)
            assert all(np.array_equal(x, y) for x, y in zip(a.coef, c.coef)), f"stream {i}: coefficients differ"
            assert len(out) < len(blobs[i])
            saved.append(1.0 - len(out) / len(blobs[i]))
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            opt.launch()
            opt.finish()
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        launches = opt.launch_count()
    # CPU figure: the oracle has no transcoder; its entropy decode + IDCT of the same frames is the nearest thing it can time
    from concurrent.futures import ThreadPoolExecutor
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda b: O.decode(b, want_rgb=False).nscans, [blobs[i % len(blobs)] for i in range(2 * threads)]))
    cpu = 2 * threads / (time.perf_counter() - t0)
    for a in pinned:
        ctx.pinned_free(a.ctypes.data)
    ctx.trim()
    comp = sum(b.size for b in batch)
    return {"row": "8(f)1 JpegOptimizer", "workload": f"{args.frames} x 3840x2160 4:2:0 SOF0 q85 DRI=240, standard -> optimised tables, inputs resident in HBM",
            "streams_per_s": args.frames / dt, "value": args.frames * W4K * H4K / 1e6 / dt, "unit": "MP/s", "ms_per_step": dt * 1e3,
            "compressed_gb_per_s": comp / dt / 1e9, "bytes_saved_fraction_mean": float(np.mean(saved)), "gpu_launches_per_step": launches,
            "parity": {"streams_checked": len(blobs), "coefficients_identical_after_transcode": True},
            "cpu_note": {"oracle_decode_only_frames_per_s": cpu, "cores": threads,
                         "meaning": "the oracle has no transcoder: entropy decode + IDCT of the same frames on all host threads, a lower bound of what the managed optimizer does per frame is its entropy decode"}}


def lossless_row(J, O, ctx, args):
    made = pool_map(_gen_lossless, args.distinct)
    blobs = [m[0] for m in made]
    batch = [np.frombuffer(blobs[i % len(blobs)], dtype=np.uint8) for i in range(args.frames)]
    with J.JpegBatchDecoder(batch, J.JB_OUT_PLANAR_I16, context=ctx, device_output=True, parse_threads=min(32, os.cpu_count() or 1)) as dec:
        dec.upload()
        for _ in range(3):
            dec.launch()
        dec.finish()
        assert dec.status() == [0] * args.frames
        for i in range(len(blobs)):
            got = dec.read_output(i)
            assert np.array_equal(got.reshape(made[i][1].shape), made[i][1]), f"frame {i}: samples differ from what was coded"
            assert np.array_equal(got.reshape(made[i][1].shape), O.decode(blobs[i], want_rgb=False).planes), f"frame {i}: oracle differs"
        dec.set_profiling(True)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dec.launch()
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        kernels = dec.profile()
        launches = dec.launch_count()
    from concurrent.futures import ThreadPoolExecutor
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda b: O.decode(b, want_rgb=False).nscans, [blobs[i % len(blobs)] for i in range(4 * threads)]))
    cpu = 4 * threads * WL * HL / 1e6 / (time.perf_counter() - t0)
    ctx.trim()
    comp = sum(b.size for b in batch)
    samples = args.frames * WL * HL * 3
    return {"row": "8(f)3 lossless SOF3", "workload": f"{args.frames} x 1024x1024 x 3 components, 8-bit, predictor 1, DRI = one row, inputs resident in HBM, int16 planes in HBM",
            "value": args.frames * WL * HL / 1e6 / dt, "unit": "MP/s", "ms_per_step": dt * 1e3, "compressed_gb_per_s": comp / dt / 1e9,
            "algorithmic_gb_per_s": (comp + 2 * samples) / dt / 1e9, "kernel_ms": {k: v for k, v in kernels}, "gpu_launches_per_step": launches,
            "parity": {"frames_checked": len(blobs), "planes_bit_exact_vs_coded_samples_and_oracle": True},
            "cpu_baseline": {"value": cpu, "unit": "MP/s", "cores": threads, "kind": "port", "sample": f"{4 * threads} frames, one per task"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--rows", default="optimizer,lossless")
    args = ap.parse_args()
    import jpeglibrary_b200 as J
    import oracle_ffi as O
    ctx = J.Context(0)
    for name, fn in (("optimizer", optimizer_row), ("lossless", lossless_row)):
        if name in args.rows.split(","):
            print(json.dumps(fn(J, O, ctx, args)), flush=True)


if __name__ == "__main__":
    main()
