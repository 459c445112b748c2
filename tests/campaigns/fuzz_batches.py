"""Damaged streams decoded as BATCHES: streams of every base type mixed in one jb_decode_batch call (good ones among them),
every image compared with the oracle -- a damaged image must fail (or decode to garbage) alone, whatever shares its
batch, its warps, its table cache and its arena neighbours.
usage (on a GPU box): python tests/campaigns/fuzz_batches.py [batches] [images per batch] [seed]"""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jpeglibrary_b200 as J
import oracle_ffi as O
import synth
import test_gpu_fuzz as F

nbatches = int(sys.argv[1]) if len(sys.argv) > 1 else 50
per = int(sys.argv[2]) if len(sys.argv) > 2 else 96
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 777
rng = np.random.default_rng(seed)
bases = dict(F.base_streams())
rgb = synth.synth_rgb(50, 176, 120)
src = synth.encode_jpeg(rgb, quality=88, subsampling="4:2:0")
bases.update({
    "422_restart_optimized": synth.encode_jpeg(rgb, quality=90, subsampling="4:2:2", restart_blocks=5, optimize=True),
    "gray_progressive": synth.encode_jpeg(rgb, quality=80, gray=True, progressive=True),
    "sequential_three_scans": synth.resequence_scans(src, O.decode(src, want_rgb=False), [[0], [1], [2]], 5),
    "progressive_420_restart": synth.encode_jpeg(rgb, quality=90, subsampling="4:2:0", progressive=True, restart_blocks=11),
    "oracle_encoded_440": O.encode_ycbcr(O.rgb_to_ycbcr(rgb), quality=88, subsampling=(1, 2)).bytes,
})
names = sorted(bases)
os.makedirs("gpurun_out", exist_ok=True)
total = bad = skipped = failed_images = 0
for bi in range(nbatches):
    blobs, wants = [], []
    while len(blobs) < per:
        name = names[int(rng.integers(len(names)))]
        t = int(rng.integers(10))
        if t == 0:
            mut = bases[name]                                   # an undamaged stream among the damaged ones
        elif t == 1:
            mut = F.mutate_header(bases[name], rng, int(rng.integers(4)))
        else:
            mut = F.mutate(bases[name], rng, F.KINDS[int(rng.integers(len(F.KINDS)))])
        want, werr = F.run_oracle(mut)
        if werr is not None and "outside the oracle's scope" in str(werr):
            continue
        try:  # what fails in the marker walk or in the planner never gets into a batch (the single-image campaigns cover it)
            p = J.Parsed(mut)
            out = (C.c_int32 * (10 * max(1, p.desc.scan_count)))()
            if J._native.cuda.jb_plan_scans(C.byref(p.desc), out, p.desc.scan_count) < 0:
                raise J.InvalidDataException("plan")
        except (J.InvalidDataException, J.InvalidOperationException, J.NotSupportedException, J.ArgumentException):
            skipped += 1
            continue
        blobs.append(mut)
        wants.append((name, want, werr))
    with J.JpegBatchDecoder(blobs, J.JB_OUT_PLANAR_I16, device_output=True) as b:
        try:
            b.run()
        except (J.InvalidDataException, J.InvalidOperationException):
            pass
        st = b.status()
        for i, (name, want, werr) in enumerate(wants):
            total += 1
            code = 0 if werr is None else werr.code
            ok = st[i] == code
            if ok and werr is None:
                got = b.read_output(i)
                wr = O.written_samples(want)
                ok = got.shape == want.planes.shape and np.array_equal(got[wr], want.planes[wr])
            failed_images += werr is not None
            if not ok:
                bad += 1
                fn = f"gpurun_out/fuzzbatch_{bi}_{i}_{name}.jpg"
                open(fn, "wb").write(blobs[i])
                print(f"batch {bi} image {i} ({name}): oracle [{werr}] GPU status {st[i]} -> {fn}", flush=True)
print(f"{nbatches} batches of {per}: {total} images ({failed_images} of them fail in the reference, {skipped} streams left to the "
      f"single-image campaigns), {bad} disagreements")
