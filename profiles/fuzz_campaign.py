"""A longer run of the corrupted-stream parity tests (tests/test_gpu_fuzz.py) with other seeds: every disagreement
between the GPU path and the oracle is printed and the stream is kept under gpurun_out/.
usage (on a GPU box): python profiles/fuzz_campaign.py [trials per stream] [seed]"""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gpu_fuzz as F

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
os.makedirs("gpurun_out", exist_ok=True)
total = bad = 0
for name, blob in F.base_streams().items():
    rng = np.random.default_rng(seed + sum(map(ord, name)))
    ok = err = 0
    for t in range(trials):
        header = t % 5 == 4
        kind = t % 4 if header else F.KINDS[t % len(F.KINDS)]
        mut = F.mutate_header(blob, rng, kind) if header else F.mutate(blob, rng, kind)
        want, werr = F.run_oracle(mut)
        if werr is not None and "outside the oracle's scope" in str(werr):
            continue
        try:
            got, gerr = F.run_gpu(mut)
        except Exception as e:  # noqa: BLE001
            got, gerr = None, e
        total += 1
        if werr is not None and gerr is not None:
            err += 1
        elif werr is None and gerr is None and got.shape == want.planes.shape and np.array_equal(got, want.planes):
            ok += 1
        else:
            bad += 1
            fn = f"gpurun_out/fuzz_{name}_{t}.jpg"
            open(fn, "wb").write(mut)
            print(f"{name} trial {t} ({'header' if header else kind}): oracle [{werr}] GPU [{type(gerr).__name__ if gerr else 'decoded'}: {gerr}] -> {fn}", flush=True)
    print(f"{name}: {ok} identical, {err} errors on both sides", flush=True)
print(f"{total} streams, {bad} disagreements")
