"""A longer run of the corrupted-stream parity tests (tests/test_gpu_fuzz.py) with other seeds: every disagreement
between the GPU path and the oracle is printed and the stream is kept under gpurun_out/.
usage (on a GPU box): python tests/campaigns/fuzz_campaign.py [trials per stream] [seed] [--more]"""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gpu_fuzz as F

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
os.makedirs("gpurun_out", exist_ok=True)
total = bad = known = 0
import synth, oracle_ffi as O
from conftest import golden_bytes
bases = dict(F.base_streams())
if "--more" in sys.argv:   # further frame types than the test suite mutates
    rgb = synth.synth_rgb(50, 176, 120)
    src = synth.encode_jpeg(rgb, quality=88, subsampling="4:2:0")
    bases = {
        "422_restart_optimized": synth.encode_jpeg(rgb, quality=90, subsampling="4:2:2", restart_blocks=5, optimize=True),
        "gray_plain": synth.encode_jpeg(rgb, quality=80, gray=True),
        "gray_progressive": synth.encode_jpeg(rgb, quality=80, gray=True, progressive=True),
        "sequential_three_scans": synth.resequence_scans(src, O.decode(src, want_rgb=False), [[0], [1], [2]], 4),
        "sequential_two_scans": synth.resequence_scans(src, O.decode(src, want_rgb=False), [[0, 1], [2]]),
        "testorig12": golden_bytes("testorig12.jpg"),
        "lossless_16bit_s22": synth.synth_lossless(51, 64, 48, precision=16, predictor=6, sampling=[(2, 2), (1, 1), (1, 1)], restart=8)[0],
        "lossless_gray_5bit": synth.synth_lossless(52, 80, 40, precision=5, predictor=7, ncomp=1)[0],
        "progressive_422": synth.encode_jpeg(rgb, quality=75, subsampling="4:2:2", progressive=True),
    }
if "--wide" in sys.argv:   # layouts and scan scripts neither of the other two sets holds
    rgb = synth.synth_rgb(53, 200, 136)
    src = synth.encode_jpeg(rgb, quality=85, subsampling="4:2:0")
    prog = synth.encode_jpeg(rgb, quality=85, subsampling="4:4:4", progressive=True)
    bases = {
        "oracle_encoded_440": O.encode_ycbcr(O.rgb_to_ycbcr(rgb), quality=88, subsampling=(1, 2)).bytes,
        "oracle_encoded_h4": O.encode_ycbcr(O.rgb_to_ycbcr(rgb), quality=80, subsampling=(4, 1)).bytes,
        "oracle_encoded_gray": O.encode_ycbcr(O.rgb_to_ycbcr(rgb), quality=70, gray=True).bytes,
        "progressive_420_restart": synth.encode_jpeg(rgb, quality=90, subsampling="4:2:0", progressive=True, restart_blocks=11),
        "progressive_444_luma_first": synth.reorder_progressive_scans(prog, [0, 1, 4, 5, 9, 2, 3, 7, 8, 6]),
        "progressive_444_scan_behind_its_refinement": synth.reorder_progressive_scans(prog, [0, 1, 4, 5, 9, 1, 2, 3, 7, 8, 6]),
        "sequential_three_scans_no_restart": synth.resequence_scans(src, O.decode(src, want_rgb=False), [[0], [1], [2]]),
        "sequential_luma_then_chroma_pair": synth.resequence_scans(src, O.decode(src, want_rgb=False), [[0], [1, 2]], 7),
        "no_restart_420_large": synth.encode_jpeg(synth.synth_rgb(54, 1280, 720), quality=85, subsampling="4:2:0"),
    }
if "--cmyk" in sys.argv:   # four components (planar int16 output only: the reference's app sink takes 1 or 3)
    import io
    from PIL import Image
    rgb = synth.synth_rgb(60, 120, 88)
    cmyk = np.concatenate([rgb, rgb[..., :1][..., ::-1]], axis=-1).astype(np.uint8)

    def enc(**kw):
        buf = io.BytesIO()
        Image.fromarray(cmyk, "CMYK").save(buf, format="JPEG", **kw)
        return buf.getvalue()
    bases = {
        "cmyk_444": enc(quality=85),
        "cmyk_444_restart": enc(quality=90, restart_marker_blocks=6),
        "cmyk_420": enc(quality=75, subsampling="4:2:0"),
        "cmyk_420_restart_optimized": enc(quality=80, subsampling="4:2:0", restart_marker_rows=1, optimize=True),
        "cmyk_progressive": enc(quality=80, progressive=True),   # (the reference refuses its scan script: quirk P6)
    }
for name, blob in bases.items():
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
        elif (werr is None and gerr is None and name.startswith("sequential_") and got.shape == want.planes.shape
              and not O.written_samples(want).all() and "--known-deviation" in sys.argv):
            known += 1  # (before the per-component MCU limits of round 2: EOI at a restart boundary of a multi-scan sequential frame)
        else:
            bad += 1
            fn = f"gpurun_out/fuzz_{name}_{t}.jpg"
            open(fn, "wb").write(mut)
            print(f"{name} trial {t} ({'header' if header else kind}): oracle [{werr}] GPU [{type(gerr).__name__ if gerr else 'decoded'}: {gerr}] -> {fn}", flush=True)
    print(f"{name}: {ok} identical, {err} errors on both sides", flush=True)
print(f"{total} streams, {bad} disagreements, {known} of the known deviation (multi-scan sequential frame, EOI at a restart boundary)")
