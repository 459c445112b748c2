"""VALID streams of random geometry through the GPU path and the oracle: decode (baseline / progressive, every Pillow
sampling, restart intervals in rows or blocks, optimised tables, grey; sizes 1..400 px) and encode (random size, sampling,
quality and content; standard and package-merge tables), plus the optimizer on the decoder's inputs.  Geometry edge
cases -- images smaller than a block, one-MCU rows, restart intervals longer than the scan -- live here.
usage (on a GPU box): python tests/campaigns/fuzz_shapes.py [trials] [seed] [largest side, default 400]"""
import os
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jpeglibrary_b200 as J
import oracle_ffi as O
import synth

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4242
maxsize = int(sys.argv[3]) if len(sys.argv) > 3 else 400
rng = np.random.default_rng(seed)
os.makedirs("gpurun_out", exist_ok=True)


def content(w, h):
    k = int(rng.integers(4))
    if k == 0:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)                       # noise
    if k == 1:
        return np.full((h, w, 3), rng.integers(0, 256, 3), dtype=np.uint8)           # flat
    if k == 2:
        return synth.synth_rgb(int(rng.integers(1 << 20)), w, h)
    g = (np.add.outer(np.arange(h), np.arange(w)) * int(rng.integers(1, 9))) & 255  # gradient
    return np.stack([g, 255 - g, g // 2], axis=-1).astype(np.uint8)


def size():
    return [int(rng.integers(1, 17)), int(rng.integers(1, 65)), int(rng.integers(1, maxsize + 1))][int(rng.integers(3))]


bad = dec_n = enc_n = opt_n = ll_n = fmt_n = batch_n = 0
pool = []  # valid streams with their RGB, decoded again as one mixed batch every 48 trials


def check_batch():
    global bad, batch_n
    if not pool:
        return
    blobs = [p[0] for p in pool]
    for device_output in (True, False):
        with J.JpegBatchDecoder(blobs, J.JB_OUT_RGB24, device_output=device_output) as b:
            b.run()
            st = b.status()
            for i, (blob, rgbw) in enumerate(pool):
                batch_n += 1
                if st[i] != 0 or not np.array_equal(b.read_output(i), rgbw):
                    bad += 1
                    open(f"gpurun_out/shape_batch_{batch_n}.jpg", "wb").write(blob)
                    print(f"batch image {i} (device_output={device_output}): status {st[i]} or pixels differ from its own single decode", flush=True)
    # the host-to-host pipeline (chunks on two contexts) over the same streams
    global pipe, pinned
    if pipe is None:
        ctx = J.Context.default()
        pipe = J.JpegPipelinedBatchDecoder([ctx, J.Context(0)], chunk=int(rng.integers(1, 9)), parse_threads=2)
        pinned = ctx.pinned_array(max(64, 48 * 3 * maxsize * maxsize // (1 << 20) // 3 + 64) * 1024 * 1024)
    pinned[:] = 0
    offs = pipe.decode(blobs, pinned)
    for i, (blob, rgbw) in enumerate(pool):
        batch_n += 1
        if not np.array_equal(pinned[offs[i]:offs[i] + rgbw.size].reshape(rgbw.shape), rgbw):
            bad += 1
            open(f"gpurun_out/shape_pipe_{batch_n}.jpg", "wb").write(blob)
            print(f"pipelined image {i}: pixels differ from its own single decode", flush=True)
    pool.clear()


pipe = pinned = None
for t in range(trials):
    w, h = size(), size()
    rgb = content(w, h)
    # ---- decode
    kw = dict(quality=int(rng.integers(1, 101)))
    gray = rng.integers(6) == 0
    if gray:
        kw["gray"] = True
    else:
        kw["subsampling"] = ["4:4:4", "4:2:2", "4:2:0"][int(rng.integers(3))]
    if rng.integers(2):
        kw["progressive"] = True
    r = int(rng.integers(4))
    if r == 1:
        kw["restart_rows"] = int(rng.integers(1, 4))
    elif r == 2:
        kw["restart_blocks"] = int(rng.integers(1, 40))
    if rng.integers(3) == 0:
        kw["optimize"] = True
    try:
        blob = synth.encode_jpeg(rgb, **kw)
    except OSError:  # (Pillow refuses some combinations, e.g. its output buffer is too small for noise at quality 100)
        continue
    what = f"decode {w}x{h} {kw}"
    try:
        want = O.decode(blob)
        werr = None
    except O.OracleError as e:
        want, werr = None, e
    try:
        dec = J.JpegDecoder(); dec.SetInput(blob); dec.Identify()
        planes = np.zeros((dec.NumberOfComponents, dec.Height, dec.Width), dtype=np.int16)
        dec.SetOutputWriter(J.CudaOutputWriter(planes, J.JB_OUT_PLANAR_I16)); dec.Decode()
        out = np.zeros((dec.Height, dec.Width, 3), dtype=np.uint8)
        dec = J.JpegDecoder(); dec.SetInput(blob); dec.Identify()
        dec.SetOutputWriter(J.CudaOutputWriter(out, J.JB_OUT_RGB24)); dec.Decode()
        gerr = None
    except (J.InvalidDataException, J.InvalidOperationException, J.NotSupportedException) as e:
        gerr = e
    dec_n += 1
    ok = (werr is None) == (gerr is None)
    if ok and werr is None:
        wr = O.written_samples(want)
        ok = np.array_equal(planes[wr], want.planes[wr]) and int(np.abs(out.astype(int) - want.rgb.astype(int))[wr.all(axis=0)].max(initial=0)) <= 1
    if not ok:
        bad += 1
        open(f"gpurun_out/shape_{t}.jpg", "wb").write(blob)
        print(f"trial {t}: {what}: oracle [{werr}] GPU [{gerr}]", flush=True)
    if werr is None and gerr is None:
        pool.append((blob, out.copy()))
        # ---- the other sinks: RGBA32, YCbCr888, a padded pitch, a device destination
        try:
            dec = J.JpegDecoder(); dec.SetInput(blob); dec.Identify()
            H, W = dec.Height, dec.Width
            rgba = np.zeros((H, W, 4), np.uint8)
            dec.SetOutputWriter(J.CudaOutputWriter(rgba, J.JB_OUT_RGBA32)); dec.Decode()
            dec = J.JpegDecoder(); dec.SetInput(blob); dec.Identify()
            ycc = np.zeros((H, W, 3), np.uint8)
            dec.SetOutputWriter(J.CudaOutputWriter(ycc, J.JB_OUT_YCBCR888)); dec.Decode()
            pitch = (3 * W + int(rng.integers(1, 40)) + 3) // 4 * 4
            padded = np.full((H, pitch), 0xA5, np.uint8)
            dec = J.JpegDecoder(); dec.SetInput(blob); dec.Identify()
            dec.SetOutputWriter(J.CudaOutputWriter(padded, J.JB_OUT_RGB24, pitch=pitch)); dec.Decode()
            same = (np.array_equal(rgba[..., :3], out) and (rgba[..., 3] == 255).all() and np.array_equal(ycc, want.ycbcr)
                    and np.array_equal(padded[:, :3 * W].reshape(H, W, 3), out) and (padded[:, 3 * W:] == 0xA5).all())
        except Exception as e:  # noqa: BLE001
            same = False
            print(f"trial {t}: sinks {what}: {type(e).__name__}: {e}", flush=True)
        fmt_n += 1
        if not same:
            bad += 1
            open(f"gpurun_out/shape_fmt_{t}.jpg", "wb").write(blob)
            print(f"trial {t}: sinks {what}: RGBA32 / YCbCr888 / padded pitch differ", flush=True)
    if len(pool) >= 48:
        check_batch()
    # ---- lossless (SOF3) of random geometry
    if t % 3 == 0:
        ncomp = [1, 3, 3][int(rng.integers(3))]
        samp = [(1, 1)] * ncomp
        if ncomp == 3 and rng.integers(2):
            samp = [[(2, 2), (1, 1), (1, 1)], [(2, 1), (1, 1), (1, 1)], [(1, 2), (1, 1), (1, 1)]][int(rng.integers(3))]
        hm, vm = max(a for a, _ in samp), max(b for _, b in samp)
        lw, lh = max(hm, size() // 4 // hm * hm), max(vm, size() // 4 // vm * vm)
        prec = int(rng.integers(2, 17))
        lkw = dict(precision=prec, predictor=int(rng.integers(1, 8)), sampling=samp, ncomp=ncomp,
                   point_transform=int(rng.integers(0, min(prec - 8, 4) + 1)) if prec > 8 and rng.integers(3) == 0 else 0,
                   restart=int(rng.integers(1, 60)) if rng.integers(2) else 0)
        try:
            lblob, coded = synth.synth_lossless(int(rng.integers(1 << 16)), lw, lh, **lkw)
            dec = J.JpegDecoder(); dec.SetInput(lblob); dec.Identify()
            lp = np.zeros((dec.NumberOfComponents, dec.Height, dec.Width), dtype=np.int16)
            dec.SetOutputWriter(J.CudaOutputWriter(lp, J.JB_OUT_PLANAR_I16)); dec.Decode()
            same = np.array_equal(lp, coded) and np.array_equal(lp, O.decode(lblob, want_rgb=False).planes)
        except Exception as e:  # noqa: BLE001
            same = False
            print(f"trial {t}: lossless {lw}x{lh} {lkw}: {type(e).__name__}: {e}", flush=True)
        ll_n += 1
        if not same:
            bad += 1
            print(f"trial {t}: lossless {lw}x{lh} {lkw}: planes differ", flush=True)
    # ---- optimizer on the same stream (sequential single-scan frames only)
    if werr is None and not kw.get("progressive"):
        try:
            opt = J.JpegOptimizer(); opt.MostOptimalCoding = bool(rng.integers(2)); opt.SetInput(blob); opt.Scan()
            o2 = bytearray(); opt.SetOutput(o2); opt.Optimize(bool(rng.integers(2)))
            back = O.decode(bytes(o2), want_rgb=False)
            same = all(np.array_equal(a, b) for a, b in zip(want.coef, back.coef))
        except Exception as e:  # noqa: BLE001
            same = False
            print(f"trial {t}: optimize {what}: {type(e).__name__}: {e}", flush=True)
        opt_n += 1
        if not same:
            bad += 1
            open(f"gpurun_out/shape_opt_{t}.jpg", "wb").write(blob)
            print(f"trial {t}: optimize {what}: coefficients differ", flush=True)
    # ---- encode
    ss = [(1, 1), (2, 1), (1, 2), (2, 2)][int(rng.integers(4))]
    q = int(rng.integers(1, 101))
    mo = bool(rng.integers(3) == 0)
    what = f"encode {w}x{h} q{q} {ss} most_optimal={mo}"
    try:
        wantb = O.encode_ycbcr(O.rgb_to_ycbcr(rgb), quality=q, subsampling=ss, optimal=mo).bytes
        got, _ = J.encode_rgb(rgb, quality=q, subsampling=ss, most_optimal=mo)
        same = got == wantb
    except Exception as e:  # noqa: BLE001
        same = False
        print(f"trial {t}: {what}: {type(e).__name__}: {e}", flush=True)
    enc_n += 1
    if not same:
        bad += 1
        np.save(f"gpurun_out/shape_enc_{t}.npy", rgb)
        print(f"trial {t}: {what}: streams differ", flush=True)
check_batch()
print(f"{dec_n} decodes, {fmt_n} x 3 other sinks, {batch_n} batch images, {ll_n} lossless frames, {opt_n} optimizer runs, {enc_n} encodes: "
      f"{bad} disagreements")
