#!/usr/bin/env python
"""bench.py -- decoded MP/s of the batched 4K 4:2:0 SOF0 decode hot path on N B200s (BASELINE.json configs[1]),
with every other config of BASELINE.json reported in the same JSON line.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.

  value        configs[1]: a step = restart scan + Huffman decode + IDCT/colour (K0..K2) over one batch of synthetic
               JPEGs that is already resident in HBM; RGB24 stays in HBM.  CUDA events on the library's stream.
  value_h2d    the same with the blueprint's timing boundary: compressed bytes start in PINNED HOST
               memory, measured interval = cudaMemcpyAsync H2D + K0..K2, two half batches double-buffered on two
               streams so that one half's copy overlaps the other half's kernels.
  e2e          the same metric through the public API with host buffers on both sides (marker walk + plan + H2D +
               kernels + D2H of the RGB), with a per-phase breakdown and the bare concurrent pinned D2H rate of the
               same ranks measured beside it (what the box's host side can take).
  workloads    (N = 1 only) configs[2] no restart markers, configs[3] progressive 1080p 4:4:4, configs[4] baseline
               encode with optimised tables, each with value, ms/step, roofline, parity gate and a CPU baseline of
               >= 10 s; configs[0], the reference's own CPU case (lake.jpg / HETissueSlide.jpg), as CPU latency.
  --impl reference   the reference algorithm's CPU restatement (oracle/; the reference's C# cannot run here: no .NET)
               on all host threads, same metric and config.
`--workload restart|norestart|progressive|encode` runs one workload alone as the headline (profiling under ncu).
"""
import argparse
import json
import os
import pickle
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

W4K, H4K = 3840, 2160
WP, HP = 1920, 1080
WORKLOAD_NAMES = {
    "restart": "configs[1]: batch of synthetic 3840x2160 4:2:0 SOF0 JPEGs, q85, DRI=240 (one MCU row)",
    "norestart": "configs[2]: the same 3840x2160 4:2:0 SOF0 q85 frames coded without restart markers (self-synchronising decode)",
    "progressive": "configs[3]: batch of synthetic 1920x1080 4:4:4 progressive SOF2 JPEGs, q85 (libjpeg's 10-scan script)",
    "encode": "configs[4]: baseline encode of synthetic 3840x2160 RGB frames, q75 4:2:0, optimised Huffman tables",
}


# ------------------------------------------------------------------------------------------------ inputs
def _gen_4k(job):
    import synth
    i, want_plain, want_rgb = job
    rgb = synth.synth_rgb(i, W4K, H4K)
    a = synth.encode_jpeg(rgb, quality=85, subsampling="4:2:0", restart_rows=1)
    b = synth.encode_jpeg(rgb, quality=85, subsampling="4:2:0") if want_plain else None
    return a, b, (rgb if want_rgb else None)


def _gen_prog(i):
    import synth
    return synth.encode_jpeg(synth.synth_rgb(i, WP, HP), quality=85, subsampling="4:4:4", progressive=True)


def make_inputs(distinct, frames, extra):
    """Synthetic inputs (seeds 1000+i), generated on a process pool (data generation only)."""
    import multiprocessing as mp
    procs = max(1, min(distinct, os.cpu_count() or 2, 32))
    jobs = [(i, extra, extra and i < frames) for i in range(distinct)]
    with mp.get_context("fork").Pool(procs) as pool:
        r = pool.map(_gen_4k, jobs)
        prog = pool.map(_gen_prog, range(distinct)) if extra else []
    return {"restart": [x[0] for x in r], "norestart": [x[1] for x in r] if extra else [],
            "frames": [x[2] for x in r if x[2] is not None], "progressive": prog}


def shared_inputs(args, rank, world, barrier, extra):
    """rank 0 generates, the other ranks of the node read the pickle from /dev/shm (N ranks x 128 4K frames would
    otherwise be generated N times on the same host cores)."""
    if world == 1:
        return make_inputs(args.distinct, 8, extra)
    path = f"/dev/shm/jb_bench_{os.environ.get('MASTER_PORT', '0')}_{args.distinct}.pkl"
    if rank == 0:
        inp = make_inputs(args.distinct, 8, extra)
        with open(path + ".tmp", "wb") as f:
            pickle.dump(inp, f)
        os.replace(path + ".tmp", path)
    barrier()
    if rank != 0:
        with open(path, "rb") as f:
            inp = pickle.load(f)
    barrier()
    if rank == 0:
        os.unlink(path)
    return inp


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, images):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json, written by
    profiles/make_traffic.py), scaled from the profiled batch to `images`; None when no capture covers it."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        doc = json.load(open(p))
        return doc["kernels"][kernel]["dram_bytes_per_image"] * images, doc.get("source")
    except Exception:
        return None, None


def oracle_rgb(blobs, width, height, threads):
    """RGB24 of every blob from the oracle (parity gate), decoded on `threads` host threads."""
    import ctypes as C
    import oracle_ffi as O
    n = len(blobs)
    arrs = [np.frombuffer(b, dtype=np.uint8) for b in blobs]
    outs = [np.empty((height, width, 3), np.uint8) for _ in range(n)]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = (C.c_size_t * n)(*[a.size for a in arrs])
    optr = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    failed = O.lib().jo_decode_batch_rgb(ptrs, lens, n, threads, optr)
    assert failed == 0, f"the oracle failed on {failed} bench inputs"
    return outs


def cpu_decode_baseline(blobs, mp_per_image, threads, target_s, what):
    """The oracle on `threads` host threads for about target_s seconds (calibrated on one image per thread)."""
    import oracle_ffi as O

    def run(images):
        sample = [blobs[i % len(blobs)] for i in range(images)]
        t0 = time.perf_counter()
        failed = O.decode_batch_rgb(sample, threads)
        dt = time.perf_counter() - t0
        assert failed == 0
        return dt

    run(min(2, threads))  # warm the library / page cache
    per_round = max(run(threads), 1e-3)
    images = int(max(threads, min(8192, threads * max(1, round(target_s / per_round)))))
    dt = run(images)
    return {"value": images * mp_per_image / dt, "unit": "MP/s", "cores": threads, "kind": "port",
            "sample": f"{images} {what}, one image per task on {threads} threads ({dt:.1f} s wall), "
                      "C restatement of the reference algorithm (reference .NET runtime unavailable)"}


def bare_d2h_rate(ctx, torch, dist, world, local_rank, pinned, seconds=1.0):
    """Aggregate GB/s of plain cudaMemcpyAsync device -> pinned host on all ranks at once: what the box's host side
    (PCIe root ports, IOMMU, host memory) can take, with none of this library on the path."""
    n = min(pinned.size, 1 << 30)
    dev = ctx.device_alloc(n)
    host = pinned[:n]
    ctx.d2h(host, dev)  # (jb_memcpy_d2h: one cudaMemcpyAsync of 1 GiB on the context's stream + a stream synchronise)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < seconds:
        ctx.d2h(host, dev)
        reps += 1
    dt = time.perf_counter() - t0
    rate = torch.tensor([n * reps / dt / 1e9], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(rate, op=dist.ReduceOp.SUM)
    ctx.device_free(dev)
    return float(rate.item())


# ------------------------------------------------------------------------------------------------ decode workloads
def decode_workload(kind, blobs, args, env, steps, warmup, full):
    """One decode workload: resident-input leg (+ parity gate over every distinct image, per-kernel times, roofline);
    with full=True also the H2D-inclusive leg and the host-to-host leg.  Returns a dict."""
    torch, dist, J = env["torch"], env["dist"], env["J"]
    ctx, world, rank, local_rank = env["ctx"], env["world"], env["rank"], env["local_rank"]
    width, height = (WP, HP) if kind == "progressive" else (W4K, H4K)
    mp_img = width * height / 1e6
    threads = os.cpu_count() or 1
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    def barrier():
        if world > 1:
            dist.barrier()

    pinned = []
    for b in blobs:
        a = ctx.pinned_array(len(b))
        a[:] = np.frombuffer(b, dtype=np.uint8)
        pinned.append(a)
    batch = args.batch
    batch_blobs = [pinned[i % len(pinned)] for i in range(batch)]
    comp_bytes = sum(b.size for b in batch_blobs)

    # ---- resident-input leg
    dec = J.JpegBatchDecoder(batch_blobs, J.JB_OUT_RGB24, context=ctx, device_output=True, parse_threads=min(32, threads))
    dec.upload()
    ctx.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()  # sampled from the warm-up on: the timed region itself lasts only a few hundred ms
    for _ in range(max(warmup, 1)):
        dec.launch()
    ctx.synchronize()
    # parity gate on the bench inputs themselves: EVERY distinct image against the oracle (rank 0)
    parity = None
    if rank == 0:
        want = oracle_rgb(blobs, width, height, threads)
        hist = np.zeros(3, dtype=np.int64)
        worst = 0
        for i in range(len(blobs)):
            diff = np.abs(dec.read_output(i).astype(np.int16) - want[i].astype(np.int16))
            worst = max(worst, int(diff.max()))
            hist += np.bincount(np.minimum(diff, 2).ravel(), minlength=3)[:3]
        del want
        assert worst <= 1, f"bench output differs from the oracle by {worst}"
        parity = {"images_checked": len(blobs), "max_abs_rgb_diff_vs_oracle": worst,
                  "abs_diff_histogram_0_1_2plus": hist.tolist()}
    dec.finish()
    assert dec.status() == [0] * batch

    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dec.set_profiling(True)
    ev0.record(stream)
    for _ in range(steps):
        dec.launch()
    ev1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    kernels = dec.profile()  # [(name, avg ms per launch)] measured with CUDA events on the launch stream
    dec.set_profiling(False)
    launches = dec.launch_count() * steps
    dec.close()

    from jpeglibrary_b200.sharding import max_over_ranks
    ms_max = max_over_ranks(ms, device=f"cuda:{local_rank}")  # whole-job time = slowest rank's device time
    value = batch * world * steps * mp_img / (ms_max / 1e3)

    # ---- roofline of the dominant kernel (ALGORITHMIC bytes per launch)
    peak, peak_src = measured_peak()
    nblocks = (width // 8) * (-(-height // 8)) * 3 if kind == "progressive" else (width // 16) * (height // 16) * 6
    store = 128 * nblocks * batch
    alg = {"jb_k0_restart_scan": comp_bytes, "jb_k0b_segment_descs": 0,
           "jb_k1_huff_segments": comp_bytes + store, "jb_k1b_selfsync_chain": comp_bytes + store,
           "jb_k1c_progressive_scans": comp_bytes + store, "jb_k2_idct_color": store + 3 * width * height * batch}
    dom = max(kernels, key=lambda kv: kv[1]) if kernels else ("none", float("nan"))
    achieved = alg.get(dom[0], 0) / (dom[1] / 1e3) / 1e9 if kernels else float("nan")
    traffic, traffic_src = measured_traffic(dom[0], batch) if kind == "restart" else (None, None)
    b_alg = comp_bytes + 3 * width * height * batch  # B_alg = C + 3WH per image
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg.get(dom[0], 0),
                "kernel_ms": {k: v for k, v in kernels},
                "kernel_gbs": {k: alg[k] / (v / 1e3) / 1e9 for k, v in kernels if k in alg and v > 0},
                "pipeline_algorithmic_bytes_per_step": b_alg,
                "pipeline_frac_of_peak": (b_alg / (ms / steps / 1e3) / 1e9) / peak,
                "pipeline_frac_of_nominal_8TBs": (b_alg / (ms / steps / 1e3) / 1e9) / 8000.0}
    out = {"workload": WORKLOAD_NAMES[kind], "value": value, "unit": "MP/s", "ms_per_step": ms_max / steps, "steps": steps,
           "images_per_gpu_per_step": batch, "distinct_images": len(blobs), "compressed_bytes_per_step": comp_bytes,
           "compressed_gb_per_s": comp_bytes * world * steps / (ms_max / 1e3) / 1e9,
           "output": "RGB24 device-resident", "parity": parity, "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
           "l2": f"inputs larger than L2: {comp_bytes / 1e9:.2f} GB compressed + {store / 1e9:.1f} GB coefficient store + "
                 f"{3 * width * height * batch / 1e9:.1f} GB RGB per step (126 MB L2)"}

    if full:
        # ---- H2D-inclusive leg (SURVEY 8d boundary): pinned host -> cudaMemcpyAsync -> K0..K2, two half batches
        # double-buffered on two streams (contexts)
        ctx2 = env.get("ctx2") or J.Context(local_rank)
        env["ctx2"] = ctx2
        half = batch // 2
        halves = [J.JpegBatchDecoder(batch_blobs[:half], J.JB_OUT_RGB24, context=ctx, device_output=True, parse_threads=min(32, threads)),
                  J.JpegBatchDecoder(batch_blobs[half:2 * half], J.JB_OUT_RGB24, context=ctx2, device_output=True, parse_threads=min(32, threads))]
        streams = [torch.cuda.ExternalStream(c.stream, device=local_rank) for c in (ctx, ctx2)]

        def h2d_step():
            for d in halves:
                d.upload()
                d.launch()

        for _ in range(max(1, min(warmup, 3))):
            h2d_step()
        ctx.synchronize(); ctx2.synchronize()
        barrier()
        torch.cuda.synchronize()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in streams]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
        for e, s in zip(e0, streams):
            e.record(s)
        for _ in range(steps):
            h2d_step()
        for e, s in zip(e1, streams):
            e.record(s)
        ctx.synchronize(); ctx2.synchronize()
        torch.cuda.synchronize()
        # both streams start together (the records above are back to back): the step ends when the later one does
        h2d_ms = max(e0[0].elapsed_time(e1[0]), e0[0].elapsed_time(e1[1]), e0[1].elapsed_time(e1[1]))
        # the copies alone (no kernels), same call pattern: what the H2D side of the leg costs by itself
        for d in halves:
            d.upload()
        ctx.synchronize(); ctx2.synchronize()
        for e, s in zip(e0, streams):
            e.record(s)
        for _ in range(steps):
            for d in halves:
                d.upload()
        for e, s in zip(e1, streams):
            e.record(s)
        ctx.synchronize(); ctx2.synchronize()
        copy_ms = max(e0[0].elapsed_time(e1[0]), e0[0].elapsed_time(e1[1]), e0[1].elapsed_time(e1[1]))
        for d in halves:
            d.launch()
            d.finish()
            assert d.status() == [0] * half
            d.close()
        h2d_ms_max = max_over_ranks(h2d_ms, device=f"cuda:{local_rank}")
        h2d_bytes = sum(b.size for b in batch_blobs[:2 * half])
        out["value_h2d"] = {"value": 2 * half * world * steps * mp_img / (h2d_ms_max / 1e3), "unit": "MP/s",
                            "ms_per_step": h2d_ms_max / steps, "h2d_bytes_per_step": h2d_bytes,
                            "h2d_gb_per_s_per_gpu": h2d_bytes * steps / (h2d_ms / 1e3) / 1e9,
                            "copies_alone_ms_per_step": copy_ms / steps,
                            "copies_alone_gb_per_s": h2d_bytes * steps / (copy_ms / 1e3) / 1e9,
                            "copies_per_step": 2 * half,
                            "boundary": "compressed bytes in pinned host memory -> cudaMemcpyAsync H2D + K0..K2 -> RGB24 in HBM "
                                        "(SURVEY.md 8d); two half batches double-buffered on two streams"}

        # ---- host-to-host leg through the public API
        eb = min(args.e2e_batch, batch)
        chunk = args.e2e_chunk
        if kind == "progressive":  # the progressive entropy kernel is serial-latency bound: only large chunks amortise it
            eb = min(max(eb, 512), batch)
            chunk = max(chunk, (eb + 1) // 2)
        # every rank pins its own RGB destination (24.9 MB per 4K image, 6.4 GB for 256): if the host refuses, halve it
        while True:
            try:
                host_out = ctx.pinned_array(eb * ((width * height * 3 + 255) // 256 * 256))
                break
            except Exception:  # noqa: BLE001 - cudaHostAlloc failure surfaces as the package's exception types
                if eb <= 32:
                    raise
                eb //= 2
        e2e_blobs = batch_blobs[:eb]
        # two contexts (= two CUDA streams): one chunk's marker walk + H2D + kernels overlap the other chunk's D2H of
        # RGB, which is what bounds a host-to-host decode.  The marker walk of a 4K frame takes 0.2 ms on one core, so a
        # rank needs few walk threads, and the N ranks of a node share its cores.
        walk_threads = max(1, min(8, threads // (2 * max(1, world))))
        pipe = J.JpegPipelinedBatchDecoder([ctx, ctx2], chunk=chunk, parse_threads=walk_threads)
        for _ in range(max(1, min(warmup, 3))):
            pipe.decode(e2e_blobs, host_out, J.JB_OUT_RGB24)
        barrier()
        torch.cuda.synchronize()
        pipe.reset_stats()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(steps, 5))
        for _ in range(e2e_steps):
            pipe.decode(e2e_blobs, host_out, J.JB_OUT_RGB24)
        torch.cuda.synchronize()
        e2e_dt = time.perf_counter() - t0
        t = torch.tensor([e2e_dt], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_val = eb * world * e2e_steps * mp_img / float(t.item())
        d2h = eb * width * height * 3
        phases = pipe.stats_summary(e2e_dt)
        ceiling = bare_d2h_rate(ctx, torch, dist, world, local_rank, host_out)
        out["e2e"] = {"value": e2e_val, "unit": "MP/s", "h2d_bytes_per_step": sum(b.size for b in e2e_blobs),
                      "d2h_bytes_per_step": d2h, "images_per_step": eb, "steps": e2e_steps,
                      "includes": "marker walk + plan + H2D + kernels + D2H of RGB24 to pinned host, chunks of "
                                  f"{chunk} images pipelined on 2 streams per rank, {walk_threads} marker-walk threads per rank",
                      "d2h_gb_per_s_all_ranks": d2h * world * e2e_steps / float(t.item()) / 1e9,
                      "bare_pinned_d2h_gb_per_s_all_ranks": ceiling,
                      "fraction_of_bare_d2h": (d2h * world * e2e_steps / float(t.item()) / 1e9) / ceiling if ceiling else None,
                      "phases_rank0": phases,
                      "host": {"cpus": threads, "ranks": world}}
        ctx.pinned_free(host_out.ctypes.data)

    if full and rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_decode_baseline(blobs, mp_img, threads, args.cpu_seconds, "of the batch's images")
    elif rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_decode_baseline(blobs, mp_img, threads, max(10.0, args.cpu_seconds * 0.6), "of the batch's images")
    for a in pinned:
        ctx.pinned_free(a.ctypes.data)
    ctx.trim()
    return out


# ------------------------------------------------------------------------------------------------ encode workload
def encode_workload(frames, args, env, steps, warmup):
    """configs[4]: a step = RGB frames resident in HBM -> scan bytes resident in HBM (K3 + histogram + on-device table
    build + bit lengths / scan / pack / stuff)."""
    torch, dist, J = env["torch"], env["dist"], env["J"]
    ctx, world, rank, local_rank = env["ctx"], env["world"], env["rank"], env["local_rank"]
    import oracle_ffi as O
    from concurrent.futures import ThreadPoolExecutor
    mp_img = W4K * H4K / 1e6
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    batch = min(args.batch, 512)  # 512 frames: 12.7 GB RGB + 12.7 GB coefficients + the stream buffers
    fbytes = W4K * H4K * 3
    dev = ctx.device_alloc(batch * fbytes)
    pinned = []
    for f in frames:
        a = ctx.pinned_array(fbytes)
        a[:] = f.reshape(-1)
        pinned.append(a)
    for i in range(batch):
        ctx.h2d(dev + i * fbytes, pinned[i % len(pinned)])
    enc = J.JpegBatchEncoder([(dev + i * fbytes, W4K, H4K) for i in range(batch)], quality=75, context=ctx)
    for _ in range(max(warmup, 1)):
        enc.launch()
    enc.finish()
    # parity gate: every distinct frame's stream equals the oracle's byte for byte ("parity unpinned": the reference
    # holds no encoder vector, the oracle's encoder is anchored on its source and on round trips through the pinned decoder)
    parity = None
    threads = os.cpu_count() or 1
    ycc = None
    if rank == 0:
        ycc = [O.rgb_to_ycbcr(f) for f in frames]
        with ThreadPoolExecutor(min(threads, len(frames))) as ex:
            want = list(ex.map(lambda y: O.encode_ycbcr(y, quality=75).bytes, ycc))
        for i in range(len(frames)):
            assert enc.stream(i) == want[i], f"GPU stream of frame {i} differs from the oracle"
        parity = {"images_checked": len(frames), "stream_equals_oracle": True,
                  "note": "encoder parity is oracle-only (unpinned): the reference ships no encoder vectors"}
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        enc.launch()
    ev1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    enc.finish()
    from jpeglibrary_b200.sharding import max_over_ranks
    ms_max = max_over_ranks(ms, device=f"cuda:{local_rank}")
    value = batch * world * steps * mp_img / (ms_max / 1e3)
    out_bytes = sum(enc.scan_length(i) for i in range(batch))
    launches = enc.launch_count()
    enc.close()
    # host-to-host: host RGB -> JPEG bytes on the host, through the public API
    eb = min(args.e2e_batch, 32)
    host_out = ctx.pinned_array(eb * 8 * 1024 * 1024)

    def e2e_step():
        with J.JpegBatchEncoder([pinned[i % len(pinned)].reshape(H4K, W4K, 3) for i in range(eb)], quality=75, context=ctx) as e2:
            e2.launch()
            e2.finish()
            e2.read_all_scans(host_out)

    e2e_step()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    e2e_dt = time.perf_counter() - t0
    peak, peak_src = measured_peak()
    alg_bytes = batch * fbytes + out_bytes  # B_alg = 3WH + C_out
    achieved = alg_bytes / (ms / steps / 1e3) / 1e9
    out = {"workload": WORKLOAD_NAMES["encode"], "metric": "encoded_megapixels_per_second", "value": value, "unit": "MP/s",
           "ms_per_step": ms_max / steps, "steps": steps, "images_per_gpu_per_step": batch, "distinct_images": len(frames),
           "us_per_frame": ms_max / steps / batch * 1e3, "scan_bytes_per_step": out_bytes, "parity": parity,
           "gpu_launches": launches * steps, "clocks": clocks,
           "l2": f"inputs larger than L2 ({batch * fbytes / 1e9:.1f} GB RGB per step)",
           "e2e": {"value": eb * world * e2e_steps * mp_img / e2e_dt, "unit": "MP/s", "h2d_bytes_per_step": eb * fbytes,
                   "d2h_bytes_per_step": out_bytes // batch * eb, "images_per_step": eb,
                   "includes": "plan + H2D of RGB + kernels + D2H of scan bytes"},
           "roofline": {"bound": "hbm", "kernel": "encode pipeline (K3..K4d)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": alg_bytes}}
    if rank == 0 and world == 1:
        def run(n):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(lambda i: len(O.encode_ycbcr(ycc[i % len(ycc)], quality=75).bytes), range(n)))
            return time.perf_counter() - t0
        per_round = max(run(threads), 1e-3)
        n = int(max(threads, min(4096, threads * max(1, round(max(10.0, args.cpu_seconds * 0.6) / per_round)))))
        dt = run(n)
        out["cpu_baseline"] = {"value": n * mp_img / dt, "unit": "MP/s", "cores": threads, "kind": "port",
                               "sample": f"{n} frames (YCbCr input) on {threads} threads ({dt:.1f} s wall), C restatement of the "
                                         "reference encoder (reference .NET runtime unavailable)"}
    ctx.pinned_free(host_out.ctypes.data)
    for a in pinned:
        ctx.pinned_free(a.ctypes.data)
    ctx.device_free(dev)
    ctx.trim()
    return out


def cpu_only_config0():
    """configs[0]: the reference's own CPU-runnable case (apps/JpegDecode on a tests/Assets baseline JPEG): single-image
    latency of the reference algorithm's restatement on one host core.  CPU only by definition."""
    import oracle_ffi as O
    res = {}
    for name in ("lake.jpg", "HETissueSlide.jpg"):
        p = os.path.join(ROOT, "tests", "golden", name)
        if not os.path.exists(p):
            continue
        blob = open(p, "rb").read()
        d = O.decode(blob)
        ts = []
        t_end = time.perf_counter() + 1.5
        while len(ts) < 3 or (time.perf_counter() < t_end and len(ts) < 50):
            t0 = time.perf_counter()
            O.decode_batch_rgb([blob], 1)
            ts.append(time.perf_counter() - t0)
        res[name] = {"width": d.width, "height": d.height, "ms_per_image_median": float(np.median(ts)) * 1e3,
                     "mp_per_s": d.width * d.height / 1e6 / float(np.median(ts)), "runs": len(ts)}
    return {"workload": "configs[0]: one baseline SOF0 4:2:0 JPEG from tests/Assets decoded to RGB on the CPU (reference path, no GPU)",
            "cores": 1, "kind": "port", "images": res}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    if rank != 0:
        return
    import oracle_ffi as O
    threads = os.cpu_count() or 1
    blobs = make_inputs(min(args.distinct, 8), 0, False)["restart"]
    mp_img = W4K * H4K / 1e6
    images = max(threads, min(16 * threads, 512))  # a few seconds of work per step on all host threads

    def run(n):
        sample = [blobs[i % len(blobs)] for i in range(n)]
        t0 = time.perf_counter()
        assert O.decode_batch_rgb(sample, threads) == 0
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        run(max(2, images // 4))
    dt = sum(run(images) for _ in range(args.steps))
    val = images * args.steps * mp_img / dt
    line = {
        "impl": "reference", "metric": "decoded_megapixels_per_second", "value": val, "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16/fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES["restart"], "images_per_step": images,
                   "note": "the reference arm's step is a bounded sample of the same workload (the GPU arm decodes "
                           "images_per_gpu_per_step images per step)"},
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": threads, "kind": "port",
                         "sample": f"{images} images per step x {args.steps} steps, one image per task, "
                                   "C restatement of the reference algorithm (reference .NET runtime unavailable)"},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU per step")
    ap.add_argument("--distinct", type=int, default=128, help="distinct synthetic images (replicated to --batch)")
    ap.add_argument("--e2e-batch", type=int, default=0,
                    help="images per step of the host-buffer (e2e) leg (0: 512 on 1-2 GPUs, 256 beyond: every rank pins its own "
                         "24.9 MB per image, and the step's pipeline fill is amortised over the batch)")
    ap.add_argument("--e2e-chunk", type=int, default=16, help="images per pipelined chunk of the e2e leg")
    ap.add_argument("--cpu-seconds", type=float, default=18.0, help="length of the headline cpu_baseline sample")
    ap.add_argument("--workload", default="all", choices=["all", "restart", "norestart", "progressive", "encode"],
                    help="all: configs[1] as the headline + every other config under `workloads` (N = 1); else that one alone")
    ap.add_argument("--no-restart", action="store_true", help="= --workload norestart")
    ap.add_argument("--progressive", action="store_true", help="= --workload progressive")
    ap.add_argument("--encode", action="store_true", help="= --workload encode")
    args = ap.parse_args()
    if args.no_restart:
        args.workload = "norestart"
    if args.progressive:
        args.workload = "progressive"
    if args.encode:
        args.workload = "encode"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.e2e_batch <= 0:
        args.e2e_batch = 512 if world <= 2 else 256
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import jpeglibrary_b200 as J

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: jpeglibrary_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    env = {"torch": torch, "dist": dist, "J": J, "ctx": J.Context(local_rank), "world": world, "rank": rank, "local_rank": local_rank}
    everything = args.workload == "all" and world == 1
    inp = shared_inputs(args, rank, world, barrier, extra=everything or args.workload in ("norestart", "progressive", "encode"))

    head_kind = "restart" if args.workload == "all" else args.workload
    if head_kind == "encode":
        head = encode_workload(inp["frames"], args, env, args.steps, args.warmup)
    else:
        head = decode_workload(head_kind, inp[head_kind], args, env, args.steps, args.warmup, full=True)
    workloads = None
    if everything:
        few = max(2, min(args.steps, 3))

        def fresh(fn, *a, **k):
            # every workload starts from a trimmed memory pool, like a `--workload X` run of its own: the pools of the
            # contexts still hold the 53 GB of the headline batch, and what a workload gets carved out of that is laid
            # out differently in HBM (the progressive scan kernel, scattered two-byte stores, ran 7 % slower on it)
            for c in (env.get("ctx"), env.get("ctx2")):
                if c is not None:
                    c.synchronize()
                    c.trim()
            return fn(*a, **k)
        workloads = {
            "configs[0]": cpu_only_config0(),
            "configs[2]": fresh(decode_workload, "norestart", inp["norestart"], args, env, few, args.warmup, full=False),
            "configs[3]": fresh(decode_workload, "progressive", inp["progressive"], args, env, few, args.warmup, full=False),
            "configs[4]": fresh(encode_workload, inp["frames"], args, env, few, args.warmup),
        }

    if rank == 0:
        line = {
            "metric": head.get("metric", "decoded_megapixels_per_second"), "value": head["value"], "unit": "MP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16/fp32",
            "data": "synthetic",
            "config": {"workload": head["workload"], "images_per_gpu_per_step": head["images_per_gpu_per_step"],
                       "distinct_images": head["distinct_images"], "l2": head["l2"],
                       **{k: head[k] for k in ("compressed_bytes_per_step", "compressed_gb_per_s", "output", "scan_bytes_per_step",
                                               "us_per_frame") if k in head},
                       "parity": head["parity"],
                       "max_abs_rgb_diff_vs_oracle": (head["parity"] or {}).get("max_abs_rgb_diff_vs_oracle")},
            "clocks": head["clocks"],
            "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"],
        }
        if "value_h2d" in head:
            line["value_h2d"] = head["value_h2d"]
        if "cpu_baseline" in head:
            line["cpu_baseline"] = head["cpu_baseline"]
        if workloads:
            line["workloads"] = workloads
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
