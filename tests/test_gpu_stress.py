"""Resource behaviour of the C-ABI under repeated use: batches are created, run and destroyed thousands of times in a
serving process, so device memory, pinned memory and mailbox slots must all come back."""
import numpy as np
import pytest

import jpeglibrary_b200 as J
import oracle_ffi as O
import synth

pytestmark = pytest.mark.gpu


def _free_device_bytes():
    import torch
    torch.cuda.synchronize()
    return torch.cuda.mem_get_info(0)[0]


def test_repeated_batches_do_not_leak_device_memory():
    rgb = synth.synth_rgb(70, 320, 240)
    blobs = [synth.encode_jpeg(rgb, restart_rows=1),
             synth.encode_jpeg(rgb),                                   # self-synchronising path
             synth.encode_jpeg(rgb, progressive=True),                 # scan list
             synth.synth_lossless(71, 64, 48, predictor=3, restart=16)[0],
             synth.encode_jpeg(rgb, gray=True, restart_blocks=5)]
    want = [O.decode(b).rgb for b in blobs]
    ctx = J.Context(0)

    def one_round(check):
        with J.JpegBatchDecoder(blobs * 4, J.JB_OUT_RGB24, context=ctx, device_output=True) as b:
            b.run()
            if check:
                for i in (0, 6, 12, 18, 19):
                    assert np.array_equal(b.read_output(i), want[i % len(blobs)])
        # single-image path, host output, every frame type
        for blob, w in zip(blobs, want):
            out = np.zeros_like(w)
            dec = J.JpegDecoder(ctx)
            dec.SetInput(blob)
            dec.SetOutputWriter(J.CudaOutputWriter(out))
            dec.Decode()
            if check:
                assert np.array_equal(out, w)
        # encoder + optimizer
        jpg, _ = J.encode_rgb(rgb, quality=80, context=ctx)
        assert len(jpg) > 1000
        opt = J.JpegOptimizer(ctx)
        opt.SetInput(blobs[0])
        opt.Scan()
        sink = bytearray()
        opt.SetOutput(sink)
        opt.Optimize()
        opt._close()

    for _ in range(5):      # pools and caches reach their steady size
        one_round(True)
    ctx.synchronize()
    before = _free_device_bytes()
    for k in range(60):
        one_round(k % 20 == 0)
    ctx.synchronize()
    after = _free_device_bytes()
    # the stream-ordered pools may keep what they have, but they must not grow round after round
    assert before - after < 64 << 20, f"device memory shrank by {(before - after) >> 20} MiB over 60 rounds"


def test_errors_do_not_leak_or_wedge_the_context():
    """A batch that fails (corrupt stream) must leave the context usable and its resources released."""
    good = synth.synth_jpeg(72, 256, 160, restart_rows=1)
    bad = bytearray(good)
    i = bad.find(b"\xff\xd1")
    bad[i + 1] = 0xC9   # a restart marker becomes another marker: "Expect restart marker."
    ctx = J.Context(0)
    want = O.decode(good).rgb
    before = None
    for k in range(40):
        with pytest.raises((J.InvalidOperationException, J.InvalidDataException)):
            with J.JpegBatchDecoder([good, bytes(bad), good], J.JB_OUT_RGB24, context=ctx, device_output=True) as b:
                b.run()
        with J.JpegBatchDecoder([good], J.JB_OUT_RGB24, context=ctx, device_output=True) as b:
            b.run()
            assert np.array_equal(b.read_output(0), want)
        if k == 5:
            before = _free_device_bytes()
    assert before - _free_device_bytes() < 32 << 20
