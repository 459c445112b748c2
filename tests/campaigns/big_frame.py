"""One very large frame (default 40000 x 36000 = 1.44 GP: RGB output and coefficient store both above 4 GiB) through
the GPU path and the oracle: 32-bit overflow check.  usage (on a GPU box): python tests/campaigns/big_frame.py [width height] [--no-restart]"""
import sys, time, hashlib, io
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jpeglibrary_b200 as J, oracle_ffi as O, synth
from PIL import Image
Image.MAX_IMAGE_PIXELS = None
W = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 36000
t0 = time.time()
tile = synth.synth_rgb(5, 2000, 1800)
rgb = np.tile(tile, ((H + 1799) // 1800, (W + 1999) // 2000, 1))[:H, :W]
rgb = np.ascontiguousarray(rgb)
rgb[::997, ::991] ^= 0x5A   # break the periodicity a little
restart = 0 if "--no-restart" in sys.argv else 1
blob = synth.encode_jpeg(rgb, quality=80, subsampling="4:2:0", restart_rows=restart)
del rgb
print(f"{W}x{H}: {len(blob) / 1e6:.1f} MB compressed, generated in {time.time() - t0:.0f} s", flush=True)
ctx = J.Context(0)
out = ctx.pinned_array(W * H * 3).reshape(H, W, 3)
t0 = time.time()
dec = J.JpegDecoder(ctx)
dec.SetInput(blob)
dec.SetOutputWriter(J.CudaOutputWriter(out))
dec.Decode()
print(f"GPU decode {time.time() - t0:.2f} s", flush=True)
t0 = time.time()
o = O.decode(blob)
print(f"oracle decode {time.time() - t0:.0f} s", flush=True)
same = True
for y0 in range(0, H, 4000):   # compare in bands (keeps peak memory down)
    a, b = out[y0:y0 + 4000], o.rgb[y0:y0 + 4000]
    if not np.array_equal(a, b):
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        print(f"rows {y0}..: max diff {int(d.max())}, {int((d > 1).sum())} samples differ by more than 1")
        same = same and d.max() <= 1
print("RGB within +-1 everywhere" if same else "MISMATCH")
