set -x
JB_DEBUG_STATUS=1 python - <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, jpeglibrary_b200 as J, oracle_ffi as O, synth
from PIL import Image
import io
blob=open('tests/fixtures/valid_420_no_restart_slow_to_synchronise.jpg','rb').read()
def run(blob, tag):
    dec = J.JpegDecoder(); dec.SetInput(blob); dec.Identify()
    planes = np.zeros((dec.NumberOfComponents, dec.Height, dec.Width), dtype=np.int16)
    dec.SetOutputWriter(J.CudaOutputWriter(planes, J.JB_OUT_PLANAR_I16))
    try:
        dec.Decode(); print(tag, "decoded", np.array_equal(planes, O.decode(blob, want_rgb=False).planes), flush=True)
    except Exception as e: print(tag, type(e).__name__, e, flush=True)
run(blob, "fixture")
# the same pixels re-encoded by Pillow with restart markers: same optimised tables, the flat path
rgb = np.array(Image.open(io.BytesIO(blob)).convert("RGB"))
run(synth.encode_jpeg(rgb, quality=96, subsampling="4:2:0", optimize=True, restart_rows=1), "restart variant")
run(synth.encode_jpeg(rgb, quality=96, subsampling="4:2:0", optimize=True), "re-encoded no restart")
PY
