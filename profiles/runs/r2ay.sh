set -x
python - <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, jpeglibrary_b200 as J, oracle_ffi as O, synth
for kw,(w,h) in ((dict(quality=34, subsampling="4:4:4"),(15,7)), (dict(quality=80, subsampling="4:2:0"),(64,48)), (dict(quality=80, gray=True),(33,20))):
    blob = synth.encode_jpeg(synth.synth_rgb(5,w,h), **kw)
    want = O.decode(blob)
    def dec(buf, fmt, **k):
        d = J.JpegDecoder(); d.SetInput(blob); d.Identify(); d.SetOutputWriter(J.CudaOutputWriter(buf, fmt, **k)); d.Decode()
    out = np.zeros((h,w,3),np.uint8); dec(out, J.JB_OUT_RGB24)
    rgba = np.zeros((h,w,4),np.uint8); dec(rgba, J.JB_OUT_RGBA32)
    ycc = np.zeros((h,w,3),np.uint8); dec(ycc, J.JB_OUT_YCBCR888)
    pitch = (3*w + 17 + 3)//4*4
    padded = np.full((h,pitch),0xA5,np.uint8); dec(padded, J.JB_OUT_RGB24, pitch=pitch)
    print(kw, w, h, 'rgba', np.array_equal(rgba[...,:3], out), (rgba[...,3]==255).all(), 'ycc', np.array_equal(ycc, want.ycbcr), want.ycbcr.shape, 'pitch', np.array_equal(padded[:, :3*w].reshape(h,w,3), out), (padded[:,3*w:]==0xA5).all())
    if not (padded[:,3*w:]==0xA5).all(): print(padded[:2, 3*w-3:3*w+8])
PY
