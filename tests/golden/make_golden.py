"""Regenerates tests/golden/*: run in the build container where /root/reference exists.

For each Huffman DCT asset of the reference's test-suite (tests/Assets, see
tests/JpegLibrary.Tests/Decoder/Huffman{Sequential,Progressive}DecodeTests.cs) this script
  1. copies the .jpg (input fixture; the GPU box has no /root/reference),
  2. loads the reference's golden 16-bit buffer from <asset>.jpg.high.png / .low-diff.png exactly
     like tests/JpegLibrary.Tests/Utils/ImageHelper.cs:12-91,
  3. checks the CPU oracle against it bit-exactly (including the test writer's (ushort) cast
     quirk, Utils/JpegExtendingOutputWriter.cs:57,77-80), and
  4. stores the sha256 of the golden buffer and of the oracle's unclamped int16 planes in
     golden.json, so that `-m gpu` tests can check the CUDA path against the reference's goldens
     without the reference tree.
Identify known-answers come from Decoder/MetadataIdentifyTests.cs:19-130.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_ffi as O  # noqa: E402

ASSETS = "/root/reference/tests/Assets"
FILES = [
    "baseline/cramps.jpg",
    "baseline/lake.jpg",
    "huffman_sequential/testorig12.jpg",
    "huffman_progressive/progress.jpg",
    "huffman_progressive/yellowcat_progressive_restart.jpg",
] + ["huffman_lossless/lossless%d_s22.jpg" % i for i in range(1, 8)]
# inputs without a golden buffer in the reference: HETissueSlide.jpg is the asset of the reference's own benchmark
# (tests/JpegLibrary.Benchmarks/DecoderBenchmark.cs:19-43) and of BASELINE.json configs[0]
INPUT_ONLY = ["baseline/HETissueSlide.jpg"]
IDENTIFY = {  # MetadataIdentifyTests.cs
    "cramps.jpg": dict(Width=800, Height=607, NumberOfComponents=1, Precision=8, JpegStreamSize=137766),
    "testorig12.jpg": dict(Width=227, Height=149, NumberOfComponents=3, Precision=12, JpegStreamSize=12394),
    "yellowcat_progressive_restart.jpg": dict(Width=720, Height=540, NumberOfComponents=3, Precision=8, JpegStreamSize=45703),
    "progress.jpg": dict(Width=341, Height=486, NumberOfComponents=3, Precision=8, JpegStreamSize=44884),
    "HETissueSlide.jpg": dict(Width=2048, Height=2048, NumberOfComponents=3, Precision=8, JpegStreamSize=783426),
}


def load_golden16(path, ncomp):
    hi = np.array(Image.open(path + ".high.png").convert("RGBA")).astype(np.uint16)
    lo = np.array(Image.open(path + ".low-diff.png").convert("RGBA")).astype(np.uint16)
    return ((hi << 8) | (hi ^ lo))[..., :ncomp]


def expected16(planes, precision):
    """What JpegExtendingOutputWriter stores for unclamped int16 samples."""
    s = planes.astype(np.int16).view(np.uint16).astype(np.uint32)  # (ushort) cast quirk
    s = np.minimum(s, (1 << precision) - 1)
    rem = 16 - precision
    e = (s << rem) | (s & ((1 << rem) - 1))
    return e.astype(np.uint16).transpose(1, 2, 0)


def main():
    out = {"assets": {}, "identify": IDENTIFY}
    for rel in FILES:
        src = os.path.join(ASSETS, rel)
        name = os.path.basename(rel)
        shutil.copyfile(src, os.path.join(HERE, name))
        os.chmod(os.path.join(HERE, name), 0o644)
        d = O.decode(open(src, "rb").read())
        gold = load_golden16(src, d.ncomp)
        mine = expected16(d.planes, d.precision)
        mism = int((gold != mine).sum())
        assert mism == 0, (rel, mism)
        out["assets"][name] = {
            "source": "tests/Assets/" + rel,
            "width": d.width, "height": d.height, "ncomp": d.ncomp, "precision": d.precision, "sof": d.sof,
            "golden16_sha256": hashlib.sha256(np.ascontiguousarray(gold).tobytes()).hexdigest(),
            "planes_i16_sha256": hashlib.sha256(np.ascontiguousarray(d.planes).tobytes()).hexdigest(),
            "negative_samples": int((d.planes < 0).sum()),
        }
        print(rel, "oracle == reference golden (0 mismatches of %d samples)" % gold.size)
    for rel in INPUT_ONLY:
        name = os.path.basename(rel)
        shutil.copyfile(os.path.join(ASSETS, rel), os.path.join(HERE, name))
        os.chmod(os.path.join(HERE, name), 0o644)
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
