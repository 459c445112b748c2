"""Deterministic synthetic inputs (SURVEY section 8d / BASELINE.md section 2).

image i <- np.random.default_rng(1000+i): sum of low-frequency sinusoids (periods 29..97 px, random
phase) + 8 random filled rectangles + Gaussian noise sigma=6, clipped to uint8 RGB.  JPEGs are
produced with Pillow/libjpeg-turbo (data generation only -- never on a measured or checked path).
"""
import io

import numpy as np


def synth_rgb(i, width, height):
    rng = np.random.default_rng(1000 + i)
    xs = np.arange(width, dtype=np.float32)
    ys = np.arange(height, dtype=np.float32)
    img = np.empty((height, width, 3), dtype=np.float32)
    for c in range(3):
        acc = np.full((height, width), 128.0, dtype=np.float32)
        for _ in range(3):
            px, py = rng.uniform(29, 97, size=2)
            ph = rng.uniform(0, 2 * np.pi)
            amp = np.float32(rng.uniform(15, 40))
            ax = (2 * np.pi / px) * xs + ph
            by = (2 * np.pi / py) * ys
            # sin(ax + by) = sin(ax)cos(by) + cos(ax)sin(by): two rank-1 updates instead of W*H sines
            acc += amp * (np.outer(np.cos(by), np.sin(ax)) + np.outer(np.sin(by), np.cos(ax))).astype(np.float32)
        img[..., c] = acc
    for _ in range(8):
        x0 = int(rng.integers(0, max(1, width - 8)))
        y0 = int(rng.integers(0, max(1, height - 8)))
        w = int(rng.integers(8, max(9, width // 4)))
        h = int(rng.integers(8, max(9, height // 4)))
        col = rng.uniform(0, 255, size=3).astype(np.float32)
        img[y0:y0 + h, x0:x0 + w, :] = col
    img += rng.standard_normal(size=img.shape, dtype=np.float32) * np.float32(6)
    return np.clip(img, 0, 255).astype(np.uint8)


def encode_jpeg(rgb, quality=85, subsampling="4:2:0", restart_rows=0, restart_blocks=0, progressive=False,
                optimize=False, gray=False):
    from PIL import Image
    im = Image.fromarray(rgb if not gray else rgb[..., 0], "L" if gray else "RGB")
    buf = io.BytesIO()
    kw = dict(format="JPEG", quality=quality, progressive=progressive, optimize=optimize)
    if not gray:
        kw["subsampling"] = subsampling
    if restart_rows:
        kw["restart_marker_rows"] = restart_rows
    if restart_blocks:
        kw["restart_marker_blocks"] = restart_blocks
    im.save(buf, **kw)
    return buf.getvalue()


def synth_jpeg(i, width, height, **kw):
    return encode_jpeg(synth_rgb(i, width, height), **kw)
