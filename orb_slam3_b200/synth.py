"""Synthetic inputs shared by tests and bench.py.

`synth_frame` avoids any OpenCV dependency so it produces the same bytes on the
build container and on the GPU box.
"""
import numpy as np


def _box_blur(a, r):
    """Separable integer box blur (float32 accumulate), reflect border."""
    k = 2 * r + 1
    p = np.pad(a.astype(np.float32), r, mode="reflect")
    c = np.cumsum(p, axis=0)
    c = np.concatenate([np.zeros((1, c.shape[1]), np.float32), c], 0)
    v = (c[k:] - c[:-k]) / k
    c = np.cumsum(v, axis=1)
    c = np.concatenate([np.zeros((c.shape[0], 1), np.float32), c], 1)
    return (c[:, k:] - c[:, :-k]) / k


def synth_frame(h, w, seed, n_rect=400, low_texture=False):
    """Noise -> blur -> `n_rect` constant rectangles; ~3.8e4 raw FAST corners over 8 levels at 720p.

    low_texture=True gives a smooth frame that exercises the minThFAST fallback
    (ORBextractor.cc:843-846)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w)).astype(np.float32)
    if low_texture:
        img = _box_blur(_box_blur(img, 4), 4)
        img = (img - img.min()) / max(float(np.ptp(img)), 1e-6) * 60 + 90
        n_rect = n_rect // 10
    else:
        img = _box_blur(img, 2)
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    for _ in range(n_rect):
        rw, rh = rng.integers(8, 40, size=2)
        x0 = rng.integers(0, max(w - rw, 1))
        y0 = rng.integers(0, max(h - rh, 1))
        img[y0:y0 + rh, x0:x0 + rw] = rng.integers(0, 256)
    return np.ascontiguousarray(img)


def shifted_frame(img, dx, dy, seed):
    """Frame t+1 of a stream: integer translation + fresh noise in the uncovered strip."""
    rng = np.random.default_rng(seed)
    out = np.roll(img, (dy, dx), axis=(0, 1)).copy()
    h, w = img.shape
    if dy > 0:
        out[:dy] = rng.integers(0, 256, size=(dy, w))
    elif dy < 0:
        out[dy:] = rng.integers(0, 256, size=(-dy, w))
    if dx > 0:
        out[:, :dx] = rng.integers(0, 256, size=(h, dx))
    elif dx < 0:
        out[:, dx:] = rng.integers(0, 256, size=(h, -dx))
    return np.ascontiguousarray(out)


def stereo_right(left, seed, disparities=(12,), noise=3):
    """Right image of a rectified pair: the left image cut into len(disparities) horizontal bands, band k
    moved left by disparities[k] px (right[y, x] = left[y, x + d]), fresh noise in the uncovered strip and
    +-`noise` grey levels of sensor noise so that the L1 window distances are not all zero."""
    rng = np.random.default_rng(seed)
    h, w = left.shape
    out = np.empty_like(left)
    edges = np.linspace(0, h, len(disparities) + 1).astype(int)
    for k, d in enumerate(disparities):
        y0, y1 = edges[k], edges[k + 1]
        band = np.roll(left[y0:y1], -int(d), axis=1).copy()
        if d > 0:
            band[:, w - d:] = rng.integers(0, 256, size=(y1 - y0, d))
        elif d < 0:
            band[:, :-d] = rng.integers(0, 256, size=(y1 - y0, -d))
        out[y0:y1] = band
    if noise:
        n = rng.integers(-noise, noise + 1, size=(h, w))
        out = np.clip(out.astype(np.int32) + n, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(out)
