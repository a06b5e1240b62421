"""Deterministic synthetic inputs: frames and sequences.

Pure numpy (PCG64 seeds) so the same bytes are produced in the build container and on the GPU box.
"""
import numpy as np


def _gauss1d(sigma):
    r = int(3 * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-x * x / (2 * sigma * sigma))
    return k / k.sum()


def _blur(img, sigma):
    k = _gauss1d(sigma)
    r = len(k) // 2
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = sum(k[i] * p[:, i:i + img.shape[1]] for i in range(len(k)))
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    return sum(k[i] * p[i:i + img.shape[0], :] for i in range(len(k)))


def synth_frame(w=640, h=480, seed=1, n_rect=60, n_seg=30):
    """Busy grayscale frame: blurred noise + filled rectangles + line segments (config 1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = _blur(rng.random((h, w)), 2.0)
    base = (base - base.min()) / (base.max() - base.min())
    img = (base * 255.0)
    for _ in range(n_rect):
        rw, rh = rng.integers(10, 120, 2)
        x0 = rng.integers(0, max(1, w - rw)); y0 = rng.integers(0, max(1, h - rh))
        img[y0:y0 + rh, x0:x0 + rw] = rng.integers(0, 256)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(n_seg):
        x0, x1 = rng.integers(0, w, 2); y0, y1 = rng.integers(0, h, 2)
        t = rng.integers(1, 4); g = rng.integers(0, 256)
        dx, dy = float(x1 - x0), float(y1 - y0)
        L2 = dx * dx + dy * dy + 1e-9
        u = np.clip(((xx - x0) * dx + (yy - y0) * dy) / L2, 0, 1)
        d2 = (xx - (x0 + u * dx)) ** 2 + (yy - (y0 + u * dy)) ** 2
        img[d2 <= (t * 0.5) ** 2 + 0.25] = g
    # mild sensor noise so the texture statistics resemble a camera image
    img = img + rng.normal(0, 3.0, (h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def warp_frame(img, seed, max_t=4.0, max_rot_deg=0.5, max_ds=0.005):
    """Similarity-warped copy of img (bilinear, border clamp) — sequence frames (config 2)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h, w = img.shape
    tx, ty = rng.uniform(-max_t, max_t, 2)
    a = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg)); s = 1 + rng.uniform(-max_ds, max_ds)
    c, sn = np.cos(a) * s, np.sin(a) * s
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    cx, cy = w / 2, h / 2
    sx = c * (xx - cx) - sn * (yy - cy) + cx + tx
    sy = sn * (xx - cx) + c * (yy - cy) + cy + ty
    sx = np.clip(sx, 0, w - 1.001); sy = np.clip(sy, 0, h - 1.001)
    x0 = np.floor(sx).astype(np.int64); y0 = np.floor(sy).astype(np.int64)
    fx = sx - x0; fy = sy - y0
    f = img.astype(np.float64)
    out = (f[y0, x0] * (1 - fx) * (1 - fy) + f[y0, x0 + 1] * fx * (1 - fy) +
           f[y0 + 1, x0] * (1 - fx) * fy + f[y0 + 1, x0 + 1] * fx * fy)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def synth_sequence(n, w=640, h=480, seed=1):
    f0 = synth_frame(w, h, seed)
    return np.stack([f0] + [warp_frame(f0, 1000 * seed + k) for k in range(1, n)])
