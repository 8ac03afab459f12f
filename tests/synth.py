"""Seeded synthetic inputs shared by the tests and bench.py."""
import numpy as np


def depth_frames(n, W=640, H=480, seed=3, holes=0.03):
    """Ground plane + 3-6 boxes/spheres, 400..4000 mm, `holes` fraction of zeros.  u16[n,H,W]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    out = np.zeros((n, H, W), np.uint16)
    for f in range(n):
        horizon = H * (0.35 + 0.1 * rng.random())
        ground = np.where(yy > horizon, 1200.0 * H / np.maximum(yy - horizon, 1.0) * 0.25 + 400.0, 4000.0)
        d = np.clip(ground, 400.0, 4000.0)
        for _ in range(int(rng.integers(3, 7))):
            cx, cy = rng.random() * W, horizon + rng.random() * (H - horizon)
            w, h = (0.05 + 0.15 * rng.random()) * W, (0.08 + 0.25 * rng.random()) * H
            z = 500.0 + 3000.0 * rng.random()
            if rng.random() < 0.5:
                m = (np.abs(xx - cx) < w / 2) & (yy < cy) & (yy > cy - h)
                d = np.where(m, np.minimum(d, z), d)
            else:
                r = 0.5 * min(w, h)
                rr = (xx - cx) ** 2 + (yy - (cy - r)) ** 2
                m = rr < r * r
                d = np.where(m, np.minimum(d, z - np.sqrt(np.maximum(r * r - rr, 0.0)) * 2.0), d)
        d = np.clip(d, 400.0, 4000.0)
        d[rng.random((H, W)) < holes] = 0.0
        out[f] = d.astype(np.uint16)
    return out


def target_frames(n, W=640, H=480, seed=4, blobs=5):
    """u16[n,H,W] = cls | id<<8 blobs (cls in 0..3, id < 100) on a terrain (0) background."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    out = np.zeros((n, H, W), np.uint16)
    for f in range(n):
        for b in range(blobs):
            cx, cy, r = rng.integers(0, W), rng.integers(0, H), rng.integers(4, max(5, H // 10))
            cls = int(rng.integers(1, 4))
            idv = int(rng.integers(0, 100)) if cls == 3 else 0
            out[f][(xx - cx) ** 2 + (yy - cy) ** 2 < r * r] = cls | (idv << 8)
    return out


def rgb_tiles(n, S=224, seed=2):
    return np.random.default_rng(seed).integers(0, 256, (n, S, S, 3), dtype=np.uint8)


def rgb_frames(n, W=640, H=480, seed=5):
    """u32[n, H*W] = r<<24|g<<16|b<<8 (scene.rs:86): smooth blobs + noise so that resampling matters."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    out = np.zeros((n, H * W), np.uint32)
    for f in range(n):
        img = rng.integers(0, 64, (H, W, 3)).astype(np.float32)
        for _ in range(8):
            cx, cy, r = rng.random() * W, rng.random() * H, 10 + rng.random() * 80
            col = rng.random(3) * 255
            img += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))[..., None] * col
        img = np.clip(img, 0, 255).astype(np.uint32)
        out[f] = ((img[..., 0] << 24) | (img[..., 1] << 16) | (img[..., 2] << 8)).reshape(-1)
    return out
