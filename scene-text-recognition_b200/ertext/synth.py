"""Synthetic inputs for tests and bench: seeded, no dataset access.

s_text_frame : "S-text" -- low-frequency colour background + ~150 glyph-like blobs (rotated boxes,
               ellipses, 3-stroke polylines) with contrast +-40..120, Gaussian blur 0.8, noise sigma 3.
s_noise_frame: i.i.d. uniform u8 BGR (worst case for the component tree: ~0.39 nodes / pixel).
"""
import numpy as np


def s_text_frame(seed, w=1920, h=1080, n_glyphs=150):
    import cv2
    rng = np.random.RandomState(seed)
    grid = rng.uniform(40, 215, (17, 30, 3)).astype(np.float32)
    img = cv2.resize(grid, (w, h), interpolation=cv2.INTER_LINEAR)
    scale = h / 1080.0
    for _ in range(n_glyphs):
        gh = rng.uniform(16, 96) * scale
        gw = gh * rng.uniform(0.3, 1.2)
        cx, cy = rng.uniform(gw, w - gw), rng.uniform(gh, h - gh)
        ang = rng.uniform(-25, 25)
        base = img[int(cy), int(cx)].copy()
        off = rng.uniform(40, 120) * (1 if rng.rand() < 0.5 else -1)
        col = tuple(float(np.clip(c + off, 0, 255)) for c in base)
        kind = rng.randint(3)
        if kind == 0:
            box = cv2.boxPoints(((cx, cy), (gw, gh), ang)).astype(np.int32)
            cv2.fillPoly(img, [box], col)
            if rng.rand() < 0.5:   # punch a hole: characters with counters
                box2 = cv2.boxPoints(((cx, cy), (gw * 0.4, gh * 0.4), ang)).astype(np.int32)
                cv2.fillPoly(img, [box2], tuple(float(c) for c in base))
        elif kind == 1:
            cv2.ellipse(img, ((cx, cy), (gw, gh), ang), col, -1)
            if rng.rand() < 0.5:
                cv2.ellipse(img, ((cx, cy), (gw * 0.45, gh * 0.45), ang), tuple(float(c) for c in base), -1)
        else:
            pts = np.stack([rng.uniform(cx - gw / 2, cx + gw / 2, 4), rng.uniform(cy - gh / 2, cy + gh / 2, 4)], 1).astype(np.int32)
            cv2.polylines(img, [pts], False, col, thickness=max(2, int(gh / 8)))
    img = cv2.GaussianBlur(img, (0, 0), 0.8)
    img += rng.normal(0, 3, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def s_noise_frame(seed, w=1920, h=1080):
    return np.random.RandomState(seed).randint(0, 256, (h, w, 3)).astype(np.uint8)


def s_text_batch(first_seed, n, w=1920, h=1080):
    return np.stack([s_text_frame(first_seed + i, w, h) for i in range(n)])


def svm_features_u8(seed, n, dims=1800):
    """Feature vectors shaped like OCR::extract_feature output: ~32 % non-zeros, values k/255 stored as k."""
    rng = np.random.RandomState(seed)
    x = np.zeros((n, dims), np.uint8)
    dens = 0.05 + 0.55 * rng.rand(n, 1)
    mask = rng.rand(n, dims) < dens
    x[mask] = rng.randint(1, 256, int(mask.sum()))
    return x


def lbp_like_hist_u8(seed, n):
    """Random multinomial histograms with the LBP layout (4 blocks x 256 bins, each block sums to 144)."""
    rng = np.random.RandomState(seed)
    out = np.zeros((n, 1024), np.uint8)
    for b in range(4):
        codes = rng.randint(0, 256, (n, 144))
        for i in range(n):
            out[i, b * 256:(b + 1) * 256] = np.bincount(codes[i], minlength=256)
    return out
