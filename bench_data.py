"""Seeded synthetic inputs of the benchmark and the tests (shapes and statistics per SURVEY.md 8d): normalised image
tiles, building-like targets [mask, distance, size] and soft probability maps.  numpy only; no reference or oracle
code.  (oracle/synthetic.py re-exports these names for the golden generator and the tests.)"""
import numpy as np


def rectangles_mask(rs, h, w, n_rect=40, lo=6, hi=30):
    """building-like binary mask: axis-aligned rectangles; returns (mask uint8, size map float32)"""
    mask = np.zeros((h, w), np.uint8)
    size = np.ones((h, w), np.float32)
    scale = max(h, w) / 300.0
    for _ in range(n_rect):
        rh, rw = int(rs.randint(lo, hi + 1) * scale) + 1, int(rs.randint(lo, hi + 1) * scale) + 1
        y0, x0 = rs.randint(0, max(1, h - rh)), rs.randint(0, max(1, w - rw))
        mask[y0:y0 + rh, x0:x0 + rw] = 1
        size[y0:y0 + rh, x0:x0 + rw] = np.round(np.sqrt(rh * rw))
    return mask, size


def train_batch(n, s, seed=1234, n_rect=40):
    """X ~ N(0,1) (n,3,s,s) f32; target (n,3,s,s) f32 = [mask, integer distance 0..255 (0 inside), sqrt-size map]"""
    rs = np.random.RandomState(seed)
    x = rs.randn(n, 3, s, s).astype(np.float32)
    t = np.zeros((n, 3, s, s), np.float32)
    for i in range(n):
        m, size = rectangles_mask(rs, s, s, n_rect)
        d = rs.randint(0, 256, (s, s)).astype(np.float32) * (1 - m)
        t[i, 0], t[i, 1], t[i, 2] = m, d, size
    return x, t


def probability_maps(n, s, seed=1234, n_rect=40):
    """soft building probabilities (n,2,s,s) f32 = [1-p, p], p = sigmoid(blurred noise + 5*rectangles - 2)"""
    from scipy import ndimage as ndi
    rs = np.random.RandomState(seed)
    out = np.zeros((n, 2, s, s), np.float32)
    for i in range(n):
        m, _ = rectangles_mask(rs, s, s, n_rect)
        z = ndi.gaussian_filter(rs.randn(s, s) * 0.5 - 2.0 + 5.0 * m, 1.0)
        p = (1.0 / (1.0 + np.exp(-z))).astype(np.float32)
        out[i, 0], out[i, 1] = 1 - p, p
    return out
