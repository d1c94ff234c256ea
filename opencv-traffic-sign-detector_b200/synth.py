"""Synthetic frames and MSER-like candidate boxes used by bench.py and the tests.

Inputs only -- no stage of the path is computed here.  numpy only (no cv2, no oracle).
"""
import numpy as np

FRAME_SEED = 1234
BOX_SEED = 4321


def _hsv_to_bgr(h, s, v):
    """h in [0,360), s,v in [0,1] -> uint8 BGR (plain textbook conversion; only used to pick rectangle colours)."""
    c = v * s
    hp = h / 60.0
    x = c * (1 - abs(hp % 2 - 1))
    r, g, b = [(c, x, 0), (x, c, 0), (0, c, x), (0, x, c), (x, 0, c), (c, 0, x)][int(hp) % 6]
    m = v - c
    return np.array([(b + m) * 255, (g + m) * 255, (r + m) * 255]).round().astype(np.uint8)


def make_frames(nframes, H=800, W=1360, seed=FRAME_SEED, nrect=64):
    """uint8[nframes,H,W,3] BGR: mid-grey background, `nrect` random rectangles with colours uniform in HSV,
    N(0, 8^2) noise, clipped."""
    rng = np.random.default_rng(seed)
    out = np.empty((nframes, H, W, 3), np.uint8)
    for f in range(nframes):
        img = np.full((H, W, 3), 128, np.int16)
        for _ in range(nrect):
            rw = int(rng.integers(8, max(9, W // 6))); rh = int(rng.integers(8, max(9, H // 4)))
            x = int(rng.integers(0, W - 4)); y = int(rng.integers(0, H - 4))
            col = _hsv_to_bgr(float(rng.uniform(0, 360)), float(rng.uniform(0, 1)), float(rng.uniform(0, 1)))
            img[y:y + rh, x:x + rw] = col
        noise = np.rint(rng.standard_normal((H, W, 3), np.float32) * 8.0).astype(np.int16)
        out[f] = np.clip(img + noise, 0, 255).astype(np.uint8)
    return out


def _crop_side(x, w, p, limit):
    """Clipped crop extent along one axis for box start x, size w, enlargement p (same f64 steps as the path)."""
    d = (w * (p - 1)) * 0.5
    lo = x - d if x - d > 0 else 0
    hi = x + w + d
    return min(int(hi), limit) - min(int(lo), limit)


def make_boxes(nframes, nboxes, H=800, W=1360, seed=BOX_SEED, enlarge=1.30, D=25):
    """-> (boxes int32[nframes*nboxes,4] (x,y,w,h), offsets int32[nframes+1]).

    w log-normal (median 28, p95 89) clipped to [3,300]; h = round(w / r), r ~ U(0.6, 1.6); 30 % are jittered
    (+-3 px) copies of an earlier box of the same frame; 2 % give exact 2Dx2D / DxD crops; 2 % touch the
    right / bottom frame edge."""
    rng = np.random.default_rng(seed)
    boxes = np.empty((nframes, nboxes, 4), np.int32)
    sigma = np.log(89.0 / 28.0) / 1.645
    for f in range(nframes):
        for i in range(nboxes):
            u = rng.random()
            if i > 0 and u < 0.30:
                src = boxes[f, int(rng.integers(0, i))]
                j = rng.integers(-3, 4, 4)
                w = int(np.clip(src[2] + j[2], 3, min(300, W - 1))); h = int(np.clip(src[3] + j[3], 3, min(300, H - 1)))
                x = int(np.clip(src[0] + j[0], 0, W - w)); y = int(np.clip(src[1] + j[1], 0, H - h))
            else:
                w = int(np.clip(round(float(np.exp(np.log(28.0) + sigma * rng.standard_normal()))), 3, min(300, W - 1)))
                h = int(np.clip(round(w / rng.uniform(0.6, 1.6)), 3, min(300, H - 1)))
                x = int(rng.integers(0, W - w + 1)); y = int(rng.integers(0, H - h + 1))
                if u > 0.98:                                   # touch the right / bottom edge
                    if rng.random() < 0.5:
                        x = W - w
                    else:
                        y = H - h
                elif u > 0.96:                                 # exact 2D x 2D or D x D crop (AREA / copy paths)
                    target = 2 * D if rng.random() < 0.5 else D
                    # interior boxes give odd/even crop sides only for some sizes; the low-side clamp at 0
                    # (DET:167-168) reaches the others, so small offsets are tried as well
                    found = False
                    for cand in range(max(3, int(target / enlarge) - 2), int(target / enlarge) + 3):
                        for xx in [int(rng.integers(8, W - cand - 8))] + list(range(0, 8)):
                            if _crop_side(xx, cand, enlarge, W) != target:
                                continue
                            for yy in [int(rng.integers(8, H - cand - 8))] + list(range(0, 8)):
                                if _crop_side(yy, cand, enlarge, H) == target:
                                    x, y, w, h = xx, yy, cand, cand
                                    found = True
                                    break
                            if found:
                                break
                        if found:
                            break
            boxes[f, i] = (x, y, w, h)
    offsets = (np.arange(nframes + 1) * nboxes).astype(np.int32)
    return boxes.reshape(-1, 4), offsets
