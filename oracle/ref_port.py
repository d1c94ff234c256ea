"""CPU port of the reference's post-MSER pipeline, calling cv2 / numpy exactly where the reference does.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): it exists so that bench.py can time "the reference's own
CPU implementation of the path" on the GPU box, where /root/reference does not exist (cpu_baseline.kind = "port",
bench.py --impl reference).  It keeps the reference's computational structure -- one cv2 call per window, both
histograms recomputed for every compared pair (DET/source.py:200-202), the 625-iteration pure-Python confusion loop
(DET/source.py:553-562) -- because that structure IS the reference's CPU cost.  It is a restatement, not a copy:
each function cites the lines it follows.  Validated against the unmodified reference in
tests/test_oracle_golden.py::test_ref_port_matches_reference_outputs (golden fixtures).
"""
import math

import cv2
import numpy as np

SIGNALLIST = ['prohibicion', 'peligro', 'stop', 'direccionProhibida', 'cedaPaso', 'direccionObligatoria']


def gray_and_enhance_contrast(image):
    """DET/source.py:135-152 + gammaCorrection :599-605: the four cv2 calls of the reference, in its order."""
    gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    eq = cv2.createCLAHE(clipLimit=2).apply(gray)
    blur = cv2.GaussianBlur(eq, (3, 3), 0)
    table = np.array([((i / 255) ** (1 / 2)) * 255 for i in range(256)], np.uint8)
    return cv2.LUT(blur, table)


def expand_or_reject(box, percentage):
    """DET/source.py:155-174."""
    x, y, w, h = box
    grow_w = w * (percentage - 1) * 0.5
    grow_h = h * (percentage - 1) * 0.5
    if not (0.8 < w / h < 1.20):
        return None
    left = x - grow_w if x - grow_w > 0 else 0
    top = y - grow_h if y - grow_h > 0 else 0
    right = x + w + grow_w if x + w + grow_w > 0 else 0
    bottom = y + h + grow_h if y + h + grow_h > 0 else 0
    return int(left), int(top), int(right), int(bottom)


def hs_histogram(window):
    """DET/source.py:575-586."""
    hsv = cv2.cvtColor(window, cv2.COLOR_BGR2HSV)
    hist = cv2.calcHist([hsv], [0, 1], None, [50, 60], [0, 180, 0, 256], accumulate=False)
    return cv2.normalize(hist, hist, alpha=0, beta=1, norm_type=cv2.NORM_MINMAX)


def corner_closeness(xa, ya, xb, yb):
    """DET/source.py:459-462."""
    dist = np.linalg.norm(np.array((xa, ya)) - np.array((xb, yb)))
    if not dist > 0:
        return 1
    return 1 / (1 + np.power(np.e, (((0.154 * np.power(dist, 1.2)) - 31.8) / (0.2 * dist))))


def fold_item(item, survivors, tolerance, by_coords):
    """DET/source.py:192-223: compare the incoming item with every survivor, in order; collect victims; merge."""
    victims = []
    for other in survivors:
        if by_coords:
            a, b = item[1], other[1]
            sim = np.sqrt(corner_closeness(a[0], a[1], b[0], b[1]) * corner_closeness(a[2], a[3], b[2], b[3]))
        else:
            sim = cv2.compareHist(hs_histogram(item[0]), hs_histogram(other[0]), cv2.HISTCMP_CORREL)
        if sim > tolerance:
            victims.append(other)
        elif tolerance * 0.8823 <= sim <= tolerance:
            a, b = item[1], other[1]
            mean_coords = ((a[0] + b[0]) // 2, (a[1] + b[1]) // 2, (a[2] + b[2]) // 2, (a[3] + b[3]) // 2)   # DET:465-468
            item = (cv2.addWeighted(item[0], 0.5, other[0], 0.5, 0.0), mean_coords) + tuple(other[2:])
            victims.append(other)
    return item, victims


def dedup(items, by_coords, tolerance):
    """DET/source.py:177-189 (+ the pop-by-pixel-equality helper :471-477)."""
    survivors = []
    for item in items:
        item, victims = fold_item(item, survivors, tolerance, by_coords)
        for v in victims:
            for idx, s in enumerate(survivors):
                if np.array_equal(s[0], v[0]):
                    survivors.pop(idx)
                    break
        survivors.append(item)
    return survivors


def frame_windows(image, boxes, name, percentage=1.30, D=25):
    """The candidate loop and both de-duplication passes of MSERTrafficSignDetector, DET/source.py:116-131
    (REC/source.py:52-62 with percentage=1.15, D=32)."""
    items = []
    for box in boxes:
        coords = expand_or_reject(box, percentage)
        if coords is not None:
            x1, y1, x2, y2 = coords
            items.append((cv2.resize(image[y1:y2, x1:x2], (D, D)), coords, name))
    items = dedup(items, False, 0.85)
    items = dedup(items, True, 0.95)
    return items


def colour_mask(window, colour):
    """DET/source.py:63-89."""
    hsv = cv2.cvtColor(cv2.resize(window, (25, 25)), cv2.COLOR_BGR2HSV)
    if colour == 'r':
        low = cv2.inRange(hsv, np.array([0, 50, 10]), np.array([10, 255, 255]))
        high = cv2.inRange(hsv, np.array([160, 50, 10]), np.array([179, 255, 255]))
        return cv2.add(low, high)
    return cv2.inRange(hsv, np.array([90, 70, 10], np.uint8), np.array([128, 255, 255], np.uint8))


def overlap_score(product, template):
    """DET/source.py:545-567 -- the per-pixel Python loop is the reference's implementation and its main CPU cost."""
    tp = fp = fn = tn = 0
    if product.shape != template.shape:
        return None
    unit = template // 255
    for row_p, row_t in zip(product, unit):
        for p, t in zip(row_p, row_t):
            if p == 1 and t == 1:
                tp += 1
            elif p == 1 and t == 0:
                fp += 1
            elif p == 0 and t == 1:
                fn += 1
            else:
                tn += 1
    size = product.shape[0] * product.shape[1]
    if size + size * 0.01 >= tn >= size - size * 0.01:
        return 0
    return round((2 * tp) / ((2 * tp) + fp + fn), 2)


def best_template(mask, templates):
    """DET/source.py:248-261."""
    best, best_id = -math.inf, ''
    for tmpl, name in templates:
        s = overlap_score(mask * tmpl, tmpl)
        if s > best:
            best, best_id = s, SIGNALLIST.index(name) + 1
    return best, best_id


def classify_window(item, red_templates, blue_templates, tolerance=0.55):
    """DET/source.py:229-245."""
    s_red, id_red = best_template(colour_mask(item[0], 'r'), red_templates)
    s_blue, id_blue = best_template(colour_mask(item[0], 'b'), blue_templates)
    x1, y1, x2, y2 = item[1]
    if s_red > s_blue:
        return (item[2], x1, y1, x2, y2, id_red, s_red) if s_red > tolerance else None
    return (item[2], x1, y1, x2, y2, id_blue, s_blue) if s_blue > tolerance else None


def detect_frame(image, boxes, name, red_templates, blue_templates):
    """DET/source.py:116-131 followed by :708-716 for one frame -> list of result tuples."""
    out = []
    for item in frame_windows(image, boxes, name):
        r = classify_window(item, red_templates, blue_templates)
        if r is not None:
            out.append(r)
    return out


def templates_as_lists(red6, blue6):
    return ([(np.asarray(m, np.uint8), n) for m, n in zip(red6, SIGNALLIST)],
            [(np.asarray(m, np.uint8), n) for m, n in zip(blue6, SIGNALLIST)])


# ---- multi-core runner used by bench.py ------------------------------------------------------------------------------
_W = {}


def _worker_init(red6, blue6):
    cv2.setNumThreads(1)
    _W["t"] = templates_as_lists(red6, blue6)


def _worker_run(job):
    frames, boxes_list = job
    red, blue = _W["t"]
    n = 0
    for f, (img, boxes) in enumerate(zip(frames, boxes_list)):
        n += len(detect_frame(img, boxes, str(f), red, blue))
    return n


def run_parallel(pool, frames, boxes, offsets, nworkers):
    """Split the frames into `nworkers` contiguous ranges and run detect_frame on all of them; returns #detections."""
    F = len(frames)
    jobs = []
    for w in range(nworkers):
        a, b = F * w // nworkers, F * (w + 1) // nworkers
        if b > a:
            jobs.append((frames[a:b], [boxes[offsets[f]:offsets[f + 1]] for f in range(a, b)]))
    return sum(pool.map(_worker_run, jobs))
