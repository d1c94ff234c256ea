"""TEST INFRASTRUCTURE (oracle): CPU restatement of the two evaluators' matching loops, plain Python.

Only tests/ may import this.  Pinned on the reference's own functions run in the build container: tests/golden/eval_golden.npz
(tests/golden/make_golden.py eval).

  * det_statistics      = generateStatistics' per (file, type) matching, `Deteción de Objetos/source.py:267-450`
  * pr_flags / voc_ap / voc_old_ap / pr_curve = precision_recall_curve, VOCap, VOColdap, draw_PR_fast,
    `Reconocimiento de Objetos/evaluar_resultados.py:199-320`
"""
import math

import numpy as np


def eucl_similarity(xa, ya, xb, yb):
    """EuclDSimilarity (DET:459-462), numpy scalar calls as in the reference."""
    d = np.linalg.norm(np.array((xa, ya)) - np.array((xb, yb)))
    if d == 0:
        return 1
    return 1 / (1 + np.power(np.e, (((0.154 * np.power(d, 1.2)) - 31.8) / (0.2 * d))))


def type_bucket(t):
    """appendResultsByTypeOnFile (DET:371-386): 1..5 -> lists 0..4, anything else (6, or None for an unlisted GTSDB class) -> the sixth."""
    return t - 1 if t in (1, 2, 3, 4, 5) else 5


def det_statistics(det, gt, nframes, tol=0.85):
    """det / gt: sequences of (frame, x1, y1, x2, y2, bucket).  Returns status [ndet], match [ndet] (index into gt or -1) and the
    tally [nframes][6][4] = (correct, incorrect, not detected, expected) of getCorrectsAndWrongByTypeOnFile (DET:401-422)."""
    status, match = [0] * len(det), [-1] * len(det)
    tally = np.zeros((nframes, 6, 4), np.int64)
    for f in range(nframes):
        for b in range(6):
            di = [i for i, d in enumerate(det) if d[0] == f and d[5] == b]
            gi = [i for i, g in enumerate(gt) if g[0] == f and g[5] == b]
            checked = set()
            for i in di:
                d = det[i]
                best, sel = -math.inf, None                  # DET:427-438: first strict maximum
                for j in gi:
                    g = gt[j]
                    s = np.sqrt(eucl_similarity(d[1], d[2], g[1], g[2]) * eucl_similarity(d[3], d[4], g[3], g[4]))
                    if s > best:
                        best, sel = s, j
                if best > tol:                               # DET:440-442 (the "duplicated" branch :443 cannot be reached)
                    status[i], match[i] = 1, sel
                    checked.add(tuple(gt[sel][1:]))          # a set of ground-truth tuples: identical rows count once
            ncorrect = sum(status[i] for i in di)
            tally[f, b] = (ncorrect, len(di) - ncorrect, len(gi) - len(checked) if di else len(gi), len(gi))
    return status, match, tally


def overlap(gt_box, dt_box, ignore):
    """bboxes_overlap (evaluar_resultados.py:53-89); boxes = (left, top, right, bottom)."""
    w = min(dt_box[2], gt_box[2]) - max(dt_box[0], gt_box[0])
    if w <= 0:
        return 0.0
    h = min(dt_box[3], gt_box[3]) - max(dt_box[1], gt_box[1])
    if h <= 0:
        return 0.0
    i = w * h
    area = lambda b: (b[2] - b[0] + 1) * (b[3] - b[1] + 1)
    u = area(dt_box) if ignore else area(dt_box) + area(gt_box) - i
    return i / u


def pr_flags(gt_by_image, det_list, ovr=0.5):
    """precision_recall_curve (evaluar_resultados.py:199-262).  gt_by_image: {image: [(l, t, r, b, cls)]}; det_list: [(image, l, t, r, b,
    score)] in file order (images sorted, as the reference concatenates them).  -> tp, fp, thr (score-descending, stable), tot."""
    tot = sum(1 for boxes in gt_by_image.values() for g in boxes if g[4] != -1)
    used = {im: [False] * len(boxes) for im, boxes in gt_by_image.items() if boxes}
    order = sorted(range(len(det_list)), key=lambda k: det_list[k][5], reverse=True)      # stable, like sorted(..., reverse=True)
    tp, fp, thr = np.zeros(len(order)), np.zeros(len(order)), np.zeros(len(order))
    for idx, k in enumerate(order):
        im, box, sc = det_list[k][0], det_list[k][1:5], det_list[k][5]
        maxovr, sel = 0, 0
        if im in used:
            for ir, g in enumerate(gt_by_image[im]):
                c = overlap(g[:4], box, g[4] == -1)
                if c >= maxovr:                              # :239 the LAST maximum
                    maxovr, sel = c, ir
        if maxovr > ovr:
            if gt_by_image[im][sel][4] != -1:
                if not used[im][sel]:
                    tp[idx] = 1
                    used[im][sel] = True
                else:
                    fp[idx] = 1
        else:
            fp[idx] = 1
        thr[idx] = sc
    return tp, fp, thr, tot


def voc_ap(rec, prec):
    """VOCap (evaluar_resultados.py:265-272)."""
    mrec = np.concatenate(([0], rec, [1]))
    mpre = np.concatenate(([0], prec, [0]))
    for i in range(len(mpre) - 2, 0, -1):
        mpre[i] = max(mpre[i], mpre[i + 1])
    i = np.where(mrec[1:] != mrec[0:-1])[0] + 1
    return np.sum((mrec[i] - mrec[i - 1]) * mpre[i])


def voc_old_ap(rec, prec):
    """VOColdap (evaluar_resultados.py:275-285): 11-point interpolation."""
    rec, prec = np.array(rec), np.array(prec)
    ap = 0.0
    for t in np.linspace(0, 1, 11):
        pr = prec[rec >= t]
        ap = ap + (np.max(pr) if pr.size else 0) / 11.0
    return ap


def pr_curve(tp, fp, tot):
    """draw_PR_fast without the plot (evaluar_resultados.py:288-307) -> rec, prec, ap (VOCap), ap11 (VOColdap)."""
    tp, fp = np.cumsum(tp), np.cumsum(fp)
    rec, prec = tp / tot, tp / (fp + tp)
    return rec, prec, voc_ap(rec, prec), voc_old_ap(rec, prec)
