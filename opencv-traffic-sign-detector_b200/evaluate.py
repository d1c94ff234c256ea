"""The reference's two evaluators with their matching loops on the GPU (reporting only).

  * generateStatistics(detections, realResultsFilePath, numberDetections)   `Deteción de Objetos/source.py:267-330`
    -- same arguments, same returned structure; the per (file, type) loops of getCorrectsAndWrongByTypeOnFile /
    checkIfDetectionByTypeOnFileIsCorrectIncorrectDuplicated (`:401-450`) are one tsd_match_detections call.
  * BoundingBox, compute_class_index, load_results_file, precision_recall_curve, VOCap, VOColdap, draw_PR_fast
    `Reconocimiento de Objetos/evaluar_resultados.py:16-307` -- precision_recall_curve's loop is one tsd_match_iou call; the
    plotting / image display of the script is not mirrored.
"""
import csv
import os

import numpy as np

from . import engine

SIGNALLIST = ['prohibicion', 'peligro', 'stop', 'direccionProhibida', 'cedaPaso', 'direccionObligatoria']       # DET/constants.py:1
_GROUPS = (('00', '01', '02', '03', '04', '05', '07', '08', '09', '10', '15', '16'),                             # DET/constants.py:2-7
           ('11', '18', '19', '20', '21', '22', '23', '24', '25', '26', '27', '28', '29', '30', '31'), ('14',), ('17',), ('13',), ('38',))
_ctx = None


def context():
    """Process-wide context of the evaluators on cuda:0 (or $TSD_DEVICE, like source_det / source_rec)."""
    global _ctx
    if _ctx is None:
        _ctx = engine.Context(device=int(os.environ.get("TSD_DEVICE", "0")), flavour="det")
    return _ctx


def calculateSignType(signType):
    """DET:518-540: GTSDB class string -> 1..6, or None for a class outside the six groups."""
    if int(signType) < 10:
        signType = '0' + signType
    for k, group in enumerate(_GROUPS):
        if signType in group:
            return k + 1
    return None


def _bucket(t):
    return t - 1 if t in (1, 2, 3, 4, 5) else 5              # DET:371-386: the final `else` takes everything that is not 1..5


def _stem(name):
    return name.split(".", 1)[0]


def generateStatistics(detections, realResultsFilePath, numberDetections):
    """DET:267-330.  detections: [(file, x1, y1, x2, y2, type, score)], numberDetections: [(file, n)]."""
    realResults = []
    with open(realResultsFilePath, "r") as fh:
        for line in fh:
            filename, x1, y1, x2, y2, signType = line.rstrip().split(';')
            realResults.append((filename, int(x1), int(y1), int(x2), int(y2), calculateSignType(signType)))
    files = [fn[0] for fn in numberDetections]
    first = {}
    for k, name in enumerate(files):
        first.setdefault(_stem(name), k)
    det_rows = [(first[_stem(d[0])], d[1], d[2], d[3], d[4], _bucket(d[5])) for d in detections if _stem(d[0]) in first]
    gt_rows = sorted(((first[_stem(r[0])], r[1], r[2], r[3], r[4], _bucket(r[5])) for r in realResults if _stem(r[0]) in first),
                     key=lambda r: r[0])                      # stable: file order inside a frame is kept
    nf = len(files)
    gt_off = np.zeros(nf + 1, np.int32)
    for r in gt_rows:
        gt_off[r[0] + 1] += 1
    gt_off = np.cumsum(gt_off).astype(np.int32)
    _, _, tally = context().match_detections(np.asarray(det_rows, np.int32).reshape(-1, 6), np.asarray(gt_rows, np.int32).reshape(-1, 6), gt_off)
    detectionsPerFileByType = []
    totals = np.zeros((6, 4), np.int64)
    for k, name in enumerate(files):
        t = tally[first[_stem(name)]]                         # (a file listed twice is counted twice, as the reference's loop does)
        perType = [(SIGNALLIST[b], int(t[b, 0]), int(t[b, 1]), int(t[b, 2]), int(t[b, 3])) for b in range(6)]
        totals += t
        detectionsPerFileByType.append((name, perType, int(t[:, 0].sum()), int(t[:, 1].sum()), int(t[:, 2].sum()), int(t[:, 3].sum())))
    totalDetectionsByType = [(SIGNALLIST[b], tuple(int(v) for v in totals[b])) for b in range(6)]
    return (detectionsPerFileByType, totalDetectionsByType, int(totals[:, 0].sum()), int(totals[:, 1].sum()), int(totals[:, 2].sum()),
            int(totals[:, 3].sum()))


# ---- evaluar_resultados.py ------------------------------------------------------------------------------------------------
class BoundingBox:
    """evaluar_resultados.py:16-50 (without the drawing helpers)."""

    def __init__(self, left, top, right, bottom, class_id=-1, score=1.0, img_idx=-1):
        self.left, self.top, self.right, self.bottom = int(left), int(top), int(right), int(bottom)
        self.class_id = int(class_id)
        self.score = score
        self.img_idx = str(img_idx)

    def area(self):
        return (self.right - self.left + 1) * (self.bottom - self.top + 1)

    def __repr__(self):
        return str((self.img_idx, self.left, self.top, self.right, self.bottom, self.class_id, self.score))


def compute_class_index(number):
    """evaluar_resultados.py:125-145: GTSDB class number -> 1..6, -1 = "ignore this sign"."""
    for k, group in enumerate(_GROUPS):
        if number in [int(v) for v in group]:
            return k + 1
    return -1


def load_results_file(file_name, test_path="", load_images=False):
    """evaluar_resultados.py:148-193 (boxes only: images are not loaded).  7 columns = detections with a score, else ground truth."""
    bboxes = dict()
    with open(file_name, 'r') as fh:
        for row in csv.reader(fh, delimiter=';', quotechar='#'):
            if len(row) == 7:
                bb = BoundingBox(row[1], row[2], row[3], row[4], class_id=str(row[5]), score=float(row[6]), img_idx=str(row[0]))
            else:
                bb = BoundingBox(row[1], row[2], row[3], row[4], class_id=compute_class_index(int(row[5])), score=float(1.0), img_idx=str(row[0]))
            bboxes.setdefault(row[0], []).append(bb)
    return dict(), bboxes


def precision_recall_curve(gt_dbboxes, det_dbboxes, show=False, ovr=0.5, images_dict=None):
    """evaluar_resultados.py:199-262 -> tp, fp, thr (float64 arrays in score-descending order), tot."""
    images = sorted(k for k, v in gt_dbboxes.items() if v)
    index = {k: i for i, k in enumerate(images)}
    tot = sum(1 for k in images for b in gt_dbboxes[k] if b.class_id != -1)
    det_list = []
    for _, boxes in sorted(det_dbboxes.items(), key=lambda x: x[0]):
        det_list = det_list + boxes
    det_list = sorted(det_list, reverse=True, key=lambda x: x.score)
    n = len(det_list)
    thr = np.array([b.score for b in det_list], np.float64) if n else np.zeros(0)
    tp, fp = np.zeros(n), np.zeros(n)
    # detections on an image without ground truth are false positives (:246-247); the others go to the GPU grouped by image
    where = np.array([index.get(b.img_idx, -1) for b in det_list], np.int64) if n else np.zeros(0, np.int64)
    fp[where < 0] = 1
    sel = np.nonzero(where >= 0)[0]
    if len(sel):
        order = sel[np.argsort(where[sel], kind="stable")]    # by image, list order inside an image
        det_rows = np.array([(det_list[k].left, det_list[k].top, det_list[k].right, det_list[k].bottom, j) for j, k in enumerate(order)], np.int32)
        det_off = np.zeros(len(images) + 1, np.int32)
        np.add.at(det_off, where[order] + 1, 1)
        det_off = np.cumsum(det_off).astype(np.int32)
        gt_rows = np.array([(b.left, b.top, b.right, b.bottom, b.class_id) for k in images for b in gt_dbboxes[k]], np.int32)
        gt_off = np.cumsum([0] + [len(gt_dbboxes[k]) for k in images]).astype(np.int32)
        t, f = context().match_iou(det_rows, det_off, gt_rows, gt_off, ovr)
        tp[order], fp[order] = t, f
    return tp, fp, thr, tot


def VOCap(rec, prec):
    """evaluar_resultados.py:265-272."""
    mrec = np.concatenate(([0], rec, [1]))
    mpre = np.concatenate(([0], prec, [0]))
    for k in range(len(mpre) - 2, 0, -1):                   # monotone envelope from the right (Python's max: a NaN neighbour is skipped)
        mpre[k] = max(mpre[k], mpre[k + 1])
    i = np.where(mrec[1:] != mrec[0:-1])[0] + 1
    return np.sum((mrec[i] - mrec[i - 1]) * mpre[i])


def VOColdap(rec, prec):
    """evaluar_resultados.py:275-285."""
    rec, prec = np.array(rec), np.array(prec)
    ap = 0.0
    for t in np.linspace(0, 1, 11):
        pr = prec[rec >= t]
        ap = ap + (np.max(pr) if pr.size else 0) / 11.0
    return ap


def draw_PR_fast(tp, fp, tot, show=False, col="g"):
    """evaluar_resultados.py:288-307 without the plot -> rec, prec, ap (VOCap)."""
    tp, fp = np.cumsum(tp), np.cumsum(fp)
    rec = tp / tot
    prec = tp / (fp + tp)
    return rec, prec, VOCap(rec, prec)
