"""GPU (-m gpu): the drop-in mirrors of the reference's own functions (source_det / source_rec), called the way the
reference's drivers call them (DET/source.py:611-853 test(), REC/source.py:646-809 testValidation()), against outputs of
the unmodified reference stored under tests/golden/."""
import types

import numpy as np
import pytest

from conftest import STORED

pytestmark = pytest.mark.gpu


def _templates_as_reference(templates, names):
    red6, blue6 = templates
    return [(red6[k], names[k]) for k in range(6)], [(blue6[k], names[k]) for k in range(6)]


def test_detection_functions_like_the_reference_driver(tsd, templates, det_frames, frames3):
    """detectSignsOnDirectory's per-frame body + the mask-correlation filter of test() (DET:101,708-716) + createDetectionsStrings
    (DET:501-508) on the three stored real frames: lists of tuples with the reference's exact conventions (Python ints, float
    score rounded to 2 dp, insertion order), and the resultado.txt lines of those files byte for byte."""
    import cv2
    S = tsd.source_det
    red, blue = _templates_as_reference(templates, S.SIGNALLIST)
    mser = cv2.MSER_create(delta=7, min_area=200, max_area=2000, max_variation=0.15)     # DET/main.py:28 default detector
    expected_lines = {}
    import os
    for ln in open(os.path.join(os.path.dirname(__file__), "golden", "det_resultado150.txt")).read().splitlines():
        expected_lines.setdefault(ln.split(";")[0], []).append(ln)
    for k in STORED:
        img, file = frames3[k], k + ".jpg"
        boxes = S.proposals(img, mser)
        assert np.array_equal(boxes, det_frames[k + "_boxes"])                            # cv2 path unchanged
        dets = S.MSERTrafficSignDetector(img, mser, file)
        assert [d[1] for d in dets] == [tuple(int(v) for v in c) for c in det_frames[k + "_p2_coords"]]
        assert all(isinstance(v, int) for d in dets for v in d[1]) and all(d[2] == file for d in dets)
        assert np.array_equal(np.stack([d[0] for d in dets]), det_frames[k + "_p2_windows"])
        out = [S.detectionsMaskCorrelation(d, red, blue, 0.55) for d in dets]
        out = [o for o in out if o is not None]
        assert [o[1:5] for o in out] == [tuple(int(v) for v in c) for c in det_frames[k + "_det_coords"]]
        assert [o[5] for o in out] == [int(v) for v in det_frames[k + "_det_ids"]]
        assert [o[6] for o in out] == [float(v) for v in det_frames[k + "_det_scores"]]
        assert S.createDetectionsStrings(out) == expected_lines.get(file, [])
        # the batched entry point gives the same tuples in one library call
        b_out, _ = S.detectBatch(img[None], boxes, np.array([0, len(boxes)], np.int32), [file], red, blue)
        assert b_out == out


def test_detection_helpers_like_the_reference(tsd, templates, det_frames, oracle):
    """makeWindowBiggerOrDiscardFakeDetections, cleanDuplicatedDetections, calculateHistAndNormalize, getColorMaskRedOrBlue,
    getSimilarSignalType, calculateScoreBetweenMatrixs with the reference's argument and return conventions."""
    S = tsd.source_det
    red, blue = _templates_as_reference(templates, S.SIGNALLIST)
    k = STORED[0]
    boxes, coords, valid = det_frames[k + "_boxes"], det_frames[k + "_coords"], det_frames[k + "_valid"]
    for i in range(0, len(boxes), 7):
        r = S.makeWindowBiggerOrDiscardFakeDetections(boxes[i], 1.30)
        assert (r is None) == (not valid[i])
        if r is not None:
            assert r == tuple(int(v) for v in coords[i]) and all(isinstance(v, int) for v in r)
    items = [(w, tuple(int(v) for v in c), "f.jpg") for w, c in zip(det_frames[k + "_windows"], coords[valid])]
    p1 = S.cleanDuplicatedDetections(items, False, 0.85)
    assert [d[1] for d in p1] == [tuple(int(v) for v in c) for c in det_frames[k + "_p1_coords"]]
    p2 = S.cleanDuplicatedDetections(p1, True, 0.95)
    assert [d[1] for d in p2] == [tuple(int(v) for v in c) for c in det_frames[k + "_p2_coords"]]
    assert S.cleanDuplicatedDetections([], False, 0.85) == []
    for i, w in enumerate(det_frames[k + "_windows"]):
        assert np.array_equal(S.calculateHistAndNormalize(w), det_frames[k + "_hists"][i])
    for i, w in enumerate(det_frames[k + "_p2_windows"]):
        mr, mb = S.getColorMaskRedOrBlue(w, 'r'), S.getColorMaskRedOrBlue(w, 'b')
        assert np.array_equal(mr, det_frames[k + "_red"][i]) and np.array_equal(mb, det_frames[k + "_blue"][i])
        assert S.getColorMaskRedOrBlue(w, 'x') is None                                    # DET:63-89 falls through
        sc_r, id_r = S.getSimilarSignalType(mr, red)
        sc_b, id_b = S.getSimilarSignalType(mb, blue)
        exp = det_frames[k + "_scores"][i]                                                # [2][6] per-template scores of the reference
        assert sc_r == float(max(exp[0])) and id_r == int(np.argmax(exp[0])) + 1
        assert sc_b == float(max(exp[1])) and id_b == int(np.argmax(exp[1])) + 1
        for t in range(6):
            m1 = mr * red[t][0]                                                           # uint8 wrap-around product, DET:254
            got = S.calculateScoreBetweenMatrixs(m1, red[t][0])
            assert got == (0 if int((red[t][0] == 255).sum()) <= 6 else float(exp[0][t]))
    assert S.calculateScoreBetweenMatrixs(np.zeros((25, 25), np.uint8), np.zeros((24, 25), np.uint8)) is None


def test_recognition_functions_like_the_reference_driver(tsd, rec_golden, rec_frames, frames3):
    """computeDescriptors / calculateDescriptors / predictProbability (REC:507-521,619-624) with stand-ins for the fitted
    scikit-learn objects that carry exactly the attributes the reference's code path reads."""
    import cv2
    R = tsd.source_rec
    g = rec_golden
    n = 200
    hog_desc = (None, 'HOG')
    d0 = R.computeDescriptors(g["gray"][0], hog_desc)
    assert d0.dtype == np.float32 and d0.shape == (324,)
    assert np.max(np.abs(d0 - g["hog"][0]) / np.maximum(np.abs(g["hog"][0]), 1e-2)) < 1e-4
    assert np.array_equal(R.computeDescriptors(g["gray"][0], (None, 'GRAY')), g["gray"][0].ravel())
    data = {t: [] for t in range(7)}
    for i in range(n):
        data[int(g["true"][i])].append((g["gray"][i], (0, 0, 32, 32), "f.jpg", int(g["true"][i])))
    desc = R.calculateDescriptors(data, hog_desc)
    flat = [d for t in range(7) for d in desc[t]]
    order = [i for t in range(7) for i in range(n) if int(g["true"][i]) == t]
    lda = [types.SimpleNamespace(coef_=g["lda_W"][:, c][None, :], intercept_=np.array([g["lda_b"][c]])) for c in range(6)]
    pred, true = R.predictProbability((lda, 'LDABAYES'), None, flat, 0.5)
    assert pred == [int(g["pred_lda"][i]) for i in order] and true == [int(g["true"][i]) for i in order]
    knn = types.SimpleNamespace(_fit_X=g["knn_Ztrain"], classes_=np.arange(7), _y=g["knn_ytrain"].astype(np.int64), n_neighbors=4)
    reducer = (types.SimpleNamespace(xbar_=g["knn_xbar"], scalings_=g["knn_scalings"]),)
    pred_k, true_k = R.predictProbability((knn, 'KNN'), reducer, flat, 0.5)
    assert [int(v) for v in pred_k] == [int(g["pred_knn"][i]) for i in order]
    # window extraction of the recognition flavour (x1.15, 32x32, 4-tuples with the label slot) on a real frame
    k = STORED[0]
    mser = cv2.MSER_create(delta=7, min_area=200, max_area=2000, max_variation=1.0)      # REC/main.py:44 default detector
    dets = R.MSERTrafficSignDetector(frames3[k], mser, k + ".jpg")
    assert all(len(d) == 4 and d[3] == 0 and d[0].shape == (32, 32, 3) for d in dets)
    assert [d[1] for d in dets] == [tuple(int(v) for v in c) for c in rec_frames[k + "_coords"]]
    assert np.array_equal(np.stack([d[0] for d in dets]), rec_frames[k + "_windows"])
    gray = R.windowsToGray(dets)
    assert np.array_equal(np.stack([d[0] for d in gray]), rec_frames[k + "_gray"])


def test_preprocessing_like_the_reference(tsd, oracle, frames3):
    """grayAndEnhanceContrast on the GPU = the reference's stored outputs on the real frames (sha1 + crop), the
    oracle on odd sizes / 4K / batches, and therefore the same cv2.MSER boxes (checked in the driver test above)."""
    import hashlib
    import os
    S = tsd.source_det
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "det_pre.npz"))
    for k in STORED:
        r = S.grayAndEnhanceContrast(frames3[k])
        assert r.dtype == np.uint8 and r.shape == (800, 1360)
        assert hashlib.sha1(np.ascontiguousarray(r).tobytes()).digest() == g[k + "_sha1"].tobytes()
        assert np.array_equal(r[300:364, 600:728], g[k + "_crop"])
    rng = np.random.default_rng(4)
    ctx = S.context()
    for (H, W, F) in ((97, 131, 3), (480, 641, 2), (481, 640, 1), (64, 64, 5), (9, 17, 1), (100, 7, 2), (2160, 3840, 1)):
        imgs = np.clip(rng.normal(120, 40, (F, H, W, 3)), 0, 255).astype(np.uint8)
        got = ctx.preprocess(imgs)
        for f in range(F):
            assert np.array_equal(got[f], oracle.preprocess(imgs[f])), (H, W, f)
    assert np.array_equal(S.gammaCorrection(np.arange(256, dtype=np.uint8), 2), oracle.gamma_table(2))


def test_training_window_extraction_like_the_reference(tsd, rec_frames, frames3, oracle, tmp_path, monkeypatch):
    """the producer of MSERTrain.val (REC:380-398) through the GPU path.  The three stored frames stand in for the
    train frames: every cache entry (grey 32x32 pixels, coords, file, label slot 0, in order) equals what the reference's
    MSERTrafficSignDetector + BGR2GRAY gave for that frame (rec_frames.npz); the pickle has the reference's layout; negatives
    follow the IoU <= 0.5 rule; positives (REC:247-258) equal cv2 on the grey frame."""
    import pickle
    import cv2
    R = tsd.source_rec
    mser = cv2.MSER_create(delta=7, min_area=200, max_area=2000, max_variation=1.0)
    train = {k + ".jpg": frames3[k] for k in STORED}
    gt = []
    for k in STORED:                                                  # "ground truth": two survivors of each frame + an odd box
        c = rec_frames[k + "_coords"]
        gt += [(k + ".jpg", int(c[0][0]), int(c[0][1]), int(c[0][2]), int(c[0][3]), 1),
               (k + ".jpg", int(c[-1][0]), int(c[-1][1]), int(c[-1][2]), int(c[-1][3]), 3), (k + ".jpg", 100, 200, 163, 251, 2)]
    pos = R.orderCroppedImagesByImageFile(train, gt)
    for k in STORED:
        g = cv2.cvtColor(frames3[k], cv2.COLOR_BGR2GRAY)
        assert len(pos[k + ".jpg"]) == 3
        for (win, coords, file, label), r in zip(pos[k + ".jpg"], [t for t in gt if t[0] == k + ".jpg"]):
            assert coords == r[1:5] and file == r[0] and label == r[5]
            x1, y1, x2, y2 = coords
            assert np.array_equal(win, cv2.resize(g[y1:y2, x1:x2], (32, 32)))
    monkeypatch.chdir(tmp_path)                                       # the cache is cwd-relative (REC:381)
    neg = R.calculateNegativeTrainResults(train, pos, mser)
    cache = pickle.load(open(tmp_path / "MSERTrain.val", "rb"))
    assert list(cache.keys()) == list(train.keys())
    for k in STORED:
        ent = cache[k + ".jpg"]
        assert [e[1] for e in ent] == [tuple(int(v) for v in c) for c in rec_frames[k + "_coords"]]
        assert all(isinstance(v, int) for e in ent for v in e[1]) and all(e[2] == k + ".jpg" and e[3] == 0 for e in ent)
        assert np.array_equal(np.stack([e[0] for e in ent]), rec_frames[k + "_gray"])
        # the two ground-truth boxes that ARE survivors overlap themselves (IoU 1 > 0.5): not negatives
        exp_neg = [e for e in ent if max(R.intersectionOverUnion(e[1], p[1]) for p in pos[k + ".jpg"]) <= 0.5]
        assert [e[1] for e in neg[k + ".jpg"]] == [e[1] for e in exp_neg] and len(exp_neg) <= len(ent) - 2
    neg2 = R.calculateNegativeTrainResults(train, pos, None)          # second call: served from the cache, no MSER needed
    assert [[e[1] for e in neg2[n]] for n in train] == [[e[1] for e in neg[n]] for n in train]



def test_evaluators_like_the_reference(tsd, eval_golden, tmp_path):
    """evaluate.generateStatistics (DET:267-450) and evaluate.precision_recall_curve / draw_PR_fast / VOColdap
    (REC/evaluar_resultados.py:199-307), matching loops on the GPU, against the outputs of the reference's own functions on its
    own ground truth and result files."""
    E, g = tsd.evaluate, eval_golden
    gt_path = tmp_path / "gt.txt"
    gt_path.write_bytes(bytes(g["gt_txt"]))
    dets = []
    for ln in bytes(g["own_txt"]).decode().split():
        f, x1, y1, x2, y2, t, sc = ln.split(";")
        dets.append((f, int(x1), int(y1), int(x2), int(y2), int(t), float(sc)))
    files = [str(f) for f in g["files"]]
    number = [(f, sum(1 for d in dets if d[0] == f)) for f in files]
    per_file, by_type, tc, ti, tn, te = E.generateStatistics(dets, str(gt_path), number)
    assert [pf[0] for pf in per_file] == files
    assert np.array_equal(np.array([[r[1:] for r in pf[1]] for pf in per_file]), g["stat_per_file"])
    assert [n for n, _ in by_type] == E.SIGNALLIST and np.array_equal(np.array([v for _, v in by_type]), g["stat_by_type"])
    assert [tc, ti, tn, te] == g["stat_totals"].tolist()
    assert all(pf[2:] == tuple(int(v) for v in np.array([r[1:] for r in pf[1]]).sum(0)) for pf in per_file)
    # precision / recall
    gt_jpg = tmp_path / "gt_jpg.txt"
    gt_jpg.write_text(bytes(g["gt_txt"]).decode().replace(".ppm", ".jpg"))
    _, gt_bb = E.load_results_file(str(gt_jpg))
    for tag in ("own", "p1", "p2"):
        p = tmp_path / (tag + ".txt")
        p.write_bytes(bytes(g[tag + "_txt"]))
        _, det_bb = E.load_results_file(str(p))
        tp, fp, thr, tot = E.precision_recall_curve(gt_bb, det_bb, show=False, ovr=0.5)
        assert tot == int(g[tag + "_tot"]) and np.array_equal(tp, g[tag + "_tp"]) and np.array_equal(fp, g[tag + "_fp"]) and np.array_equal(thr, g[tag + "_thr"])
        rec, prec, ap = E.draw_PR_fast(tp, fp, tot, show=False)
        assert np.array_equal(rec, g[tag + "_rec"]) and np.array_equal(prec, g[tag + "_prec"], equal_nan=True)
        assert ap == float(g[tag + "_ap"]) and E.VOColdap(rec, prec) == float(g[tag + "_ap11"])
    _, gt_asis = E.load_results_file(str(gt_path))
    p = tmp_path / "own.txt"
    _, det_bb = E.load_results_file(str(p))
    tp, fp, _, tot = E.precision_recall_curve(gt_asis, det_bb)
    assert tp.sum() == 0 and np.array_equal(fp, g["asis_fp"]) and tot == int(g["asis_tot"])


def test_evaluators_random_vs_oracle(tsd):
    """The GPU matching loops against oracle/evaluate.py on random boxes: several detections per ground truth, identical ground-truth
    rows, score ties, overlap ties, ignore regions (class -1), images without ground truth, empty inputs."""
    from oracle import evaluate as O
    E = tsd.evaluate
    rng = np.random.default_rng(29)
    for trial in range(8):
        nimg = 20
        gt, det = {}, {}
        for k in range(nimg):
            name = "%05d.jpg" % k
            boxes = []
            for _ in range(int(rng.integers(0, 6))):
                x, y, s = int(rng.integers(0, 1200)), int(rng.integers(0, 700)), int(rng.integers(16, 90))
                boxes.append((x, y, x + s, y + s, int(rng.choice([1, 2, 3, 4, 5, 6, -1]))))
            if boxes and rng.random() < 0.3:
                boxes.append(boxes[0])
            if boxes or rng.random() < 0.5:
                gt[name] = boxes
            rows = []
            for _ in range(int(rng.integers(0, 10))):
                if boxes and rng.random() < 0.7:
                    b = boxes[int(rng.integers(0, len(boxes)))]
                    j = rng.integers(-12, 13, 4)
                    rows.append((name, b[0] + int(j[0]), b[1] + int(j[1]), b[2] + int(j[2]), b[3] + int(j[3]), float(rng.choice([0.5, 0.6, 0.7, 0.8, 0.9]))))
                else:
                    x, y, s = int(rng.integers(0, 1200)), int(rng.integers(0, 700)), int(rng.integers(16, 90))
                    rows.append((name, x, y, x + s, y + s, float(rng.choice([0.5, 0.6, 0.7]))))
            if rows:
                det[name] = rows
        if trial == 7:
            det = {}
        gt_bb = {k: [E.BoundingBox(b[0], b[1], b[2], b[3], class_id=b[4], img_idx=k) for b in v] for k, v in gt.items()}
        det_bb = {k: [E.BoundingBox(r[1], r[2], r[3], r[4], class_id=1, score=r[5], img_idx=k) for r in v] for k, v in det.items()}
        tp, fp, thr, tot = E.precision_recall_curve(gt_bb, det_bb, ovr=0.5)
        det_list = [r for k in sorted(det) for r in det[k]]
        otp, ofp, othr, otot = O.pr_flags(gt, det_list)
        assert tot == otot and np.array_equal(tp, otp) and np.array_equal(fp, ofp) and np.array_equal(thr, othr)
        # similarity-based matching of generateStatistics
        dets = [(int(r[0][:5]), r[1], r[2], r[3], r[4], int(rng.integers(0, 6))) for r in det_list]
        gts = sorted(((int(k[:5]), b[0], b[1], b[2], b[3], int(rng.integers(0, 6))) for k in gt for b in gt[k]), key=lambda r: r[0])
        off = np.zeros(nimg + 1, np.int32)
        for g in gts:
            off[g[0] + 1] += 1
        off = np.cumsum(off).astype(np.int32)
        st, mt, tally = E.context().match_detections(np.array(dets, np.int32).reshape(-1, 6), np.array(gts, np.int32).reshape(-1, 6), off)
        ost, omt, otally = O.det_statistics(dets, gts, nimg)
        assert st.tolist() == ost and mt.tolist() == omt and np.array_equal(tally, otally)
