"""Drop-in mirror of the hot-path functions of "Reconocimiento de Objetos/source.py" (REC).

Window extraction (x1.15, 32x32 -- REC:47-64), grey conversion (REC:388), HOG (REC:517-521) and the LDA / KNN
decisions (REC:565-641, 592-596) run on the GPU.  Fitting stays scikit-learn on the host (REC:526-616, out of scope);
its fitted objects are read for their weights only.
"""
import os

import numpy as np

from . import engine
from . import source_det as _det
from ._capi import TsdError

_ctx = None
_lda_key = None
_knn_key = None


def context():
    """Process-wide recognition-flavour context (x1.15, 32x32)."""
    global _ctx
    if _ctx is None:
        _ctx = engine.Context(device=int(os.environ.get("TSD_DEVICE", "0")), flavour="rec")
    return _ctx


grayAndEnhanceContrast = _det.grayAndEnhanceContrast     # REC:67-84 is the same code as DET:135-152
proposals = _det.proposals
cropImageByCoords = _det.cropImageByCoords


def makeWindowBiggerOrDiscardFakeDetections(window, percentage):
    """REC:88-107."""
    coords, valid = context().expand_boxes(np.asarray(window, np.int32).reshape(1, 4), percentage)
    return tuple(int(v) for v in coords[0]) if valid[0] else None


def cleanDuplicatedDetections(imageDetections, isSimilarityByEuclideanDistanceON, tolerance):
    """REC:110-122 (4-tuples: the label slot of the absorbed detection is kept, REC:153)."""
    items = list(imageDetections)
    if not items:
        return []
    wins = np.stack([np.asarray(i[0], np.uint8) for i in items])
    coords = np.array([i[1] for i in items], np.int32)
    ow, oc, _ = context().dedup(wins, coords, np.array([0, len(items)], np.int32), isSimilarityByEuclideanDistanceON, tolerance)
    tail = tuple(items[0][2:])
    return [(ow[i], tuple(int(v) for v in oc[i])) + tail for i in range(len(oc))]


def MSERTrafficSignDetector(image, mser, file):
    """REC:47-64 -> list of (uint8[32,32,3], (x1,y1,x2,y2), file, 0)."""
    ctx = context()
    boxes = proposals(image, mser)
    wins, coords, _ = ctx.windows(image, boxes, np.array([0, len(boxes)], np.int32))
    dets = [(wins[i], tuple(int(v) for v in coords[i]), file, 0) for i in range(len(coords))]
    dets = cleanDuplicatedDetections(dets, False, 0.85)
    dets = cleanDuplicatedDetections(dets, True, 0.95)
    return dets


def windowsToGray(detections):
    """The per-window cv2.cvtColor(BGR2GRAY) loop of REC:386-389, batched."""
    if not detections:
        return []
    g = context().bgr2gray(np.stack([np.asarray(d[0], np.uint8) for d in detections]))
    return [(g[i],) + tuple(d[1:]) for i, d in enumerate(detections)]


# ---- descriptors ---------------------------------------------------------------------------------------------------
def computeDescriptors(image, featureDescriptor):
    """REC:517-521: 'HOG' -> float32[324] (cv2.HOGDescriptor.compute), 'GRAY' -> image.ravel()."""
    if featureDescriptor[1] == 'HOG':
        return context().hog(np.asarray(image, np.uint8)[None])[0]
    elif featureDescriptor[1] == 'GRAY':
        return np.asarray(image).ravel()


def calculateDescriptors(trainImages, featureDescriptor):
    """REC:507-514, one GPU call per class list instead of one cv2 call per window."""
    out = dict((signType, []) for signType in range(0, 7))
    for signType in trainImages.keys():
        dets = trainImages[signType]
        if not dets:
            continue
        if featureDescriptor[1] == 'HOG':
            descs = context().hog(np.stack([np.asarray(d[0], np.uint8) for d in dets]))
        else:
            descs = [np.asarray(d[0]).ravel() for d in dets]
        out[signType] = [(descs[i], d[1], d[2], d[3]) for i, d in enumerate(dets)]
    return out


# ---- classifiers -----------------------------------------------------------------------------------------------------
def extractDescriptorsAndRealSignTypes(detectionsDescriptors):
    """REC:307-313."""
    return [d[0] for d in detectionsDescriptors], [d[3] for d in detectionsDescriptors]


def _use_lda(LDAClassifiers):
    global _lda_key
    W = np.stack([np.asarray(c.coef_, np.float64)[0] for c in LDAClassifiers], 1)
    b = np.array([np.asarray(c.intercept_, np.float64)[0] for c in LDAClassifiers])
    key = (W.tobytes(), b.tobytes())
    if key != _lda_key:
        context().set_lda(W, b)
        _lda_key = key


def predictProbabilityLDAClassifiers(LDAClassifiers, detectionsDescriptors, tolerance):
    """REC:565-577 + extractBestPredictions REC:627-641 -> (predicted labels list, true labels list)."""
    X, true = extractDescriptorsAndRealSignTypes(detectionsDescriptors)
    if not X:
        return [], true
    _use_lda(LDAClassifiers)
    _, labels = context().lda_predict(np.stack(X).astype(np.float32), tol=tolerance, want_logits=False)
    return [int(v) for v in labels], true


def predictProbabilityKNNClassifiers(KNNClassifier, reducer, detectionDescriptors):
    """REC:592-596 -> (predicted labels ndarray, true labels list)."""
    global _knn_key
    X, true = extractDescriptorsAndRealSignTypes(detectionDescriptors)
    lda = reducer[0]
    Zt = np.asarray(KNNClassifier._fit_X, np.float64)
    yt = np.asarray(KNNClassifier.classes_)[np.asarray(KNNClassifier._y)].astype(np.int32)
    key = (lda.xbar_.tobytes(), Zt.tobytes(), yt.tobytes(), KNNClassifier.n_neighbors)
    if key != _knn_key:
        ncomp = Zt.shape[1]
        if ncomp != 6:
            raise TsdError("KNN path expects the 6-component LDA reducer of REC:586-589")
        context().set_knn(lda.xbar_, lda.scalings_[:, :ncomp], Zt, yt, KNNClassifier.n_neighbors)
        _knn_key = key
    if not X:
        return np.zeros(0, np.int64), true
    _, labels = context().knn_predict(np.stack(X).astype(np.float32), want_Z=False)
    return labels.astype(np.int64), true


def predictProbability(classifiers, reducer, testDataDescriptors, tolerance):
    """REC:619-624 dispatcher."""
    if classifiers[1] == 'LDABAYES':
        return predictProbabilityLDAClassifiers(classifiers[0], testDataDescriptors, tolerance)
    elif classifiers[1] == 'KNN':
        return predictProbabilityKNNClassifiers(classifiers[0], reducer, testDataDescriptors)


_PATCHED = ("makeWindowBiggerOrDiscardFakeDetections", "cleanDuplicatedDetections", "MSERTrafficSignDetector",
            "computeDescriptors", "calculateDescriptors", "predictProbabilityLDAClassifiers",
            "predictProbabilityKNNClassifiers", "predictProbability")


def install(reference_source_module):
    """Monkey-patch the reference's REC `source` module so `testValidation()` runs the GPU path."""
    for name in _PATCHED:
        setattr(reference_source_module, name, globals()[name])
    return reference_source_module
