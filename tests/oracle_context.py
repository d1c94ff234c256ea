"""TEST DOUBLE: engine.Context's detection-flavour methods answered by the CPU oracle.

Only for tests/test_oracle_full.py::test_reference_driver_through_install: the reference's own `test()` driver is run through
`source_det.install()` in the GPU-less container, where the reference tree exists but no CUDA device does.  It checks the
host-side glue of the drop-in mirrors (argument / return conventions, the patched names, the driver's control flow) end to end
against the reference's own resultado.txt; the kernels themselves are checked against the same oracle by the -m gpu tests.
Nothing on the product path imports this file.
"""
import types

import numpy as np

from oracle import oracle as O


class OracleContext:
    def __init__(self, flavour="det"):
        self.D = 25 if flavour == "det" else 32
        self.cfg = types.SimpleNamespace(enlarge=1.30 if flavour == "det" else 1.15, proba_tol=0.5)
        self.red6 = self.blue6 = None
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def set_templates(self, red6, blue6):
        self._count("set_templates")
        self.red6 = np.asarray(red6, np.uint8).reshape(6, self.D, self.D)
        self.blue6 = np.asarray(blue6, np.uint8).reshape(6, self.D, self.D)

    def preprocess(self, frames, clip_limit=2.0, tiles=(8, 8), gamma=2):
        self._count("preprocess")
        frames = np.asarray(frames, np.uint8)
        if frames.ndim == 3:
            return O.preprocess(frames, clip_limit, tiles, gamma)
        return np.stack([O.preprocess(f, clip_limit, tiles, gamma) for f in frames])

    def expand_boxes(self, boxes, enlarge=None):
        self._count("expand_boxes")
        return O.expand_boxes(np.asarray(boxes, np.int32).reshape(-1, 4), self.cfg.enlarge if enlarge is None else enlarge)

    def crop_resize(self, frames, coords, win_frame=None, D=None):
        self._count("crop_resize")
        frames = np.asarray(frames, np.uint8)
        grey = frames.ndim == 2 or (frames.ndim == 3 and frames.shape[-1] != 3)
        if frames.ndim == 2 or (frames.ndim == 3 and not grey):
            frames = frames[None]
        coords = np.asarray(coords, np.int32).reshape(-1, 4)
        wf = np.zeros(len(coords), np.int32) if win_frame is None else np.asarray(win_frame, np.int32)
        D = self.D if D is None else int(D)
        return np.stack([O.crop_resize(frames[f], c, D) for c, f in zip(coords, wf)])

    def windows(self, frames, boxes, box_offsets):
        self._count("windows")
        frames = np.asarray(frames, np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        F, H, W = frames.shape[:3]
        boxes = np.asarray(boxes, np.int32).reshape(-1, 4); box_offsets = np.asarray(box_offsets, np.int32)
        wins, coords, woff = [], [], [0]
        for f in range(F):
            c, v = O.expand_boxes(boxes[box_offsets[f]:box_offsets[f + 1]], self.cfg.enlarge)
            for cc, ok in zip(c, v):
                # K1 also drops a box whose crop is empty after clipping to the frame (cv2.resize would raise on it)
                if ok and min(cc[2], W) > min(cc[0], W) and min(cc[3], H) > min(cc[1], H):
                    wins.append(O.crop_resize(frames[f], cc, self.D)); coords.append(cc)
            woff.append(len(coords))
        wins = np.stack(wins) if wins else np.zeros((0, self.D, self.D, 3), np.uint8)
        return wins, np.array(coords, np.int32).reshape(-1, 4), np.array(woff, np.int32)

    def dedup(self, windows, coords, offsets, by_coords, tol):
        self._count("dedup")
        windows = np.asarray(windows, np.uint8); coords = np.asarray(coords, np.int32).reshape(-1, 4)
        ow, oc, ooff = [], [], [0]
        for a, b in zip(offsets[:-1], offsets[1:]):
            w, c = O.dedup(windows[a:b], coords[a:b], by_coords, tol)
            ow.append(w); oc.append(c); ooff.append(ooff[-1] + len(c))
        return np.concatenate(ow), np.concatenate(oc), np.array(ooff, np.int32)

    def hist(self, windows):
        self._count("hist")
        return np.stack([O.hist_normalized(w) for w in np.asarray(windows, np.uint8)])

    def color_masks(self, windows):
        self._count("color_masks")
        rb = [O.color_masks(w) for w in np.asarray(windows, np.uint8)]
        return np.stack([r for r, _ in rb]), np.stack([b for _, b in rb])

    def mean_windows(self, windows, group_offsets):
        self._count("mean_windows")
        out = []
        for a, b in zip(group_offsets[:-1], group_offsets[1:]):
            mask = np.zeros(windows.shape[1:], np.uint8)
            for k, r in enumerate(windows[a:b].astype(np.int32)):
                if k == 0:
                    mask = r.astype(np.uint8)
                else:                                         # cv2.addWeighted(img, .5, mask, .5, 0): round-half-even of the mean (DET:52)
                    s = r + mask.astype(np.int32)
                    mask = ((s >> 1) + ((s & 1) & ((s >> 1) & 1))).astype(np.uint8)
            out.append(mask)
        return np.stack(out)

    def score_masks(self, red, blue, want_scores=True):
        """DET:229-261 on masks: per colour the first maximum over the six templates; red wins only if strictly larger; > 0.55."""
        self._count("score_masks")
        red, blue = np.asarray(red, np.uint8), np.asarray(blue, np.uint8)
        n = len(red)
        sc = np.zeros((n, 2, 6), np.int32); ids = np.zeros(n, np.int32); hs = np.zeros(n, np.int32); em = np.zeros(n, bool)
        for i in range(n):
            for k in range(6):
                sc[i, 0, k] = O.score_hundredths(red[i], self.red6[k])[0]
                sc[i, 1, k] = O.score_hundredths(blue[i], self.blue6[k])[0]
            kr, kb = int(np.argmax(sc[i, 0])), int(np.argmax(sc[i, 1]))      # (argmax = first maximum)
            if sc[i, 0, kr] > sc[i, 1, kb]:
                ids[i], hs[i] = kr + 1, sc[i, 0, kr]
            else:
                ids[i], hs[i] = kb + 1, sc[i, 1, kb]
            em[i] = hs[i] > 55
        return dict(scores=sc if want_scores else None, id=ids, hundredths=hs, emit=em)

    # ---- recognition flavour ------------------------------------------------------------------------------------------
    def bgr2gray(self, bgr):
        self._count("bgr2gray")
        return O.bgr2gray(np.asarray(bgr, np.uint8))

    def hog(self, gray):
        self._count("hog")
        gray = np.asarray(gray, np.uint8).reshape(-1, 32, 32)
        return np.stack([O.hog32(g) for g in gray]) if len(gray) else np.zeros((0, 324), np.float32)

    def set_lda(self, W, b):
        self._count("set_lda")
        self.W, self.b = np.asarray(W, np.float64), np.asarray(b, np.float64)

    def lda_predict(self, X, tol=None, want_logits=True):
        self._count("lda_predict")
        lg, lab = O.lda_predict(np.asarray(X, np.float32), self.W, self.b, self.cfg.proba_tol if tol is None else tol)
        return (lg if want_logits else None), lab

    def set_knn(self, xbar, scalings, Ztrain, ytrain, k=4):
        self._count("set_knn")
        self.knn = (np.asarray(xbar, np.float64), np.asarray(scalings, np.float64), np.asarray(Ztrain, np.float64), np.asarray(ytrain, np.int32), int(k))

    def knn_predict(self, X, want_Z=True):
        self._count("knn_predict")
        Z, lab = O.knn_predict(np.asarray(X, np.float32), *self.knn)
        return (Z if want_Z else None), lab
