"""One-off large-scale parity fuzz (not part of the test suite): the device-resident chain against the oracle, record by record,
on many seeded synthetic frames.  det: K1 K2 K5 K5 K3 K4; rec: K1 K2 K5 K5 K6 K7 K8 (labels of all survivors via the stage calls).

    python tools/fuzz_chain.py --frames 1024 --seed 5 [--mode rec]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tsd_b200
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--boxes", type=int, default=200)
ap.add_argument("--mode", default="det", choices=["det", "rec"])
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
U = 32
uniq = tsd_b200.synth.make_frames(U, seed=tsd_b200.synth.FRAME_SEED + a.seed)
boxes, off = tsd_b200.synth.make_boxes(a.frames, a.boxes, seed=tsd_b200.synth.BOX_SEED + 100 + a.seed)
dev = torch.device("cuda", 0)
d_frames = torch.from_numpy(uniq).to(dev)[torch.arange(a.frames, device=dev) % U].contiguous()
d_boxes, d_off = torch.from_numpy(boxes).to(dev), torch.from_numpy(off).to(dev)
t0 = time.time()
bad = 0
if a.mode == "det":
    g = np.load(os.path.join(root, "tests", "golden", "det_templates.npz"))
    ctx = tsd_b200.Context(0, "det")
    ctx.set_templates(g["red6"], g["blue6"])
    ctx.enqueue_frames(d_frames.data_ptr(), a.frames, 800, 1360, d_boxes.data_ptr(), d_off.data_ptr(), int(off[-1]), max_boxes_per_frame=a.boxes)
    det, counts = ctx.fetch_detections(int(off[-1]))
    got = {}
    for d in det:
        got.setdefault(int(d["frame"]), []).append((int(d["x1"]), int(d["y1"]), int(d["x2"]), int(d["y2"]), int(d["id"]), int(d["hundredths"])))
    tot = np.zeros(4, np.int64)
    for f in range(a.frames):
        o = O.detect_frame(uniq[f % U], boxes[off[f]:off[f + 1]], g["red6"], g["blue6"])
        tot += o["stage_counts"]
        exp = [tuple(int(v) for v in c) + (int(i), int(h)) for c, i, h in zip(o["coords"], o["ids"], o["hundredths"])]
        if got.get(f, []) != exp:
            bad += 1
            print("MISMATCH frame", f)
    print("det: frames", a.frames, "seed", a.seed, "counts gpu", counts.tolist(), "oracle", tot.tolist(), "bad frames", bad, "%.1fs" % (time.time() - t0))
    bad += int(counts.tolist() != tot.tolist())
else:
    r = np.load(os.path.join(root, "tests", "golden", "rec_golden.npz"))
    ctx = tsd_b200.Context(0, "rec")
    ctx.set_lda(r["lda_W"], r["lda_b"])
    ctx.enqueue_frames(d_frames.data_ptr(), a.frames, 800, 1360, d_boxes.data_ptr(), d_off.data_ptr(), int(off[-1]), mode=tsd_b200.RUN_RECOGNIZE,
                       max_boxes_per_frame=a.boxes)
    det, counts = ctx.fetch_detections(int(off[-1]))
    got = {}
    for d in det:
        got.setdefault(int(d["frame"]), []).append((int(d["x1"]), int(d["y1"]), int(d["x2"]), int(d["y2"]), int(d["id"])))
    nsurv = 0
    for f in range(a.frames):
        b = boxes[off[f]:off[f + 1]]
        c, v = O.expand_boxes(b, 1.15)
        c = c[v]
        wins = np.stack([O.crop_resize(uniq[f % U], cc, 32) for cc in c]) if len(c) else np.zeros((0, 32, 32, 3), np.uint8)
        w1, c1 = O.dedup(wins, c, False, 0.85)
        w2, c2 = O.dedup(w1, c1, True, 0.95)
        nsurv += len(c2)
        exp = []
        if len(c2):
            hog = np.stack([O.hog32(O.bgr2gray(w)) for w in w2])
            _, lab = O.lda_predict(hog, r["lda_W"], r["lda_b"], 0.5)
            exp = [tuple(int(x) for x in cc) + (int(l),) for cc, l in zip(c2, lab) if l != 0]
        if got.get(f, []) != exp:
            bad += 1
            print("MISMATCH frame", f)
    print("rec: frames", a.frames, "seed", a.seed, "counts gpu", counts.tolist(), "oracle survivors", nsurv, "bad frames", bad, "%.1fs" % (time.time() - t0))
    bad += int(int(counts[2]) != nsurv)
sys.exit(1 if bad else 0)
