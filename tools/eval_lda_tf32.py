"""north_star: "Only the batched LDA projection ... is evaluated for tensor cores (TF32) against FP32 FMA".  Measures, on the
reference's held-out descriptors (tests/golden/rec_golden.npz, 1 623 windows) and on 35 059 perturbed copies (the survivor count
of the 1024-frame recognition chain), label agreement and logit error of plain TF32 and 3xTF32 mma.sync against the product's f64
FMA kernel, and the kernel times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsd_b200

r = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "rec_golden.npz"))
ctx = tsd_b200.Context(0, "rec")
ctx.set_lda(r["lda_W"], r["lda_b"])
rng = np.random.default_rng(0)
sets = {"golden_1623": r["hog"].astype(np.float32)}
big = np.tile(r["hog"], (22, 1))[:35059].astype(np.float32)
big = np.clip(big + rng.normal(0, 0.01, big.shape).astype(np.float32), 0, None)
sets["perturbed_35059"] = big
for name, X in sets.items():
    z64, lab64 = ctx.lda_predict(X)
    if name == "golden_1623":
        assert np.array_equal(lab64, r["pred_lda"]), "f64 kernel must reproduce the reference's labels"
    margin = np.abs(z64).min()
    for split in (1, 3):
        z, lab, ms = ctx.lda_predict_tf32(X, split=split)
        err = np.abs(z.astype(np.float64) - z64)
        print("%-16s n=%6d  %s: label flips %5d  max |dz| %.3e  mean |dz| %.3e  (smallest |z| in set %.4f)  kernel %.4f ms" % (
            name, len(X), "TF32  " if split == 1 else "3xTF32", int((lab != lab64).sum()), err.max(), err.mean(), margin, ms))
# f64 kernel time: device-resident recognition chain stage time is reported by tools/prof_step.py --mode rec (k8_lda)
