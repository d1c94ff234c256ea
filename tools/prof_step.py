"""Small driver for ncu / timing experiments: N steps of the device-resident detection chain on synthetic frames."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tsd_b200

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--boxes", type=int, default=200)
ap.add_argument("--H", type=int, default=800)
ap.add_argument("--W", type=int, default=1360)
ap.add_argument("--times", action="store_true")
ap.add_argument("--timeline", action="store_true", help="stage boundaries of the overlapped batches (ms after the first)")
ap.add_argument("--wall", action="store_true", help="CUDA-event time of the steps enqueued back to back (no per-stage events)")
ap.add_argument("--real", action="store_true", help="the three stored real test frames with their real cv2.MSER boxes, tiled to --frames (bench.py's real_mser_frames)")
ap.add_argument("--mode", default="det", choices=["det", "rec"], help="det: K1 K2 K5 K3 K4 (x1.30, 25x25); rec: K1 K2 K5 K6 K7 K8 (x1.15, 32x32)")
a = ap.parse_args()
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "det_templates.npz"))
if a.real:
    import cv2
    gd = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    gf = np.load(os.path.join(gd, "det_frames.npz"))
    names = ["00604", "00639", "00719"]
    uniq = np.stack([cv2.imread(os.path.join(gd, "det_frame_%s.png" % k)) for k in names])
    bl = [gf[k + "_boxes"].astype(np.int32) for k in names]
    U = 3
    boxes = np.concatenate([bl[f % 3] for f in range(a.frames)])
    off = np.concatenate([[0], np.cumsum([len(bl[f % 3]) for f in range(a.frames)])]).astype(np.int32)
    a.boxes = int(max(len(b) for b in bl))
else:
    U = min(16, a.frames)
    uniq = tsd_b200.synth.make_frames(U, a.H, a.W)
    boxes, off = tsd_b200.synth.make_boxes(a.frames, a.boxes, a.H, a.W)
dev = torch.device("cuda", 0)
d_frames = torch.from_numpy(uniq).to(dev)[torch.arange(a.frames, device=dev) % U].contiguous()
d_boxes, d_off = torch.from_numpy(boxes).to(dev), torch.from_numpy(off).to(dev)
mode = tsd_b200.RUN_DETECT if a.mode == "det" else tsd_b200.RUN_RECOGNIZE
ctx = tsd_b200.Context(0, a.mode)
if a.mode == "det":
    ctx.set_templates(g["red6"], g["blue6"])
else:
    r = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "rec_golden.npz"))
    ctx.set_lda(r["lda_W"], r["lda_b"])
def one():
    ctx.enqueue_frames(d_frames.data_ptr(), a.frames, a.H, a.W, d_boxes.data_ptr(), d_off.data_ptr(), int(off[-1]), mode=mode, max_boxes_per_frame=a.boxes)
if a.times:
    for _ in range(3):
        one()
    ctx.synchronize()
    ctx.set_profiling(True)
if a.timeline:
    for _ in range(4):
        one()
    ctx.synchronize()
    ctx.set_profiling(2)
if a.wall:
    for _ in range(3):
        one()
    ctx.synchronize()
    st = torch.cuda.ExternalStream(ctx.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
for _ in range(a.steps):
    one()
if a.wall:
    ctx.flush()
    e1.record(st)
ctx.synchronize()
if a.wall:
    print("ms_per_step %.4f" % (e0.elapsed_time(e1) / a.steps))
if a.times:
    print({k: round(v / a.steps, 4) for k, v in ctx.stage_times()})
if a.timeline:
    tl = ctx.timeline()
    b = -1
    for name, t in tl:
        if name == "start":
            b += 1
            print("\nbatch %d (slot %d): start %.3f" % (b, b & 1, t), end="")
        else:
            print("  %s %.3f" % (name.replace("k5_", "").replace("_crop_resize", "").replace("_expand_filter", ""), t), end="")
    print()
det, counts = ctx.fetch_detections(int(off[-1]))
print("counts", counts.tolist(), "ndet", len(det))
