"""Frame sharding across ranks (one process per GPU) and the single small gather of detections for reporting.

Frames are independent units: every stage is per window except the fold, which couples
windows of ONE frame only.  So the hot path needs no collective; each rank owns a contiguous block of frames chosen to
balance the candidate count.  The only exchange is an end-of-batch gather of fixed-size detection records.
Works with torch.distributed backends `nccl` (device tensors) and `gloo` (CPU tests).
"""
import numpy as np

from ._capi import DET_DTYPE


def shard_bounds(box_offsets, world_size):
    """Contiguous frame ranges [lo, hi) per rank, balanced by the number of candidate boxes (not by frame count).

    box_offsets int[F+1] (CSR).  Every frame goes to exactly one rank; ranks may be empty when F < world_size."""
    off = np.asarray(box_offsets, np.int64)
    F = len(off) - 1
    total = int(off[-1] - off[0])
    bounds = [0]
    for r in range(1, world_size):
        if total == 0:
            cut = F * r // world_size
        else:
            target = off[0] + total * r / world_size
            cut = int(np.searchsorted(off, target, side="left"))
            # choose the nearer of the two neighbouring frame boundaries
            if cut > 0 and abs(off[cut - 1] - target) <= abs(off[min(cut, F)] - target):
                cut -= 1
        bounds.append(min(max(cut, bounds[-1]), F))
    bounds.append(F)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def local_shard(frames, boxes, box_offsets, rank, world_size):
    """-> (frames[lo:hi], boxes of those frames, rebased offsets, lo)."""
    lo, hi = shard_bounds(box_offsets, world_size)[rank]
    off = np.asarray(box_offsets)
    b0, b1 = int(off[lo]), int(off[hi])
    return frames[lo:hi], np.asarray(boxes)[b0:b1], (off[lo:hi + 1] - off[lo]).astype(np.int32), lo


def gather_detections(det, frame_base, dist, device=None, dst=0):
    """Gather every rank's detection records on rank `dst` (variable length: one count exchange, one padded
    all_gather of 32-byte records).  `det` is this rank's structured array (DET_DTYPE) with LOCAL frame indices;
    `frame_base` is the global index of the rank's first frame.  Returns the concatenated records in global frame
    order on every rank (all_gather), which is what the report needs."""
    import torch
    world = dist.get_world_size()
    det = np.ascontiguousarray(det, DET_DTYPE).copy()
    det["frame"] += int(frame_base)
    n = torch.tensor([len(det)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = np.zeros((cap, 8), np.int32)
    buf[:len(det)] = det.view(np.int32).reshape(-1, 8)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = [p.cpu().numpy()[:c] for p, c in zip(parts, counts)]
    allrec = np.concatenate(out).reshape(-1, 8)
    return np.ascontiguousarray(allrec).view(DET_DTYPE).reshape(-1)
