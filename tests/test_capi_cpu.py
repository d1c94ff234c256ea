"""CPU (-m "not gpu"): the C-ABI library loads, exports every symbol include/tsd_b200.h declares, and the host
logic around it behaves (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "tsd_b200.h")).read()
    return sorted(set(re.findall(r"TSD_API\s+[\w\s\*]+?\b(tsd_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(tsd):
    lib = tsd._capi.lib()
    syms = _header_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(tsd._capi.EXPORTS)


def test_struct_layouts_match_header(tsd):
    assert C.sizeof(tsd._capi.Detection) == 32 and tsd.DET_DTYPE.itemsize == 32
    cfg = tsd.default_config("det")
    assert (cfg.enlarge, cfg.window, cfg.score_tol_hundredths) == (1.30, 25, 55)
    assert (cfg.hist_tol, cfg.coord_tol, cfg.merge_factor, cfg.proba_tol, cfg.knn_k) == (0.85, 0.95, 0.8823, 0.5, 4)
    assert [list(r) for r in cfg.red_lo] == [[0, 50, 10], [160, 50, 10]] and list(cfg.blue_hi) == [128, 255, 255]
    rec = tsd.default_config("rec")
    assert (rec.enlarge, rec.window) == (1.15, 32)
    assert cfg.enlarge - 1 == 0.30000000000000004 and rec.enlarge - 1 == 0.1499999999999999   # SURVEY A.1 literals


def test_no_cpu_fallback(tsd):
    """Without a CUDA device the product path must fail loudly."""
    if tsd._capi.lib().tsd_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(tsd.TsdError, match="no CUDA device"):
        tsd.Context()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "opencv-traffic-sign-detector_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "tsd_oracle" not in text and "orc_" not in text, f


def test_similarity_table_matches_reference_expression(tsd, oracle):
    t = tsd.similarity_table(4096)
    assert t[0] == 1.0 and np.all(np.diff(t) <= 0) and t[-1] < t[1]       # monotone non-increasing
    for d2 in (1, 2, 25, 1000, 3721, 4095):
        assert abs(t[d2] - oracle.eucl_similarity_d2(d2)) < 1e-15
    full = tsd.similarity_table()
    assert np.sqrt(full[-1]) < 0.95 * 0.8823                              # beyond the table nothing can act


def test_synth_is_deterministic_and_exercises_paths(tsd, oracle):
    fr = tsd.synth.make_frames(2, H=200, W=300)
    assert fr.shape == (2, 200, 300, 3) and fr.dtype == np.uint8
    assert np.array_equal(fr, tsd.synth.make_frames(2, H=200, W=300))
    bx, off = tsd.synth.make_boxes(16, 200)
    assert np.array_equal(bx, tsd.synth.make_boxes(16, 200)[0]) and off[-1] == 3200
    assert np.all(bx[:, 0] >= 0) and np.all(bx[:, 0] + bx[:, 2] <= 1360) and np.all(bx[:, 1] + bx[:, 3] <= 800)
    coords, valid = oracle.expand_boxes(bx, 1.30)
    frac = valid.mean()
    assert 0.25 < frac < 0.6
    cw = np.minimum(coords[valid][:, 2], 1360) - coords[valid][:, 0]
    ch = np.minimum(coords[valid][:, 3], 800) - coords[valid][:, 1]
    assert np.any((cw == 50) & (ch == 50)) and np.any((cw == 25) & (ch == 25))      # AREA and copy paths
    assert np.any(coords[valid][:, 2] > 1360) or np.any(coords[valid][:, 3] > 800)    # clipped crops


def test_evaluator_host_logic(tsd, eval_golden, tmp_path):
    """The host side of evaluate.py that needs no GPU: class tables, the results-file parser, VOCap / VOColdap / draw_PR_fast on the
    reference's own tp / fp flags (tests/golden/eval_golden.npz)."""
    E, g = tsd.evaluate, eval_golden
    assert [E.calculateSignType(s) for s in ("0", "7", "11", "14", "17", "13", "38", "6", "12", "42")] == [1, 1, 2, 3, 4, 5, 6, None, None, None]
    assert [E.compute_class_index(n) for n in (0, 16, 11, 31, 14, 17, 13, 38, 6, 12, 42)] == [1, 1, 2, 2, 3, 4, 5, 6, -1, -1, -1]
    p = tmp_path / "own.txt"
    p.write_bytes(bytes(g["own_txt"]))
    _, boxes = E.load_results_file(str(p))
    assert sum(len(v) for v in boxes.values()) == len(g["own_tp"]) and all(isinstance(b.score, float) for v in boxes.values() for b in v)
    gt = tmp_path / "gt.txt"
    gt.write_bytes(bytes(g["gt_txt"]))
    _, gtb = E.load_results_file(str(gt))
    assert sum(1 for v in gtb.values() for b in v if b.class_id != -1) == int(g["own_tot"])
    for tag in ("own", "p1", "p2"):
        rec, prec, ap = E.draw_PR_fast(g[tag + "_tp"], g[tag + "_fp"], int(g[tag + "_tot"]), show=False)
        assert np.array_equal(rec, g[tag + "_rec"]) and np.array_equal(prec, g[tag + "_prec"], equal_nan=True)
        assert ap == float(g[tag + "_ap"]) and E.VOColdap(rec, prec) == float(g[tag + "_ap11"])


def test_gram_low_byte_correction_identity():
    """The arithmetic identity k5_gram relies on (tsd_k5.cuh): with counts kept modulo 256 in the u8 tile, adding
    c_r*c_i - (c_r & 255)*(c_i & 255) for every entry (r, bin) whose count exceeds 255 and every other row i holding that bin --
    skipping the visit from the HIGHER row when both counts exceed 255 in the same bin -- restores the exact Gram matrix C C^T."""
    rng = np.random.default_rng(3)
    for trial in range(20):
        n, nbins = int(rng.integers(2, 40)), 300
        Cm = np.zeros((n, nbins), np.int64)
        for r in range(n):
            k = int(rng.integers(1, 60))
            Cm[r, rng.choice(nbins, k, replace=False)] = rng.integers(1, 40, k)
            if rng.random() < 0.4:                                       # a flat window: one or two bins far above 255
                Cm[r, rng.choice(8, int(rng.integers(1, 3)), replace=False)] = rng.integers(256, 1025, 1)
        tile = (Cm & 255) @ (Cm & 255).T
        for r in range(n):
            for b in np.nonzero(Cm[r] > 255)[0]:
                for i in range(n):
                    if i == r or Cm[i, b] == 0 or (Cm[i, b] > 255 and i < r):
                        continue
                    corr = Cm[r, b] * Cm[i, b] - (Cm[r, b] & 255) * (Cm[i, b] & 255)
                    tile[max(r, i), min(r, i)] += corr
        exact = Cm @ Cm.T
        low = np.tril_indices(n, -1)
        assert np.array_equal(tile[low], exact[low])
