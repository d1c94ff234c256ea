import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
STORED = ("00604", "00639", "00719")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; oracle/tsd_oracle.c)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def templates():
    g = np.load(os.path.join(GOLDEN, "det_templates.npz"))
    return g["red6"], g["blue6"]


@pytest.fixture(scope="session")
def det_crops():
    """The class crops calculateMeanMasks reads, in the reference's iteration order -> 6 lists of BGR images."""
    g = np.load(os.path.join(GOLDEN, "det_crops.npz"))
    pix, shapes, off = g["pixels"], g["shapes"], g["type_offsets"]
    crops, pos = [], 0
    for h, w in shapes:
        crops.append(pix[pos:pos + h * w * 3].reshape(h, w, 3)); pos += h * w * 3
    return [crops[off[t]:off[t + 1]] for t in range(6)]


@pytest.fixture(scope="session")
def det_frames():
    return np.load(os.path.join(GOLDEN, "det_frames.npz"))


@pytest.fixture(scope="session")
def det_windows50():
    return np.load(os.path.join(GOLDEN, "det_windows50.npz"))


@pytest.fixture(scope="session")
def det_full150():
    """Every stage output of the reference for ALL 150 test frames (tests/golden/make_golden.py full)."""
    return np.load(os.path.join(GOLDEN, "det_full150.npz"))


@pytest.fixture(scope="session")
def resultado150():
    """The reference's resultado.txt for the 150 test frames, sorted file order (192 lines)."""
    return open(os.path.join(GOLDEN, "det_resultado150.txt")).read().split()


@pytest.fixture(scope="session")
def jpeg24():
    """24 real test frames stored as their original JPEG bytes -> dict(index int32[24] into the 150 sorted files, files, frames
    uint8 [24,800,1360,3]).  The pixels must be the ones the reference decoded when the goldens were made (sha1 stored beside the
    bytes); a cv2 build that decodes differently cannot check K2 on these frames and skips."""
    import hashlib
    cv2 = pytest.importorskip("cv2")
    g = np.load(os.path.join(GOLDEN, "det_jpeg24.npz"))
    frames = []
    for i in range(len(g["index"])):
        img = cv2.imdecode(g["jpeg"][g["jpeg_offsets"][i]:g["jpeg_offsets"][i + 1]], cv2.IMREAD_COLOR)
        if hashlib.sha1(img.tobytes()).digest() != g["sha1"][i].tobytes():
            pytest.skip("this cv2 build decodes %s differently from the container the goldens were made in" % g["files"][i])
        frames.append(img)
    return dict(index=g["index"], files=[str(f) for f in g["files"]], frames=np.stack(frames))


@pytest.fixture(scope="session")
def rec_gray_golden():
    return np.load(os.path.join(GOLDEN, "rec_gray_golden.npz"))


@pytest.fixture(scope="session")
def rec_golden():
    return np.load(os.path.join(GOLDEN, "rec_golden.npz"))


@pytest.fixture(scope="session")
def eval_golden():
    return np.load(os.path.join(GOLDEN, "eval_golden.npz"))


@pytest.fixture(scope="session")
def rec_frames():
    return np.load(os.path.join(GOLDEN, "rec_frames.npz"))


def load_frame(name):
    """Stored real test frame (decoded BGR, lossless PNG).  Decoded with cv2 if present, else a tiny PNG reader."""
    path = os.path.join(GOLDEN, "det_frame_%s.png" % name)
    try:
        import cv2
        img = cv2.imread(path)
        assert img is not None
        return img
    except ImportError:
        return _read_png_rgb8(path)[:, :, ::-1].copy()


def _read_png_rgb8(path):
    import struct
    import zlib
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        ln, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + ln]
        if typ == b"IHDR":
            w, h, bd, ct = struct.unpack(">IIBB", body[:10])
            assert bd == 8 and ct == 2
        elif typ == b"IDAT":
            idat += body
        pos += 12 + ln
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * 3)
    out = np.zeros((h, w * 3), np.uint8)
    prev = np.zeros(w * 3, np.int32)
    for y in range(h):
        ft, line = raw[y, 0], raw[y, 1:].astype(np.int32)
        cur = np.zeros(w * 3, np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:
            for x in range(w * 3):
                a = cur[x - 3] if x >= 3 else 0
                b = prev[x]
                c = prev[x - 3] if x >= 3 else 0
                if ft == 1:
                    p = a
                elif ft == 3:
                    p = (a + b) >> 1
                else:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[x] = (line[x] + p) & 255
        out[y] = cur
        prev = cur
    return out.reshape(h, w, 3)


@pytest.fixture(scope="session")
def frames3():
    return {k: load_frame(k) for k in STORED}


@pytest.fixture(scope="session")
def tsd():
    import tsd_b200
    return tsd_b200


@pytest.fixture(scope="session")
def ctx_det(tsd, templates):
    c = tsd.Context(device=0, flavour="det")
    c.set_templates(*templates)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx_rec(tsd, rec_golden):
    c = tsd.Context(device=0, flavour="rec")
    c.set_lda(rec_golden["lda_W"], rec_golden["lda_b"])
    c.set_knn(rec_golden["knn_xbar"], rec_golden["knn_scalings"], rec_golden["knn_Ztrain"], rec_golden["knn_ytrain"], 4)
    yield c
    c.close()
