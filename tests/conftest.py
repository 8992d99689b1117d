import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scene-text-recognition_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests", "model")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.refbind import PortOracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "libert_port.so")):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libert_port.so"])
    return PortOracle(with_svm=True)


@pytest.fixture(scope="session")
def ref():
    """The reference's own code; only where oracle/_ref/libref_oracle.so was built (needs /root/reference at build time)."""
    from oracle.refbind import RefOracle
    try:
        return RefOracle(with_svm=True)
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref/libref_oracle.so not built here")


@pytest.fixture(scope="session")
def golden_frames():
    return np.load(os.path.join(GOLDEN, "frames.npz"))["frames"]


@pytest.fixture(scope="session")
def golden_planes():
    return np.load(os.path.join(GOLDEN, "ref_planes.npz"))


@pytest.fixture(scope="session")
def golden_feats():
    return np.load(os.path.join(GOLDEN, "ref_feats.npz"))


@pytest.fixture(scope="session")
def golden_svm():
    return np.load(os.path.join(GOLDEN, "ref_svm.npz"))


@pytest.fixture(scope="session")
def ert():
    import ertext
    e = ertext.ErText(device=0, load_svm=True)
    yield e
    e.close()


def blur(img, k):
    a = img.astype(np.float32)
    for _ in range(k):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5.0
    return np.clip(a, 0, 255).astype(np.uint8)


def make_plane(seed, h, w, kind):
    """Seeded synthetic planes covering the component-tree edge cases."""
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (h, w)).astype(np.uint8)
    if kind == "noise":
        return img
    if kind == "smooth":
        return np.clip((blur(img, 6).astype(int) - 128) * 6 + 128, 0, 255).astype(np.uint8)
    if kind == "blobs":
        b = np.clip((blur(img, 10).astype(int) - 128) * 12 + 128, 0, 255).astype(np.uint8)
        return b
    if kind == "walls":      # 255-walls split the plane; only the flood's component counts
        b = np.clip((blur(img, 3).astype(int) - 128) * 5 + 128, 0, 255).astype(np.uint8)
        b[h // 3, : max(w - 3, 0)] = 255
        b[:, w // 2] = 255
        b[0, 0] = 255
        return b
    if kind == "wall0":      # start pixel is a wall, flood escapes to the right neighbour
        b = blur(img, 2)
        b[0, 0] = 255
        return b
    if kind == "wall01":     # ... or to the bottom neighbour
        b = blur(img, 2)
        b[0, 0] = 255; b[0, min(1, w - 1)] = 255
        return b
    if kind == "allwall":    # pixels 0, 1 and W are all walls: lone component of area 2
        b = blur(img, 2)
        b[0, 0] = 255; b[0, min(1, w - 1)] = 254; b[min(1, h - 1), 0] = 253
        return b
    if kind == "flat":
        return np.full((h, w), 77, np.uint8)
    if kind == "checker":    # every pixel its own node
        yy, xx = np.mgrid[0:h, 0:w]
        return (((yy + xx) % 2) * 120 + (yy % 3) * 16 + (xx % 5) * 8).astype(np.uint8)
    if kind == "ramp":
        yy, xx = np.mgrid[0:h, 0:w]
        return ((xx * 255) // max(w - 1, 1)).astype(np.uint8)
    raise ValueError(kind)
