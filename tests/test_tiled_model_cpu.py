"""CPU: the serial model of the tiled pipeline (tests/model/tiled_model.py: tile-local forests, the seam-aware BORDER rule,
on-chip folding, seam records, alias folding, refit, start rule) yields exactly the oracle's canonical node set -- for several
tile shapes, with walls, with MIN_AREA 0 / 5 / 120 -- and the seam-aware rule sends fewer nodes to the global forest."""
import numpy as np
import pytest
from conftest import make_plane
from tiled_model import tiled_nodes


@pytest.mark.parametrize("kind", ["noise", "smooth", "blobs", "walls", "wall0", "wall01", "allwall", "flat", "checker", "ramp"])
def test_tiled_model_matches_oracle(port, kind):
    for seed, (h, w), (th, tw) in [(0, (37, 53), (8, 16)), (1, (40, 70), (16, 16)), (2, (33, 64), (32, 64)), (3, (50, 45), (7, 11))]:
        img = make_plane(40 + seed, h, w, kind)
        lev = port.quantize(img, 8).reshape(img.shape)
        for ma in (0, 5, 120):
            exp = sorted(map(tuple, port.canonical_nodes(img, ma)))
            for aware in (True, False):
                got = sorted(tiled_nodes(lev, ma, TH=th, TW=tw, seam_aware=aware))
                assert got == exp, (kind, h, w, th, tw, ma, aware)


def test_seam_aware_rule_sends_fewer_nodes_global(port):
    rng = np.random.RandomState(7)
    base = rng.randint(60, 200, (12, 16)).astype(np.float32)
    import cv2
    img = cv2.resize(base, (192, 128), interpolation=cv2.INTER_CUBIC) + rng.normal(0, 3, (128, 192))
    img = np.clip(img, 0, 255).astype(np.uint8)
    lev = port.quantize(img, 8).reshape(img.shape)
    a, b = {}, {}
    na = sorted(tiled_nodes(lev, 120, TH=32, TW=64, seam_aware=True, stats=a))
    nb = sorted(tiled_nodes(lev, 120, TH=32, TW=64, seam_aware=False, stats=b))
    assert na == nb == sorted(map(tuple, port.canonical_nodes(img, 120)))
    assert a["local_nodes"] == b["local_nodes"] and a["global_nodes"] < b["global_nodes"]


@pytest.mark.parametrize("kind", ["noise", "smooth", "blobs", "walls", "wall0", "allwall", "checker", "ramp"])
def test_supertile_merge_model_matches_oracle(port, kind):
    """the round-2 design (unite the BORDER graphs of a group of tiles before the global kernels) leaves the result unchanged"""
    for seed, (h, w), (th, tw), sup in [(0, (40, 70), (8, 16), (2, 2)), (1, (50, 45), (7, 11), (4, 2)), (2, (64, 96), (16, 16), (2, 3))]:
        img = make_plane(60 + seed, h, w, kind)
        lev = port.quantize(img, 8).reshape(img.shape)
        for ma in (0, 5, 120):
            exp = sorted(map(tuple, port.canonical_nodes(img, ma)))
            st = {}
            got = sorted(tiled_nodes(lev, ma, TH=th, TW=tw, supertile=sup, stats=st))
            assert got == exp, (kind, h, w, th, tw, sup, ma)
            assert st["global_nodes_after_supertile"] <= st["global_border_nodes"]
