"""GPU: er_tree_extract + NMS + classify through the C ABI against the oracle -- bit-exact.

The oracle is run with the canonical sibling order (see test_sibling_order.py); with that, node
arrays (DFS order, parents, child counts), pool order, labels and scores must be IDENTICAL."""
import numpy as np
import pytest
from conftest import make_plane

pytestmark = pytest.mark.gpu


def _check(got, exp, tag):
    assert got.nodes.shape == exp["nodes"].shape and (got.nodes == exp["nodes"]).all(), ("nodes", tag)
    assert got.pool.shape == exp["pool"].shape and (got.pool == exp["pool"]).all(), ("pool", tag)
    assert (got.label == exp["label"]).all(), ("label", tag)
    assert (got.strong_score == exp["strong_score"]).all() and (got.weak_score == exp["weak_score"]).all(), ("score", tag)


KINDS = ["noise", "smooth", "blobs", "walls", "wall0", "wall01", "allwall", "flat", "checker", "ramp"]
SIZES = [(1, 1), (1, 9), (7, 1), (31, 33), (32, 64), (33, 65), (64, 128), (100, 130), (257, 191)]


@pytest.mark.parametrize("kind", KINDS)
def test_planes_match_oracle(ert, port, kind):
    try:
        for si, (h, w) in enumerate(SIZES):
            img = make_plane(si, h, w, kind)
            for ma in (3, 120):
                ert.set_min_area(ma); port.params["min_area"] = ma
                exp = port.plane(img, scores=True, canonical_order=True)
                got = ert.planes_detect(img)
                assert got.status == 0
                _check(got.planes[0], exp, (kind, h, w, ma))
    finally:
        ert.set_min_area(120); port.params["min_area"] = 120


@pytest.mark.parametrize("step", [5, 8, 13, 16, 32])
def test_other_threshold_steps(ert, port, step):
    img = make_plane(5, 120, 160, "smooth")
    port.params["thresh_step"] = step
    ert.set_thresh_step(step)
    try:
        _check(ert.planes_detect(img).planes[0], port.plane(img, scores=True, canonical_order=True), step)
    finally:
        port.params["thresh_step"] = 8; ert.set_thresh_step(8)


def test_batch_of_planes_is_independent_of_position(ert, port):
    planes = np.stack([make_plane(s, 96, 200, k) for s, k in enumerate(["noise", "smooth", "walls", "blobs", "flat", "checker", "allwall"])])
    res = ert.planes_detect(planes)
    for i in range(planes.shape[0]):
        _check(res.planes[i], port.plane(planes[i], scores=True, canonical_order=True), i)
    res2 = ert.planes_detect(planes[::-1].copy())
    for i in range(planes.shape[0]):
        assert (res2.planes[planes.shape[0] - 1 - i].nodes == res.planes[i].nodes).all()


def test_full_hd_plane_and_mode_equivalence(ert, port):
    """1080p: oracle parity on one natural-like and one noise plane; the seam kernels (compacted lists / one thread per
    seam position) agree."""
    for kind in ("blobs", "noise"):
        img = make_plane(42, 1080, 1920, kind)
        exp = port.plane(img, scores=True, canonical_order=True)
        got = ert.planes_detect(img)
        assert got.status == 0
        _check(got.planes[0], exp, kind)
        ert.set_seam_list(0)
        try:
            got0 = ert.planes_detect(img)
        finally:
            ert.set_seam_list(1)
        assert (got0.planes[0].nodes == got.planes[0].nodes).all() and (got0.planes[0].pool == got.planes[0].pool).all()
        root = got.planes[0].nodes[0]
        assert root[6] == -1 and root[1] > 1080 * 1920 - (img >= 252).sum()   # area = pixels + nodes (Q2)


def test_kept_capacity_overflow_is_reported(ert):
    img = make_plane(1, 300, 300, "noise")
    ert.set_capacity(64, 64)
    ert.set_min_area(3)
    try:
        res = ert.planes_detect(img)
        assert res.status & 2
    finally:
        ert.set_capacity(16384, 2048); ert.set_min_area(120)


def test_node_capacity_overflow_is_reported(ert):
    """more tile-local nodes leave their tiles than the plane's node array holds: flagged, and fine again with the default"""
    img = make_plane(1, 300, 300, "noise")
    ert.set_node_capacity(64)
    try:
        assert ert.planes_detect(img).status & 16
    finally:
        ert.set_node_capacity(0)
    assert ert.planes_detect(img).status == 0


def test_status_word_belongs_to_one_batch(ert):
    """an overflow reported for one batch must not poison the next batch on the same context (the status word is per batch)"""
    img = make_plane(1, 300, 300, "noise")
    ert.set_capacity(64, 64)
    ert.set_min_area(3)
    try:
        assert ert.planes_detect(img).status & 2
        ert.set_min_area(120)                      # does not touch the workspace: same context, same status word
        res = ert.planes_detect(np.full((64, 64), 77, np.uint8))
        assert res.status == 0 and len(res.planes[0].nodes) == 1
    finally:
        ert.set_capacity(16384, 2048); ert.set_min_area(120)


@pytest.mark.parametrize("cfg", [1, 2, 3])
def test_tile_kernel_variants_agree(ert, cfg):
    """the A/B variants of the tile kernel (per-level fold / arrival counters, with / without the horizontal edge skip)
    produce byte-identical results"""
    imgs = [make_plane(7, 257, 191, k) for k in ("noise", "blobs", "walls", "checker")] 
    base = [ert.planes_detect(im) for im in imgs]
    ert.set_tile_config(cfg)
    try:
        for im, b in zip(imgs, base):
            g = ert.planes_detect(im)
            assert g.planes[0].nodes.tobytes() == b.planes[0].nodes.tobytes() and g.planes[0].pool.tobytes() == b.planes[0].pool.tobytes()
    finally:
        ert.set_tile_config(0)


@pytest.mark.gpu
def test_enqueue_planes_matches_planes_detect_and_overlaps_scales(ert):
    """asynchronous plane entry point: two contexts, two plane sizes in flight at once == the synchronous calls"""
    import ertext
    from conftest import make_plane
    a = np.stack([make_plane(3, 120, 200, "blobs"), make_plane(4, 120, 200, "smooth")])
    b = np.stack([make_plane(5, 60, 100, "blobs")])
    ra, rb = ert.planes_detect(a), ert.planes_detect(b)
    c1, c2 = ertext.ErText(), ertext.ErText()
    c1.enqueue_planes(a); c2.enqueue_planes(b)
    qa, qb = c1.fetch(), c2.fetch()
    for x, y in ((ra, qa), (rb, qb)):
        assert x.status == y.status == 0
        for p, q in zip(x.planes, y.planes):
            assert p.nodes.tobytes() == q.nodes.tobytes() and p.pool.tobytes() == q.pool.tobytes() and p.label.tobytes() == q.label.tobytes()
    c1.close(); c2.close()


@pytest.mark.gpu
def test_maximum_plane_size_matches_oracle(port):
    """the largest plane the key layout supports: 8191 x 8191 (67 092 481 pixels < 2^26), with a wall band and a wall start pixel;
    one pixel more per side is refused"""
    import ertext
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    big = cv2.resize(rng.randint(0, 256, (64, 64)).astype(np.uint8), (8191, 8191), interpolation=cv2.INTER_CUBIC)
    big[4000:4100, 100:8000] = 255
    big[0, 0] = 255
    e = ertext.ErText()
    res = e.planes_detect(big, upto=ertext.STAGE_NMS)
    assert res.status == 0
    exp = port.plane(big, classify=False, canonical_order=True)
    got = res.planes[0]
    assert got.nodes.shape == exp["nodes"].shape and (got.nodes == exp["nodes"]).all()
    assert (got.pool == exp["pool"]).all()
    with pytest.raises(ertext.ErtError):
        e.planes_detect(np.zeros((8, 8192), np.uint8))
    e.close()
