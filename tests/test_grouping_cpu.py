"""CPU: the facade's er_grouping / overlap_suppression / inner_suppression / fitline_avgslope / er_ocr duplicate removal
(host-only C++ in scene-text-recognition_b200/host/ERFilter.hpp; SURVEY 8f rank 3 stays on the CPU) against the
reference's own code (oracle/_ref: src/ER.cpp:612-692, 893-964, 1362-1389, 702-724 compiled verbatim)."""
import os
import subprocess
import numpy as np
import pytest
from conftest import GOLDEN, ROOT, PKG


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("grp") / "grouping_demo")
    subprocess.check_call(["g++", "-std=c++11", "-O1", os.path.join(ROOT, "tests", "cpp", "grouping_demo.cpp"), "-o", exe,
                           "-L", PKG, "-l:libertext.so", "-Wl,-rpath," + PKG])
    return exe


def _rows_from_golden(g, tag):
    """all_er exactly as er_track leaves it: tracked order, with centre and colours"""
    S, W = g[tag + "_strong"], g[tag + "_weak"]
    rows = []
    for kind, idx in g[tag + "_tracked"]:
        r = (S if kind == 0 else W)[idx]
        col = (g[tag + "_strong_color"] if kind == 0 else g[tag + "_weak_color"])[idx]
        cen = (g[tag + "_strong_center"] if kind == 0 else g[tag + "_weak_center"])[idx]
        rows.append([r[0], r[1], r[2], r[3], r[4], r[5], cen[0], cen[1], col[0], col[1], col[2]])
    return np.array(rows, np.float64).reshape(-1, 11)


def _run_demo(exe, rows, osup, isup, dedupe, tmp_path):
    p = tmp_path / "rows.txt"
    with open(p, "w") as f:
        f.write("%d %d %d %d\n" % (len(rows), osup, isup, dedupe))
        for r in rows:
            f.write(" ".join(str(int(v)) for v in r[:8]) + " " + " ".join(repr(float(v)) for v in r[8:]) + "\n")
    out = subprocess.run([exe, str(p)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    after, bounds, texts = [], [], []
    for line in out.stdout.splitlines():
        t = line.split()
        if t[0] == "AFTER":
            after = [int(v) for v in t[1:]]
        elif t[0] == "B":
            bounds.append([int(v) for v in t[1:]])
        elif t[0] == "T":
            texts.append((float(t[1]), [int(v) for v in t[2:]]))
    return after, np.array(bounds, np.int32).reshape(-1, 6), texts


def _same(a, b):
    return a == b or (np.isnan(a) and np.isnan(b))


def _check(ref, exe, rows, osup, isup, dedupe, tmp_path, tag):
    exp = ref.er_grouping(rows, osup, isup, dedupe)
    after, bounds, texts = _run_demo(exe, rows, osup, isup, dedupe, tmp_path)
    assert after == exp["after"], tag
    assert (bounds == exp["bounds"]).all(), tag
    assert len(texts) == len(exp["texts"]), tag
    for (s1, m1), (s2, m2) in zip(texts, exp["texts"]):
        assert m1 == m2 and _same(s1, s2), tag          # same members in the same order, bit-identical slope


def test_grouping_matches_reference_on_golden_tracked_lists(ref, demo, tmp_path):
    g = np.load(os.path.join(GOLDEN, "ref_next.npz"))
    n_texts = 0
    for tag in ("f0", "f1", "f2", "s1", "s3", "s4", "s5"):
        rows = _rows_from_golden(g, tag)
        for (osup, isup, dedupe) in ((0, 1, 0), (0, 1, 1), (0, 0, 0), (1, 1, 1)):     # text_detect uses (false, true) / (false, false)
            _check(ref, demo, rows, osup, isup, dedupe, tmp_path, (tag, osup, isup, dedupe))
        n_texts += len(ref.er_grouping(rows, False, True)["texts"])
    assert n_texts > 10


def test_grouping_matches_reference_on_seeded_lines(ref, demo, tmp_path):
    """synthetic text lines (slanted, overlapping duplicates, nested boxes, ties in center.x, NaN colours)"""
    rng = np.random.RandomState(21)
    for case in range(12):
        rows = []
        for line in range(rng.randint(1, 5)):
            x0, y0, h, sl = rng.randint(20, 300), rng.randint(20, 400), rng.randint(12, 60), rng.uniform(-0.3, 0.3)
            col = rng.uniform(40, 200, 3)
            for k in range(rng.randint(2, 9)):
                w = int(h * rng.uniform(0.4, 1.0)); x = x0 + int(k * h * 0.9); y = y0 + int(sl * (x - x0)) + rng.randint(-2, 3)
                hh = h + rng.randint(-3, 4)
                rows.append([rng.randint(0, 6), x, y, w, hh, int(w * hh * 0.5) + 121, x + w // 2, y + hh // 2] + list(col + rng.uniform(-6, 6, 3)))
                if rng.rand() < 0.3:       # near-duplicate from another channel
                    rows.append([rng.randint(0, 6), x + rng.randint(0, 2), y, w, hh, int(w * hh * 0.5) + 130, x + w // 2, y + hh // 2] + list(col))
                if rng.rand() < 0.2:       # a small box nested inside
                    rows.append([rng.randint(0, 6), x + w // 4, y + hh // 4, w // 2, hh // 3, 125, x + w // 2, y + hh // 2 - 1] + list(col))
        if case % 4 == 3 and rows:
            rows[0][8] = float("nan")
        rows = np.array(rows, np.float64).reshape(-1, 11)
        rows = rows[rng.permutation(len(rows))]
        for (osup, isup, dedupe) in ((0, 1, 1), (1, 0, 0), (1, 1, 1)):
            _check(ref, demo, rows, osup, isup, dedupe, tmp_path, (case, osup, isup, dedupe))
