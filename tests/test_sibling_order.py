"""CPU: the one place where the parallel path cannot follow the reference literally is the order in
which siblings are visited by NMS (the reference's order is the sequential flood's completion order).
This test measures what the canonical order costs on the golden frames and pins it."""
import numpy as np


def test_canonical_sibling_order_costs_at_most_a_handful_of_pool_entries(port, golden_frames):
    tot = diff = 0
    for f in range(golden_frames.shape[0]):
        ch = port.channels(golden_frames[f])
        for k in range(6):
            a = port.plane(ch[k], classify=False)
            b = port.plane(ch[k], classify=False, canonical_order=True)
            # node sets are identical, only the child order may differ
            assert sorted(map(tuple, a["nodes"][:, :6])) == sorted(map(tuple, b["nodes"][:, :6]))
            pa = set(map(tuple, a["nodes"][a["pool"]][:, :6])); pb = set(map(tuple, b["nodes"][b["pool"]][:, :6]))
            tot += len(pa); diff += len(pa ^ pb)
    # measured on 211 ICDAR frames: 1 of 14 668; on these 3 frames: 0
    assert tot > 150 and diff <= 2, (tot, diff)
