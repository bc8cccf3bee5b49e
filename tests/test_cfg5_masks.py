"""cfg5 inputs: the graph-cut seam masks "precomputed by the reference" (BASELINE.json configs[4]) -- gcut::graph_cut
restated in oracle/graph_cut.py on the reference's own vendored max-flow, stored bit-packed in tests/golden/cfg5_masks.npz.
They are an INPUT of the hot path; this test pins that the fixture belongs to the cfg2 layout, that the masks partition the
preview canvas (no pixel owned twice) with irregular seams, and -- where the reference tree is present -- that the
restatement still reproduces the fixture for one pair of tiles."""
import os

import numpy as np
import pytest


def _layout():
    from oracle import ref_bench
    synth = ref_bench.load_synth()
    full = synth.config("cfg2")
    K, R, g = synth.cameras(full)
    corners, sizes, W, H, T = ref_bench.job_geometry(full, K, R)
    return synth, full, K, R, g, corners, sizes


def test_fixture_matches_layout_and_partitions_the_canvas():
    from oracle import graph_cut, ref_bench
    synth, full, K, R, g, corners, sizes = _layout()
    masks, how = graph_cut.seam_masks_for(full, K, R, g, corners, sizes)
    assert len(masks) == full.n and "graph-cut" in how
    small = synth.config("cfg2", 1.0 / graph_cut.PREVIEW)
    Ks, Rs, _ = synth.cameras(small)
    pc, ps, Wp, Hp, _ = ref_bench.job_geometry(small, Ks, Rs)
    assert [m.shape for m in masks] == [(h, w) for (w, h) in ps]
    mx, my = min(c[0] for c in pc), min(c[1] for c in pc)
    cnt = np.zeros((Hp, Wp), np.uint8)
    for c, m in zip(pc, masks):
        assert set(np.unique(m)) <= {0, 255}
        cnt[c[1] - my:c[1] - my + m.shape[0], c[0] - mx:c[0] - mx + m.shape[1]] += (m != 0)
    assert cnt.max() == 1 and (cnt > 0).mean() > 0.97
    # irregular seams: the left edge of a tile's kept region is not a straight column
    m = masks[5]
    first = np.array([np.flatnonzero(r)[0] for r in m if r.any()])
    assert first.std() > 3.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/max_flow"), reason="needs the reference's vendored max_flow sources")
def test_restated_compute_cut_runs_on_the_vendored_maxflow():
    from oracle import graph_cut
    graph_cut.build()
    rng = np.random.default_rng(3)
    h, w = 60, 90
    a = rng.integers(0, 255, (h, w), dtype=np.uint8)
    b = a.copy(); b[:, 40:] = rng.integers(0, 255, (h, w - 40), dtype=np.uint8)   # images agree on the left part
    scene = np.zeros((h, w), np.uint8); scene[:, :60] = 255                       # what is already pasted
    elem = np.zeros((h, w), np.uint8); elem[:, 20:] = 255                         # the new image
    cut = graph_cut.compute_cut(a, b, scene, elem)
    assert cut.shape == elem.shape and set(np.unique(cut)) <= {0, 255}
    assert (cut[:, 60:] == 255).all() and (cut[:, :20] == 0).all()                # outside the overlap the element mask is kept
    seam = np.array([np.flatnonzero(r[20:60])[0] if r[20:60].any() else 40 for r in cut])
    assert seam.min() >= 0 and (cut[:, 20:60] != 0).any() and (cut[:, 20:60] == 0).any()   # the cut runs through the overlap
