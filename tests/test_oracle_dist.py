"""CPU: the oracle's restatement of cv::distanceTransform(DIST_L2, DIST_MASK_5) and dcut::dist_cut against the golden
vectors OpenCV 4.13 (+IPP) produced (oracle/gen_golden_dist.py -> tests/golden/dist.npz).  Bit-exact."""
import numpy as np


def test_distance_transform_matches_opencv(oracle, golden):
    g = golden("dist.npz")
    for name in g["dt_names"]:
        got = oracle.distance_transform(g[f"dt_mask_{name}"])
        ref = g[f"dt_ref_{name}"]
        assert got.shape == ref.shape
        if name == "far_corner":
            # Distances beyond 32 px: where `left neighbour + 1` is an exact float tie, this OpenCV build (IPP) ends up
            # one ulp above the plain two-pass minimum on a thin band of pixels (its internal evaluation order is not
            # published).  Bounded here: <= 1 ulp (1.2e-7 relative, the path's float bar is 1e-5), < 1 % of the pixels.
            ulp = np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32))
            assert ulp.max() <= 1 and (ulp != 0).mean() < 0.01
            continue
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), name


def test_dist_cut_matches_reference_restatement(oracle, golden):
    g = golden("dist.npz")
    corners = [tuple(int(v) for v in c) for c in g["cut_corners"]]
    masks = [g[f"cut_mask_{i}"] for i in range(len(corners))]
    cuts = oracle.dist_cut(masks, corners)
    for i, c in enumerate(cuts):
        assert np.array_equal(c, g[f"cut_ref_{i}"]), i
    # the cut only ever removes pixels, and the disjoint image is untouched
    assert all(np.all((c == m) | (c == 0)) for c, m in zip(cuts, masks))
    assert np.array_equal(cuts[4], masks[4])


def _blend_case(g):
    corners = [tuple(int(v) for v in c) for c in g["cut_corners"]]
    tiles = [g[f"blend_tile_{i}"] for i in range(len(corners))]
    masks = [g[f"blend_mask_{i}"] for i in range(len(corners))]
    return tiles, masks, corners


def test_simple_and_no_blend_match_opencv(oracle, golden):
    """blnd::simple_blend / blnd::no_blend restated in C against the same loops driven through cv2 (golden)."""
    g = golden("dist.npz")
    tiles, masks, corners = _blend_case(g)
    s = oracle.simple_blend(tiles, masks, corners)
    assert np.abs(s.astype(int) - g["simple_ref"].astype(int)).max() <= 1      # <= 1 LSB on the 8-bit canvas
    assert np.array_equal(oracle.no_blend(tiles, masks, corners), g["noblend_ref"])


def test_overlap_intensity_matches_opencv(oracle, golden):
    """gain::get_overlapp_intensity: exact integer sums, same pair order as the reference's push_back."""
    g = golden("dist.npz")
    corners = [tuple(int(v) for v in c) for c in g["cut_corners"]]
    tiles = [g[f"ov_tile_{i}"] for i in range(len(corners))]
    got = np.array(oracle.overlap_intensity(tiles, corners, g["ov_adj"]), np.float64)
    assert got.shape == g["ov_ref"].shape and np.array_equal(got, g["ov_ref"])
    assert (got[:, 2] > 0).sum() >= 6 and (got[:, 2] == 0).sum() >= 1      # overlapping pairs and the disjoint one
