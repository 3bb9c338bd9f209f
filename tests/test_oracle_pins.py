"""CPU tests that PIN the oracle: the NumPy restatements (oracle/npops.py) and the C++ shim against the real
OpenCV (cv2 = the reference's dependency) and against the committed golden vectors (tests/golden/ops_golden.npz,
made by tests/golden/make_golden.py).  The reference ships no tests or vectors of its own."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import cvops, npops, frontend as ofe  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ops_golden.npz"))


def _imgs():
    rng = np.random.default_rng(9)
    base = cv2.GaussianBlur(rng.integers(0, 256, (150, 222), dtype=np.uint8), (0, 0), 1.2)
    return [GOLD["img_a"], base, rng.integers(0, 256, (64, 97), dtype=np.uint8), np.full((40, 40), 9, np.uint8)]


def test_equalize_hist_bit_exact():
    for img in _imgs():
        assert np.array_equal(npops.equalize_hist(img), cvops.equalize_hist(img))
    assert np.array_equal(npops.equalize_hist(GOLD["img_a"]), GOLD["eq_a"])


def test_clahe_bit_exact():
    """NumPy restatement of cv::createCLAHE(10, 8x8) against the real one, incl. sizes that need the reflect extension."""
    rng = np.random.default_rng(2)
    imgs = [rng.integers(0, 256, (70, 160), dtype=np.uint8), rng.integers(90, 140, (283, 645), dtype=np.uint8),
            rng.integers(0, 256, (283, 640), dtype=np.uint8), np.full((64, 64), 7, np.uint8),
            (np.add.outer(np.arange(120), np.arange(200)) % 256).astype(np.uint8)]
    for im in imgs:
        assert np.array_equal(npops.clahe(im), cvops.clahe(im)), im.shape
    assert np.array_equal(npops.clahe(GOLD["img_a"]), GOLD["clahe_a"])


def test_downsample_bit_exact():
    """UpdaterCamera's pyrDown to the truncated half size (image and mask), incl. odd sides."""
    rng = np.random.default_rng(4)
    for shape in [(560, 1280), (563, 1285), (101, 64), (64, 99)]:
        im = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(npops.downsample(im), cvops.downsample(im)), shape
        m = (rng.random(shape) > 0.7).astype(np.uint8) * 255
        assert np.array_equal(npops.downsample(m), cvops.downsample(m)), shape


def test_pyramid_and_half_bit_exact():
    for img in _imgs()[:3]:
        a, b = npops.build_pyramid(img, 3, 4), cvops.build_pyramid(img, 3, 4)
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        ev = img[: img.shape[0] // 2 * 2, : img.shape[1] // 2 * 2]
        assert np.array_equal(npops.half_res(ev), cvops.half_res(ev))
    for l in range(4):
        assert np.array_equal(npops.build_pyramid(GOLD["eq_a"], 15, 3)[l], GOLD["pyr_a_%d" % l])
    # level count rule of buildOpticalFlowPyramid: stop when the next level is <= win
    assert len(npops.build_pyramid(GOLD["eq_a"], 15, 5)) == len(cvops.build_pyramid(GOLD["eq_a"], 15, 5))


def test_resize_nearest_mask():
    rng = np.random.default_rng(2)
    m = (rng.random((560, 1280)) > 0.5).astype(np.uint8) * 255
    for gx, gy in ((5, 5), (15, 15), (10, 6), (5, 3)):
        assert np.array_equal(npops.resize_nearest(m, gx, gy), cvops.resize_nearest(m, gx, gy))


def test_fast_bit_exact_incl_order():
    for img in _imgs()[:3]:
        for thr in (10, 20, 30):
            xy, r = npops.fast_cell(img, thr)
            xy2, r2 = cvops.fast_cell(img, thr)
            assert np.array_equal(xy, xy2) and np.array_equal(r, r2)
    xy, r = npops.fast_cell(GOLD["fast_roi"], 20)
    assert np.array_equal(xy, GOLD["fast_xy"]) and np.array_equal(r, GOLD["fast_resp"])


def test_std_sort_is_unstable_and_pinned():
    resp = GOLD["fast_resp"]
    perm = cvops.sort_perm(resp)
    assert np.array_equal(perm, GOLD["fast_perm"])
    assert np.all(np.diff(resp[perm]) <= 0)
    stable = np.argsort(-resp, kind="stable")
    if len(resp) > 16 and len(np.unique(resp)) < len(resp):
        assert not np.array_equal(perm, stable), "introsort must differ from a stable sort on tied responses"


def test_corner_subpix_close():
    got = npops.corner_subpix(GOLD["eq_a"], GOLD["subpix_in"])
    assert np.abs(got - GOLD["subpix_out"]).max() < 1e-3
    assert np.abs(cvops.corner_subpix(GOLD["eq_a"], GOLD["subpix_in"]) - GOLD["subpix_out"]).max() < 1e-6


def test_lk_close_and_status_equal():
    eq_b = cvops.equalize_hist(GOLD["img_b"])
    p1, st = npops.lk(GOLD["eq_a"], eq_b, GOLD["subpix_out"], GOLD["subpix_out"], 15, 3)
    assert np.array_equal(st, GOLD["lk_status"])
    ok = st == 1
    assert np.abs(p1[ok] - GOLD["lk_p1"][ok]).max() < 2e-3
    # border-hugging points: statuses must still agree with OpenCV
    rng = np.random.default_rng(4)
    h, w = GOLD["eq_a"].shape
    pts = np.stack([rng.choice([0.2, 2.0, 7.5, w - 8.0, w - 1.2], 60), rng.uniform(0, h - 1, 60)], 1).astype(np.float32)
    r_p, r_s = cvops.lk(GOLD["eq_a"], eq_b, pts, pts, 15, 3)
    g_p, g_s = npops.lk(GOLD["eq_a"], eq_b, pts, pts, 15, 3)
    assert np.array_equal(r_s, g_s)
    ok = r_s == 1
    assert np.abs(g_p[ok] - r_p[ok]).max() < 5e-3


def test_undistort_equidistant_restatement_exact():
    """The next camera model (cam/CamEqui.h:108-129, cv::fisheye::undistortPoints): the NumPy restatement against cv2 on an
    EuRoC-like calibration, points over the whole image and beyond its border.  (The library refuses FE_CAM_EQUI so far.)"""
    Ke, De = [458.654, 457.296, 367.215, 248.375], [-0.0348, 0.0123, -0.0071, 0.0021]
    rng = np.random.default_rng(5)
    pts = np.stack([rng.uniform(-20, 772, 3000), rng.uniform(-20, 500, 3000)], 1).astype(np.float32)
    pts[0] = (Ke[2], Ke[3])          # the principal point: theta_d = 0
    assert np.array_equal(npops.undistort_equi(pts, Ke, De), cvops.undistort_equi(pts, Ke, De))
    strong = [0.35, -0.6, 0.9, -0.4]  # a strongly distorting lens
    assert np.array_equal(npops.undistort_equi(pts, Ke, strong), cvops.undistort_equi(pts, Ke, strong))


def test_undistort_exact():
    K, D = tuple(GOLD["K"]), tuple(GOLD["D"])
    assert np.array_equal(npops.undistort(GOLD["subpix_out"], K, D), GOLD["und_p0"])
    assert np.array_equal(npops.undistort(GOLD["lk_p1"], K, D), GOLD["und_p1"])


def test_canny_bit_exact():
    for img in _imgs()[:3]:
        assert np.array_equal(npops.canny(img), cvops.canny(img))
    assert np.array_equal(npops.canny(GOLD["half_a"]), GOLD["canny_half_a"])


def test_fitline_shim_matches_cv2():
    pts = np.ascontiguousarray(GOLD["fitline_pts"], np.int32)
    out = np.zeros(4, np.float32)
    cvops.shim().oracle_fit_line(pts.ctypes.data, len(pts), out.ctypes.data)
    assert np.abs(out - GOLD["fitline_out"]).max() < 1e-6
    ref = cv2.fitLine(pts.astype(np.float32), cv2.DIST_L2, 0, 0.01, 0.01).reshape(-1)
    assert np.abs(out - ref).max() < 1e-6


def test_ransac_restatement_mask_exact():
    K = tuple(GOLD["K"])
    m = npops.find_fundamental_mask(GOLD["und_p0"], GOLD["und_p1"], 2.0 / max(K[0], K[1]))
    assert np.array_equal(m, GOLD["ransac_mask"])


def test_fld_restatement_runs_on_both_cannys():
    a = cvops.fld_detect(GOLD["half_a"])
    b = npops.fld_detect(GOLD["half_a"])
    assert a.shape == b.shape and np.array_equal(a, b)
    # a synthetic image with known straight strokes: every returned segment lies on a stroke
    img = np.full((200, 300), 40, np.uint8)
    cv2.line(img, (20, 30), (280, 60), 220, 3)
    cv2.line(img, (50, 180), (250, 100), 220, 2)
    segs = cvops.fld_detect(img)
    assert len(segs) >= 2
    for s in segs:
        d1 = abs((s[1] - 30) - (s[0] - 20) * 30 / 260.0)
        d2 = abs((s[1] - 180) - (s[0] - 50) * (-80) / 200.0)
        assert min(d1, d2) < 4.0


def test_frontend_oracle_matches_golden_rows():
    import plviwo_b200  # noqa: F401
    from plviwo_b200 import synth
    seq = synth.SynthSequence(seed=77, width=320, height=192, n_frames=4)
    assert np.array_equal(seq.frame(0), GOLD["img_a"]), "synthetic sequence generator is not reproducible"
    for ops in (cvops, npops):
        fe = ofe.FrontEnd(ofe.FeConfig(num_features=60, grid_x=4, grid_y=3, pyr_levels=3, K=tuple(GOLD["K"]), D=tuple(GOLD["D"])), ops)
        for t in range(3):
            prow, lrow = fe.feed(seq.timestamp(t), seq.frame(t), None, seq.vanishing_points(t))
            g = GOLD["fe_rows_%d" % t]
            assert [r.id for r in prow] == list(g[:, 0].astype(int)), (ops.__name__, t)
            if len(prow):
                uv = np.array([[r.u, r.v, r.un, r.vn] for r in prow])
                assert np.abs(uv[:, :2] - g[:, 1:3]).max() < (1e-6 if ops is cvops else 2e-2)
            assert list(fe.klt.get_last_ids()) == list(GOLD["fe_last_ids_%d" % t])
            assert [r.id for r in lrow] == list(GOLD["fe_line_ids_%d" % t])


def _fld_golden():
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gpath, ipath = os.path.join(here, "fld_golden.npz"), os.path.join(here, "fld_inputs.npz")
    if not os.path.exists(gpath):
        pytest.skip("tests/golden/fld_golden.npz absent: run tests/golden/make_golden_fld.py where cv2.ximgproc exists "
                    "(FastLineDetector parity stays unpinned until then)")
    return dict(np.load(ipath)), dict(np.load(gpath))


def test_fld_inputs_are_committed():
    """The images the FastLineDetector pin is generated from travel with the repo (tests/golden/make_golden_fld.py)."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fld_inputs.npz"))
    assert set(d.files) == {"kaist_320x192", "kaist_640x280", "lines_640x280"}
    assert d["kaist_640x280"].shape == (280, 640) and d["kaist_320x192"].dtype == np.uint8


def test_fld_restatement_against_contrib_golden():
    """oracle/csrc/oracle_shim.cpp (FastLineDetector restated) against the real cv::ximgproc::FastLineDetector output:
    same number of segments, same order, end points to 1e-3 px."""
    from oracle import cvops
    inputs, golden = _fld_golden()
    for name, img in inputs.items():
        got = cvops.fld_detect(img)
        want = golden[name]
        assert got.shape == want.shape, (name, got.shape, want.shape)
        assert np.abs(got - want).max() <= 1e-3 if len(want) else True, name
