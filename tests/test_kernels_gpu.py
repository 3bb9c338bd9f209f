"""GPU parity tests, kernel by kernel, through the C ABI (plviwo_op_*), against the oracle (the reference's
OpenCV calls executed by cv2 — oracle/cvops.py).  Bit-exact for integer work; float tolerances are stated."""
import numpy as np
import pytest

from oracle import cvops

pytestmark = pytest.mark.gpu


def _images(synth):
    seq = synth.SynthSequence(seed=11, n_frames=4)
    rng = np.random.default_rng(5)
    return {
        "kaist": seq.frame(0),
        "noise": rng.integers(0, 256, (560, 1280), dtype=np.uint8),
        "odd": rng.integers(30, 200, (141, 333), dtype=np.uint8),
        "flat": np.full((128, 256), 77, np.uint8),
    }


@pytest.mark.parametrize("name", ["kaist", "noise", "odd", "flat"])
def test_equalize_pyramid_bit_exact(fe, synth, name):
    img = _images(synth)[name]
    levels = 5
    lv, half = fe.op_equalize_pyramid(img, levels)
    eq = cvops.equalize_hist(img)
    assert np.array_equal(lv[0], eq), "equalizeHist differs"
    ref = cvops.build_pyramid(eq, 3, levels)  # win 3: never stops early
    for l in range(1, min(len(ref), len(lv))):
        assert lv[l].shape == ref[l].shape
        assert np.array_equal(lv[l], ref[l]), "pyrDown level %d differs" % l
    if img.shape[0] % 2 == 0 and img.shape[1] % 2 == 0:
        assert np.array_equal(half, cvops.half_res(eq)), "half-res differs"


@pytest.mark.parametrize("shape", [(112, 256), (37, 85), (180, 192), (64, 64)])
def test_fast_cell_bit_exact(fe, synth, shape):
    img = cvops.equalize_hist(_images(synth)["kaist"])
    rng = np.random.default_rng(3)
    for trial in range(4):
        y0 = int(rng.integers(0, img.shape[0] - shape[0]))
        x0 = int(rng.integers(0, img.shape[1] - shape[1]))
        roi = np.ascontiguousarray(img[y0:y0 + shape[0], x0:x0 + shape[1]])
        xy, resp = cvops.fast_cell(roi, 20)
        got = fe.op_fast_cell(roi, 20)
        assert len(got) == len(xy)
        assert np.array_equal(got[:, :2], xy), "FAST keypoint list / order differs"
        assert np.array_equal(got[:, 2].astype(np.float32), resp), "FAST scores differ"


def test_corner_subpix(fe, synth):
    img = cvops.equalize_hist(_images(synth)["kaist"])
    xy, resp = cvops.fast_cell(img, 30)
    pts = xy[np.argsort(-resp, kind="stable")[:400]].astype(np.float32)
    # include border cases
    pts = np.concatenate([pts, np.array([[3, 3], [1276, 556], [640, 3], [3, 280]], np.float32)], 0)
    ref = cvops.corner_subpix(img, pts)
    got = fe.op_corner_subpix(img, pts)
    err = np.abs(got - ref).max(1)
    assert np.percentile(err, 99) < 1e-3, err.max()
    assert err.max() < 2e-2, err.max()   # iteration-count flips at the eps boundary stay well inside 0.05 px


@pytest.mark.parametrize("win,levels", [(15, 3), (15, 5), (21, 4)])
def test_lk_vs_opencv(fe, synth, win, levels):
    seq = synth.SynthSequence(seed=21, n_frames=6)
    a = cvops.equalize_hist(seq.frame(2))
    b = cvops.equalize_hist(seq.frame(3))
    xy, resp = cvops.fast_cell(a, 25)
    rng = np.random.default_rng(0)
    sel = rng.choice(len(xy), size=min(600, len(xy)), replace=False)
    pts0 = xy[sel].astype(np.float32) + rng.uniform(-0.5, 0.5, (len(sel), 2)).astype(np.float32)
    # points hugging the borders exercise the out-of-frame window paths
    border = np.stack([rng.uniform(0, 1279, 200), rng.choice([1.5, 4.0, 9.0, 550.0, 556.5, 558.9], 200)], 1).astype(np.float32)
    border2 = np.stack([rng.choice([0.5, 3.0, 8.0, 1270.0, 1277.5, 1279.0], 100), rng.uniform(0, 559, 100)], 1).astype(np.float32)
    pts0 = np.concatenate([pts0, border, border2], 0)
    ref_p, ref_s = cvops.lk(a, b, pts0, pts0, win, levels)
    got_p, got_s = fe.op_lk(a, b, pts0, pts0, win, levels)
    agree = ref_s == got_s
    assert agree.mean() >= 0.995, "status agreement %.4f" % agree.mean()
    both = agree & (ref_s == 1)
    err = np.abs(got_p[both] - ref_p[both]).max(1)
    assert err.max() < 0.05, "max |duv| = %g px" % err.max()      # north_star tolerance: 0.05 px
    assert np.percentile(err, 99) < 5e-3


def test_undistort(fe):
    from plviwo_b200 import synth as s
    rng = np.random.default_rng(1)
    pts = np.stack([rng.uniform(0, 1280, 500), rng.uniform(0, 560, 500)], 1).astype(np.float32)
    ref = np.concatenate([cvops.undistort(pts[i:i + 1], s.KAIST_K, s.KAIST_D) for i in range(len(pts))], 0)
    got = fe.op_undistort(pts, s.KAIST_K, s.KAIST_D)
    assert np.array_equal(got, ref) or np.abs(got - ref).max() < 1e-7


def test_canny_bit_exact(fe, synth):
    for name in ("kaist", "noise", "odd"):
        img = _images(synth)[name]
        small = img[: img.shape[0] // 2 * 2, : img.shape[1] // 2 * 2]
        small = cvops.half_res(cvops.equalize_hist(small))
        ref = cvops.canny(small)
        ref[:6, :6] = 0          # FastLineDetector clears the two corner blocks before walking
        ref[-5:, -5:] = 0
        got = fe.op_canny(small, 50.0)
        assert np.array_equal(got, ref), "%s: %d pixels differ" % (name, int((got != ref).sum()))


@pytest.mark.parametrize("thread_walk", [False, True])
def test_fld_vs_restatement(fe, synth, thread_walk, monkeypatch):
    """The line extractor has no executable reference (opencv_contrib absent): compared with the oracle's C++
    restatement.  Transcendental rounding (atan2/cos/sin) differs between device and host libm, so endpoints are
    compared to 1e-3 px and segment counts must be equal.  thread_walk: the components whose bounding box fits 62 x 44 pixels
    are walked by one thread each (k_fld_walk_thread; by default only launches that carry >= 16 frames do that)."""
    if thread_walk:
        monkeypatch.setenv("PLVIWO_WALK_THREAD_MIN", "1")
    else:
        monkeypatch.delenv("PLVIWO_WALK_THREAD_MIN", raising=False)
    for lh in (False, True):
        seq = synth.SynthSequence(seed=31, n_frames=4, line_heavy=lh)
        small = cvops.half_res(cvops.equalize_hist(seq.frame(1)))
        ref = cvops.fld_detect(small)
        got = fe.op_fld(small)
        assert len(got) == len(ref), (len(got), len(ref))
        assert np.abs(got - ref).max() < 1e-3


def test_select_kernel_equals_host_std_sort(fe):
    """k_fast_select (device introsort + top num_features_grid) against the host std::sort restatement, including a cell
    larger than the shared-memory capacity (global scratch path)."""
    rng = np.random.default_rng(11)
    for n, span in [(0, 5), (1, 5), (17, 1), (400, 30), (3000, 8), (8192, 3), (9000, 4), (20000, 40)]:
        resp = rng.integers(21, 21 + span, n).astype(np.uint32)
        x = rng.integers(0, 4000, n).astype(np.uint32)
        y = rng.integers(0, 4000, n).astype(np.uint32)
        packed = (resp << 24) | (y << 12) | x
        for nfg in (9, 17):
            ref = fe.op_sort_corners(packed, nfg, device=-1)[:nfg]
            got = fe.op_sort_corners(packed, nfg, device=0)
            exp = np.stack([(ref & 0xfff).astype(np.float32), ((ref >> 12) & 0xfff).astype(np.float32)], 1)
            assert got.shape == exp.shape and np.array_equal(got, exp), (n, span, nfg)


@pytest.mark.parametrize("shape", [(560, 1280), (283, 645), (280, 645), (283, 640), (64, 64), (1080, 1920)])
def test_clahe_bit_exact(fe, synth, shape):
    """k_clahe_lut + the CLAHE blend in k_eq_pyr1 against cv::createCLAHE(10.0, 8x8)->apply (TrackKLT.cpp:60-64)."""
    from oracle import cvops
    rng = np.random.default_rng(shape[0] + shape[1])
    if shape == (560, 1280):
        img = synth.SynthSequence(seed=5, n_frames=1).frame(0)
    elif shape == (64, 64):
        img = np.full(shape, 9, np.uint8)
    else:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(fe.op_clahe(img), cvops.clahe(img))


def test_fld_restatement_against_contrib_golden(fe):
    """The CUDA line extractor against the real cv::ximgproc::FastLineDetector output (tests/golden/make_golden_fld.py);
    skipped only while the golden file is absent (no opencv_contrib in the authoring image)."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gpath = os.path.join(here, "fld_golden.npz")
    if not os.path.exists(gpath):
        pytest.skip("tests/golden/fld_golden.npz absent: run tests/golden/make_golden_fld.py where cv2.ximgproc exists")
    inputs, golden = dict(np.load(os.path.join(here, "fld_inputs.npz"))), dict(np.load(gpath))
    for name, img in inputs.items():
        got = fe.op_fld(img)
        want = golden[name]
        assert got.shape == want.shape, (name, got.shape, want.shape)
        if len(want):
            assert np.abs(got - want).max() <= 2e-3, name
