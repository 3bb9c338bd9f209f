"""Stereo point front end (TrackKLT::feed_stereo / perform_detection_stereo, TrackKLT.cpp:202-393, 530-827) through the
C ABI against the oracle (oracle/stereo.py), same bars as the monocular path: feature ids bit-exact, tracked UVs within
0.05 px, status flags equal on >= 99.5 % of features.  Teacher-forced (state loaded from the oracle before every pair)
and free-running; golden rows from cv2 (tests/golden/stereo_golden.npz)."""
import os

import numpy as np
import pytest

from oracle import frontend as ofe
from oracle import stereo as ost

pytestmark = pytest.mark.gpu

CFG1 = dict(num_features=200, fast_threshold=20, grid_x=5, grid_y=5, min_px_dist=10, pyr_levels=3, win_size=15)
CFG2 = dict(num_features=400, fast_threshold=20, grid_x=5, grid_y=5, min_px_dist=10, pyr_levels=4, win_size=15)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "stereo_golden.npz")


def _blob(fe, o, W, H):
    st = o.get_state()
    cams = [fe.pack_state(W, H, st["currid"], st["pts_last"][c], st["ids_last"][c], st["img_last"][c], st["mask_last"][c])
            for c in (0, 1)]
    return fe.pack_stereo_state(st["currid"], cams[0], cams[1])


def _run(fe, synth, n_frames, kw, seed=1000, width=1280, height=560, teacher_forced=True, hard=True, moving_mask=False,
         K_right=None, lookahead=0):
    seq = synth.SynthSequence(seed=seed, width=width, height=height, n_frames=n_frames, hard=hard, moving_mask=moving_mask)
    o = ost.TrackKLTStereo(ofe.FeConfig(K=seq.K, D=seq.D, **kw), K_right=K_right)
    g = fe.StereoFrontEnd(fe.default_config(width=width, height=height, K=seq.K, D=seq.D, lookahead=lookahead, use_lines=0, **kw),
                          K_right=K_right)
    s = dict(frames=0, rows=0, rows_sym=0, order_equal=0, frames_equal=0, stereo_rows=0, det_left=0, det_right=0, new_stereo=0,
             first_divergence=None)
    duv, dun = [0.0], [0.0]
    for t in range(n_frames):
        il, ir = seq.frame(t, 0), seq.frame(t, 1)
        ml = seq.mask(t, 0) if moving_mask else np.zeros_like(il)
        mr = seq.mask(t, 1) if moving_mask else np.zeros_like(il)
        if teacher_forced and t > 0:
            g.set_state(_blob(fe, o, width, height))
        rows_o = o.feed_new_camera(seq.timestamp(t), il, ir, ml, mr)
        info = g.feed_new_camera(seq.timestamp(t), il, ir, ml if moving_mask else None, mr if moving_mask else None)
        det = o.trace["det"]
        assert bool(info.reset) == bool(o.trace["reset"]), t
        assert bool(info.first_frame) == bool(o.trace["first"]), t
        assert [bool(info.detection_ran[0]), bool(info.detection_ran[1])] == [det["ran_left"], det["ran_right"]], t
        s["frames"] += 1
        s["det_left"] += int(det["ran_left"])
        s["det_right"] += int(det["ran_right"])
        s["new_stereo"] += int(info.n_stereo_new)
        s["stereo_rows"] += int(info.n_stereo_rows)
        frame_equal = True
        for cam in (0, 1):
            rows = g.point_rows(cam)
            ids_o = {r.id: r for r in rows_o[cam]}
            ids_g = {int(r["id"]): r for r in rows}
            sym = set(ids_o) ^ set(ids_g)
            s["rows"] += len(ids_o)
            s["rows_sym"] += len(sym)
            if sym:
                frame_equal = False
            else:
                s["order_equal"] += int([r.id for r in rows_o[cam]] == [int(v) for v in rows["id"]])
            for fid in set(ids_o) & set(ids_g):
                a, b = ids_g[fid], ids_o[fid]
                duv.append(max(abs(float(a["u"]) - b.u), abs(float(a["v"]) - b.v)))
                dun.append(max(abs(float(a["un"]) - b.un), abs(float(a["vn"]) - b.vn)))
        if frame_equal:
            s["frames_equal"] += 1
            last_ids = g.get_last_ids()
            for cam in (0, 1):
                assert np.array_equal(last_ids[cam], np.array(o.ids_last[cam], np.uint64)), (t, cam)
            last = g.get_last_obs()     # pts_last are the rows' positions (their differences are counted in duv)
            for cam in (0, 1):
                assert last[cam].shape == o.pts_last[cam].shape, (t, cam)
        elif s["first_divergence"] is None:
            s["first_divergence"] = t
            if not teacher_forced:
                break
    g.close()
    duv, dun = np.array(duv), np.array(dun)
    s.update(max_duv=float(duv.max()), duv_p99=float(np.percentile(duv, 99)), n_duv_gt_005=int((duv > 0.05).sum()),
             max_dun=float(dun.max()), dun_p99=float(np.percentile(dun, 99)))
    print(s)
    return s


def _assert_parity(s):
    assert s["rows"] > 0, s
    assert s["rows_sym"] <= 0.005 * s["rows"], s              # a flipped status flag changes one row
    assert s["order_equal"] >= 2 * s["frames_equal"], s       # and where the sets agree the row order is the reference's
    assert s["duv_p99"] < 0.01, s
    assert s["n_duv_gt_005"] <= max(1, int(0.001 * s["rows"])), s
    assert s["dun_p99"] < 0.01 / 500.0, s                    # normalised coordinates: the same bar through the focal length


@pytest.mark.parametrize("kw,seed", [(CFG1, 1000), (CFG2, 1001)])
def test_stereo_teacher_forced(fe, synth, kw, seed):
    s = _run(fe, synth, 30, kw, seed=seed)
    _assert_parity(s)
    assert s["det_left"] >= 5 and s["det_right"] >= 1 and s["new_stereo"] > 100 and s["stereo_rows"] > 1000, s


def test_stereo_teacher_forced_moving_mask_and_right_calib(fe, synth):
    """Masks (the right working mask is a clone of the LEFT mask, TrackKLT.cpp:691) and a right camera with its own
    intrinsics (undistortion and RANSAC threshold per camera, :866-872)."""
    kr = tuple(v * 1.01 for v in synth.KAIST_K)
    _assert_parity(_run(fe, synth, 20, dict(CFG1, grid_y=3, pyr_levels=5), seed=1002, moving_mask=True, K_right=kr))


@pytest.mark.parametrize("hist", [0, 2])
def test_stereo_teacher_forced_other_preprocessing(fe, synth, hist):
    """histogram_method NONE and CLAHE (TrackKLT.cpp:57-67 runs per image of the message)."""
    _assert_parity(_run(fe, synth, 8, dict(CFG1, histogram_method=hist), seed=1006 + hist))


def test_stereo_teacher_forced_odd_size_and_wide_window(fe, synth):
    """641 x 361 (odd sides: ragged pyramid levels, REFLECT_101 borders), 21 x 21 window = the generic LK kernel."""
    _assert_parity(_run(fe, synth, 8, dict(CFG1, num_features=150, pyr_levels=3, win_size=21), seed=1008, width=641, height=361, hard=False))


def test_stereo_bad_arguments_and_setters(fe, synth):
    with pytest.raises(fe.FrontEndError):
        fe.StereoFrontEnd(fe.default_config(win_size=14))
    with pytest.raises(fe.FrontEndError):
        fe.StereoFrontEnd(fe.default_config(width=641, height=361, use_lines=1))   # line tracker needs even sides
    seq = synth.SynthSequence(seed=1015, width=640, height=280, n_frames=4, hard=False)
    kw = dict(CFG1, num_features=100)
    g = fe.StereoFrontEnd(fe.default_config(width=640, height=280, K=seq.K, D=seq.D, lookahead=1, use_lines=0, **kw))
    with pytest.raises(fe.FrontEndError):
        g.feed_new_camera(0.0, np.zeros((100, 100), np.uint8), np.zeros((100, 100), np.uint8))   # the reference exit()s here
    with pytest.raises(fe.FrontEndError):
        g.feed_new_camera(0.0, seq.frame(0), np.zeros((100, 100), np.uint8))
    with pytest.raises(fe.FrontEndError):
        g.collect()                                                                   # nothing submitted
    g.submit(1.0, seq.frame(0, 0), seq.frame(0, 1))
    g.submit(1.1, seq.frame(1, 0), seq.frame(1, 1))
    g.submit(1.2, seq.frame(2, 0), seq.frame(2, 1))                                    # lookahead + 2 slots, none holds a "last" pair yet
    with pytest.raises(fe.FrontEndError):
        g.submit(1.3, seq.frame(3, 0), seq.frame(3, 1))                                # lookahead window full
    with pytest.raises(fe.FrontEndError):
        g.set_num_features(50)                                                        # pairs pending
    g.collect()
    g.collect()
    g.collect()
    ids = g.get_last_ids()
    shared = sorted(set(ids[0].tolist()) & set(ids[1].tolist()))
    assert shared, "no stereo features"
    g.change_feat_id(shared[0], 10 ** 6)                                              # TrackBase.cpp:267-285, both cameras
    ids2 = g.get_last_ids()
    assert 10 ** 6 in ids2[0] and 10 ** 6 in ids2[1] and shared[0] not in ids2[0] and shared[0] not in ids2[1]
    # set_num_features between pairs: the oracle with the same change stays in step
    o = ost.TrackKLTStereo(ofe.FeConfig(K=seq.K, D=seq.D, **kw))
    g2 = fe.StereoFrontEnd(fe.default_config(width=640, height=280, K=seq.K, D=seq.D, **kw))
    z = np.zeros((280, 640), np.uint8)
    for t in range(4):
        if t == 2:
            g2.set_num_features(160)
            o.cfg.num_features = 160
        ro = o.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1), z, z)
        g2.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1))
        for cam in (0, 1):
            assert [int(v) for v in g2.point_rows(cam)["id"]] == [r.id for r in ro[cam]], (t, cam)
            assert np.array_equal(g2.get_last_ids()[cam], np.array(o.ids_last[cam], np.uint64)), (t, cam)
    g.close()
    g2.close()


def test_stereo_online_calibration_update(fe, synth):
    """Intrinsics are refined online (StateHelper.cpp:166): set_calib per camera between pairs changes the normalised
    coordinates of the rows and the RANSAC threshold from the next pair on, exactly as in the oracle."""
    seq = synth.SynthSequence(seed=1018, width=640, height=280, n_frames=6, hard=False)
    kw = dict(CFG1, num_features=120)
    o = ost.TrackKLTStereo(ofe.FeConfig(K=seq.K, D=seq.D, **kw))
    g = fe.StereoFrontEnd(fe.default_config(width=640, height=280, K=seq.K, D=seq.D, use_lines=0, **kw))
    z = np.zeros((280, 640), np.uint8)
    worst = 0.0
    for t in range(6):
        if t in (2, 4):
            for cam in (0, 1):
                K = tuple(v * (1.0 + 0.01 * (t + cam)) for v in seq.K)
                D = tuple(v * (1.0 - 0.05 * (t + cam)) for v in seq.D)
                o.set_calib(cam, K, D)
                g.set_calib(cam, K, D)
        ro = o.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1), z, z)
        g.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1))
        for cam in (0, 1):
            rows = g.point_rows(cam)
            assert [int(v) for v in rows["id"]] == [r.id for r in ro[cam]], (t, cam)
            if len(rows):
                un_o = np.array([[r.un, r.vn] for r in ro[cam]], np.float32)
                worst = max(worst, float(np.abs(np.stack([rows["un"], rows["vn"]], 1) - un_o).max()))
    g.close()
    assert worst < 1e-4, worst


def test_stereo_free_running(fe, synth):
    s = _run(fe, synth, 25, CFG1, seed=1003, teacher_forced=False)
    assert s["rows_sym"] <= 0.005 * s["rows"], s
    assert s["duv_p99"] < 0.05, s


def test_stereo_small_image_and_reset(fe, synth):
    """640x280, then a black pair: every track dies in both cameras, the tracker resets and re-detects identically."""
    seq = synth.SynthSequence(seed=1012, width=640, height=280, n_frames=6, hard=False)
    kw = dict(CFG1, num_features=120)
    o = ost.TrackKLTStereo(ofe.FeConfig(K=seq.K, D=seq.D, **kw))
    g = fe.StereoFrontEnd(fe.default_config(width=640, height=280, K=seq.K, D=seq.D, **kw))
    z = np.zeros((280, 640), np.uint8)
    frames = [(seq.frame(t, 0), seq.frame(t, 1)) for t in range(4)] + [(z, z), (seq.frame(4, 0), seq.frame(4, 1)),
                                                                         (seq.frame(5, 0), seq.frame(5, 1))]
    for t, (il, ir) in enumerate(frames):
        ro = o.feed_new_camera(1.0 + t, il, ir, z, z)
        info = g.feed_new_camera(1.0 + t, il, ir)
        for cam in (0, 1):
            assert [int(v) for v in g.point_rows(cam)["id"]] == [r.id for r in ro[cam]], (t, cam)
            assert np.array_equal(g.get_last_ids()[cam], np.array(o.ids_last[cam], np.uint64)), (t, cam)
        assert bool(info.reset) == bool(o.trace["reset"])
    g.close()


def test_stereo_state_roundtrip_and_pipelined(fe, synth):
    """get_state / set_state round trip; submit / collect with lookahead gives the rows of feed()."""
    seq = synth.SynthSequence(seed=1013, n_frames=10)
    cfg = dict(width=1280, height=560, K=seq.K, D=seq.D, **CFG1)
    a = fe.StereoFrontEnd(fe.default_config(**cfg))
    b = fe.StereoFrontEnd(fe.default_config(lookahead=3, **cfg))
    ref = []
    for t in range(8):
        a.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1))
        ref.append((a.point_rows(0).copy(), a.point_rows(1).copy()))
        if t == 3:
            blob = a.get_state()
    # pipelined
    frames = [(seq.frame(t, 0), seq.frame(t, 1)) for t in range(8)]
    sub = 0
    for t in range(8):
        while sub < 8 and sub <= t + 3:
            b.submit(seq.timestamp(sub), frames[sub][0], frames[sub][1])
            sub += 1
        b.collect()
        for cam in (0, 1):
            assert np.array_equal(b.point_rows(cam), ref[t][cam]), (t, cam)
    # resume from the checkpoint taken after pair 3
    c = fe.StereoFrontEnd(fe.default_config(**cfg))
    c.set_state(blob)
    assert c.get_state() == blob
    for t in range(4, 8):
        c.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1))
        for cam in (0, 1):
            assert np.array_equal(c.point_rows(cam), ref[t][cam]), (t, cam)
    for h in (a, b, c):
        h.close()


def test_stereo_with_left_image_lines(fe, synth):
    """Stereo rig + line tracker: TrackLSD runs its monocular code on the LEFT image against the stereo tracker's left
    points (TrackLSD.cpp:57-60, :127-129).  Teacher-forced; line ids / classes / attached point ids identical."""
    from test_frontend_gpu import _compare_lines
    W, H, n = 1280, 560, 12
    seq = synth.SynthSequence(seed=1014, width=W, height=H, n_frames=n)
    o = ost.StereoFrontEnd(ofe.FeConfig(K=seq.K, D=seq.D, use_lines=True, **CFG1))
    g = fe.StereoFrontEnd(fe.default_config(width=W, height=H, K=seq.K, D=seq.D, use_lines=1, **CFG1))
    n_line_rows = line_frames = line_equal = 0
    for t in range(n):
        il, ir, vps = seq.frame(t, 0), seq.frame(t, 1), seq.vanishing_points(t)
        if t > 0:
            st = o.klt.get_state()
            l = o.lsd.get_state()
            left = fe.pack_state(W, H, st["currid"], st["pts_last"][0], st["ids_last"][0], st["img_last"][0], None, l["currid"],
                                 l["lines_last"], l["ids_last"], l["pol_last"])
            right = fe.pack_state(W, H, st["currid"], st["pts_last"][1], st["ids_last"][1], st["img_last"][1], None)
            g.set_state(fe.pack_stereo_state(st["currid"], left, right))
        rl, rr, lrow_o = o.feed(seq.timestamp(t), il, ir, None, None, vps)
        info = g.feed_new_camera(seq.timestamp(t), il, ir, vanishing_points=vps)
        same_points = all([int(v) for v in g.point_rows(c)["id"]] == [r.id for r in rows] for c, rows in ((0, rl), (1, rr)))
        if same_points:    # a flipped point status legitimately changes the point-on-line sets of that frame
            lrows, lpts = g.line_rows()
            assert info.n_line_rows == len(lrows)
            line_frames += 1
            n_line_rows += len(lrow_o)
            line_equal += int(_compare_lines(lrows, lpts, lrow_o))
    g.close()
    assert line_frames >= n - 2 and n_line_rows > 20, (line_frames, n_line_rows)
    assert line_equal == line_frames, (line_equal, line_frames)


def test_stereo_pipelined_with_batched_lines(fe, synth):
    """Stereo rig + left-image line tracker on a pipelined handle (lookahead 8: the line paths of 4 consecutive left images
    share their launches, speculative LK is on for neither camera — the stereo state machine launches its own): rows equal
    the synchronous handle's bit for bit, including when collect() arrives before a line batch is full."""
    n = 10
    seq = synth.SynthSequence(seed=1017, n_frames=n)
    cfg = dict(width=1280, height=560, K=seq.K, D=seq.D, use_lines=1, **CFG1)
    a = fe.StereoFrontEnd(fe.default_config(**cfg))
    ref = []
    for t in range(n):
        a.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1), vanishing_points=seq.vanishing_points(t))
        lr, lp = a.line_rows()
        ref.append((a.point_rows(0).copy(), a.point_rows(1).copy(), lr.copy(), lp.copy()))
    a.close()
    assert sum(len(r[2]) for r in ref) > 30
    b = fe.StereoFrontEnd(fe.default_config(lookahead=8, **cfg))
    frames = [(seq.frame(t, 0), seq.frame(t, 1)) for t in range(n)]
    sub = col = 0
    for burst in (2, 1, 3, 9, 9, 9, 9, 9, 9, 9):
        for _ in range(burst):
            if sub < n and sub - col <= 8:
                b.submit(seq.timestamp(sub), frames[sub][0], frames[sub][1], vanishing_points=seq.vanishing_points(sub))
                sub += 1
        if col < sub:
            info = b.collect()
            lr, lp = b.line_rows()
            assert np.array_equal(b.point_rows(0), ref[col][0]) and np.array_equal(b.point_rows(1), ref[col][1]), col
            assert info.n_line_rows == len(ref[col][2]) and np.array_equal(lr, ref[col][2]) and np.array_equal(lp, ref[col][3]), col
            col += 1
    while col < n:
        b.collect()
        assert np.array_equal(b.point_rows(0), ref[col][0]), col
        assert np.array_equal(b.line_rows()[0], ref[col][2]), col
        col += 1
    b.close()


def test_stereo_golden_rows(fe, synth):
    """Rows of the cv2-driven oracle committed as a fixture (tests/golden/make_golden_stereo.py)."""
    gold = np.load(GOLDEN)
    n, W, H = int(gold["n_frames"]), int(gold["width"]), int(gold["height"])
    seq = synth.SynthSequence(seed=int(gold["seed"]), width=W, height=H, n_frames=n, hard=False)
    kw = dict(CFG1, num_features=int(gold["num_features"]))
    g = fe.StereoFrontEnd(fe.default_config(width=W, height=H, K=seq.K, D=seq.D, **kw))
    for t in range(n):
        g.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1))
        for cam in (0, 1):
            rows = g.point_rows(cam)
            ids = gold["ids_%d_%d" % (t, cam)]
            assert np.array_equal(rows["id"], ids), (t, cam)
            if len(ids):
                uv = gold["uv_%d_%d" % (t, cam)]
                assert np.abs(np.stack([rows["u"], rows["v"]], 1) - uv).max() < 0.05, (t, cam)
    g.close()
