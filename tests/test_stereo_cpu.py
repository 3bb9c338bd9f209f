"""CPU checks of the stereo oracle (oracle/stereo.py): the committed golden rows, the reference's quirks, and the
structural invariants of feed_stereo (TrackKLT.cpp:202-393)."""
import os

import numpy as np

from oracle import frontend as ofe
from oracle import stereo as ost
from plviwo_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "stereo_golden.npz")
KW = dict(num_features=120, fast_threshold=20, grid_x=5, grid_y=5, min_px_dist=10, pyr_levels=3, win_size=15)


def _seq(gold):
    return synth.SynthSequence(seed=int(gold["seed"]), width=int(gold["width"]), height=int(gold["height"]),
                               n_frames=int(gold["n_frames"]), hard=False)


def test_oracle_matches_golden_rows():
    gold = np.load(GOLDEN)
    seq = _seq(gold)
    o = ost.TrackKLTStereo(ofe.FeConfig(K=seq.K, D=seq.D, **dict(KW, num_features=int(gold["num_features"]))))
    z = np.zeros((seq.H, seq.W), np.uint8)
    for t in range(int(gold["n_frames"])):
        rows = o.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1), z, z)
        for cam in (0, 1):
            assert np.array_equal(np.array([r.id for r in rows[cam]], np.uint64), gold["ids_%d_%d" % (t, cam)]), (t, cam)
            uv = np.array([[r.u, r.v] for r in rows[cam]], np.float32).reshape(-1, 2)
            assert np.array_equal(uv, gold["uv_%d_%d" % (t, cam)]), (t, cam)


def test_stereo_invariants():
    gold = np.load(GOLDEN)
    seq = _seq(gold)
    o = ost.TrackKLTStereo(ofe.FeConfig(K=seq.K, D=seq.D, **KW))
    z = np.zeros((seq.H, seq.W), np.uint8)
    first_id = o.currid + 1
    n_stereo = 0
    for t in range(5):
        rows = o.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1), z, z)
        for cam in (0, 1):
            ids = [r.id for r in rows[cam]]
            assert len(ids) == len(set(ids))                       # one observation per feature and camera
            assert t == 0 or ids == o.ids_last[cam]              # rows are exactly the new pts_last (:352-378)
            assert all(first_id <= i <= o.currid for i in ids)     # one id counter for both cameras
        n_stereo += len(set(r.id for r in rows[0]) & set(r.id for r in rows[1]))
    assert n_stereo > 100
    # the synthetic rig has a 9 px disparity (at 1280 px width): a stereo feature's right observation sits left of the left one
    l = {r.id: r for r in rows[0]}
    dx = [l[r.id].u - r.u for r in rows[1] if r.id in l]
    assert abs(float(np.median(dx)) - 9.0 * seq.sc) < 1.0


def test_right_working_mask_is_a_clone_of_the_left_mask():
    """TrackKLT.cpp:691 `mask1_updated = mask0.clone()`: a region masked in the LEFT image yields no new right-only
    corners even though the right mask is empty (while existing right points are tested against the right mask)."""
    seq = synth.SynthSequence(seed=1021, width=640, height=280, n_frames=2, hard=False)
    left_mask = np.zeros((280, 640), np.uint8)
    left_mask[:, 320:] = 255
    z = np.zeros_like(left_mask)
    o = ost.TrackKLTStereo(ofe.FeConfig(K=seq.K, D=seq.D, **KW))
    o.feed_new_camera(1.0, seq.frame(0, 0), seq.frame(0, 1), left_mask, z)
    # every left point avoids the masked half; right points come from left->right tracks (x < 320) or from the right
    # detection, which must avoid x >= 320 as well because it works on the left mask's clone
    assert len(o.pts_last[0]) > 10 and float(o.pts_last[0][:, 0].max()) < 323
    assert len(o.pts_last[1]) > 10 and float(o.pts_last[1][:, 0].max()) < 323   # cornerSubPix may move a corner by a pixel or two
    ref = ost.TrackKLTStereo(ofe.FeConfig(K=seq.K, D=seq.D, **KW))
    ref.feed_new_camera(1.0, seq.frame(0, 0), seq.frame(0, 1), z, z)
    assert float(ref.pts_last[1][:, 0].max()) > 330
