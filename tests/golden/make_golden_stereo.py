"""Generates tests/golden/stereo_golden.npz: the database rows of the stereo point front end on a small synthetic
sequence, computed by the oracle (oracle/stereo.py) driving the real OpenCV kernels (cv2) in this container.
Run from the repo root:  python tests/golden/make_golden_stereo.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import plviwo_b200  # noqa: E402
from plviwo_b200 import synth  # noqa: E402
from oracle import frontend as ofe, stereo as ost  # noqa: E402

SEED, W, H, N, NF = 1020, 640, 280, 8, 120
seq = synth.SynthSequence(seed=SEED, width=W, height=H, n_frames=N, hard=False)
kw = dict(num_features=NF, fast_threshold=20, grid_x=5, grid_y=5, min_px_dist=10, pyr_levels=3, win_size=15)
o = ost.TrackKLTStereo(ofe.FeConfig(K=seq.K, D=seq.D, **kw))
out = dict(seed=SEED, width=W, height=H, n_frames=N, num_features=NF)
z = np.zeros((H, W), np.uint8)
for t in range(N):
    rows = o.feed_new_camera(seq.timestamp(t), seq.frame(t, 0), seq.frame(t, 1), z, z)
    for cam in (0, 1):
        out["ids_%d_%d" % (t, cam)] = np.array([r.id for r in rows[cam]], np.uint64)
        out["uv_%d_%d" % (t, cam)] = np.array([[r.u, r.v] for r in rows[cam]], np.float32).reshape(-1, 2)
        out["un_%d_%d" % (t, cam)] = np.array([[r.un, r.vn] for r in rows[cam]], np.float32).reshape(-1, 2)
    print(t, len(rows[0]), len(rows[1]))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stereo_golden.npz"), **out)
