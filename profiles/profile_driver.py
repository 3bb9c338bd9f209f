#!/usr/bin/env python
"""Short front-end run for ncu: `frames` frames of BASELINE.json configs[1] (1280x560, 400 pts, maxLevel 4, lines on)
through plviwo_fe_submit / plviwo_fe_collect.  Used by profiles/capture.sh; never a source of bench numbers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plviwo_b200 as fe  # noqa: E402
from plviwo_b200 import synth  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
lookahead = int(sys.argv[2]) if len(sys.argv) > 2 else 2
seq = synth.SynthSequence(seed=1000, width=1280, height=560, n_frames=frames)
h = fe.FrontEnd(fe.default_config(width=1280, height=560, K=seq.K, D=seq.D, num_features=400, fast_threshold=20, grid_x=5,
                                  grid_y=5, min_px_dist=10, pyr_levels=4, win_size=15, lookahead=lookahead))
imgs = [seq.frame(t) for t in range(frames)]
sub = 0
for i in range(frames):
    while sub < frames and sub <= i + lookahead:
        h.submit(seq.timestamp(sub), imgs[sub], vanishing_points=seq.vanishing_points(sub))
        sub += 1
    info = h.collect()
print("profile_driver: %d frames, last frame %d point rows, %d line rows" % (frames, info.n_point_rows, info.n_line_rows))
