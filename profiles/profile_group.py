#!/usr/bin/env python
"""Short stream-group run for ncu: `ticks` ticks of `streams` camera streams (BASELINE.json configs[1] front end each)
through plviwo_fe_group_submit / _collect.  Used for the launch list and the --set full captures; never a source of bench
numbers.  Streams replay the frames of a few synthetic sequences from different start frames."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plviwo_b200 as fe  # noqa: E402
from plviwo_b200 import synth  # noqa: E402

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 6
lookahead = int(sys.argv[3]) if len(sys.argv) > 3 else 2
nseq = min(streams, 4)
nfr = ticks + 8
seqs = [synth.SynthSequence(seed=1000 + k, width=1280, height=560, n_frames=nfr) for k in range(nseq)]
frames = [[q.frame(t) for t in range(nfr)] for q in seqs]
cfg = fe.default_config(width=1280, height=560, num_features=400, fast_threshold=20, grid_x=5, grid_y=5, min_px_dist=10,
                        pyr_levels=4, win_size=15, lookahead=lookahead)
g = fe.GroupFrontEnd(cfg, streams, calibs=[(seqs[s % nseq].K, seqs[s % nseq].D) for s in range(streams)])
vps = [seqs[s % nseq].vanishing_points(0) for s in range(streams)]
sub = 0
for i in range(ticks):
    while sub < ticks and sub <= i + lookahead:
        g.submit([1.0 + 0.1 * sub] * streams, [frames[s % nseq][sub + (s // nseq) % 8] for s in range(streams)], vanishing_points=vps)
        sub += 1
    infos = g.collect()
print("profile_group: %d streams x %d ticks, last tick: %d point rows, %d line rows on stream 0" %
      (streams, ticks, infos[0].n_point_rows, infos[0].n_line_rows))
