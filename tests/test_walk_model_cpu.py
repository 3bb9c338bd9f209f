"""CPU model of the chain walk's parallelisation (no GPU, no CUDA code involved): the two observations DESIGN.md §4 builds on,
and the decision tables of pl-viwo_b200/csrc/kernels_lines.cu (init_fld_constants), checked against the sequential walk of
FastLineDetector as oracle/csrc/oracle_shim.cpp restates it (Fld::detect / get_point_chain).

  1. a walk never leaves the 8-connected component of its seed, so "walk every component on its own, list the chains by the
     raster index of their seed" gives the chains of the sequential raster-order walk;
  2. one step of getPointChain is a function of the 3 x 3 neighbourhood, the running direction and min(step, 7): the two
     tables (neighbourhood key x direction -> neighbour, step x direction x neighbour -> direction).

The component walk below is the form k_fld_walk_thread uses: one bit-row per image row of the component's bounding box (a zero
column on either side, a zero row above and below), seeds = lowest set bit of the first non-empty row."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
ndimage = pytest.importorskip("scipy.ndimage")

import plviwo_b200  # noqa: F401  (import shim)
from plviwo_b200 import synth
from oracle import cvops

T = 20  # length_threshold (TrackLSD.h:269)
NB = [(1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1), (0, 1)]   # neighbour i -> (dr, dc), oracle_shim.cpp:94


def sequential_chains(edges):
    """Fld::detect's walk (oracle_shim.cpp:212-256), chains only."""
    img = (edges > 0).astype(np.uint8)
    H, W = img.shape
    chains = []
    for r in range(H):
        for c in range(W):
            if not img[r, c]:
                continue
            pts = [(c, r)]
            img[r, c] = 0
            x, y, direction, step = c, r, 0, 0
            while True:
                best, best_dir, min_diff = None, 0, 7.0
                found_first = None
                for i, (dr, dc) in enumerate(NB):
                    ci, ri = x + dc, y + dr
                    if ri < 0 or ri == H or ci < 0 or ci == W or not img[ri, ci]:
                        continue
                    cur = i - 8 if i > 4 else i
                    if step == 0:
                        found_first = (ci, ri, cur)
                        break
                    d = abs(cur - direction)
                    d = 8 - d if d > 4 else d
                    if d <= min_diff:
                        min_diff, best, best_dir = float(d), (ci, ri), cur
                if step == 0:
                    if found_first is None:
                        break
                    x, y, direction = found_first
                elif min_diff < 2:
                    x, y = best
                    direction = int((direction * step + best_dir) / (step + 1))   # C integer division truncates towards zero
                else:
                    break
                pts.append((x, y))
                step += 1
                img[y, x] = 0
            if len(pts) >= T + 1:
                chains.append(pts)
    return chains


def tables():
    """init_fld_constants (kernels_lines.cu): lut1[first][key][direction + 3] -> i | (dr + 1) << 4 | (dc + 1) << 6 (i == 8: stop),
    lut2[min(step, 7)][direction + 3][i] -> new direction + 3."""
    key_bit = [7, 6, 5, 3, 0, 1, 2, 4]

    def choose(mask, direction):
        best, min_diff = 8, 7
        for i in range(8):
            if not (mask >> i) & 1:
                continue
            cur = i - 8 if i > 4 else i
            d = abs(cur - direction)
            d = 8 - d if d > 4 else d
            if d <= min_diff:
                min_diff, best = d, i
        return best if min_diff < 2 else 8

    lut1 = np.zeros((2, 256, 8), np.uint8)
    for first in range(2):
        for key in range(256):
            mask = 0
            for i in range(8):
                mask |= ((key >> key_bit[i]) & 1) << i
            for d in range(8):
                i = 8
                if mask:
                    i = (mask & -mask).bit_length() - 1 if first else choose(mask, d - 3)
                dr, dc = NB[i] if i < 8 else (0, 0)
                lut1[first, key, d] = i | ((dr + 1) << 4) | ((dc + 1) << 6)
    lut2 = np.zeros((8, 8, 8), np.uint8)
    for sc in range(8):
        for d in range(8):
            for i in range(8):
                cd = i - 8 if i > 4 else i
                nd = cd if sc == 0 else int(((d - 3) * sc + cd) / (sc + 1))
                lut2[sc, d, i] = nd + 3
    return lut1, lut2


def component_chains(edges):
    """Every component walked on its own through the tables, chains listed by the raster index of their seed."""
    lut1, lut2 = tables()
    H, W = edges.shape
    lab, n = ndimage.label(edges > 0, structure=np.ones((3, 3)))
    out = []
    for k, sl in enumerate(ndimage.find_objects(lab), start=1):
        comp = lab[sl] == k
        if comp.sum() < T + 1:
            continue
        y0, x0 = sl[0].start, sl[1].start
        bh = comp.shape[0]
        rows = [0] * (bh + 2)   # bit b of row r = pixel (x0 - 1 + b, y0 - 1 + r)
        for r in range(bh):
            v = 0
            for b in np.nonzero(comp[r])[0]:
                v |= 1 << (int(b) + 1)
            rows[r + 1] = v
        sy = 1
        while sy <= bh:
            R = rows[sy]
            if R == 0:
                sy += 1
                continue
            cx, cy = (R & -R).bit_length() - 1, sy
            rows[sy] = R & ~(1 << cx)
            seed = (y0 - 1 + cy) * W + x0 - 1 + cx
            pts, step, dsel, first = [], 0, 0, 1
            while True:
                pts.append((x0 - 1 + cx, y0 - 1 + cy))
                t3, c3, b3 = (rows[cy - 1] >> (cx - 1)) & 7, (rows[cy] >> (cx - 1)) & 7, (rows[cy + 1] >> (cx - 1)) & 7
                key = t3 | ((c3 & 1) << 3) | ((c3 >> 2) << 4) | (b3 << 5)
                e = int(lut1[first, key, dsel])
                i = e & 15
                if i == 8:
                    break
                dsel = int(lut2[min(step, 7), dsel, i])
                step += 1
                first = 0
                cy += ((e >> 4) & 3) - 1
                cx += ((e >> 6) & 3) - 1
                rows[cy] &= ~(1 << cx)
            if len(pts) >= T + 1:
                out.append((seed, pts))
    out.sort(key=lambda sp: sp[0])
    return [p for _, p in out]


@pytest.mark.parametrize("seed,line_heavy", [(31, False), (31, True), (1000, False)])
def test_component_walk_equals_sequential_walk(seed, line_heavy):
    seq = synth.SynthSequence(seed=seed, n_frames=3, line_heavy=line_heavy)
    half = cvops.half_res(cvops.equalize_hist(seq.frame(1)))
    edges = cvops.canny(half).copy()
    edges[:6, :6] = 0          # FastLineDetector clears the two corner blocks before walking
    edges[-5:, -5:] = 0
    a = sequential_chains(edges)
    b = component_chains(edges)
    assert len(a) > 20
    assert len(a) == len(b)
    for k, (ca, cb) in enumerate(zip(a, b)):
        assert ca == cb, "chain %d differs (seed %r vs %r, lengths %d / %d)" % (k, ca[0], cb[0], len(ca), len(cb))
