"""ORACLE (test infrastructure): line-by-line CPU restatement of the reference's visual front end glue.

Follows (all paths under /root/reference):
  open_vins/ov_core/src/track/TrackKLT.cpp:34-200   feed_new_camera / feed_monocular
  open_vins/ov_core/src/track/TrackKLT.cpp:395-528  perform_detection_monocular
  open_vins/ov_core/src/track/TrackKLT.cpp:829-886  perform_matching
  open_vins/ov_core/src/track/Grider_GRID.h:74-180  perform_griding (+ Grider_FAST.h:57 comparator)
  open_vins/ov_core/src/track/TrackBase.cpp:30-41   currid initialisation
  open_vins/ov_core/src/cam/CamBase.h:108-136, cam/CamRadtan.h:99-120   undistort_cv / undistort_line
  PL-VIWO/src/update/cam/TrackLSD.cpp:30-37, 70-192, 194-236, 318-366, 368-407, 435-448, 744-830
  open_vins/ov_core/src/feat/FeatureDatabase.cpp:60-85 and
  PL-VIWO/src/update/cam/linefeat/LineFeatureDatabase.cpp:40-76 (the rows the trackers emit)

GUI calls (imshow / waitKey / drawing, TrackLSD.cpp:171-172, 243-282) are not part of the contract and are left
out.  The arithmetic of the OpenCV calls comes from an ``ops`` module: ``oracle.cvops`` (real OpenCV through cv2)
or ``oracle.npops`` (NumPy restatement).  float32 intermediate types are kept where the C++ has ``float``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import cvops

f32 = np.float32

HIST_NONE, HIST_HISTOGRAM, HIST_CLAHE = 0, 1, 2  # TrackBase.h:78


@dataclass
class FeConfig:
    """Front-end knobs: options/OptionsCamera.h:77-108 plus the constants hard-coded in TrackKLT.h:143-144,
    TrackKLT.cpp:416,468,857,873, Grider_GRID.h:163-165 and TrackLSD.h:269-273 / TrackLSD.cpp:231,780,824,361."""
    num_features: int = 200
    fast_threshold: int = 20
    grid_x: int = 5
    grid_y: int = 5
    min_px_dist: int = 10
    pyr_levels: int = 5          # OpenCV maxLevel => pyr_levels + 1 images
    win_size: int = 15
    histogram_method: int = HIST_HISTOGRAM
    K: Tuple[float, float, float, float] = (816.90378992770002, 811.56803828490001, 608.50726281690004, 263.47599764440002)
    D: Tuple[float, float, float, float] = (-5.6143027800000002e-02, 1.3952563200000001e-01,
                                            -1.2155906999999999e-03, -9.7281389999999998e-04)
    use_lines: bool = True
    numaruco: int = 0
    # line detector (TrackLSD.h:269-273) and association thresholds
    fld_length_threshold: int = 20
    fld_distance_threshold: float = 1.414213562
    canny_th1: float = 50.0
    canny_th2: float = 50.0
    line_min_length: float = 40.0      # TrackLSD.cpp:231
    # extension (BASELINE.json config 3; NOT in the reference): LK over points sampled along last frame's segments
    line_samples: int = 0
    # UpdaterCamera::feed_measurement's pre-step (UpdaterCamera.cpp:86-95, options/OptionsCamera.h:74)
    downsample: bool = False


@dataclass
class PointRow:
    """One FeatureDatabase::update_feature call (TrackKLT.cpp:176-179)."""
    id: int
    u: float
    v: float
    un: float
    vn: float


@dataclass
class LineRow:
    """One LineFeatureDatabase::update_feature call (TrackLSD.cpp:163-167)."""
    id: int
    line: np.ndarray        # (4,) f32 x1,y1,x2,y2 full-res px
    line_n: np.ndarray      # (4,) f32 normalised
    pids: List[int]         # keys of point_on_lines (ascending)
    dists: List[float]      # values of point_on_lines
    pts: np.ndarray         # (k,2) f32 point_position
    D: int


class TrackKLT:
    """ov_core::TrackKLT (monocular), restated."""

    def __init__(self, cfg: FeConfig, ops=cvops):
        self.cfg = cfg
        self.ops = ops
        self.currid = 4 * cfg.numaruco + 1          # TrackBase.cpp:34
        self.pts_last = np.zeros((0, 2), f32)
        self.ids_last: List[int] = []
        self.img_last: Optional[np.ndarray] = None   # equalised level-0 image of the previous frame
        self.mask_last: Optional[np.ndarray] = None
        self.trace: Dict[str, object] = {}
        self.K, self.D = tuple(cfg.K), tuple(cfg.D)

    # -- state (teacher forcing / checkpoint)
    def get_state(self):
        return dict(currid=self.currid, pts_last=self.pts_last.copy(), ids_last=list(self.ids_last),
                    img_last=None if self.img_last is None else self.img_last.copy(),
                    mask_last=None if self.mask_last is None else self.mask_last.copy())

    def set_state(self, st):
        self.currid = int(st["currid"])
        self.pts_last = np.asarray(st["pts_last"], f32).reshape(-1, 2).copy()
        self.ids_last = [int(i) for i in st["ids_last"]]
        self.img_last = None if st["img_last"] is None else st["img_last"].copy()
        self.mask_last = None if st["mask_last"] is None else st["mask_last"].copy()

    def set_calib(self, K, D):
        self.K, self.D = tuple(K), tuple(D)

    def get_last_obs(self):
        return self.pts_last.copy()

    def get_last_ids(self):
        return list(self.ids_last)

    # -- TrackKLT.cpp:34-94 + 96-200
    def feed_new_camera(self, timestamp: float, img: np.ndarray, mask: np.ndarray) -> List[PointRow]:
        cfg = self.cfg
        self.trace = {}
        if cfg.histogram_method == HIST_HISTOGRAM:
            img_eq = self.ops.equalize_hist(img)
        elif cfg.histogram_method == HIST_CLAHE:
            img_eq = self.ops.clahe(img)
        else:
            img_eq = img
        self.trace["img_eq"] = img_eq
        rows: List[PointRow] = []

        if len(self.pts_last) == 0:                               # :110-123
            pts, ids = self._perform_detection(img_eq, mask, np.zeros((0, 2), f32), [])
            self.img_last, self.mask_last = img_eq, mask
            self.pts_last, self.ids_last = pts, ids
            self.trace["reset"] = False
            return rows

        # top-off on the PREVIOUS image (:127-130)
        pts_old, ids_old = self._perform_detection(self.img_last, self.mask_last, self.pts_last.copy(), list(self.ids_last))
        self.trace["pts_old"], self.trace["ids_old"] = pts_old.copy(), list(ids_old)
        pts_new, mask_ll = self._perform_matching(self.img_last, img_eq, pts_old, pts_old.copy())
        if mask_ll is None:                                       # :143-152 (only when there are no points)
            self.img_last, self.mask_last = img_eq, mask
            self.pts_last, self.ids_last = np.zeros((0, 2), f32), []
            self.trace["reset"] = True
            return rows
        H, W = img_eq.shape
        # :159-173, element-wise on arrays (int() of the C++ = truncation towards zero = astype(int))
        xs, ys = pts_new[:, 0], pts_new[:, 1]
        xi, yi = xs.astype(np.int64), ys.astype(np.int64)
        inb = ~((xs < 0) | (ys < 0) | (xi >= W) | (yi >= H))
        keep = inb.copy()
        keep[inb] &= ~(mask[yi[inb], xi[inb]] > 127)
        keep &= np.asarray(mask_ll, bool)
        sel = np.nonzero(keep)[0]
        good = np.ascontiguousarray(pts_new[sel], f32).reshape(-1, 2)
        good_ids = [ids_old[i] for i in sel]
        und = self.ops.undistort(good, self.K, self.D)            # :176-179
        for i in range(len(good)):
            rows.append(PointRow(int(good_ids[i]), float(good[i, 0]), float(good[i, 1]), float(und[i, 0]), float(und[i, 1])))
        self.img_last, self.mask_last = img_eq, mask              # :182-189
        self.pts_last, self.ids_last = good, good_ids
        self.trace["reset"] = False
        return rows

    # -- TrackKLT.cpp:395-528
    def _perform_detection(self, img0: np.ndarray, mask0: np.ndarray, pts0: np.ndarray, ids0: List[int]):
        cfg = self.cfg
        H, W = img0.shape
        d = cfg.min_px_dist
        close_w = int(f32(W) / f32(d))
        close_h = int(f32(H) / f32(d))
        grid_close = np.zeros((close_h, close_w), np.uint8)
        size_x = f32(W) / f32(cfg.grid_x)
        size_y = f32(H) / f32(cfg.grid_y)
        grid_grid = np.zeros((cfg.grid_y, cfg.grid_x), np.uint8)
        mask_upd = mask0.copy()
        keep_idx = []
        # :411-464.  The per-point quantities are float32 / int expressions of the point alone and are evaluated for
        # all points at once (same element-wise arithmetic); only the order-dependent occupancy tests stay in the loop.
        P = np.asarray(pts0, f32).reshape(-1, 2)
        PX, PY = P[:, 0], P[:, 1]
        X, Y = PX.astype(np.int64), PY.astype(np.int64)
        edge = 10
        XC, YC = (PX / f32(d)).astype(np.int64), (PY / f32(d)).astype(np.int64)
        XG, YG = np.floor(PX / size_x).astype(np.int64), np.floor(PY / size_y).astype(np.int64)
        ok = ~((X < edge) | (X >= W - edge) | (Y < edge) | (Y >= H - edge))
        ok &= ~((XC < 0) | (XC >= close_w) | (YC < 0) | (YC >= close_h))
        ok &= ~((XG < 0) | (XG >= cfg.grid_x) | (YG < 0) | (YG >= cfg.grid_y))
        Xl, Yl, XCl, YCl, XGl, YGl = X.tolist(), Y.tolist(), XC.tolist(), YC.tolist(), XG.tolist(), YG.tolist()
        for k in np.nonzero(ok)[0].tolist():
            x, y, x_close, y_close = Xl[k], Yl[k], XCl[k], YCl[k]
            if grid_close[y_close, x_close] > 127:
                continue
            if mask0[y, x] > 127:
                continue
            grid_close[y_close, x_close] = 255
            x_grid, y_grid = XGl[k], YGl[k]
            if grid_grid[y_grid, x_grid] < 255:
                grid_grid[y_grid, x_grid] += 1
            if x - d >= 0 and x + d < W and y - d >= 0 and y + d < H:
                mask_upd[y - d:y + d + 1, x - d:x + d + 1] = 255   # cv::rectangle FILLED, both corners inclusive
            keep_idx.append(k)
        ids0 = [ids0[k] for k in keep_idx]
        pts0 = np.ascontiguousarray(P[keep_idx], f32).reshape(-1, 2)
        self.trace.setdefault("det", {})
        det = self.trace["det"] = {"pts_kept": pts0.copy(), "ran": False}

        num_featsneeded = cfg.num_features - len(pts0)            # :468-471
        if num_featsneeded < min(20, int(0.5 * cfg.num_features)):
            return pts0, ids0
        det["ran"] = True
        mask0_grid = self.ops.resize_nearest(mask0, cfg.grid_x, cfg.grid_y)   # :479-480
        nfg = int(float(cfg.num_features) / float(cfg.grid_x * cfg.grid_y)) + 1
        nfg_req = max(1, int(0.5 * nfg))
        valid_locs = []
        for x in range(cfg.grid_x):                               # x-major! (:486-492)
            for y in range(cfg.grid_y):
                if int(grid_grid[y, x]) < nfg_req and int(mask0_grid[y, x]) != 255:
                    valid_locs.append((x, y))
        det["valid_locs"] = list(valid_locs)
        ext = self._perform_griding(img0, mask_upd, valid_locs, det)
        new_pts = []
        for k in range(len(ext)):                                 # :497-512
            px, py = f32(ext[k, 0]), f32(ext[k, 1])
            xg = int(px / f32(d))
            yg = int(py / f32(d))
            if xg < 0 or xg >= close_w or yg < 0 or yg >= close_h:
                continue
            if grid_close[yg, xg] > 127:
                continue
            new_pts.append((px, py))
            grid_close[yg, xg] = 255
        new_ids = []
        for _ in new_pts:                                         # :519-527
            self.currid += 1
            new_ids.append(self.currid)
        det["new_pts"] = np.asarray(new_pts, f32).reshape(-1, 2)
        det["new_ids"] = list(new_ids)
        pts_out = np.concatenate([pts0, np.asarray(new_pts, f32).reshape(-1, 2)], 0)
        return pts_out, ids0 + new_ids

    # -- Grider_GRID.h:74-180
    def _perform_griding(self, img: np.ndarray, mask: np.ndarray, valid_locs, det) -> np.ndarray:
        cfg = self.cfg
        if not valid_locs:
            det["cells"] = []
            return np.zeros((0, 2), f32)
        H, W = img.shape
        gx, gy = cfg.grid_x, cfg.grid_y
        nf = cfg.num_features
        if nf < gx * gy:                                          # :88-92
            ratio = float(gx) / float(gy)
            gy = int(math.ceil(math.sqrt(nf / ratio)))
            gx = int(math.ceil(gy * ratio))
        nfg = int(float(nf) / float(gx * gy)) + 1
        size_x = W // gx
        size_y = H // gy
        out = []
        cells = []
        for (cx, cy) in valid_locs:                               # :108-151
            x, y = cx * size_x, cy * size_y
            if x + size_x > W or y + size_y > H:
                cells.append(None)
                continue
            xy, resp = self.ops.fast_cell(img[y:y + size_y, x:x + size_x], cfg.fast_threshold)
            perm = self.ops.sort_perm(resp)
            cells.append(dict(origin=(x, y), xy=xy, resp=resp, perm=perm))
            for i in range(min(nfg, len(perm))):
                p = perm[i]
                fx = f32(xy[p, 0]) + f32(x)
                fy = f32(xy[p, 1]) + f32(y)
                if int(fx) < 0 or int(fx) > W or int(fy) < 0 or int(fy) > H:
                    continue
                if mask[int(fy), int(fx)] > 127:
                    continue
                out.append((fx, fy))
        det["cells"] = cells
        pts = np.asarray(out, f32).reshape(-1, 2)
        det["selected"] = pts.copy()
        if len(pts) == 0:
            return pts
        ref = self.ops.corner_subpix(img, pts)                    # :163-179
        det["refined"] = ref.copy()
        return ref

    # -- TrackKLT.cpp:829-886
    def _perform_matching(self, img0, img1, pts0: np.ndarray, pts1: np.ndarray):
        cfg = self.cfg
        n = len(pts0)
        if n == 0:
            return pts1, None
        if n < 10:                                                # :848-852
            return pts1, np.zeros((n,), np.uint8)
        p1, mask_klt = self.ops.lk(img0, img1, pts0, pts1, cfg.win_size, cfg.pyr_levels)
        p0n = self.ops.undistort(pts0, self.K, self.D)
        p1n = self.ops.undistort(p1, self.K, self.D)
        maxf = max(self.K[0], self.K[1])
        mask_rsc = self.ops.find_fundamental_mask(p0n, p1n, 2.0 / maxf)
        rsc = np.zeros((n,), bool)
        rsc[:len(mask_rsc)] = np.asarray(mask_rsc, bool)[:n]
        out = (np.asarray(mask_klt, bool) & rsc).astype(np.uint8)
        self.trace["lk_pts1"] = p1.copy()
        self.trace["mask_klt"] = mask_klt.copy()
        self.trace["mask_rsc"] = mask_rsc.copy()
        self.trace["p0n"], self.trace["p1n"] = p0n, p1n
        return p1, out


# =============================================================================================== lines
def point_line_distance(line: np.ndarray, x0, y0) -> np.float32:
    """TrackLSD::PointLineDistance (TrackLSD.cpp:794-814): point-to-SEGMENT distance, float arithmetic; the last
    branch mixes in double through std::pow(float, int)."""
    x0, y0 = f32(x0), f32(y0)
    x1, y1, x2, y2 = f32(line[0]), f32(line[1]), f32(line[2]), f32(line[3])
    cross = (x2 - x1) * (x0 - x1) + (y2 - y1) * (y0 - y1)
    if cross <= 0:
        return f32(np.sqrt((x0 - x1) * (x0 - x1) + (y0 - y1) * (y0 - y1)))
    d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1)
    if cross > d:
        return f32(np.sqrt((x0 - x2) * (x0 - x2) + (y0 - y2) * (y0 - y2)))
    num = f32(abs((y2 - y1) * x0 + (x1 - x2) * y0 + ((x2 * y1) - (x1 * y2))))
    den = math.sqrt(float(y2 - y1) ** 2 + float(x1 - x2) ** 2)   # std::pow(float,int) -> double
    return f32(abs(float(num) / den))


def point_line_distance_vec(line: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """point_line_distance over arrays of points: the same float32 operations, element-wise (speed only)."""
    x0, y0 = xs.astype(f32), ys.astype(f32)
    x1, y1, x2, y2 = f32(line[0]), f32(line[1]), f32(line[2]), f32(line[3])
    cross = (x2 - x1) * (x0 - x1) + (y2 - y1) * (y0 - y1)
    d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1)
    d1 = np.sqrt((x0 - x1) * (x0 - x1) + (y0 - y1) * (y0 - y1)).astype(f32)
    d2 = np.sqrt((x0 - x2) * (x0 - x2) + (y0 - y2) * (y0 - y2)).astype(f32)
    num = np.abs((y2 - y1) * x0 + (x1 - x2) * y0 + ((x2 * y1) - (x1 * y2))).astype(f32)
    den = math.sqrt(float(y2 - y1) ** 2 + float(x1 - x2) ** 2)
    with np.errstate(all="ignore"):
        d3 = np.abs(num.astype(np.float64) / den).astype(f32)
    return np.where(cross <= 0, d1, np.where(cross > d, d2, d3)).astype(f32)


def line_similar(line2: np.ndarray, line1: np.ndarray) -> bool:
    """TrackLSD::LineSimilar (TrackLSD.cpp:816-830)."""
    mx = (f32(line1[0]) + f32(line1[2])) / f32(2)
    my = (f32(line1[1]) + f32(line1[3])) / f32(2)
    return bool(point_line_distance(line2, mx, my) <= 6)


def line_class(line: np.ndarray, vp) -> bool:
    """TrackLSD::LineClass (TrackLSD.cpp:335-366), including the ``atan(dy) / dx`` quirk at :350-351."""
    # Eigen::Vector3d arithmetic written out in Python doubles (the interpreter cost of tiny NumPy arrays used to
    # dominate the CPU baseline): s, e homogeneous end points, mid = (s + e) / 2, ln = mid x v3
    s0, s1, e0, e1 = float(line[0]), float(line[1]), float(line[2]), float(line[3])
    m0, m1, m2 = (s0 + e0) / 2, (s1 + e1) / 2, (1.0 + 1.0) / 2
    v0, v1 = float(vp[0]), float(vp[1])
    l0, l1, l2 = m1 * 1.0 - m2 * v1, m2 * v0 - m0 * 1.0, m0 * v1 - m1 * v0
    with np.errstate(all="ignore"):
        if l0 != 0 or l1 != 0:
            dis = (abs(l0 * s0 + l1 * s1 + l2 * 1.0) + abs(l0 * e0 + l1 * e1 + l2 * 1.0)) / (2 * math.sqrt(l0 * l0 + l1 * l1))
        else:
            dis = float("nan")
        dis = abs(dis)
        a1 = float(np.arctan(f32(line[1]) - f32(line[3])) / (f32(line[0]) - f32(line[2])))   # float arithmetic
        den = m0 - v0
        num = math.atan(m1 - v1)
        a2 = num / den if den != 0 else (math.copysign(math.inf, num) * math.copysign(1.0, den) if num != 0 else float("nan"))
        err = abs(a1 - a2)
    return bool(dis <= 5.0 and err <= 0.35)


def line_classification(line: np.ndarray, vps) -> int:
    """TrackLSD::LineClassification (TrackLSD.cpp:318-333): z first, then y, then x."""
    if line_class(line, vps[2]):
        return 3
    if line_class(line, vps[1]):
        return 2
    if line_class(line, vps[0]):
        return 1
    return 0


class TrackLSD:
    """viw::TrackLSD (monocular), restated.  Needs the point tracker it shares IDs with (TrackLSD.cpp:30-31)."""

    def __init__(self, cfg: FeConfig, track_feats: TrackKLT, ops=cvops):
        self.cfg = cfg
        self.ops = ops
        self.track_feats = track_feats
        self.currid = 1                                           # TrackLSD.cpp:32
        self.lines_last = np.zeros((0, 4), f32)
        self.lines_det_last = np.zeros((0, 4), f32)   # extension (cfg.line_samples): segments detected in the last frame
        self.ids_last: List[int] = []
        self.pol_last: List[Dict[int, float]] = []
        self.trace: Dict[str, object] = {}

    def get_state(self):
        return dict(currid=self.currid, lines_last=self.lines_last.copy(), ids_last=list(self.ids_last),
                    pol_last=[dict(m) for m in self.pol_last])

    def set_state(self, st):
        self.currid = int(st["currid"])
        self.lines_last = np.asarray(st["lines_last"], f32).reshape(-1, 4).copy()
        self.ids_last = [int(i) for i in st["ids_last"]]
        self.pol_last = [dict(m) for m in st["pol_last"]]

    # -- TrackLSD.cpp:194-236 (GUI at :243-282 omitted)
    def _perform_detection(self, img_eq: np.ndarray):
        cfg = self.cfg
        small = self.ops.half_res(img_eq)                         # :204
        self.trace["small"] = small
        cv_lines = self.ops.fld_detect(small, cfg.fld_length_threshold, cfg.fld_distance_threshold,
                                       cfg.canny_th1, cfg.canny_th2, 3)            # :200-205
        self.trace["fld"] = cv_lines.copy()
        lines = (cv_lines * f32(2)).astype(f32)                   # :218-221
        if len(lines):                                            # FilterShortLines(lines0, 40)  :435-448
            dx = lines[:, 2] - lines[:, 0]
            dy = lines[:, 3] - lines[:, 1]
            lsq = dx * dx + dy * dy
            thr = f32(cfg.line_min_length) * f32(cfg.line_min_length)
            lines = lines[lsq > thr]
        ids = []
        for _ in range(len(lines)):                               # :233-236
            self.currid += 1
            ids.append(self.currid)
        self.lines_det_new = lines.copy()
        return lines, ids

    # -- TrackLSD.cpp:744-792 (bbox index mix-up reproduced)
    @staticmethod
    def assign_points_to_lines(lines, line_ids, points, pids):
        """AssignPointToLines.  The double loop over (line, point) is evaluated as NumPy arrays — the same float32 /
        float64 comparisons and PointLineDistance operations element-wise — so that the CPU baseline is not dominated
        by interpreter overhead; the dictionaries are then filled in the reference's (line, point) order."""
        relation, positions, new_lines, new_ids = [], [], [], []
        lines = np.asarray(lines, f32).reshape(-1, 4)
        points = np.asarray(points, f32).reshape(-1, 2)
        if len(lines) == 0 or len(points) == 0:
            return relation, positions, np.zeros((0, 4), f32), new_ids
        L = lines.astype(np.float64)
        lx1, lx2, ly1, ly2 = L[:, 0], L[:, 1], L[:, 2], L[:, 3]               # :754-757 (sic)
        min_lx, max_lx = np.minimum(lx1, lx2)[:, None], np.maximum(lx1, lx2)[:, None]
        min_ly, max_ly = np.minimum(ly1, ly2)[:, None], np.maximum(ly1, ly2)[:, None]
        px, py = points[:, 0].astype(np.float64)[None, :], points[:, 1].astype(np.float64)[None, :]
        inside = ~((px < min_lx) | (px > max_lx) | (py < min_ly) | (py > max_ly))   # :775
        # PointLineDistance (:794-814), float32 like the C++, evaluated only for the pairs that pass the box test (the
        # reference calls it after the box test too, :775-780)
        li, pj = np.nonzero(inside)
        if len(li) == 0:
            return relation, positions, np.zeros((0, 4), f32), new_ids
        x0, y0 = points[pj, 0], points[pj, 1]
        x1, y1, x2, y2 = lines[li, 0], lines[li, 1], lines[li, 2], lines[li, 3]
        cross = (x2 - x1) * (x0 - x1) + (y2 - y1) * (y0 - y1)
        d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1)
        d1 = np.sqrt((x0 - x1) * (x0 - x1) + (y0 - y1) * (y0 - y1))
        d2 = np.sqrt((x0 - x2) * (x0 - x2) + (y0 - y2) * (y0 - y2))
        num = np.abs((y2 - y1) * x0 + (x1 - x2) * y0 + ((x2 * y1) - (x1 * y2)))
        den = np.sqrt((y2 - y1).astype(np.float64) ** 2 + (x1 - x2).astype(np.float64) ** 2)
        with np.errstate(all="ignore"):
            d3 = np.abs(num.astype(np.float64) / den).astype(f32)
        dist = np.where(cross <= 0, d1, np.where(cross > d, d2, d3)).astype(f32)
        hit = ~(dist > 5)                                                      # :780
        li, pj, dist = li[hit], pj[hit], dist[hit]
        # pairs are in (line, point) order; group them by line
        if len(li) == 0:
            return relation, positions, np.zeros((0, 4), f32), new_ids
        starts = np.nonzero(np.r_[True, li[1:] != li[:-1]])[0].tolist() + [len(li)]
        pj_l, dist_l, li_l = pj.tolist(), dist.tolist(), li.tolist()
        for a0, a1 in zip(starts[:-1], starts[1:]):
            i = li_l[a0]
            pol = {int(pids[j]): dv for j, dv in zip(pj_l[a0:a1], dist_l[a0:a1])}
            relation.append(dict(sorted(pol.items())))
            new_lines.append(lines[i])
            new_ids.append(line_ids[i])
            positions.append(points[pj[a0:a1]].copy())
        return relation, positions, np.asarray(new_lines, f32).reshape(-1, 4), new_ids

    # -- TrackLSD.cpp:368-407
    @staticmethod
    def line_match(lines_new, lines_last, pol_last, pol_new) -> Dict[int, int]:
        matches: Dict[int, int] = {}
        n0, n1 = len(pol_last), len(pol_new)
        if n0 == 0 or n1 == 0:
            return matches
        for i in range(n1):
            if len(pol_new[i]) < 1:
                continue
            for j in range(n0):
                if len(pol_last[j]) < 1:
                    continue
                m = 0
                for pid in sorted(pol_last[j].keys()):
                    if pid not in pol_new[i]:
                        continue
                    m += 1
                    if m >= 2:
                        matches[i] = j
                        break
                    elif m == 1 and line_similar(lines_new[i], lines_last[j]):
                        matches[i] = j
                        break
        return matches

    # -- TrackLSD.cpp:70-192
    def feed_new_camera(self, timestamp: float, img: np.ndarray, mask: np.ndarray, vps, img_eq=None) -> List[LineRow]:
        cfg = self.cfg
        self.trace = {}
        if img_eq is None:
            if cfg.histogram_method == HIST_HISTOGRAM:
                img_eq = self.ops.equalize_hist(img)               # :83 (the reference recomputes it)
            elif cfg.histogram_method == HIST_CLAHE:
                img_eq = self.ops.clahe(img)
            else:
                img_eq = img
        rows: List[LineRow] = []
        first = len(self.lines_last) == 0                         # :95
        lines_new, ids_new = self._perform_detection(img_eq)
        points_new = self.track_feats.get_last_obs()              # :127-129 (point tracker already ran)
        pids_new = self.track_feats.get_last_ids()
        pol_new, positions, filt_lines, filt_ids = self.assign_points_to_lines(lines_new, ids_new, points_new, pids_new)
        self.trace.update(lines_det=lines_new.copy(), ids_det=list(ids_new), filt_lines=filt_lines.copy(),
                          filt_ids=list(filt_ids), pol_new=[dict(m) for m in pol_new])
        if first:                                                 # :95-115
            self.lines_last, self.ids_last, self.pol_last = filt_lines, list(filt_ids), pol_new
            self.trace["matches"] = {}
            return rows
        matches = self.line_match(filt_lines, self.lines_last, self.pol_last, pol_new)   # :138
        self.trace["matches"] = dict(matches)
        good_ids = []
        for i in range(len(filt_lines)):                          # :146-158
            if i in matches:
                good_ids.append(int(np.int32(self.ids_last[matches[i]])))
            else:
                good_ids.append(int(np.int32(filt_ids[i])))
        K, D = self.track_feats.K, self.track_feats.D
        for i in range(len(filt_lines)):                          # :163-167
            l = filt_lines[i]
            e = self.ops.undistort(np.array([[l[0], l[1]], [l[2], l[3]]], f32), K, D)
            line_n = np.array([e[0, 0], e[0, 1], e[1, 0], e[1, 1]], f32)
            Dcls = line_classification(l, vps)
            rows.append(LineRow(good_ids[i], l.copy(), line_n, list(pol_new[i].keys()), list(pol_new[i].values()),
                                positions[i], Dcls))
        self.lines_last, self.ids_last, self.pol_last = filt_lines, good_ids, pol_new    # :175-182
        return rows


class FrontEnd:
    """UpdaterCamera::feed_measurement's tracker calls (UpdaterCamera.cpp:105-110): points, then lines."""

    def __init__(self, cfg: FeConfig, ops=cvops):
        self.cfg = cfg
        self.klt = TrackKLT(cfg, ops)
        self.lsd = TrackLSD(cfg, self.klt, ops) if cfg.use_lines else None

    def feed(self, timestamp, img, mask=None, vps=None):
        if mask is None:
            mask = np.zeros_like(img)
        if self.cfg.downsample:          # UpdaterCamera.cpp:86-95
            img, mask = self.klt.ops.downsample(img), self.klt.ops.downsample(mask)
        prev_eq = self.klt.img_last
        prows = self.klt.feed_new_camera(timestamp, img, mask)
        # extension (BASELINE.json configs[2]; NOT in the reference): LK over cfg.line_samples points sampled evenly along
        # every segment the detector kept in the previous frame, previous -> current equalised image, with the point
        # tracker's LK parameters (cv::calcOpticalFlowPyrLK, initial flow = the sample itself).  Rides in the point
        # tracker's LK call, so it only exists when that call is made (>= 10 points to track).
        self.sample_uv, self.sample_status = np.zeros((0, 4), f32), np.zeros((0,), np.uint8)
        S = int(self.cfg.line_samples)
        if S > 0 and self.lsd is not None and prev_eq is not None and len(self.lsd.lines_det_last) and "mask_klt" in self.klt.trace:
            L = self.lsd.lines_det_last.astype(f32)
            a = (np.arange(S, dtype=f32) / f32(S - 1)) if S > 1 else np.full((1,), 0.5, f32)
            p0 = np.stack([(L[:, None, 0] + (L[:, None, 2] - L[:, None, 0]) * a[None, :]).astype(f32),
                           (L[:, None, 1] + (L[:, None, 3] - L[:, None, 1]) * a[None, :]).astype(f32)], -1).reshape(-1, 2)
            p1, st = self.klt.ops.lk(prev_eq, self.klt.trace["img_eq"], p0, p0.copy(), self.cfg.win_size, self.cfg.pyr_levels)
            self.sample_uv = np.concatenate([p0, p1], 1).astype(f32)
            self.sample_status = st
        lrows = []
        if self.lsd is not None:
            if vps is None:
                vps = [(1e5, 263.0), (608.0, -1e5), (608.0, 263.0)]
            lrows = self.lsd.feed_new_camera(timestamp, img, mask, vps, img_eq=self.klt.trace["img_eq"])
            self.lsd.lines_det_last = self.lsd.lines_det_new
        return prows, lrows
