"""ORACLE (test infrastructure): line-by-line CPU restatement of the reference's STEREO point front end.

Follows (all paths under /root/reference/open_vins/ov_core/src/track):
  TrackKLT.cpp:34-94    feed_new_camera (two images: equalise + pyramid per camera, then feed_stereo)
  TrackKLT.cpp:202-393  feed_stereo
  TrackKLT.cpp:530-827  perform_detection_stereo
  TrackKLT.cpp:829-886  perform_matching (per-camera calibration for the undistortion / RANSAC threshold)
  Grider_GRID.h:74-180  perform_griding (shared with the monocular restatement, oracle/frontend.py)

Quirks kept on purpose (they change results): the right image's working mask starts as a clone of the LEFT mask
(`mask1_updated = mask0.clone()`, :691); the left bounds test after the temporal LK uses `>` where the right one uses
`>=` (:302-303 vs :322-323); no mask test after tracking (the monocular path has one); new left points are tracked
left->right WITHOUT the RANSAC gate and count as stereo when status == 1 and both are in bounds (:646-676).

The line tracker has no stereo path (`TrackLSD.cpp:57-60` falls back to feed_monocular on camera 0), so nothing of
PL-VIWO/src is restated here.  Arithmetic comes from an ``ops`` module exactly as in oracle/frontend.py.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import cvops
from .frontend import FeConfig, PointRow, TrackKLT, TrackLSD, HIST_HISTOGRAM, HIST_CLAHE, f32


class TrackKLTStereo:
    """ov_core::TrackKLT with use_stereo = true fed two images per message (cam 0 = left, cam 1 = right)."""

    def __init__(self, cfg: FeConfig, K_right=None, D_right=None, ops=cvops):
        self.cfg = cfg
        self.ops = ops
        self.currid = 4 * cfg.numaruco + 1           # TrackBase.cpp:34 (one atomic currid for both cameras)
        self.pts_last = [np.zeros((0, 2), f32), np.zeros((0, 2), f32)]
        self.ids_last: List[List[int]] = [[], []]
        self.img_last: List[Optional[np.ndarray]] = [None, None]
        self.mask_last: List[Optional[np.ndarray]] = [None, None]
        self.K = [tuple(cfg.K), tuple(K_right if K_right is not None else cfg.K)]
        self.D = [tuple(cfg.D), tuple(D_right if D_right is not None else cfg.D)]
        self.trace = {}
        # Grider_GRID is shared with the monocular restatement: borrow its method through a helper instance
        self._mono = TrackKLT(cfg, ops)

    # -- state (teacher forcing / checkpoint)
    def get_state(self):
        return dict(currid=self.currid, pts_last=[p.copy() for p in self.pts_last], ids_last=[list(i) for i in self.ids_last],
                    img_last=[None if i is None else i.copy() for i in self.img_last],
                    mask_last=[None if m is None else m.copy() for m in self.mask_last])

    def set_calib(self, cam, K, D):
        self.K[cam], self.D[cam] = tuple(K), tuple(D)

    def _equalise(self, img):
        if self.cfg.histogram_method == HIST_HISTOGRAM:
            return self.ops.equalize_hist(img)
        if self.cfg.histogram_method == HIST_CLAHE:
            return self.ops.clahe(img)
        return img

    # -- TrackKLT.cpp:34-94 + 202-393.  Returns (rows_left, rows_right) in database order (left rows first, :360-370)
    def feed_new_camera(self, timestamp, img_left, img_right, mask_left, mask_right):
        self.trace = {}
        eq = [self._equalise(img_left), self._equalise(img_right)]
        masks = [mask_left, mask_right]
        self.trace["img_eq"] = eq
        rows = ([], [])
        if len(self.pts_last[0]) == 0 and len(self.pts_last[1]) == 0:          # :221-241
            gl, gr, il, ir = self._perform_detection_stereo(eq[0], eq[1], masks[0], masks[1], np.zeros((0, 2), f32),
                                                            np.zeros((0, 2), f32), [], [])
            self.img_last, self.mask_last = eq, masks
            self.pts_last, self.ids_last = [gl, gr], [il, ir]
            self.trace["reset"] = False
            self.trace["first"] = True
            return rows
        self.trace["first"] = False
        # top-off on the PREVIOUS pair (:245-252)
        pl_old, pr_old, il_old, ir_old = self._perform_detection_stereo(
            self.img_last[0], self.img_last[1], self.mask_last[0], self.mask_last[1], self.pts_last[0].copy(),
            self.pts_last[1].copy(), list(self.ids_last[0]), list(self.ids_last[1]))
        self.trace["pts_old"] = [pl_old.copy(), pr_old.copy()]
        self.trace["ids_old"] = [list(il_old), list(ir_old)]
        # temporal tracking of both cameras (:261-270)
        pl_new, mask_ll = self._perform_matching(self.img_last[0], eq[0], pl_old, pl_old.copy(), 0, 0, "l")
        pr_new, mask_rr = self._perform_matching(self.img_last[1], eq[1], pr_old, pr_old.copy(), 1, 1, "r")
        if mask_ll is None and mask_rr is None:                                # :286-300
            self.img_last, self.mask_last = eq, masks
            self.pts_last, self.ids_last = [np.zeros((0, 2), f32), np.zeros((0, 2), f32)], [[], []]
            self.trace["reset"] = True
            return rows
        H, W = eq[0].shape
        Hr, Wr = eq[1].shape
        good_l, good_r, gid_l, gid_r = [], [], [], []
        idx_r = {}
        for n, i in enumerate(ir_old):                                       # first occurrence wins (:309-315)
            idx_r.setdefault(i, n)
        for i in range(len(pl_new)):                                           # :298-334
            x, y = pl_new[i]
            if x < 0 or y < 0 or int(x) > W or int(y) > H:
                continue
            n = idx_r.get(il_old[i], -1)
            if mask_ll[i] and n >= 0 and mask_rr is not None and mask_rr[n]:
                xr, yr = pr_new[n]
                if xr < 0 or yr < 0 or int(xr) >= Wr or int(yr) >= Hr:
                    continue
                good_l.append((x, y)); good_r.append((xr, yr))
                gid_l.append(il_old[i]); gid_r.append(ir_old[n])
            elif mask_ll[i]:
                good_l.append((x, y)); gid_l.append(il_old[i])
        added = set(gid_r)
        for i in range(len(pr_new)):                                           # :337-349
            x, y = pr_new[i]
            if x < 0 or y < 0 or int(x) >= Wr or int(y) >= Hr:
                continue
            if mask_rr[i] and ir_old[i] not in added:
                good_r.append((x, y)); gid_r.append(ir_old[i])
                added.add(ir_old[i])
        good_l = np.asarray(good_l, f32).reshape(-1, 2)
        good_r = np.asarray(good_r, f32).reshape(-1, 2)
        for cam, (g, ids) in enumerate(((good_l, gid_l), (good_r, gid_r))):    # :352-363
            und = self.ops.undistort(g, self.K[cam], self.D[cam]) if len(g) else np.zeros((0, 2), f32)
            for i in range(len(g)):
                rows[cam].append(PointRow(int(ids[i]), float(g[i, 0]), float(g[i, 1]), float(und[i, 0]), float(und[i, 1])))
        self.img_last, self.mask_last = eq, masks                              # :366-378
        self.pts_last, self.ids_last = [good_l, good_r], [gid_l, gid_r]
        self.trace["reset"] = False
        return rows

    # -- first loops of perform_detection_stereo (:545-598 left, :692-748 right): returns the surviving indices, the two
    # occupancy grids and the working mask
    def _filter_existing(self, pts, ids, mask_test, mask_clone, W, H, stereo_ids=None):
        cfg = self.cfg
        d = cfg.min_px_dist
        close_w, close_h = int(f32(W) / f32(d)), int(f32(H) / f32(d))
        grid_close = np.zeros((close_h, close_w), np.uint8)
        size_x, size_y = f32(W) / f32(cfg.grid_x), f32(H) / f32(cfg.grid_y)
        grid_grid = np.zeros((cfg.grid_y, cfg.grid_x), np.uint8)
        mask_upd = mask_clone.copy()
        keep = []
        P = np.asarray(pts, f32).reshape(-1, 2)
        for k in range(len(P)):
            px, py = P[k, 0], P[k, 1]
            x, y = int(px), int(py)
            if x < 10 or x >= W - 10 or y < 10 or y >= H - 10:
                continue
            xc, yc = int(px / f32(d)), int(py / f32(d))
            if xc < 0 or xc >= close_w or yc < 0 or yc >= close_h:
                continue
            xg, yg = int(np.floor(px / size_x)), int(np.floor(py / size_y))
            if xg < 0 or xg >= cfg.grid_x or yg < 0 or yg >= cfg.grid_y:
                continue
            if grid_close[yc, xc] > 127 and not (stereo_ids is not None and ids[k] in stereo_ids):   # :720-726
                continue
            if mask_test[y, x] > 127:
                continue
            grid_close[yc, xc] = 255
            if grid_grid[yg, xg] < 255:
                grid_grid[yg, xg] += 1
            if x - d >= 0 and x + d < W and y - d >= 0 and y + d < H:
                mask_upd[y - d:y + d + 1, x - d:x + d + 1] = 255
            keep.append(k)
        return keep, grid_close, grid_grid, mask_upd

    def _valid_locs(self, grid_grid, mask):
        cfg = self.cfg
        mask_grid = self.ops.resize_nearest(mask, cfg.grid_x, cfg.grid_y)
        nfg = int(float(cfg.num_features) / float(cfg.grid_x * cfg.grid_y)) + 1
        nfg_req = max(1, int(0.5 * nfg))
        return [(x, y) for x in range(cfg.grid_x) for y in range(cfg.grid_y)
                if int(grid_grid[y, x]) < nfg_req and int(mask_grid[y, x]) != 255]

    # -- TrackKLT.cpp:530-827
    def _perform_detection_stereo(self, img0, img1, mask0, mask1, pts0, pts1, ids0, ids1):
        cfg = self.cfg
        d = cfg.min_px_dist
        H0, W0 = img0.shape
        H1, W1 = img1.shape
        det = self.trace["det"] = {"ran_left": False, "ran_right": False}
        keep, grid_close0, grid_grid0, mask0_upd = self._filter_existing(pts0, ids0, mask0, mask0, W0, H0)
        pts0 = np.ascontiguousarray(np.asarray(pts0, f32).reshape(-1, 2)[keep])
        ids0 = [ids0[k] for k in keep]
        pts1 = np.asarray(pts1, f32).reshape(-1, 2)
        ids1 = list(ids1)
        thr = min(20, int(0.5 * cfg.num_features))
        if cfg.num_features - len(pts0) > thr:                                 # :606 (strictly greater, unlike the mono path)
            det["ran_left"] = True
            valid = self._valid_locs(grid_grid0, mask0)
            det["valid_left"] = list(valid)
            sub = {}
            ext = self._mono._perform_griding(img0, mask0_upd, valid, sub)
            det["left"] = sub
            close_w, close_h = grid_close0.shape[1], grid_close0.shape[0]
            new0 = []
            for k in range(len(ext)):                                          # :631-645
                px, py = f32(ext[k, 0]), f32(ext[k, 1])
                xg, yg = int(px / f32(d)), int(py / f32(d))
                if xg < 0 or xg >= close_w or yg < 0 or yg >= close_h:
                    continue
                if grid_close0[yg, xg] > 127:
                    continue
                grid_close0[yg, xg] = 255
                new0.append((px, py))
            new0 = np.asarray(new0, f32).reshape(-1, 2)
            det["new_left"] = new0.copy()
            if len(new0):                                                      # :658-699
                new1, status = self.ops.lk(img0, img1, new0, new0.copy(), cfg.win_size, cfg.pyr_levels)
                det["lr_pts1"], det["lr_status"] = new1.copy(), np.asarray(status).copy()
                add0, add1 = [], []
                for i in range(len(new0)):
                    x0, y0 = new0[i]
                    x1, y1 = new1[i]
                    oob_l = int(x0) < 0 or int(x0) >= W0 or int(y0) < 0 or int(y0) >= H0
                    oob_r = int(x1) < 0 or int(x1) >= W1 or int(y1) < 0 or int(y1) >= H1
                    if not oob_l and not oob_r and status[i] == 1:
                        self.currid += 1
                        add0.append(((x0, y0), self.currid))
                        add1.append(((x1, y1), self.currid))
                    elif not oob_l:
                        self.currid += 1
                        add0.append(((x0, y0), self.currid))
                if add0:
                    pts0 = np.concatenate([pts0, np.asarray([a[0] for a in add0], f32).reshape(-1, 2)], 0)
                    ids0 = ids0 + [a[1] for a in add0]
                if add1:
                    pts1 = np.concatenate([pts1, np.asarray([a[0] for a in add1], f32).reshape(-1, 2)], 0)
                    ids1 = ids1 + [a[1] for a in add1]
        # RIGHT (:684-826).  mask1_updated starts as a clone of the LEFT mask (:691)
        keep, grid_close1, grid_grid1, mask1_upd = self._filter_existing(pts1, ids1, mask1, mask0, W1, H1, stereo_ids=set(ids0))
        pts1 = np.ascontiguousarray(pts1[keep]).reshape(-1, 2)
        ids1 = [ids1[k] for k in keep]
        if cfg.num_features - len(pts1) > thr:                                 # :753
            det["ran_right"] = True
            valid = self._valid_locs(grid_grid1, mask1)
            det["valid_right"] = list(valid)
            sub = {}
            ext = self._mono._perform_griding(img1, mask1_upd, valid, sub)
            det["right"] = sub
            close_w, close_h = grid_close1.shape[1], grid_close1.shape[0]
            new1 = []
            for k in range(len(ext)):                                          # :779-793
                px, py = f32(ext[k, 0]), f32(ext[k, 1])
                xg, yg = int(px / f32(d)), int(py / f32(d))
                if xg < 0 or xg >= close_w or yg < 0 or yg >= close_h:
                    continue
                if grid_close1[yg, xg] > 127:
                    continue
                self.currid += 1
                new1.append(((px, py), self.currid))
                grid_close1[yg, xg] = 255
            if new1:
                pts1 = np.concatenate([pts1, np.asarray([a[0] for a in new1], f32).reshape(-1, 2)], 0)
                ids1 = ids1 + [a[1] for a in new1]
        return pts0, pts1, ids0, ids1

    # -- TrackKLT.cpp:829-886 with per-camera calibration
    def _perform_matching(self, img0, img1, pts0, pts1, id0, id1, tag):
        cfg = self.cfg
        n = len(pts0)
        if n == 0:
            return pts1, None
        if n < 10:
            return pts1, np.zeros((n,), np.uint8)
        p1, mask_klt = self.ops.lk(img0, img1, pts0, pts1, cfg.win_size, cfg.pyr_levels)
        p0n = self.ops.undistort(pts0, self.K[id0], self.D[id0])
        p1n = self.ops.undistort(p1, self.K[id1], self.D[id1])
        maxf = max(max(self.K[id0][0], self.K[id0][1]), max(self.K[id1][0], self.K[id1][1]))
        mask_rsc = self.ops.find_fundamental_mask(p0n, p1n, 2.0 / maxf)
        rsc = np.zeros((n,), bool)
        rsc[:len(mask_rsc)] = np.asarray(mask_rsc, bool)[:n]
        out = (np.asarray(mask_klt, bool) & rsc).astype(np.uint8)
        self.trace["lk_pts1_" + tag] = p1.copy()
        self.trace["mask_klt_" + tag] = np.asarray(mask_klt).copy()
        self.trace["mask_rsc_" + tag] = np.asarray(mask_rsc).copy()
        return p1, out


class _LeftView:
    """What viw::TrackLSD reads from `trackFEATS.at(cam_id)` for camera 0 (TrackLSD.cpp:127-129): the stereo tracker's
    LEFT observations and the left camera's calibration."""

    def __init__(self, klt: TrackKLTStereo):
        self._klt = klt

    K = property(lambda self: self._klt.K[0])
    D = property(lambda self: self._klt.D[0])

    def get_last_obs(self):
        return self._klt.pts_last[0].copy()

    def get_last_ids(self):
        return list(self._klt.ids_last[0])


class StereoFrontEnd:
    """UpdaterCamera::feed_measurement's tracker calls for a stereo rig (UpdaterCamera.cpp:105-110): the point tracker
    gets both images, then the line tracker — which has no stereo path and runs its monocular code on the LEFT image
    (TrackLSD.cpp:57-60) against the left points the point tracker has just produced."""

    def __init__(self, cfg: FeConfig, K_right=None, D_right=None, ops=cvops):
        self.cfg = cfg
        self.klt = TrackKLTStereo(cfg, K_right, D_right, ops)
        self.lsd = TrackLSD(cfg, _LeftView(self.klt), ops) if cfg.use_lines else None

    def feed(self, timestamp, img_left, img_right, mask_left=None, mask_right=None, vps=None):
        if mask_left is None:
            mask_left = np.zeros_like(img_left)
        if mask_right is None:
            mask_right = np.zeros_like(img_right)
        rows_l, rows_r = self.klt.feed_new_camera(timestamp, img_left, img_right, mask_left, mask_right)
        lrows = []
        if self.lsd is not None and vps is not None:
            lrows = self.lsd.feed_new_camera(timestamp, img_left, mask_left, vps, img_eq=self.klt.trace["img_eq"][0])
        return rows_l, rows_r, lrows
