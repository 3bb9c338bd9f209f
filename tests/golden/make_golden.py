"""Generates tests/golden/ops_golden.npz: small inputs and the outputs of the REAL OpenCV (cv2, the reference's
third-party dependency) for every operation on the hot path, plus two frames of the oracle front end.  Run in the
authoring container (python tests/golden/make_golden.py); the .npz is committed so that the checks also run where
cv2 differs or is absent.  cv2 version is recorded inside."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import plviwo_b200  # noqa: E402,F401
from plviwo_b200 import synth  # noqa: E402
from oracle import cvops, frontend as ofe  # noqa: E402


def main():
    seq = synth.SynthSequence(seed=77, width=320, height=192, n_frames=4)
    a, b = seq.frame(0), seq.frame(1)
    out = {"cv2_version": np.array(cv2.__version__), "img_a": a, "img_b": b}
    eq_a, eq_b = cvops.equalize_hist(a), cvops.equalize_hist(b)
    out["eq_a"] = eq_a
    out["clahe_a"] = cvops.clahe(a)
    pyr = cvops.build_pyramid(eq_a, 15, 3)
    for l, p in enumerate(pyr):
        out["pyr_a_%d" % l] = p
    out["half_a"] = cvops.half_res(eq_a)
    out["canny_half_a"] = cvops.canny(out["half_a"])
    roi = np.ascontiguousarray(eq_a[16:16 + 112, 32:32 + 200])
    xy, resp = cvops.fast_cell(roi, 20)
    out["fast_roi"], out["fast_xy"], out["fast_resp"] = roi, xy, resp
    out["fast_perm"] = cvops.sort_perm(resp)
    pts = xy[out["fast_perm"][:40]].astype(np.float32) + np.array([32, 16], np.float32)
    out["subpix_in"], out["subpix_out"] = pts, cvops.corner_subpix(eq_a, pts)
    p1, st = cvops.lk(eq_a, eq_b, out["subpix_out"], out["subpix_out"], 15, 3)
    out["lk_p1"], out["lk_status"] = p1, st
    K, D = seq.K, seq.D
    out["K"], out["D"] = np.array(K), np.array(D)
    out["und_p0"] = cvops.undistort(out["subpix_out"], K, D)
    out["und_p1"] = cvops.undistort(p1, K, D)
    out["ransac_mask"] = cvops.find_fundamental_mask(out["und_p0"], out["und_p1"], 2.0 / max(K[0], K[1]))
    rng = np.random.default_rng(0)
    chain = np.stack([np.arange(30), np.rint(0.37 * np.arange(30) + rng.normal(0, 0.4, 30))], 1).astype(np.int32)
    out["fitline_pts"] = chain
    out["fitline_out"] = cv2.fitLine(chain.astype(np.float32), cv2.DIST_L2, 0, 0.01, 0.01).reshape(-1)
    # two frames of the whole front end (320x192, 60 points): the rows the reference would write to its databases
    fe = ofe.FrontEnd(ofe.FeConfig(num_features=60, grid_x=4, grid_y=3, pyr_levels=3, K=K, D=D))
    for t in range(3):
        prow, lrow = fe.feed(seq.timestamp(t), seq.frame(t), None, seq.vanishing_points(t))
        out["fe_rows_%d" % t] = np.array([[r.id, r.u, r.v, r.un, r.vn] for r in prow], np.float64).reshape(-1, 5)
        out["fe_last_ids_%d" % t] = np.array(fe.klt.get_last_ids(), np.int64)
        out["fe_line_ids_%d" % t] = np.array([r.id for r in lrow], np.int64)
        out["fe_lines_%d" % t] = np.array([r.line for r in lrow], np.float32).reshape(-1, 4)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ops_golden.npz"), **out)
    print("wrote ops_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
