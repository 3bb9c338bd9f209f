"""ORACLE (test infrastructure): the reference's OpenCV call sites, executed by the real OpenCV (cv2 4.13.0).

Each function names the reference call it stands for.  OpenCV is the reference's third-party dependency
(README.md:19 "OpenCV 4.2"; not vendored under /root/reference); here it is the same library through its
Python binding, so these ARE the reference's CPU kernels.  ``oracle.npops`` restates the same arithmetic in
NumPy and tests/test_oracle_pins.py checks one against the other.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import cv2

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = None


def shim():
    """ctypes handle of oracle/liboracle_shim.so (std::sort + FastLineDetector restatement); builds it if absent."""
    global _SHIM
    if _SHIM is None:
        path = os.path.join(_HERE, "liboracle_shim.so")
        src = os.path.join(_HERE, "csrc", "oracle_shim.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        lib = ctypes.CDLL(path)
        lib.oracle_sort_perm.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.oracle_sort_perm.restype = None
        lib.oracle_fld_detect.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int]
        lib.oracle_fld_detect.restype = ctypes.c_int
        lib.oracle_fit_line.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.oracle_fit_line.restype = None
        _SHIM = lib
    return _SHIM


# ----------------------------------------------------------------------------- point path
def equalize_hist(img: np.ndarray) -> np.ndarray:
    """cv::equalizeHist — TrackKLT.cpp:59, TrackLSD.cpp:83."""
    return cv2.equalizeHist(img)


def clahe(img: np.ndarray) -> np.ndarray:
    """cv::createCLAHE(10.0, 8x8)->apply — TrackKLT.cpp:60-64."""
    return cv2.createCLAHE(10.0, (8, 8)).apply(img)


def downsample(img: np.ndarray) -> np.ndarray:
    """cv::pyrDown(img, out, Size(cols / 2.0, rows / 2.0)) — UpdaterCamera.cpp:90-93 (image and mask)."""
    return cv2.pyrDown(img, dstsize=(int(img.shape[1] / 2.0), int(img.shape[0] / 2.0)))


def build_pyramid(img: np.ndarray, win: int, max_level: int):
    """cv::buildOpticalFlowPyramid(img, pyr, win, maxLevel) — TrackKLT.cpp:71.  Returns the image levels only
    (derivative planes are an implementation artefact of OpenCV's LK)."""
    _, pyr = cv2.buildOpticalFlowPyramid(img, (win, win), max_level, withDerivatives=False)
    return [np.ascontiguousarray(p) for p in pyr]


def fast_cell(roi: np.ndarray, threshold: int):
    """cv::FAST(img(roi), pts, threshold, true) — Grider_GRID.h:125.  Returns (xy int32 (n,2), response float32 (n,))
    in OpenCV's output order (row-major)."""
    det = cv2.FastFeatureDetector_create(int(threshold), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(np.ascontiguousarray(roi), None)
    n = len(kps)
    xy = np.empty((n, 2), np.int32)
    resp = np.empty((n,), np.float32)
    for i, k in enumerate(kps):
        xy[i, 0] = int(k.pt[0])
        xy[i, 1] = int(k.pt[1])
        resp[i] = k.response
    return xy, resp


def sort_perm(resp: np.ndarray) -> np.ndarray:
    """std::sort(pts.begin(), pts.end(), compare_response) — Grider_GRID.h:128 (libstdc++ introsort, unstable)."""
    resp = np.ascontiguousarray(resp, np.float32)
    perm = np.empty((len(resp),), np.int32)
    if len(resp):
        shim().oracle_sort_perm(resp.ctypes.data, len(resp), perm.ctypes.data)
    return perm


def corner_subpix(img: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """cv::cornerSubPix(img, pts, (5,5), (-1,-1), {COUNT+EPS, 20, 0.001}) — Grider_GRID.h:163-174."""
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2).copy()
    if len(p) == 0:
        return p.reshape(-1, 2)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.001)
    cv2.cornerSubPix(img, p, (5, 5), (-1, -1), crit)
    return p.reshape(-1, 2)


def lk(img0: np.ndarray, img1: np.ndarray, pts0: np.ndarray, pts1_init: np.ndarray, win: int, max_level: int):
    """cv::calcOpticalFlowPyrLK(pyr0, pyr1, pts0, pts1, status, err, win, maxLevel, {COUNT|EPS,30,0.01},
    OPTFLOW_USE_INITIAL_FLOW) — TrackKLT.cpp:855-858.  The Python binding only takes images; OpenCV then builds the
    same pyramid internally."""
    p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 1, 2)
    p1 = np.ascontiguousarray(pts1_init, np.float32).reshape(-1, 1, 2).copy()
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 30, 0.01)
    p1, st, _err = cv2.calcOpticalFlowPyrLK(img0, img1, p0, p1, winSize=(win, win), maxLevel=max_level, criteria=crit,
                                            flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    return p1.reshape(-1, 2), st.reshape(-1).astype(np.uint8)


def undistort(pts: np.ndarray, K, D) -> np.ndarray:
    """CamRadtan::undistort_f -> cv::undistortPoints(1 point, K, D4) — cam/CamRadtan.h:99-120 (one call per point)."""
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float64)
    Dm = np.array(D, np.float64)
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2)
    if len(p) == 0:
        return p.reshape(-1, 2)
    return cv2.undistortPoints(p, Km, Dm).reshape(-1, 2).astype(np.float32)


def undistort_equi(pts: np.ndarray, K, D) -> np.ndarray:
    """CamEqui::undistort_f -> cv::fisheye::undistortPoints(1 point, K, D4) — cam/CamEqui.h:108-129 (one call per point).
    Not on the product path yet (the library refuses FE_CAM_EQUI): the oracle side of the next camera model."""
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float64)
    Dm = np.array(D, np.float64)
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2)
    if len(p) == 0:
        return p.reshape(-1, 2)
    return cv2.fisheye.undistortPoints(p, Km, Dm).reshape(-1, 2).astype(np.float32)


def find_fundamental_mask(p0n: np.ndarray, p1n: np.ndarray, thr: float):
    """cv::findFundamentalMat(p0, p1, FM_RANSAC, thr, 0.999, mask) — TrackKLT.cpp:873.  Returns the uint8 mask, or an
    empty array when OpenCV returns no model."""
    if len(p0n) < 7:
        return np.zeros((0,), np.uint8)
    _F, mask = cv2.findFundamentalMat(np.ascontiguousarray(p0n, np.float32), np.ascontiguousarray(p1n, np.float32),
                                      cv2.FM_RANSAC, thr, 0.999)
    if mask is None:
        return np.zeros((0,), np.uint8)
    return mask.reshape(-1).astype(np.uint8)


def resize_nearest(mask: np.ndarray, gx: int, gy: int) -> np.ndarray:
    """cv::resize(mask0, mask0_grid, Size(grid_x, grid_y), 0, 0, INTER_NEAREST) — TrackKLT.cpp:480."""
    return cv2.resize(mask, (gx, gy), interpolation=cv2.INTER_NEAREST)


# ----------------------------------------------------------------------------- line path
def half_res(img: np.ndarray) -> np.ndarray:
    """cv::resize(img0, smaller, Size(), 0.5, 0.5, INTER_LINEAR) — TrackLSD.cpp:204."""
    return cv2.resize(img, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR)


def canny(img: np.ndarray, th1: float = 50.0, th2: float = 50.0, aperture: int = 3) -> np.ndarray:
    """cv::Canny inside FastLineDetector::lineDetection (parameters TrackLSD.h:269-273)."""
    return cv2.Canny(img, th1, th2, apertureSize=aperture)


def fld_detect(img_small: np.ndarray, length_threshold: int = 20, distance_threshold: float = 1.414213562,
               th1: float = 50.0, th2: float = 50.0, aperture: int = 3) -> np.ndarray:
    """cv::ximgproc::createFastLineDetector(20, 1.414213562f, 50, 50, 3, false)->detect — TrackLSD.cpp:200-205.
    Canny by cv2; chain walk + segment fit by the restated shim (PARITY UNPINNED, see oracle/__init__.py)."""
    img_small = np.ascontiguousarray(img_small)
    edges = np.ascontiguousarray(canny(img_small, th1, th2, aperture))
    h, w = img_small.shape
    cap = 8192
    out = np.empty((cap, 4), np.float32)
    n = shim().oracle_fld_detect(img_small.ctypes.data, w, h, img_small.strides[0], edges.ctypes.data,
                                 int(length_threshold), float(distance_threshold), out.ctypes.data, cap)
    return out[:min(n, cap)].copy()
