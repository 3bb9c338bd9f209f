"""ORACLE (test infrastructure): NumPy restatement of the arithmetic of the OpenCV calls on the hot path.

Same function signatures as ``oracle.cvops`` so that ``oracle.frontend`` can run on either.  Each function states
the published algorithm it follows and is pinned against the real OpenCV (cv2 4.13.0) by
tests/test_oracle_pins.py and by the golden vectors in tests/golden/.  Integer ops are bit-exact; float ops state
their tolerance.  Pure-Python loops: small cases only.
"""
from __future__ import annotations

import numpy as np

from . import cvops

f32 = np.float32


def _reflect101(i, n):
    i = np.where(i < 0, -i, i)
    i = np.where(i >= n, 2 * n - 2 - i, i)
    return np.clip(i, 0, n - 1)


# ---------------------------------------------------------------------------------------------- A1
def equalize_hist(img: np.ndarray) -> np.ndarray:
    """cv::equalizeHist: 256-bin histogram -> LUT with float32 scale and round-half-even."""
    hist = np.bincount(img.reshape(-1), minlength=256).astype(np.int64)
    total = img.size
    i0 = int(np.nonzero(hist)[0][0])
    if hist[i0] == total:
        return np.full_like(img, i0)
    scale = f32(255.0) / f32(total - hist[i0])
    lut = np.zeros(256, np.uint8)
    s = 0
    for i in range(i0 + 1, 256):
        s += int(hist[i])
        lut[i] = np.clip(np.rint(f32(s) * scale), 0, 255).astype(np.uint8)
    return lut[img]


# ---------------------------------------------------------------------------------------------- A2
def clahe(img: np.ndarray, clip: float = 10.0, tiles=(8, 8)) -> np.ndarray:
    """cv::createCLAHE(clip, tiles)->apply on 8UC1 (TrackKLT.cpp:60-64), following OpenCV's clahe.cpp: extend to a
    multiple of the tile grid with BORDER_REFLECT_101 (both directions as soon as one is not divisible), per tile:
    histogram, clip at int(clip * area / 256) (>= 1), spread the excess (remainder: one count every 256 // remainder
    bins), LUT = cvRound(cumsum * (255 / area)) in float32; output = bilinear blend of the 4 nearest tile LUTs in float32,
    (l11 * xa1 + l12 * xa) * ya1 + (l21 * xa1 + l22 * xa) * ya.  Bit-exact against cv2 4.13 (tests/test_oracle_pins.py)."""
    H, W = img.shape
    tx, ty = tiles
    if W % tx == 0 and H % ty == 0:
        ext = img
    else:
        eh, ew = H + (ty - H % ty), W + (tx - W % tx)
        ext = img[np.ix_(_reflect101(np.arange(eh), H), _reflect101(np.arange(ew), W))]
    tw, th = ext.shape[1] // tx, ext.shape[0] // ty
    area = tw * th
    lut_scale = f32(255) / f32(area)
    cl = max(int(clip * area / 256), 1)
    luts = np.zeros((ty, tx, 256), np.uint8)
    for j in range(ty):
        for i in range(tx):
            hist = np.bincount(ext[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int64)
            clipped = int(np.maximum(hist - cl, 0).sum())
            hist = np.minimum(hist, cl)
            batch = clipped // 256
            residual = clipped - batch * 256
            hist += batch
            if residual != 0:
                step = max(256 // residual, 1)
                k = 0
                while k < 256 and residual > 0:
                    hist[k] += 1
                    k += step
                    residual -= 1
            luts[j, i] = np.clip(np.rint(np.cumsum(hist).astype(f32) * lut_scale), 0, 255).astype(np.uint8)
    inv_tw, inv_th = f32(1.0) / f32(tw), f32(1.0) / f32(th)
    xf = np.arange(W).astype(f32) * inv_tw - f32(0.5)
    tx1 = np.floor(xf).astype(np.int32)
    xa = (xf - tx1.astype(f32)).astype(f32)
    xa1 = (f32(1) - xa).astype(f32)
    tx2, tx1 = np.minimum(tx1 + 1, tx - 1), np.maximum(tx1, 0)
    yf = np.arange(H).astype(f32) * inv_th - f32(0.5)
    ty1 = np.floor(yf).astype(np.int32)
    ya = (yf - ty1.astype(f32)).astype(f32)
    ya1 = (f32(1) - ya).astype(f32)
    ty2, ty1 = np.minimum(ty1 + 1, ty - 1), np.maximum(ty1, 0)
    v = img.astype(np.int32)
    l11, l12 = luts[ty1[:, None], tx1[None, :], v].astype(f32), luts[ty1[:, None], tx2[None, :], v].astype(f32)
    l21, l22 = luts[ty2[:, None], tx1[None, :], v].astype(f32), luts[ty2[:, None], tx2[None, :], v].astype(f32)
    top = ((l11 * xa1[None, :]).astype(f32) + (l12 * xa[None, :]).astype(f32)).astype(f32)
    bot = ((l21 * xa1[None, :]).astype(f32) + (l22 * xa[None, :]).astype(f32)).astype(f32)
    res = ((top * ya1[:, None]).astype(f32) + (bot * ya[:, None]).astype(f32)).astype(f32)
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


def pyr_down(img: np.ndarray) -> np.ndarray:
    """cv::pyrDown: separable [1 4 6 4 1], BORDER_REFLECT_101, (v + 128) >> 8, size ((w+1)/2, (h+1)/2)."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int32)
    src = img.astype(np.int32)
    cols = _reflect101(2 * np.arange(ow)[:, None] + np.arange(-2, 3)[None, :], w)     # (ow, 5)
    hor = (src[:, cols] * k[None, None, :]).sum(2)                                        # (h, ow)
    rows = _reflect101(2 * np.arange(oh)[:, None] + np.arange(-2, 3)[None, :], h)     # (oh, 5)
    ver = (hor[rows, :] * k[None, :, None]).sum(1)                                        # (oh, ow)
    return ((ver + 128) >> 8).astype(np.uint8)


def downsample(img: np.ndarray) -> np.ndarray:
    """cv::pyrDown(img, out, Size(cols / 2.0, rows / 2.0)) — UpdaterCamera.cpp:90-93: the same filter, but the output size
    is the TRUNCATED half (an odd side loses its last sample compared with the default (n + 1) / 2)."""
    h, w = img.shape
    return pyr_down(img)[:int(h / 2.0), :int(w / 2.0)].copy()


def build_pyramid(img: np.ndarray, win: int, max_level: int):
    """cv::buildOpticalFlowPyramid image levels: stops when the next level is <= win in either dimension."""
    pyr = [img]
    for _ in range(max_level):
        h, w = pyr[-1].shape
        if (w + 1) // 2 <= win or (h + 1) // 2 <= win:
            break
        pyr.append(pyr_down(pyr[-1]))
    return pyr


def half_res(img: np.ndarray) -> np.ndarray:
    """cv::resize(.., 0.5, 0.5, INTER_LINEAR) on even sizes == INTER_AREA: (a + b + c + d + 2) >> 2."""
    s = img.astype(np.int32)
    return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)


def resize_nearest(mask: np.ndarray, gx: int, gy: int) -> np.ndarray:
    """cv::resize INTER_NEAREST: sx = min(floor(x * (1 / (gx / W))), W - 1)."""
    h, w = mask.shape
    ifx, ify = 1.0 / (float(gx) / float(w)), 1.0 / (float(gy) / float(h))
    xs = np.minimum(np.floor(np.arange(gx) * ifx).astype(int), w - 1)
    ys = np.minimum(np.floor(np.arange(gy) * ify).astype(int), h - 1)
    return mask[np.ix_(ys, xs)]


# ---------------------------------------------------------------------------------------------- A4
_RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0),
         (-3, 1), (-2, 2), (-1, 3)]


def fast_cell(roi: np.ndarray, threshold: int):
    """cv::FAST(roi, thr, nonmax=true), TYPE_9_16: corner test, score = max(a0, -b0) - 1, strict 3x3 NMS, row-major."""
    h, w = roi.shape
    if h < 7 or w < 7:
        return np.zeros((0, 2), np.int32), np.zeros((0,), f32)
    I = roi.astype(np.int32)
    c = I[3:h - 3, 3:w - 3]
    d = np.stack([c - I[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in _RING], 0)   # (16, H, W)
    dd = np.concatenate([d, d[:8]], 0)                                                     # circular
    mn = np.stack([dd[k:k + 9].min(0) for k in range(16)], 0)
    mx = np.stack([dd[k:k + 9].max(0) for k in range(16)], 0)
    a0, b0 = mn.max(0), mx.min(0)
    corner = (a0 > threshold) | (b0 < -threshold)
    score = np.zeros((h, w), np.int32)
    score[3:h - 3, 3:w - 3] = np.where(corner, np.maximum(a0, -b0) - 1, 0)
    p = np.pad(score, 1)
    keep = score > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= score > p[1 + dy:h + 1 + dy, 1 + dx:w + 1 + dx]
    ys, xs = np.nonzero(keep)
    return np.stack([xs, ys], 1).astype(np.int32), score[ys, xs].astype(f32)


sort_perm = cvops.sort_perm   # libstdc++ std::sort, executed for real (oracle_shim.cpp)


# ---------------------------------------------------------------------------------------------- A5
def _subpix_mask():
    m = np.zeros((11, 11), f32)
    for i in range(11):
        y = f32(i - 5) / f32(5)
        vy = np.exp(-y * y, dtype=f32)
        for j in range(11):
            x = f32(j - 5) / f32(5)
            m[i, j] = f32(vy * np.exp(-x * x, dtype=f32))
    return m


def corner_subpix(img: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """cv::cornerSubPix(win (5,5), zero zone (-1,-1), {COUNT+EPS, 20, 0.001}); agrees with cv2 to ~1e-4 px."""
    h, w = img.shape
    mask = _subpix_mask().astype(np.float64)
    src = img.astype(f32)
    out = np.array(pts, f32).reshape(-1, 2).copy()
    py, px = np.mgrid[-5:6, -5:6].astype(np.float64)
    for n in range(len(out)):
        cT = out[n].copy()
        cI = cT.copy()
        it = 0
        while True:
            cx, cy = f32(cI[0]) - f32(6.0), f32(cI[1]) - f32(6.0)
            ipx, ipy = int(np.floor(cx)), int(np.floor(cy))
            a, b = f32(cx - f32(ipx)), f32(cy - f32(ipy))
            a11, a12 = f32((f32(1) - a) * (f32(1) - b)), f32(a * (f32(1) - b))
            a21, a22 = f32((f32(1) - a) * b), f32(a * b)
            x0 = np.clip(ipx + np.arange(13), 0, w - 1)
            x1 = np.clip(ipx + 1 + np.arange(13), 0, w - 1)
            y0 = np.clip(ipy + np.arange(13), 0, h - 1)
            y1 = np.clip(ipy + 1 + np.arange(13), 0, h - 1)
            patch = (((src[np.ix_(y0, x0)] * a11 + src[np.ix_(y0, x1)] * a12) + src[np.ix_(y1, x0)] * a21) +
                     src[np.ix_(y1, x1)] * a22).astype(f32)
            tgx = (patch[1:12, 2:13] - patch[1:12, 0:11]).astype(np.float64)
            tgy = (patch[2:13, 1:12] - patch[0:11, 1:12]).astype(np.float64)
            gxx, gxy, gyy = tgx * tgx * mask, tgx * tgy * mask, tgy * tgy * mask
            A, B, Cc = gxx.sum(), gxy.sum(), gyy.sum()
            bb1 = (gxx * px + gxy * py).sum()
            bb2 = (gxy * px + gyy * py).sum()
            det = A * Cc - B * B
            if abs(det) <= np.finfo(np.float64).eps ** 2:
                break
            scale = 1.0 / det
            c2 = np.array([f32(cI[0] + Cc * scale * bb1 - B * scale * bb2), f32(cI[1] - B * scale * bb1 + A * scale * bb2)], f32)
            err = float((c2[0] - cI[0]) * (c2[0] - cI[0]) + (c2[1] - cI[1]) * (c2[1] - cI[1]))
            cI = c2
            if cI[0] < 0 or cI[0] >= w or cI[1] < 0 or cI[1] >= h:
                break
            it += 1
            if not (it < 20 and err > 1e-6):
                break
        if abs(cI[0] - cT[0]) > 5 or abs(cI[1] - cT[1]) > 5:
            cI = cT
        out[n] = cI
    return out


# ---------------------------------------------------------------------------------------------- A3
def _scharr(I: np.ndarray):
    """calcSharrDeriv: int16 (Ix, Iy), REFLECT_101 at the image border."""
    h, w = I.shape
    P = np.pad(I.astype(np.int32), 1, mode="reflect")
    S = 3 * (P[0:h, :] + P[2:h + 2, :]) + 10 * P[1:h + 1, :]          # (h, w+2)
    Ix = S[:, 2:] - S[:, :-2]
    Dv = P[2:h + 2, :] - P[0:h, :]
    Iy = 3 * (Dv[:, :-2] + Dv[:, 2:]) + 10 * Dv[:, 1:-1]
    return Ix.astype(np.int32), Iy.astype(np.int32)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def lk(img0, img1, pts0, pts1_init, win: int, max_level: int, max_count: int = 30, eps: float = 0.01,
       min_eig: float = 1e-4):
    """cv::calcOpticalFlowPyrLK with OPTFLOW_USE_INITIAL_FLOW, error vector requested (as the reference does), scalar
    restatement: W_BITS = 14 fixed-point bilinear patches, float32 normal equations.  Agrees with cv2 to ~1e-4 px
    (OpenCV's SIMD accumulation order differs); statuses equal."""
    pyr0 = build_pyramid(img0, win, max_level)
    pyr1 = build_pyramid(img1, win, max_level)
    L = min(len(pyr0), len(pyr1)) - 1
    pad = win + 1
    P0 = [np.pad(p.astype(np.int32), pad, mode="reflect") for p in pyr0]
    P1 = [np.pad(p.astype(np.int32), pad, mode="reflect") for p in pyr1]
    DER = []
    for p in pyr0:
        ix, iy = _scharr(p)
        DER.append((np.pad(ix, pad), np.pad(iy, pad)))
    n = len(pts0)
    prev_all = np.asarray(pts0, f32).reshape(-1, 2)
    nxt_all = np.asarray(pts1_init, f32).reshape(-1, 2).copy()
    status = np.ones(n, np.uint8)
    half = f32((win - 1) * 0.5)
    eps_sq = float(eps) * float(eps)
    FLT_SCALE = f32(1.0 / (1 << 20))
    ar = np.arange(win)
    for i in range(n):
        nextp = nxt_all[i].copy()
        for level in range(L, -1, -1):
            rows, cols = pyr0[level].shape
            sc = f32(1.0 / (1 << level))
            prevPt = prev_all[i] * sc
            nextp = nextp * sc if level == L else nextp * f32(2)
            prevPt = (prevPt - half).astype(f32)
            ipx, ipy = int(np.floor(prevPt[0])), int(np.floor(prevPt[1]))
            if ipx < -win or ipx >= cols or ipy < -win or ipy >= rows:
                if level == 0:
                    status[i] = 0
                continue
            a, b = f32(prevPt[0] - f32(ipx)), f32(prevPt[1] - f32(ipy))
            iw00 = int(np.rint(f32(f32((f32(1) - a) * (f32(1) - b)) * f32(16384))))
            iw01 = int(np.rint(f32(f32(a * (f32(1) - b)) * f32(16384))))
            iw10 = int(np.rint(f32(f32((f32(1) - a) * b) * f32(16384))))
            iw11 = 16384 - iw00 - iw01 - iw10
            ys, xs = ipy + pad + ar, ipx + pad + ar
            I = P0[level]
            Dx, Dy = DER[level]

            def bil(M, n_bits):
                return _descale(M[np.ix_(ys, xs)] * iw00 + M[np.ix_(ys, xs + 1)] * iw01 + M[np.ix_(ys + 1, xs)] * iw10 +
                                M[np.ix_(ys + 1, xs + 1)] * iw11, n_bits)
            Iw, Ixw, Iyw = bil(I, 9), bil(Dx, 14), bil(Dy, 14)
            A11 = f32((Ixw * Ixw).astype(f32).sum(dtype=f32)) * FLT_SCALE
            A12 = f32((Ixw * Iyw).astype(f32).sum(dtype=f32)) * FLT_SCALE
            A22 = f32((Iyw * Iyw).astype(f32).sum(dtype=f32)) * FLT_SCALE
            D = f32(A11 * A22 - A12 * A12)
            minEig = f32((A22 + A11 - np.sqrt(f32((A11 - A22) * (A11 - A22) + f32(4) * A12 * A12), dtype=f32)) / f32(2 * win * win))
            if minEig < min_eig or D < np.finfo(f32).eps:
                if level == 0:
                    status[i] = 0
                continue
            D = f32(1) / D
            stored = nextp.copy()
            nextp = (nextp - half).astype(f32)
            prevDelta = np.zeros(2, f32)
            J = P1[level]
            for j in range(max_count):
                inx, iny = int(np.floor(nextp[0])), int(np.floor(nextp[1]))
                if inx < -win or inx >= cols or iny < -win or iny >= rows:
                    if level == 0:
                        status[i] = 0
                    break
                a, b = f32(nextp[0] - f32(inx)), f32(nextp[1] - f32(iny))
                jw00 = int(np.rint(f32(f32((f32(1) - a) * (f32(1) - b)) * f32(16384))))
                jw01 = int(np.rint(f32(f32(a * (f32(1) - b)) * f32(16384))))
                jw10 = int(np.rint(f32(f32((f32(1) - a) * b) * f32(16384))))
                jw11 = 16384 - jw00 - jw01 - jw10
                yj, xj = iny + pad + ar, inx + pad + ar
                Jw = _descale(J[np.ix_(yj, xj)] * jw00 + J[np.ix_(yj, xj + 1)] * jw01 + J[np.ix_(yj + 1, xj)] * jw10 +
                              J[np.ix_(yj + 1, xj + 1)] * jw11, 9)
                diff = Jw - Iw
                b1 = f32((diff * Ixw).astype(f32).sum(dtype=f32)) * FLT_SCALE
                b2 = f32((diff * Iyw).astype(f32).sum(dtype=f32)) * FLT_SCALE
                delta = np.array([f32(f32(A12 * b2 - A22 * b1) * D), f32(f32(A12 * b1 - A11 * b2) * D)], f32)
                nextp = (nextp + delta).astype(f32)
                stored = (nextp + half).astype(f32)
                if float(delta[0]) * float(delta[0]) + float(delta[1]) * float(delta[1]) <= eps_sq:
                    break
                if j > 0 and abs(float(f32(delta[0] + prevDelta[0]))) < 0.01 and abs(float(f32(delta[1] + prevDelta[1]))) < 0.01:
                    stored = (stored - delta * f32(0.5)).astype(f32)
                    break
                prevDelta = delta
            nextp = stored
            if level == 0 and status[i]:
                fx, fy = int(np.floor(f32(nextp[0] - half))), int(np.floor(f32(nextp[1] - half)))
                if fx < -win or fx >= cols or fy < -win or fy >= rows:
                    status[i] = 0
        nxt_all[i] = nextp
    return nxt_all, status


# ---------------------------------------------------------------------------------------------- A6
def undistort_equi(pts, K, D):
    """cv::fisheye::undistortPoints (OpenCV 4.x fisheye.cpp), identity R / P, default criteria (10 iterations, eps 1e-8):
    theta_d = |(x - c) / f| clipped to [-pi/2, pi/2]; Newton on theta (1 + k1 theta^2 + k2 theta^4 + k3 theta^6 + k4 theta^8) =
    theta_d; the point is pw tan(theta) / theta_d, or (-1e6, -1e6) when the iteration did not converge or theta changed sign.
    Bit-exact against cv2 4.13 (tests/test_oracle_pins.py)."""
    p = np.asarray(pts, f32).reshape(-1, 2).astype(np.float64)
    pwx = (p[:, 0] - K[2]) / K[0]
    pwy = (p[:, 1] - K[3]) / K[1]
    theta_d = np.minimum(np.maximum(-np.pi / 2.0, np.sqrt(pwx * pwx + pwy * pwy)), np.pi / 2.0)
    out = np.zeros_like(p)
    for i in range(len(p)):
        td = float(theta_d[i])
        converged, theta, scale = False, td, 0.0
        if abs(td) > 1e-8:
            for _ in range(10):
                t2 = theta * theta
                t4 = t2 * t2
                t6 = t4 * t2
                t8 = t6 * t2
                k0, k1, k2, k3 = D[0] * t2, D[1] * t4, D[2] * t6, D[3] * t8
                fix = (theta * (1 + k0 + k1 + k2 + k3) - td) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3)
                theta = theta - fix
                if abs(fix) < 1e-8:
                    converged = True
                    break
            scale = np.tan(theta) / td
        else:
            converged = True
        flipped = (td < 0 and theta > 0) or (td > 0 and theta < 0)
        out[i] = (pwx[i] * scale, pwy[i] * scale) if (converged and not flipped) else (-1000000.0, -1000000.0)
    return out.astype(f32)


def undistort(pts, K, D):
    """cv::undistortPoints, radtan 4 coefficients: 5 fixed-point iterations in double."""
    p = np.asarray(pts, f32).reshape(-1, 2).astype(np.float64)
    x0 = (p[:, 0] - K[2]) / K[0]
    y0 = (p[:, 1] - K[3]) / K[1]
    x, y = x0.copy(), y0.copy()
    k1, k2, p1, p2 = D
    for _ in range(5):
        r2 = x * x + y * y
        icd = 1.0 / (1.0 + (k2 * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x = (x0 - dx) * icd
        y = (y0 - dy) * icd
    return np.stack([x, y], 1).astype(f32)


# ---------------------------------------------------------------------------------------------- A8
def canny(img: np.ndarray, th1: float = 50.0, th2: float = 50.0, aperture: int = 3) -> np.ndarray:
    """cv::Canny(L1 gradient) for low == high: Sobel 3x3 (BORDER_REPLICATE), direction NMS, no hysteresis walk."""
    assert th1 == th2 and aperture == 3
    low = int(np.floor(th1))
    h, w = img.shape
    P = np.pad(img.astype(np.int32), 1, mode="edge")
    dx = (P[0:h, 2:] + 2 * P[1:h + 1, 2:] + P[2:h + 2, 2:]) - (P[0:h, :-2] + 2 * P[1:h + 1, :-2] + P[2:h + 2, :-2])
    dy = (P[2:h + 2, 0:w] + 2 * P[2:h + 2, 1:w + 1] + P[2:h + 2, 2:]) - (P[0:h, 0:w] + 2 * P[0:h, 1:w + 1] + P[0:h, 2:])
    mag = np.abs(dx) + np.abs(dy)
    M = np.pad(mag, 1)
    m = mag
    x, y = np.abs(dx), np.abs(dy) << 15
    tg22x = x * 13573
    tg67x = tg22x + (x << 16)
    c = M[1:h + 1, 1:w + 1]
    horiz = (c > M[1:h + 1, 0:w]) & (c >= M[1:h + 1, 2:])
    vert = (c > M[0:h, 1:w + 1]) & (c >= M[2:h + 2, 1:w + 1])
    s_neg = (dx ^ dy) < 0
    diag_pos = (c > M[0:h, 0:w]) & (c > M[2:h + 2, 2:])          # s = +1: (y-1, x-1), (y+1, x+1)
    diag_neg = (c > M[0:h, 2:]) & (c > M[2:h + 2, 0:w])          # s = -1: (y-1, x+1), (y+1, x-1)
    is_max = np.where(y < tg22x, horiz, np.where(y > tg67x, vert, np.where(s_neg, diag_neg, diag_pos)))
    return np.where((m > low) & is_max, 255, 0).astype(np.uint8)


def fld_detect(img_small, length_threshold=20, distance_threshold=1.414213562, th1=50.0, th2=50.0, aperture=3):
    """FastLineDetector restatement (oracle_shim.cpp) on this module's own Canny."""
    import ctypes  # noqa: F401
    img_small = np.ascontiguousarray(img_small)
    edges = np.ascontiguousarray(canny(img_small, th1, th2, aperture))
    h, w = img_small.shape
    cap = 8192
    out = np.empty((cap, 4), f32)
    n = cvops.shim().oracle_fld_detect(img_small.ctypes.data, w, h, img_small.strides[0], edges.ctypes.data,
                                       int(length_threshold), float(distance_threshold), out.ctypes.data, cap)
    return out[:min(n, cap)].copy()


# ---------------------------------------------------------------------------------------------- A9
class _CvRng:
    def __init__(self, state):
        self.state = state & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def _collinear(m):
    i = len(m) - 1
    for j in range(i):
        dx1, dy1 = float(m[j, 0]) - float(m[i, 0]), float(m[j, 1]) - float(m[i, 1])
        for k in range(j):
            dx2, dy2 = float(m[k, 0]) - float(m[i, 0]), float(m[k, 1]) - float(m[i, 1])
            if abs(dx2 * dy1 - dy2 * dx1) <= np.finfo(f32).eps * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _seven_point(m1, m2):
    import cv2
    A = np.zeros((7, 9))
    for i in range(7):
        x0, y0, x1, y1 = float(m1[i, 0]), float(m1[i, 1]), float(m2[i, 0]), float(m2[i, 1])
        A[i] = [x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1]
    _w, _u, vt = cv2.SVDecomp(A, flags=cv2.SVD_FULL_UV)      # the library's own null-space basis
    f1, f2 = vt[7].copy(), vt[8].copy()
    f1 -= f2
    t0, t1, t2 = f2[4] * f2[8] - f2[5] * f2[7], f2[3] * f2[8] - f2[5] * f2[6], f2[3] * f2[7] - f2[4] * f2[6]
    c = np.zeros(4)
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2
    c[2] = (f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
            f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
            f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]))
    t0, t1, t2 = f1[4] * f1[8] - f1[5] * f1[7], f1[3] * f1[8] - f1[5] * f1[6], f1[3] * f1[7] - f1[4] * f1[6]
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2
    c[1] = (f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
            f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
            f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]))
    nroots, roots = cv2.solveCubic(c.reshape(1, 4))
    roots = roots.reshape(-1)
    Fs = []
    for k in range(nroots):
        lam, mu = roots[k], 1.0
        s = f1[8] * roots[k] + f2[8]
        F = np.zeros(9)
        if abs(s) > np.finfo(np.float64).eps:
            mu = 1.0 / s
            lam *= mu
            F[8] = 1.0
        F[:8] = f1[:8] * lam + f2[:8] * mu
        Fs.append(F)
    return Fs


def find_fundamental_mask(p0n, p1n, thr: float, conf: float = 0.999):
    """cv::findFundamentalMat(FM_RANSAC) for n >= 15 (RANSACPointSetRegistrator + 7-point); n < 15 defers to cv2 (LMedS)."""
    m1 = np.ascontiguousarray(p0n, f32).reshape(-1, 2)
    m2 = np.ascontiguousarray(p1n, f32).reshape(-1, 2)
    n = len(m1)
    if n < 15:
        return cvops.find_fundamental_mask(m1, m2, thr)
    rng = _CvRng(0xFFFFFFFFFFFFFFFF)
    niters, max_good, best = 1000, 0, None
    t = f32(thr * thr)
    x1, y1, x2, y2 = (m1[:, 0].astype(np.float64), m1[:, 1].astype(np.float64), m2[:, 0].astype(np.float64),
                      m2[:, 1].astype(np.float64))
    it = 0
    while it < niters:
        idx, attempts, ok = [], 0, False
        while attempts < 10000:
            idx = []
            while len(idx) < 7:
                k = rng.uniform(0, n)
                while k in idx:
                    k = rng.uniform(0, n)
                idx.append(k)
            if _collinear(m1[idx]) or _collinear(m2[idx]):
                attempts += 1
                continue
            ok = True
            break
        if not ok:
            break
        for F in _seven_point(m1[idx], m2[idx]):
            a = F[0] * x1 + F[1] * y1 + F[2]
            b = F[3] * x1 + F[4] * y1 + F[5]
            c = F[6] * x1 + F[7] * y1 + F[8]
            s2 = 1.0 / (a * a + b * b)
            d2 = x2 * a + y2 * b + c
            a = F[0] * x2 + F[3] * y2 + F[6]
            b = F[1] * x2 + F[4] * y2 + F[7]
            c = F[2] * x2 + F[5] * y2 + F[8]
            s1 = 1.0 / (a * a + b * b)
            d1 = x1 * a + y1 * b + c
            err = np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(f32)
            mask = err <= t
            good = int(mask.sum())
            if good > max(max_good, 6):
                best, max_good = mask.astype(np.uint8), good
                ep = (n - good) / n
                num = max(1.0 - conf, np.finfo(np.float64).tiny)
                den = 1.0 - (1.0 - ep) ** 7
                if den < np.finfo(np.float64).tiny:
                    niters = 0
                else:
                    num, den = np.log(num), np.log(den)
                    niters = niters if (den >= 0 or -num >= niters * (-den)) else int(np.rint(num / den))
        it += 1
    return best if best is not None else np.zeros((n,), np.uint8)
