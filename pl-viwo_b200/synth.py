"""Seeded synthetic KAIST-urban-shaped camera sequences (test / bench tooling, not the product path).

SURVEY.md section 8(d): W x H = 1280 x 560 8-bit mono, a textured world canvas with straight strokes and
"facade" rectangles, compressed histogram (so equalisation matters), forward-driving zoom about a vanishing
point plus lateral sway and jitter, additive pixel noise, and deliberate failure sources (independently
moving occluders, gain/offset flicker, a few sudden jumps, an optional moving circular mask as in the
reference's test_tracking.cpp:288-302) so that a few percent of the tracks fail KLT or RANSAC each frame.

Intrinsics default to KAIST cam0 (PL-VIWO/config/kaist/kaist_C/config_camera.yaml:47-50).
"""
from __future__ import annotations

import numpy as np
import cv2

KAIST_K = (8.1690378992770002e02, 8.1156803828490001e02, 6.0850726281690004e02, 2.6347599764440002e02)
KAIST_D = (-5.6143027800000002e-02, 1.3952563200000001e-01, -1.2155906999999999e-03, -9.7281389999999998e-04)


class SynthSequence:
    """Deterministic image sequence; ``frame(t)`` returns an (H, W) uint8 array."""

    def __init__(self, seed: int = 1000, width: int = 1280, height: int = 560, n_frames: int = 300,
                 line_heavy: bool = False, moving_mask: bool = False, hard: bool = True):
        self.seed = int(seed)
        self.W, self.H = int(width), int(height)
        self.n_frames = int(n_frames)
        self.line_heavy = bool(line_heavy)
        self.moving_mask = bool(moving_mask)
        self.hard = bool(hard)
        self.noise_sigma = 0.6 if line_heavy else 1.2
        sc = self.W / 1280.0
        self.sc = sc
        # intrinsics scale with resolution (config 4 uses x1.5)
        self.K = tuple(v * sc for v in KAIST_K)
        self.D = KAIST_D
        rng = np.random.default_rng(self.seed)
        self._build_canvas(rng)
        self._build_motion(rng)
        self._build_occluders(rng)

    # ------------------------------------------------------------------ canvas
    def _build_canvas(self, rng):
        cw, ch = int(self.W * 1.9), int(self.H * 1.9)
        self.cw, self.ch = cw, ch
        acc = np.zeros((ch, cw), np.float32)
        amps = (0.0, 0.0, 1.0) if self.line_heavy else (0.35, 0.8, 0.8)
        for sigma, amp in zip((1.5, 3.0, 6.0), amps):
            n = rng.standard_normal((ch, cw)).astype(np.float32)
            n = cv2.GaussianBlur(n, (0, 0), sigma * max(self.sc, 1.0))
            n /= n.std() + 1e-6
            acc += amp * n
        acc /= acc.std() + 1e-6
        # texture strength varies over the scene (road / sky are smooth, vegetation / facades are busy), like a real
        # urban frame: about a third of the canvas is strongly textured
        amp = cv2.GaussianBlur(rng.standard_normal((ch // 8 + 1, cw // 8 + 1)).astype(np.float32), (0, 0), 6)
        amp = cv2.resize(amp / (amp.std() + 1e-6), (cw, ch), interpolation=cv2.INTER_CUBIC)
        amp = 1.0 / (1.0 + np.exp(-(amp - 0.9) * 4.0))
        img = 120.0 + (3.0 if self.line_heavy else 26.0 * amp + 1.5) * acc
        # broad large-scale shading: keeps the histogram wide so equalisation does not amplify noise in flat regions
        shade = cv2.GaussianBlur(rng.standard_normal((ch // 8 + 1, cw // 8 + 1)).astype(np.float32), (0, 0), 10)
        shade = cv2.resize(shade, (cw, ch), interpolation=cv2.INTER_CUBIC)
        img += (38.0 if self.line_heavy else 30.0) * shade / (shade.std() + 1e-6)
        # piecewise-constant regions ("buildings", "road") give long crisp edges
        n_rect = 36 if self.line_heavy else 28
        for _ in range(n_rect):
            w = int(rng.integers(60, 420) * self.sc)
            h = int(rng.integers(40, 300) * self.sc)
            x = int(rng.integers(0, cw - w))
            y = int(rng.integers(0, ch - h))
            delta = float(rng.choice([-1, 1]) * rng.uniform(18, 55))
            img[y:y + h, x:x + w] += delta
            if rng.random() < (0.0 if self.line_heavy else 0.6):  # window grid (the line-heavy scene keeps its edges long)
                nx, ny = int(rng.integers(2, 6)), int(rng.integers(2, 5))
                for i in range(nx):
                    for j in range(ny):
                        wx = x + int((i + 0.2) * w / nx)
                        wy = y + int((j + 0.2) * h / ny)
                        ww, wh = max(3, int(0.55 * w / nx)), max(3, int(0.55 * h / ny))
                        img[wy:wy + wh, wx:wx + ww] -= delta * 1.3
        n_strokes = 150 if self.line_heavy else 300
        for _ in range(n_strokes):
            L = rng.uniform(60, 400) * self.sc
            ang = rng.uniform(0, np.pi)
            if rng.random() < 0.5:  # favour near-vertical / near-horizontal / towards-vp strokes
                ang = float(rng.choice([0.0, np.pi / 2])) + rng.normal(0, 0.05)
            x0, y0 = rng.uniform(0, cw), rng.uniform(0, ch)
            x1, y1 = x0 + L * np.cos(ang), y0 + L * np.sin(ang)
            val = float(rng.choice([-1, 1]) * rng.uniform(35, 80))
            th = int(rng.integers(1, 4))
            layer = np.zeros((ch, cw), np.float32)
            cv2.line(layer, (int(x0), int(y0)), (int(x1), int(y1)), 1.0, th, cv2.LINE_AA)
            img += val * layer
        if self.line_heavy:
            # BASELINE.json configs[2] wants >= 300 segments longer than TrackLSD's 40 px cut: a jittered lattice of long,
            # slightly tilted bars ("street grid / facade lines").  Crossings are >= 50 px apart in the frame at every zoom
            # of the sequence, so each bar contributes pieces longer than the cut on both of its sides.
            P = 80.0 * self.sc
            for k in range(int(cw / P) + 2):
                x = (k + rng.uniform(-0.2, 0.2)) * P
                tilt = rng.normal(0, 0.03) * ch
                val = float(rng.choice([-1, 1]) * rng.uniform(40, 70))
                layer = np.zeros((ch, cw), np.float32)
                cv2.line(layer, (int(x - tilt / 2), 0), (int(x + tilt / 2), ch - 1), 1.0, int(rng.integers(3, 6)), cv2.LINE_AA)
                img += val * layer
            for k in range(int(ch / P) + 2):
                y = (k + rng.uniform(-0.2, 0.2)) * P
                tilt = rng.normal(0, 0.03) * cw
                val = float(rng.choice([-1, 1]) * rng.uniform(40, 70))
                layer = np.zeros((ch, cw), np.float32)
                cv2.line(layer, (0, int(y - tilt / 2)), (cw - 1, int(y + tilt / 2)), 1.0, int(rng.integers(3, 6)), cv2.LINE_AA)
                img += val * layer
        self.canvas = np.clip(img, 40, 200).astype(np.float32)

    # ------------------------------------------------------------------ motion
    def _build_motion(self, rng):
        n = self.n_frames
        zoom = rng.uniform(0.002, 0.005, n)
        s = np.empty(n)
        s[0] = 1.55
        for t in range(1, n):
            s[t] = s[t - 1] * (1.0 - zoom[t])
            if s[t] < 0.75:  # wrap: "turn a corner" and start over (a jump, tracks are lost)
                s[t] = 1.55
        self.scale = s
        tt = np.arange(n)
        amp = 70.0 * self.sc
        self.cx = self.cw / 2 + amp * np.sin(2 * np.pi * tt / 110.0 + rng.uniform(0, 6.28))
        self.cy = self.ch / 2 + 0.25 * amp * np.sin(2 * np.pi * tt / 170.0 + rng.uniform(0, 6.28))
        self.cx = self.cx + rng.uniform(-0.3, 0.3, n)
        self.cy = self.cy + rng.uniform(-0.3, 0.3, n)
        self.gain = np.ones(n)
        self.offset = np.zeros(n)
        if self.hard:
            self.gain = rng.uniform(0.92, 1.08, n)
            self.offset = rng.uniform(-6, 6, n)
            for t in rng.choice(np.arange(20, max(21, n - 5)), size=max(1, n // 75), replace=False):
                self.cx[t:] += rng.choice([-1, 1]) * rng.uniform(20, 40) * self.sc
        self.vp = (608.0 * self.sc, 263.0 * self.sc)

    def _build_occluders(self, rng):
        self.occ = []
        if not self.hard:
            return
        for _ in range(5):
            w, h = int(rng.integers(50, 140) * self.sc), int(rng.integers(40, 110) * self.sc)
            tex = rng.standard_normal((h, w)).astype(np.float32)
            tex = cv2.GaussianBlur(tex, (0, 0), 1.8)
            tex = 120 + 40 * tex / (tex.std() + 1e-6)
            x0, y0 = rng.uniform(0, self.W), rng.uniform(0.2 * self.H, 0.8 * self.H)
            vx = rng.choice([-1, 1]) * rng.uniform(5, 14) * self.sc
            vy = rng.uniform(-1.5, 1.5) * self.sc
            self.occ.append((np.clip(tex, 40, 200), x0, y0, vx, vy))

    # ------------------------------------------------------------------ frames
    def frame(self, t: int, cam: int = 0) -> np.ndarray:
        """cam 0: the (left) camera; cam 1: the right camera of a stereo pair (KAIST-like rig, stereo_pairs 0-1 in
        config_camera.yaml:5-7): the same scene seen with a constant horizontal disparity of 9 px and a small vertical
        offset, its own pixel noise and gain."""
        t = int(t) % self.n_frames
        s = self.scale[t]
        # frame pixel p -> canvas pixel c = centre + s * (p - vp)
        M = np.array([[s, 0, self.cx[t] - s * self.vp[0]], [0, s, self.cy[t] - s * self.vp[1]]], np.float64)
        if cam:
            M[0, 2] += s * 9.0 * self.sc
            M[1, 2] += s * 0.4 * self.sc
        img = cv2.warpAffine(self.canvas, M, (self.W, self.H), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                             borderMode=cv2.BORDER_REFLECT_101)
        for tex, x0, y0, vx, vy in self.occ:
            h, w = tex.shape
            x = int((x0 + vx * t) % (self.W + w)) - w
            y = int(y0 + vy * t)
            xa, xb = max(x, 0), min(x + w, self.W)
            ya, yb = max(y, 0), min(y + h, self.H)
            if xb > xa and yb > ya:
                img[ya:yb, xa:xb] = tex[ya - y:yb - y, xa - x:xb - x]
        rng = np.random.default_rng(self.seed * 100003 + t + (50000017 if cam else 0))
        noise = rng.standard_normal(img.shape).astype(np.float32)
        # correlated (demosaic-like) noise keeps Canny(50,50) after equalisation readable
        noise = cv2.GaussianBlur(noise, (0, 0), 1.2 if self.line_heavy else 0.8) * (1.0 if self.line_heavy else 1.6)
        img = (self.gain[t] * (1.03 if cam else 1.0)) * img + self.offset[t] + self.noise_sigma * noise
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)

    def mask(self, t: int, cam: int = 0) -> np.ndarray:
        """All-zero mask, or the moving circular mask of test_tracking.cpp:288-302 (r=100 px, 2.5 px/frame)."""
        m = np.zeros((self.H, self.W), np.uint8)
        if self.moving_mask:
            cx = int((100 + 2.5 * t - (9 if cam else 0)) % self.W)
            cy = int(self.H / 2 + 0.3 * self.H * np.sin(t / 40.0))
            cv2.circle(m, (cx, cy), int(100 * self.sc), 255, -1)
        return m

    def timestamp(self, t: int) -> float:
        return 1.0 + 0.1 * t

    def vanishing_points(self, t: int):
        """Three vanishing points (x, y, z) in pixels: stand-in for LineHelper::Vanishing_Points
        (PL-VIWO/src/update/cam/linefeat/LineHelper.cpp:1026-1056), fixed for the synthetic camera."""
        return [(1.0e5, self.vp[1]), (self.vp[0], -1.0e5), (self.vp[0], self.vp[1])]
