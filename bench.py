#!/usr/bin/env python
"""bench.py — front-end frames/s of the B200-native PL-VIWO visual front end (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one frame of one camera stream through the whole point + line front end (equalise, pyramid, top-off
FAST detection + sub-pixel, pyramidal LK, RANSAC gate, Canny + line segments, line/point association).
Workload at every N: BASELINE.json configs[1] — synthetic KAIST-shaped 1280x560 sequence, 400 points, maxLevel 4,
15x15 window, 5x5 grid, lines on — one independent stream per GPU (weak scaling: streams never exchange data, so
there is no collective on the data path; NCCL is only used for the barrier and the max-over-ranks reduction).

  value  : frames/s with the frames already resident in HBM (device pointers handed to plviwo_fe_submit), the
           sequence (215 MB) is larger than L2, so no frame is served from cache
  e2e    : the same frames from pinned HOST memory through the C ABI (H2D of every frame and D2H of every result inside
           the timed region)
  roofline / cpu_baseline / clocks : see DESIGN.md "Measurement"
--impl reference times the reference's CPU path: the oracle's restatement of TrackKLT/TrackLSD driving the real
OpenCV kernels (cv2) on the host cores — the reference itself cannot be compiled in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# One hardware work queue per CUDA stream (default 8): otherwise the latency-critical LK launch can be queued behind a
# multi-millisecond line-walk kernel of another stream that happens to share its queue.  Must be set before the
# CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

WORKLOAD = dict(width=1280, height=560, num_features=400, fast_threshold=20, grid_x=5, grid_y=5, min_px_dist=10,
                pyr_levels=4, win_size=15)
WORKLOAD_NAME = "BASELINE.json configs[1]: synthetic KAIST-shaped 1280x560 mono, point+line front end, 400 pts, 5x5 grid, maxLevel 4, win 15"
SEQ_FRAMES = 300
LOOKAHEAD = 48
PIPELINE_FILL = 2 * (LOOKAHEAD + 2) + 4   # untimed frames a fresh handle needs before it is in steady state
METRIC = "front-end frames/sec @1280x560"
KERNEL_OF_STAGE = {"hist": "k_hist", "eq_pyr1": "k_eq_pyr1", "pyr_rest": "k_pyr_down", "fast": "k_fast", "subpix": "k_corner_subpix",
                   "lk": "k_lk15", "canny": "k_canny", "fld_walk": "k_fld_walk_cc", "fld_ccl": "k_ccl_merge", "fld_seg": "k_fld_segments"}


def algorithmic_bytes(n_lk_pts: float, cfg=WORKLOAD) -> dict:
    """every stage reads its input once and writes its output once."""
    N = cfg["width"] * cfg["height"]
    L, w = cfg["pyr_levels"], cfg["win_size"]
    sizes, ww, hh = [], cfg["width"], cfg["height"]
    for _ in range(L + 1):
        sizes.append(ww * hh)
        ww, hh = (ww + 1) // 2, (hh + 1) // 2
    b = {
        "hist": N,                                            # read the frame
        "eq_pyr1": N + N + sizes[1] + N // 4,                 # read frame, write level 0, level 1, half-res image
        "pyr_rest": sum(sizes[l - 1] + sizes[l] for l in range(2, L + 1)),
        "fast": N,                                            # worst case: every cell valid
        "canny": N // 4 + N // 4,                             # Canny read + write (one byte per half-res pixel)
        "fld_walk": N // 4,                                   # chain walk: every half-res pixel read once
        "fld_ccl": 0, "fld_seg": 0,                           # implementation artefacts of the parallel walk: no SURVEY bytes
        "lk": n_lk_pts * (L + 1) * ((w + 3) ** 2 + (w + 1) ** 2),
        "subpix": 0,
    }
    b["frame_total"] = 3 * N + sum(sizes[l - 1] + sizes[l] for l in range(1, L + 1)) + N + (N + 4 * (N // 4)) + b["lk"]
    return b


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for nme, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def make_frames(seed: int, n: int):
    import plviwo_b200  # noqa: F401
    from plviwo_b200 import synth
    seq = synth.SynthSequence(seed=seed, width=WORKLOAD["width"], height=WORKLOAD["height"], n_frames=n)
    return seq, [seq.frame(t) for t in range(n)]


# ----------------------------------------------------------------------------------------------- CPU arm
def run_cpu(seq, frames, steps, warmup, threads=None, rows_out=None):
    """The reference's CPU path: oracle restatement of the glue + the real OpenCV kernels (cv2)."""
    import cv2
    from oracle import frontend as ofe
    if threads:
        cv2.setNumThreads(threads)
    kw = {k: v for k, v in WORKLOAD.items() if k not in ("width", "height")}
    from oracle import cvops
    in_kernels = [0.0]

    class TimedOps:
        """oracle.cvops with the time spent INSIDE the OpenCV calls (and the FastLineDetector / std::sort shim) added up:
        the rest of a frame is the Python restatement of the reference's glue, which the C++ reference does much faster."""

        def __getattr__(self, name):
            f = getattr(cvops, name)
            if not callable(f):
                return f

            def timed(*a, **k):
                t = time.perf_counter()
                r = f(*a, **k)
                in_kernels[0] += time.perf_counter() - t
                return r
            return timed

    fe = ofe.FrontEnd(ofe.FeConfig(K=seq.K, D=seq.D, **kw), ops=TimedOps())
    n = len(frames)
    def keep(prow):
        if rows_out is not None:
            rows_out.append(np.array([(r.id, r.u, r.v) for r in prow], np.float64).reshape(-1, 3))

    for t in range(warmup):
        keep(fe.feed(seq.timestamp(t), frames[t % n], None, seq.vanishing_points(t % n))[0])
    per = []
    in_kernels[0] = 0.0
    t0 = time.perf_counter()
    for k in range(steps):
        t = warmup + k
        a = time.perf_counter()
        prow, _ = fe.feed(seq.timestamp(t), frames[t % n], None, seq.vanishing_points(t % n))
        per.append(time.perf_counter() - a)
        keep(prow)
    dt = time.perf_counter() - t0
    return dict(fps=steps / dt, ms_per_step=1e3 * dt / steps, p50_ms=1e3 * float(np.median(per)), cores=cv2.getNumThreads(),
                kernel_ms=1e3 * in_kernels[0] / steps, glue_ms=1e3 * (dt - in_kernels[0]) / steps,
                kernels_only_fps=steps / in_kernels[0] if in_kernels[0] > 0 else None)


def parity_against(fe_mod, seq, frames, rows_cpu, kw, dev):
    """BASELINE.json's third metric ("KLT px err"): the frames the CPU baseline has just processed, through the synchronous
    drop-in call on the GPU from the same initial state, both free-running.  The oracle is the checker here, nothing of it
    is timed.  A single flipped status flag changes every later feature id, so id equality is reported
    up to the first frame where the row sets differ and the pixel error over the rows both sides have."""
    h = fe_mod.FrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=0, **kw), device=dev)
    n = len(frames)
    first_div, eq_frames, duv, rows_total, sym = None, 0, [], 0, 0
    for t, want in enumerate(rows_cpu):
        h.feed_new_camera(seq.timestamp(t), frames[t % n], None, seq.vanishing_points(t % n), update_db=False)
        got = h.point_rows()
        ids_g = {int(i): k for k, i in enumerate(got["id"])}
        ids_c = {int(i): k for k, i in enumerate(want[:, 0])}
        d = set(ids_g) ^ set(ids_c)
        rows_total += len(ids_c)
        if first_div is None:
            if d:
                first_div = t
            else:
                eq_frames += 1
        if first_div is None or t == first_div:   # same tracks on both sides: compare positions
            sym += len(d)
            for i in set(ids_g) & set(ids_c):
                a, b = got[ids_g[i]], want[ids_c[i]]
                duv.append(max(abs(float(a["u"]) - b[1]), abs(float(a["v"]) - b[2])))
    h.close()
    duv = np.array(duv) if duv else np.zeros(1)
    # the per-frame bar: the GPU tracker is loaded with the oracle's state before every frame (teacher forcing), so each
    # frame is compared on identical inputs — what tests/test_frontend_gpu.py asserts, here over 40 frames as a number
    from oracle import frontend as ofe
    okw = {k: v for k, v in kw.items() if k not in ("width", "height")}
    o = ofe.FrontEnd(ofe.FeConfig(K=seq.K, D=seq.D, **okw))
    h = fe_mod.FrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=0, **kw), device=dev)
    W, H = kw["width"], kw["height"]
    tf_duv, tf_rows, tf_sym, tf_frames_equal = [], 0, 0, 0
    for t in range(40):
        if t > 0:
            k, l = o.klt.get_state(), o.lsd.get_state()
            h.set_state(fe_mod.pack_state(W, H, k["currid"], k["pts_last"], k["ids_last"], k["img_last"], k["mask_last"], l["currid"],
                                          l["lines_last"], l["ids_last"], l["pol_last"]))
        prow, _ = o.feed(seq.timestamp(t), frames[t % n], None, seq.vanishing_points(t % n))
        h.feed_new_camera(seq.timestamp(t), frames[t % n], None, seq.vanishing_points(t % n), update_db=False)
        got = h.point_rows()
        ids_g = {int(i): k2 for k2, i in enumerate(got["id"])}
        ids_c = {r.id: r for r in prow}
        d = set(ids_g) ^ set(ids_c)
        tf_rows += len(ids_c)
        tf_sym += len(d)
        tf_frames_equal += int(not d and [r.id for r in prow] == [int(i) for i in got["id"]])
        for i in set(ids_g) & set(ids_c):
            a, b = got[ids_g[i]], ids_c[i]
            tf_duv.append(max(abs(float(a["u"]) - b.u), abs(float(a["v"]) - b.v)))
    h.close()
    tf_duv = np.array(tf_duv) if tf_duv else np.zeros(1)
    teacher = {"frames": 40, "frames_rows_identical_ids_and_order": tf_frames_equal, "rows": tf_rows,
               "rows_with_flipped_status": tf_sym, "status_agreement": 1.0 - tf_sym / max(tf_rows, 1),
               "max_duv_px": float(tf_duv.max()), "p99_duv_px": float(np.percentile(tf_duv, 99)),
               "rows_over_0.05px": int((tf_duv > 0.05).sum())}
    return {"teacher_forced": teacher, "frames": len(rows_cpu), "frames_ids_identical": eq_frames, "first_frame_with_different_rows": first_div,
            "rows_compared": int(len(duv)), "max_duv_px": float(duv.max()), "p99_duv_px": float(np.percentile(duv, 99)),
            "rows_only_on_one_side_at_divergence": sym,
            "note": "free-running GPU (plviwo_fe_feed) vs the CPU baseline's oracle on the same frames; the teacher-forced "
                    "per-frame bars (0.05 px, ids bit-exact, status >= 99.5 %) are asserted in tests/test_frontend_gpu.py"}


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu_pass(fe_mod, torch, handle, seq, srcs, steps, warmup, on_device, pitch, dist, timing):
    """Pipelined submit/collect over `steps` frames after `warmup` frames; returns (elapsed_ms by CUDA events, infos)."""
    n = len(srcs)
    # every frame slot of the handle runs its first frame with direct launches and captures its CUDA graphs on its second:
    # the untimed part covers two rounds over the lookahead + 2 slots so that neither lands in the timed region
    warmup = max(warmup, PIPELINE_FILL)
    tot = warmup + steps
    sub = 0
    rows = 0

    def submit(i):
        t = i % n
        if on_device:
            handle.submit(seq.timestamp(i), srcs[t], stride=pitch, on_device=True, vanishing_points=seq.vanishing_points(t))
        else:
            handle.submit(seq.timestamp(i), srcs[t], vanishing_points=seq.vanishing_points(t))

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per = []
    for i in range(tot):
        if i == warmup:
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            handle.enable_timing(timing)
            handle.stage_times(reset=True)
            ev0.record()
            t_wall = time.perf_counter()
        while sub < tot and sub <= i + LOOKAHEAD:
            submit(sub)
            sub += 1
        info = handle.collect()
        if i >= warmup:
            per.append(time.perf_counter())   # completion times: the per-frame period is their difference
            rows += info.n_point_rows + info.n_line_rows
    torch.cuda.synchronize()
    ev1.record()
    ev1.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    if dist is not None:
        dist.barrier()
    st = handle.stage_times(reset=False)
    handle.enable_timing(False)
    period = np.diff(np.array(per)) if len(per) > 1 else np.zeros(1)
    return dict(ms=float(ev0.elapsed_time(ev1)), wall_ms=wall_ms, stage=st, rows=rows, p50_ms=1e3 * float(np.median(period)),
                p95_ms=1e3 * float(np.percentile(period, 95)))


def run_gpu_multi(fe_mod, torch, seq, d_ptrs, pitch, n_streams, steps, warmup, cfg_kw, dev):
    """configs[4] shape on one GPU: n_streams independent handles (own CUDA streams, own tracker threads), each driven by
    one host thread through plviwo_fe_play (the submit/collect loop inside the library, so the Python GIL is not part of
    the measurement).  All streams replay the same device-resident sequence from different start frames (they never
    exchange data, so this is n_streams times the single-stream work)."""
    import threading
    n = len(d_ptrs)
    la = int(os.environ.get("PLVIWO_BENCH_LA", "16"))
    warmup = max(warmup, 2 * (la + 2) + 4)
    handles = [fe_mod.FrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=la, **cfg_kw), device=dev) for _ in range(n_streams)]
    gate = threading.Barrier(n_streams + 1)
    frames_done = [0] * n_streams

    def drive(k):
        h = handles[k]
        off = (k * 37) % n
        idx = [(off + i) % n for i in range(warmup + steps)]
        ts = [seq.timestamp(i) for i in range(warmup + steps)]
        ptrs = [d_ptrs[t] for t in idx]
        vps = [seq.vanishing_points(t) for t in idx] if not os.environ.get("PLVIWO_BENCH_NOLINES") else None
        h.play(ts[:warmup], ptrs[:warmup], stride=pitch, on_device=True, vanishing_points=vps[:warmup] if vps else None)
        gate.wait()      # everybody warmed up
        gate.wait()      # timed region starts
        if os.environ.get("PLVIWO_BENCH_TIMING"):
            h.enable_timing(True)
            h.stage_times(reset=True)
        st = h.play(ts[warmup:], ptrs[warmup:], stride=pitch, on_device=True, vanishing_points=vps[warmup:] if vps else None)
        frames_done[k] = int(st.frames)
        if os.environ.get("PLVIWO_BENCH_TIMING") and k == 0:
            tt = h.stage_times()
            sys.stderr.write("stream0 stage ms/frame: %s\n" % {a: round(b / max(tt["frames"], 1), 4) for a, b in tt["ms"].items()})
            sys.stderr.write("stream0 host ms/frame: %s\n" % {a: round(b / max(tt["frames"], 1), 4) for a, b in tt["host_ms"].items()})

    th = [threading.Thread(target=drive, args=(k,)) for k in range(n_streams)]
    for t in th:
        t.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gate.wait()
    torch.cuda.synchronize()
    ev0.record()
    t0 = time.perf_counter()
    gate.wait()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    ev1.record()
    ev1.synchronize()
    wall = time.perf_counter() - t0
    for h in handles:
        h.close()
    ms = float(ev0.elapsed_time(ev1))
    return {"streams": n_streams, "frames": sum(frames_done), "ms": ms, "value": sum(frames_done) / (ms * 1e-3), "unit": "frames/s",
            "wall_fps": sum(frames_done) / wall, "lookahead": la,
            "note": "BASELINE.json configs[4] shape on ONE GPU: independent streams, one host driver thread each "
                    "(plviwo_fe_play), frames resident in HBM"}


def run_stereo(fe_mod, torch, seq, h_left, steps, warmup, kw, dev):
    """the stereo rig (TrackKLT::feed_stereo + the left-image line tracker) through
    plviwo_fe_stereo_submit / _collect from pinned HOST pairs; pairs/s by wall clock around a device synchronise."""
    n = min(len(h_left), 60)
    H, W = h_left[0].shape
    h_r = torch.empty((n, H, W), dtype=torch.uint8).pin_memory()
    for t in range(n):
        h_r[t].copy_(torch.from_numpy(seq.frame(t, 1)))
    right = [h_r[t].numpy() for t in range(n)]
    g = fe_mod.StereoFrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=LOOKAHEAD, **kw), device=dev)
    warmup = max(warmup, PIPELINE_FILL)
    tot, sub, rows = warmup + steps, 0, 0
    for i in range(tot):
        if i == warmup:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        while sub < tot and sub <= i + LOOKAHEAD:
            t = sub % n
            g.submit(seq.timestamp(sub), h_left[t], right[t], vanishing_points=seq.vanishing_points(t))
            sub += 1
        info = g.collect()
        if i >= warmup:
            rows += info.n_point_rows[0] + info.n_point_rows[1] + info.n_line_rows
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = g.stage_times()
    g.close()
    return {"value": steps / dt, "unit": "stereo pairs/s", "rows_per_pair": rows / max(steps, 1),
            "h2d_bytes_per_pair": 2 * H * W, "gpu_launches": st["kernel_launches_total"],
            "api": "plviwo_fe_stereo_submit/_collect from pinned host pairs (60-pair loop), left-image line tracker on, lookahead %d" % LOOKAHEAD}


def pin_to_gpu_numa_node(torch, dev):
    """The handle's host threads poll flags and exchange the (small) feature arrays with the LK kernel through pinned host
    memory: keep the rank on the cores of the NUMA node its GPU hangs off.  Returns the node, or None when the platform
    does not say (single socket, virtualised PCI topology)."""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 4:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=150, help="frames of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stereo", action="store_true", help="skip the extra stereo-rig measurement")
    ap.add_argument("--multi-streams", type=int, default=8, help="streams per GPU of the extra multi-stream measurement (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # exactly ONE line on stdout: libraries that print banners there (NCCL's version line) are sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        seq, frames = make_frames(1000, SEQ_FRAMES)
        steps = min(args.steps, 400)   # bounded sample of the same workload
        r = run_cpu(seq, frames, steps, min(args.warmup, 20))
        line = {"metric": METRIC, "value": r["fps"], "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 20), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/int32 fixed-point + f32/f64 (OpenCV)", "data": "synthetic", "impl": "reference",
                "config": {"workload": WORKLOAD_NAME, "streams": 1},
                "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port",
                                 "sample": "%d frames of the same synthetic sequence; reference glue restated in Python "
                                           "(oracle/frontend.py) driving the real OpenCV kernels through cv2; per frame %.2f ms "
                                           "inside OpenCV / the FLD shim + %.2f ms Python glue" % (steps, r["kernel_ms"], r["glue_ms"]),
                                 "kernels_only_value": r["kernels_only_fps"],
                                 "kernels_only_note": "frames/s if the reference's glue cost nothing (upper bound for any "
                                                      "CPU implementation built on these OpenCV kernels)"},
                "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "p50_ms_per_frame": r["p50_ms"]}
        emit(line)
        return 0

    import torch
    import plviwo_b200 as fe_mod
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the front end has no CPU fallback (use --impl reference for the CPU arm)")
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    numa = pin_to_gpu_numa_node(torch, dev)

    seq, frames = make_frames(1000 + rank, SEQ_FRAMES)   # stream s = seed 1000 + s
    H, W = frames[0].shape
    # device-resident copy of the sequence (215 MB > 126 MB L2) and a pinned host copy
    d_seq = torch.empty((len(frames), H, W), dtype=torch.uint8, device="cuda")
    h_seq = torch.empty((len(frames), H, W), dtype=torch.uint8).pin_memory()
    for t, f in enumerate(frames):
        h_seq[t].copy_(torch.from_numpy(f))
    d_seq.copy_(h_seq)
    torch.cuda.synchronize()
    d_ptrs = [d_seq[t].data_ptr() for t in range(len(frames))]
    h_np = [h_seq[t].numpy() for t in range(len(frames))]

    kw = {k: v for k, v in WORKLOAD.items()}
    handle = fe_mod.FrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=LOOKAHEAD, **kw), device=dev)

    # untimed dry pass first: a fresh process's GPU is still ramping its clocks during the first ~100 ms of work, longer
    # than the 104 untimed frames (8 ms) in front of the timed region
    run_gpu_pass(fe_mod, torch, handle, seq, d_ptrs, min(args.steps, 1000), args.warmup, True, W, dist, timing=False)
    sampler = ClockSampler(dev) if rank == 0 else None
    res = run_gpu_pass(fe_mod, torch, handle, seq, d_ptrs, args.steps, args.warmup, True, W, dist, timing=False)
    clocks = sampler.stop() if sampler else {}
    # per-kernel durations: the same pass again with CUDA-event stage timing on (direct launches instead of graph replays)
    res_t = run_gpu_pass(fe_mod, torch, handle, seq, d_ptrs, min(args.steps, 200), args.warmup, True, W, None, timing=True)
    handle.close()
    handle = fe_mod.FrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=LOOKAHEAD, **kw), device=dev)
    run_gpu_pass(fe_mod, torch, handle, seq, h_np, min(args.steps, 1000), args.warmup, False, W, dist, timing=False)   # untimed dry pass
    res_e2e = run_gpu_pass(fe_mod, torch, handle, seq, h_np, args.steps, args.warmup, False, W, dist, timing=False)
    handle.close()
    # strict drop-in: synchronous plviwo_fe_feed per frame from host memory
    handle = fe_mod.FrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=0, **kw), device=dev)
    n_sync = min(args.steps, 300)
    for t in range(args.warmup):
        handle.feed_new_camera(seq.timestamp(t), h_np[t % len(h_np)], None, seq.vanishing_points(t % len(h_np)), update_db=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n_sync):
        t = args.warmup + k
        handle.feed_new_camera(seq.timestamp(t), h_np[t % len(h_np)], None, seq.vanishing_points(t % len(h_np)), update_db=False)
    sync_fps = n_sync / (time.perf_counter() - t0)
    handle.close()

    stereo = None
    if world == 1 and not args.no_stereo:
        try:
            stereo = run_stereo(fe_mod, torch, seq, h_np, min(args.steps, 300), args.warmup, kw, dev)
        except Exception as e:   # an extra, never fatal for the bench line
            stereo = {"error": str(e)}
    multi = None
    if world == 1 and args.multi_streams > 1:
        multi = run_gpu_multi(fe_mod, torch, seq, d_ptrs, W, args.multi_streams, min(args.steps, 400), args.warmup, kw, dev)

    ms, ms_e2e = res["ms"], res_e2e["ms"]
    per_rank = [[ms, ms_e2e]]
    total_frames = args.steps * world
    if dist is not None:
        # whole-job numbers: SUM of frames, MAX of elapsed time over the ranks (pl-viwo_b200/shard.py)
        tt = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        allms = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allms, tt)
        per_rank = [[float(v[0]), float(v[1])] for v in allms]
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])
        total_frames = int(round(fe_mod.shard.distributed_throughput(dist, torch, args.steps, ms, device="cuda") * ms * 1e-3))
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    st = res_t["stage"]
    nfr = max(st["frames"], 1)
    # roofline of the dominant kernel (largest share of the timed region), live CUDA-event durations
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    lk_pts = 0.0
    if st["launches"]["lk"]:
        lk_pts = res_t["rows"] / max(st["launches"]["lk"], 1)  # lower bound: rows written; LK input is slightly larger
    ab = algorithmic_bytes(max(lk_pts, 1.0))
    # single kernels only: "fld" is the sum of fld_ccl + fld_walk + fld_seg, pyr_rest / fld_ccl / fld_seg are 3-4 launches
    stage_ms = {k: v for k, v in st["ms"].items() if k not in ("h2d", "fld", "line_frames") and st["launches"][k]}
    # line-path kernels are launched once per BATCH of frames (grid.y = frame): frames carried / launches per launch
    LINE_STAGES = ("canny", "fld", "fld_ccl", "fld_walk", "fld_seg")
    line_frames = st["launches"].get("line_frames", 0)
    fpl = {k: (line_frames / max(st["launches"][k], 1) if k in LINE_STAGES and line_frames else 1.0) for k in st["ms"]}
    frames_of = {k: (line_frames if k in LINE_STAGES and line_frames else nfr) for k in st["ms"]}
    dom = max(stage_ms, key=lambda k: stage_ms[k] / max(frames_of[k], 1)) if stage_ms else "lk"   # largest cost per frame
    avg_ms = stage_ms.get(dom, 0.0) / max(st["launches"][dom], 1)
    achieved = (ab[dom] * fpl.get(dom, 1.0) / (avg_ms * 1e-3)) / 1e9 if avg_ms > 0 else 0.0
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full summary
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["kernels"].get(KERNEL_OF_STAGE.get(dom, dom))
    except Exception:
        pass
    if traffic is not None and fpl.get(dom, 1.0) != 1.0:
        traffic = traffic * fpl[dom]   # the ncu --set full capture profiles single-frame launches: scaled to the frames a launch carries
    per_kernel = {}
    for k, v in stage_ms.items():
        ms_k = v / max(st["launches"][k], 1)
        per_kernel[k] = {"avg_ms": ms_k, "frames_per_launch": round(fpl[k], 2), "algorithmic_bytes": ab.get(k, 0) * fpl[k],
                         "GBps": (ab.get(k, 0) * fpl[k] / (ms_k * 1e-3)) / 1e9 if ms_k > 0 else 0.0}
        per_kernel[k]["frac"] = per_kernel[k]["GBps"] / peak if peak else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ab[dom] * fpl.get(dom, 1.0), "frames_per_launch": round(fpl.get(dom, 1.0), 2),
                "avg_launch_ms": avg_ms,
                "note": "the dominant kernel is the sequential chain walk of the line detector (one launch per batch of "
                        "frames): latency bound, not HBM bound (DESIGN.md 'Kernels'); per_kernel lists every kernel of the "
                        "frame, stage_ms_per_frame their cost per frame",
                "per_kernel": per_kernel,
                "stage_ms_per_frame": {k: v / max(frames_of[k], 1) for k, v in st["ms"].items() if k != "line_frames"},
                "host_ms_per_frame": {k: v / max(res["stage"]["frames"], 1) for k, v in res["stage"]["host_ms"].items()
                                      if not k.startswith("unused")},
                "whole_frame": {"algorithmic_bytes": ab["frame_total"],
                                "achieved_GBps": ab["frame_total"] * (total_frames / (ms * 1e-3)) / 1e9 / world}}
    # what the image-domain kernels reach when ONE launch carries a batch of frames' worth of pixels (134 MB > L2) instead
    # of one 0.7 MB frame: same kernels, device-resident synthetic image, CUDA events around single launches
    at_scale = None
    try:
        AW, AH = 16384, 8192
        t = fe_mod.op_image_kernels_time(AW, AH, 5, device=dev)
        NA = AW * AH
        bytes_ = {"hist": NA, "eq_pyr1": NA + NA + NA // 4 + NA // 4, "fast": NA, "canny": NA // 4 + NA // 4}
        at_scale = {"image": "%dx%d (= %.0f frames of 1280x560 per launch)" % (AW, AH, NA / (1280 * 560.0)),
                    "kernels": {k: {"ms": t[k], "algorithmic_bytes": bytes_[k], "GBps": bytes_[k] / (t[k] * 1e-3) / 1e9,
                                    "frac": bytes_[k] / (t[k] * 1e-3) / 1e9 / peak} for k in t if t[k] > 0}}
    except Exception as e:   # never fatal for the bench line
        at_scale = {"error": str(e)}
    roofline["at_scale"] = at_scale
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        rows_cpu = []
        r = run_cpu(seq, frames, args.cpu_sample, 10, rows_out=rows_cpu)
        try:
            parity = parity_against(fe_mod, seq, h_np, rows_cpu, kw, dev)
        except Exception as e:   # an extra, never fatal for the bench line
            parity = {"error": str(e)}
        cpu = {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port",
               "sample": "%d frames of the same sequence; oracle/frontend.py (reference glue restated) driving real OpenCV "
                         "kernels via cv2, %d OpenCV threads; p50 %.2f ms/frame = %.2f ms inside OpenCV / the FLD shim + %.2f ms "
                         "Python glue" % (args.cpu_sample, r["cores"], r["p50_ms"], r["kernel_ms"], r["glue_ms"]),
               "kernels_only_value": r["kernels_only_fps"],
               "kernels_only_note": "frames/s if the reference's glue cost nothing (upper bound for any CPU implementation "
                                    "built on these OpenCV kernels)"}
    line = {
        "metric": METRIC, "value": total_frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32 fixed-point + f32 (LK), f64 (sub-pixel, undistort, RANSAC)", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "streams_per_gpu": 1, "streams": world, "stream_to_gpu": "stream s -> rank s mod N, seed 1000 + s",
                   "lookahead": LOOKAHEAD, "sequence_frames": SEQ_FRAMES,
                   "untimed_frames": max(args.warmup, PIPELINE_FILL),
                   "host": {"cores": os.cpu_count(), "rank0_numa_node": numa, "rank0_cpus": len(os.sched_getaffinity(0))},
                   "per_rank_ms": {"resident": [round(v[0], 2) for v in per_rank], "e2e": [round(v[1], 2) for v in per_rank]},
                   "cache": "inputs larger than L2 (215 MB device-resident sequence, every frame read once per pass)"},
        "p50_ms_per_frame": res["p50_ms"], "p95_ms_per_frame": res["p95_ms"],
        "p50_note": "median time between consecutive frame completions (plviwo_fe_collect returns) in the timed region; the "
                    "strictly synchronous per-frame latency is 1000 / e2e.sync_feed_fps ms",
        "klt_parity": parity,
        "e2e": {"value": total_frames / (ms_e2e * 1e-3), "unit": "frames/s",
                "h2d_bytes_per_step": res_e2e["stage"]["h2d_bytes"] / max(res_e2e["stage"]["frames"], 1),
                "d2h_bytes_per_step": res_e2e["stage"]["d2h_bytes"] / max(res_e2e["stage"]["frames"], 1),
                "api": "plviwo_fe_submit/plviwo_fe_collect from pinned host frames, lookahead %d" % LOOKAHEAD,
                "sync_feed_fps": sync_fps},
        "gpu_launches": res["stage"]["kernel_launches_total"],
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "multi_stream": multi, "stereo": stereo,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
