"""Experiment runner (not the bench contract): one short device-resident pass and one host-fed pass of the bench workload
per configuration, each configuration in its own process because the knobs are environment variables read at library load.

  python profiles/exp_bench.py --out gpurun_out/exp.jsonl  CFG [CFG ...]
  CFG = comma-separated KEY=VALUE pairs put into the child's environment; the pseudo keys TASKSET=a-b (cpu list),
        STREAMS=n (n > 1: the multi-stream measurement instead of the single-stream one) and STEREO=1 are consumed here;
        EXP_LINES=0 switches the line tracker off, EXP_LA=n sets the lookahead.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/tmp/plviwo_exp_seq.npy"


def child(args):
    import numpy as np
    import torch
    import bench
    import plviwo_b200 as fe_mod
    from plviwo_b200 import synth
    seq = synth.SynthSequence(seed=1000, width=1280, height=560, n_frames=bench.SEQ_FRAMES)
    frames = np.load(CACHE)
    n, H, W = frames.shape
    d_seq = torch.from_numpy(frames).cuda()
    h_seq = torch.from_numpy(frames).pin_memory()
    d_ptrs = [d_seq[t].data_ptr() for t in range(n)]
    h_np = [h_seq[t].numpy() for t in range(n)]
    kw = dict(bench.WORKLOAD)
    if os.environ.get("EXP_LINES") == "0":
        kw["use_lines"] = 0
    if os.environ.get("EXP_LA"):
        bench.LOOKAHEAD = int(os.environ["EXP_LA"])
    out = {"cfg": args.tag}
    if args.stereo:
        right = np.load(CACHE.replace(".npy", "_r.npy"))
        d_r = torch.from_numpy(right).cuda()
        kw.pop("use_lines", None)
        for lines in (0, 1):
            g = fe_mod.StereoFrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=bench.LOOKAHEAD, use_lines=lines, **kw))
            tot, sub = args.warmup + args.steps, 0
            nr = right.shape[0]
            for i in range(tot):
                if i == args.warmup:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                while sub < tot and sub <= i + bench.LOOKAHEAD:
                    t = sub % nr
                    g.submit(seq.timestamp(sub), d_seq[t].data_ptr(), d_r[t].data_ptr(), stride=W, on_device=True,
                             vanishing_points=seq.vanishing_points(t) if lines else None)
                    sub += 1
                g.collect()
            torch.cuda.synchronize()
            out["stereo_pairs_per_s_lines%d" % lines] = args.steps / (time.perf_counter() - t0)
            g.close()
    elif args.streams > 1:
        out["multi"] = bench.run_gpu_multi(fe_mod, torch, seq, d_ptrs, W, args.streams, args.steps, args.warmup, kw, 0)
    else:
        h = fe_mod.FrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=bench.LOOKAHEAD, **kw))
        r = bench.run_gpu_pass(fe_mod, torch, h, seq, d_ptrs, args.steps, args.warmup, True, W, None, timing=False)
        out["fps"] = args.steps / (r["ms"] * 1e-3)
        out["host_ms"] = {k: round(v / max(r["stage"]["frames"], 1), 4) for k, v in r["stage"]["host_ms"].items() if v}
        h.close()
        h = fe_mod.FrontEnd(fe_mod.default_config(K=seq.K, D=seq.D, lookahead=bench.LOOKAHEAD, **kw))
        r = bench.run_gpu_pass(fe_mod, torch, h, seq, h_np, args.steps, args.warmup, False, W, None, timing=False)
        out["e2e_fps"] = args.steps / (r["ms"] * 1e-3)
        h.close()
    print("EXP " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cfgs", nargs="*")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "exp.jsonl"))
    ap.add_argument("--steps", type=int, default=900)
    ap.add_argument("--warmup", type=int, default=60)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--tag", default="")
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--stereo", type=int, default=0)
    args = ap.parse_args()
    if args.child:
        return child(args)
    import numpy as np
    import bench
    if not os.path.exists(CACHE):
        seq, frames = bench.make_frames(1000, bench.SEQ_FRAMES)
        np.save(CACHE, np.stack(frames))
        np.save(CACHE.replace(".npy", "_r.npy"), np.stack([seq.frame(t, 1) for t in range(60)]))
    with open(args.out, "a") as f:
        for cfg in args.cfgs:
            env = dict(os.environ)
            prefix, streams, stereo = [], 1, 0
            for kv in [c for c in cfg.split(",") if c and c != "default"]:
                k, v = kv.split("=", 1)
                if k == "TASKSET":
                    prefix = ["taskset", "-c", v]
                elif k == "STREAMS":
                    streams = int(v)
                elif k == "STEREO":
                    stereo = int(v)
                else:
                    env[k] = v
            cmd = prefix + [sys.executable, os.path.abspath(__file__), "--child", "--tag", cfg, "--steps", str(args.steps),
                            "--warmup", str(args.warmup), "--streams", str(streams), "--stereo", str(stereo)]
            p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
            lines = [l for l in p.stdout.splitlines() if l.startswith("EXP ")]
            rec = json.loads(lines[-1][4:]) if lines else {"cfg": cfg, "error": (p.stderr or p.stdout)[-800:]}
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec)[:400], flush=True)


if __name__ == "__main__":
    main()
