import json,sys
d=json.loads(sys.stdin.read())
r=d["roofline"]
print(round(d["value"]), round(d["e2e"]["value"]))
print(r["kernel"], r["achieved"], r["frac"], r["algorithmic_bytes_per_launch"], r["frames_per_launch"], r["avg_launch_ms"], r["traffic"])
print({k:(round(v["avg_ms"],4), v["frames_per_launch"]) for k,v in r["per_kernel"].items()})
print({k:round(v,4) for k,v in r["stage_ms_per_frame"].items()})
