#!/usr/bin/env python3
"""Per-launch device times (CUDA events on the launching stream) of one warm frame of a bench
workload, in issue order, with the queue sizes per depth.  GPU box only.
usage: tools/launch_times.py [workload] [spp] [params JSON]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import kiraray_b200 as krr
import torch

key = sys.argv[1] if len(sys.argv) > 1 else "cbox"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
extra = json.loads(sys.argv[3]) if len(sys.argv) > 3 else {}
wl = bench.Workload(key, spp)
gpu = krr.Wfpt(params=dict(wl.params, spp=spp, debug_taps=False, **extra))
gpu.set_scene(wl.desc)
gpu.resize(wl.W, wl.H)
film = torch.empty((wl.H, wl.W, 4), dtype=torch.float32, device="cuda")
for f in (1, 2, 3):
    gpu.begin_frame(f, wl.camera(f))
    gpu.render(film.data_ptr())
torch.cuda.synchronize()
gpu.set_profiling(True)
gpu.begin_frame(4, wl.camera(4))
gpu.render(film.data_ptr())
torch.cuda.synchronize()
lt = gpu.launch_times()
st = gpu.stats()
print("closest_by_depth", st["closest_by_depth"][:wl.max_depth + 2])
print("shadow_by_depth ", st["shadow_by_depth"][:wl.max_depth + 1])
tot = {}
for name, ms in lt:
    tot[name] = tot.get(name, 0) + ms
print("totals(ms)", {k: round(v, 3) for k, v in tot.items()}, "sum", round(sum(tot.values()), 3))
print(" ".join(f"{n[:3]}:{ms * 1e3:.0f}" for n, ms in lt))
