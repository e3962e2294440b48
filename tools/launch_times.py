#!/usr/bin/env python3
"""Per-launch device times (CUDA events on the launching stream) of one warm frame of the bench
workload, in issue order, with the queue sizes per depth.  GPU box only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import kiraray_b200 as krr
import torch

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1
app = bench.make_app(spp)
cam = app.camera()
gpu = krr.Wfpt(params=dict(app.wfpt_params()))
gpu.set_scene(app.scene_desc())
gpu.resize(bench.W, bench.H)
film = torch.empty((bench.H, bench.W, 4), dtype=torch.float32, device="cuda")
for f in (1, 2, 3):
    gpu.begin_frame(f, cam)
    gpu.render(film.data_ptr())
torch.cuda.synchronize()
gpu.set_profiling(True)
gpu.begin_frame(4, cam)
gpu.render(film.data_ptr())
torch.cuda.synchronize()
lt = gpu.launch_times()
st = gpu.stats()
print("closest_by_depth", st["closest_by_depth"][:12])
print("shadow_by_depth ", st["shadow_by_depth"][:12])
tot = {}
for name, ms in lt:
    tot[name] = tot.get(name, 0) + ms
print("totals(ms)", {k: round(v, 3) for k, v in tot.items()}, "sum", round(sum(tot.values()), 3))
print(" ".join(f"{n[:3]}:{ms * 1e3:.0f}" for n, ms in lt))
