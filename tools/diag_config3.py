#!/usr/bin/env python3
"""GPU box: per-launch device times of one frame of config 3 (tools/bench_configs.py) -- finds the launch
behind an outlier frame.  usage: tools/diag_config3.py [scale] [frame]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import kiraray_b200 as krr
from kiraray_b200 import scenes
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
frame = int(sys.argv[2]) if len(sys.argv) > 2 else 10
b = scenes.tessellated_scene(n_objects=max(8, int(200 * scale)), tris_per_object=max(2000, int(100_000 * scale)), n_emissive=1000)
cam = scenes.look_at_camera((0.4, 0.5, 3.4), (0, -0.1, 0), 16 / 9)
gpu = krr.Wfpt(params=dict(spp=2, max_depth=10))
gpu.set_scene(b.build())
gpu.resize(1920, 1080)
film = torch.empty((1080, 1920, 4), dtype=torch.float32, device="cuda")
gpu.begin_frame(1, cam); gpu.render(film.data_ptr()); torch.cuda.synchronize()
for f in (frame, frame + 1):
    gpu.set_profiling(True)
    gpu.begin_frame(f, cam); gpu.render(film.data_ptr()); torch.cuda.synchronize()
    lt = gpu.launch_times()
    s = gpu.stats()
    print("frame", f, "launches", len(lt), "sum ms", round(sum(ms for _, ms in lt), 2), [(i, n, round(ms, 2)) for i, (n, ms) in enumerate(lt) if ms > 1.0])
    print("  closest_by_depth", s["closest_by_depth"][:11], "shadow_by_depth", s["shadow_by_depth"][:10], flush=True)
    gpu.set_profiling(False)
