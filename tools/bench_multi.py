#!/usr/bin/env python3
"""GPU box with N GPUs: the C++ MultiDeviceRenderApp (one process, one host thread per device, NCCL inside the product
library) on a bench workload.  Prints one JSON line.  usage: tools/bench_multi.py <workload> <n_gpus> <tiles> [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from kiraray_b200.binding import MultiDeviceApp

key, n, tiles = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
wl = bench.Workload(key)
params = {**wl.params, "spp": wl.spp, "debug_taps": False, "frame_batch": wl.batch}
app = MultiDeviceApp(wl.desc, params, wl.W, wl.H, list(range(n)), tiles)
app.render(wl.camera(0), 1, 3)                       # warm-up
film, ms, _ = app.render(wl.camera(0), 100, steps)
# rays: one more step per slice at the same frame indices as the timed run's last step
_, _, rays_last = app.render(wl.camera(0), 100 + (steps - 1) * (n // tiles) * wl.batch, 1)
print(json.dumps({"impl": "MultiDeviceRenderApp (C++ host layer, one process)", "workload": wl.spec["name"], "n_gpus": n, "tiles": tiles, "spp_slices": n // tiles,
                  "nccl": app.uses_nccl, "steps": steps, "ms_per_step": ms / steps, "Mrays_per_s_estimate": rays_last / (ms / steps * 1e-3) / 1e6,
                  "note": "wall clock of the slowest rank thread, film read-back pipelined on rank 0; rays of the last step x steps"}))
