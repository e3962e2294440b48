#!/usr/bin/env python3
"""GPU box: is the frame host-bound?  Times the host side of begin_frame + render (asynchronous launches,
no sync) against the device time of the same frames, per `bands` setting.
usage: python tools/host_issue.py [bands ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import kiraray_b200 as krr

app = bench.make_app(8)
cam = app.camera()
film = torch.empty((bench.H, bench.W, 4), dtype=torch.float32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for bands in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    gpu = krr.Wfpt(params=dict(app.wfpt_params(), debug_taps=False, bands=bands))
    gpu.set_scene(app.scene_desc())
    gpu.resize(bench.W, bench.H)
    for i in range(3):
        gpu.begin_frame(i + 1, cam, s)
        gpu.render(film.data_ptr(), s)
    torch.cuda.synchronize()
    n = 5
    t0 = time.perf_counter()
    for i in range(n):
        gpu.begin_frame(i + 4, cam, s)
        gpu.render(film.data_ptr(), s)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    launches = gpu.stats()["kernel_launches"]
    print(json.dumps({"bands": bands, "host_issue_ms_per_frame": 1e3 * (t1 - t0) / n, "wall_ms_per_frame": 1e3 * (t2 - t0) / n,
                      "launches_per_frame": launches, "host_us_per_launch": 1e6 * (t1 - t0) / n / launches}))
    del gpu
