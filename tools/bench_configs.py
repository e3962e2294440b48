#!/usr/bin/env python3
"""GPU box: Mrays/s of the BASELINE.json configs other than the headline one (configs 3, 4, 5 -- the
parity-test scenes at their full sizes), one GPU, device-timed like bench.py.  Prints one JSON line each.

  python tools/bench_configs.py [3] [4] [5] [--scale S] [--params JSON]
      --scale < 1 shrinks triangle / instance counts; --params adds pass parameters (e.g. '{"bands": 1}')
The parity taps are off (the C ABI's default), as in bench.py.
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import kiraray_b200 as krr
from kiraray_b200 import scenes

scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
extra_params = json.loads(sys.argv[sys.argv.index("--params") + 1]) if "--params" in sys.argv else {}
args = [a for a in sys.argv[1:] if a.isdigit()]
which = [int(a) for a in args] or [3, 4, 5]


def run(name, desc, cam, w, h, params, frames=3, warm=1, extra=None):
    params = dict(params, debug_taps=False, **extra_params)
    gpu = krr.Wfpt(params=params)
    t0 = time.time()
    gpu.set_scene(desc)
    build_s = time.time() - t0
    gpu.resize(w, h)
    film = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream()
    for f in range(warm):
        gpu.begin_frame(1 + f, cam, st.cuda_stream)
        gpu.render(film.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = 0
    ms = 0.0
    per_frame = []
    for f in range(frames):
        e0.record(st)
        gpu.begin_frame(10 + f, cam, st.cuda_stream)
        gpu.render(film.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
        per_frame.append(round(e0.elapsed_time(e1), 2))
        s = gpu.stats()
        rays += s["closest_rays"] + s["shadow_rays"]
    gpu.set_profiling(True)
    gpu.begin_frame(20, cam, st.cuda_stream)
    gpu.render(film.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    stages = {k: round(v["ms"], 3) for k, v in gpu.stage_times().items() if v["launches"]}
    s = gpu.stats()
    line = {"workload": name, "width": w, "height": h, "params": params, "Mrays_per_s": rays / (ms * 1e-3) / 1e6, "ms_per_frame": ms / frames, "ms_frames": per_frame,
            "rays_per_frame": rays // frames, "accel_build_s": round(build_s, 3), "bvh_nodes": s["bvh_nodes"], "bvh_triangles": s["bvh_triangles"],
            "tlas_nodes": s["tlas_nodes"], "stage_ms": stages, "finite": bool(torch.isfinite(film).all())}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


if 3 in which:
    b = scenes.tessellated_scene(n_objects=max(8, int(200 * scale)), tris_per_object=max(2000, int(100_000 * scale)), n_emissive=1000)
    cam = scenes.look_at_camera((0.4, 0.5, 3.4), (0, -0.1, 0), 16 / 9)
    run("config3_tessellated_%dtris" % b.triangle_count(), b.build(), cam, 1920, 1080, dict(spp=2, max_depth=10))
    del b
if 4 in which:
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", "cbox_smoke.json"), asset_root=ROOT)
    app.set_resolution(1920, 1080)
    app.set_wfpt_params(spp=1, max_depth=15, nee=True)
    run("config4_cbox_smoke_grid_in_mist", app.scene_desc(), app.camera(), 1920, 1080, dict(app.wfpt_params()))
if 5 in which:
    ng = max(4, int(100 * scale ** 0.5))
    b, info = scenes.instanced_scene(n_blas=16, tris_per_blas=max(500, int(20_000 * scale)), n_groups=ng, per_group=ng, motion=True, time=0.5)
    cam = scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 16 / 9, shutter_open=0.5, shutter_time=0.05)
    run("config5_instanced_%dinst_motionblur" % (ng * ng), b.build(), cam, 3840, 2160, dict(spp=1, max_depth=5), extra={"instances": ng * ng})
