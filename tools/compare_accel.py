#!/usr/bin/env python3
"""GPU box: renders the bench workload with every acceleration-structure layout (tree / flat list,
merged / per-instance) and reports where ray counts or film bits differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import kiraray_b200 as krr

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1
app = bench.make_app(spp)
cam = app.camera()
res = {}
for merge in (False, True):
    for flat in (0, 48):
        gpu = krr.Wfpt(params=dict(app.wfpt_params(), merge_static=merge, flat_blas_max=flat))
        gpu.set_scene(app.scene_desc())
        gpu.resize(bench.W, bench.H)
        gpu.begin_frame(4, cam)
        film = gpu.render_to_host().copy()
        st = gpu.stats()
        res[(merge, flat)] = (film, st, gpu.first_hits())
        print(merge, flat, st["closest_by_depth"][:11], st["shadow_by_depth"][:10], flush=True)
ref = res[(True, 48)]
for k, v in res.items():
    d = (v[0].view(np.uint32) != ref[0].view(np.uint32)).any(axis=-1)
    fh = (v[2][0] != ref[2][0]) | (v[2][1] != ref[2][1])
    print(k, "pixels differing from merged+flat:", int(d.sum()), "first hits differing:", int(fh.sum()), np.argwhere(d)[:5].tolist())
