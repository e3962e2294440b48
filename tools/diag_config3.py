import os, sys, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import kiraray_b200 as krr
from kiraray_b200 import scenes
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
b = scenes.tessellated_scene(n_objects=max(8, int(200 * scale)), tris_per_object=max(2000, int(100_000 * scale)), n_emissive=1000)
cam = scenes.look_at_camera((0.4, 0.5, 3.4), (0, -0.1, 0), 16 / 9)
desc = b.build()
for params in (dict(spp=2, max_depth=10), dict(spp=2, max_depth=10, fuse_stages=False), dict(spp=2, max_depth=10, merge_static=False)):
    gpu = krr.Wfpt(params=params)
    gpu.set_scene(desc)
    gpu.resize(1920, 1080)
    film = torch.empty((1080, 1920, 4), dtype=torch.float32, device="cuda")
    for f in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        gpu.begin_frame(10 + f, cam)
        t1 = time.time()
        gpu.render(film.data_ptr())
        t2 = time.time()
        torch.cuda.synchronize(); t3 = time.time()
        s = gpu.stats()
        print(params, "frame", f, "begin %.1f ms, render issue %.1f ms, sync %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), s["closest_by_depth"][:11], flush=True)
    gpu.set_profiling(True)
    gpu.begin_frame(9, cam); gpu.render(film.data_ptr()); torch.cuda.synchronize()
    lt = gpu.launch_times(); print(len(lt), sum(ms for _, ms in lt), [(n, round(ms, 2)) for n, ms in lt if ms > 5], flush=True)
