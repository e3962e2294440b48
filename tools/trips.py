import sys, os
sys.path.insert(0, "/root/repo")
os.chdir("/root/repo")
import bench, torch
import kiraray_b200 as krr
key = sys.argv[1]
wl = bench.Workload(key, 1)
gpu = krr.Wfpt(params=dict(wl.params, spp=1, debug_taps=False))
gpu.set_scene(wl.desc); gpu.resize(wl.W, wl.H)
film = torch.empty((wl.H, wl.W, 4), dtype=torch.float32, device="cuda")
gpu.begin_frame(1, wl.camera(1)); gpu.render(film.data_ptr()); torch.cuda.synchronize(); gpu.stats()
