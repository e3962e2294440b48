#!/usr/bin/env python3
"""GPU box: which first hits of the textured test scene differ between the GPU and the oracle (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import kiraray_b200 as krr
import oracle_binding as ob
import test_gpu_textured as T

b, cam = T.textured_scene()
desc = b.build()
w = h = 96
for spp in (1, 4):
    gpu = krr.Wfpt(params=dict(spp=spp, max_depth=4))
    gpu.set_scene(desc); gpu.resize(w, h); gpu.begin_frame(1, cam)
    gpu.render_to_host()
    inst, prim = gpu.first_hits()
    orc = ob.Oracle(desc, "reference")
    ref = orc.render(cam, w, h, frame_index=1, spp=spp, max_depth=4, use_bvh=False)
    refb = orc.render(cam, w, h, frame_index=1, spp=spp, max_depth=4, use_bvh=True)
    ri, rp = ref["first_hits"][:, 0], ref["first_hits"][:, 1]
    print("spp", spp, "oracle brute vs bvh differ:", int(((ri != refb["first_hits"][:, 0]) | (rp != refb["first_hits"][:, 1])).sum()))
    bad = np.nonzero((inst != ri) | (prim != rp))[0]
    print("spp", spp, "differ", len(bad))
    for k in bad[:40]:
        print("  pixel", k % w, k // w, "gpu", inst[k], prim[k], "oracle", ri[k], rp[k])
    orc.close()
