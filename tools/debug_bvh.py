#!/usr/bin/env python3
"""GPU box: first-hit ids of the wide-BVH traversal against the oracle's brute-force loop on scenes of growing
complexity; prints how the mismatches split (GPU miss / oracle miss / different primitive)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import kiraray_b200 as krr, oracle_binding as ob
from kiraray_b200 import scenes

def check(name, b, cam, w=96, h=96, **kw):
    desc = b.build()
    gpu = krr.Wfpt(params=dict(spp=1, max_depth=2, **kw)); gpu.set_scene(desc); gpu.resize(w, h); gpu.begin_frame(1, cam)
    gpu.render_to_host()
    inst, prim = gpu.first_hits()
    orc = ob.Oracle(desc, "reference"); ref = orc.render(cam, w, h, frame_index=1, spp=1, max_depth=2, use_bvh=True); orc.close()
    ri, rp = ref["first_hits"][:, 0], ref["first_hits"][:, 1]
    bad = (inst != ri) | (prim != rp)
    st = gpu.stats()
    print(f"{name:28s} tris {b.triangle_count():8d} nodes {st['bvh_nodes']:7d} tlas {st['tlas_nodes']:4d} mismatches {int(bad.sum()):6d} / {w*h}: gpu-miss {int((bad & (inst < 0)).sum())} "
          f"oracle-miss {int((bad & (ri < 0)).sum())} other {int((bad & (inst >= 0) & (ri >= 0)).sum())}  hits {int((ri >= 0).sum())}", flush=True)
    return bad

rng = np.random.Generator(np.random.PCG64(7272))
cam = scenes.look_at_camera((0, 0.5, 4.5), (0, 0, 0), 1.0)
for nu, nv in ((8, 5), (12, 8), (24, 16), (64, 40), (200, 120)):
    sph = scenes.displaced_sphere(nu, nv, rng, amplitude=0.05)
    for xf, label in ((scenes.IDENTITY, "identity(merged)"), (scenes.translation((0.1, 0, 0), 1.0), "translated(tlas)")):
        b = scenes.SceneBuilder()
        m = b.add_material(diffuse=(0.7, 0.4, 0.3))
        b.add_instance(b.add_mesh(sph[0], sph[2], sph[1], m), xf)
        p, n, idx = scenes.quad((-1, 3.0, -1), (2, 0, 0), (0, 0, 2))
        b.add_instance(b.add_mesh(p, idx, n, b.add_material(diffuse=(0, 0, 0), emissive=(17, 12, 4))), scenes.translation((0, 0.01, 0), 1.0))
        check(f"sphere {nu}x{nv} {label}", b, cam, flat_blas_max=0)
b = scenes.tessellated_scene(n_objects=12, tris_per_object=3000, n_emissive=40)
check("tessellated 12x3000", b, scenes.look_at_camera((0.4, 0.5, 3.4), (0, -0.1, 0), 1.0))
b, info = scenes.instanced_scene(n_blas=3, tris_per_blas=300, n_groups=4, per_group=6, motion=False)
check("instanced static", b, scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0))
b, info = scenes.instanced_scene(n_blas=3, tris_per_blas=300, n_groups=4, per_group=6, motion=True)
check("instanced motion", b, scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0, shutter_open=0.5, shutter_time=0.05))
