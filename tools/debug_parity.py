import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import kiraray_b200 as krr, oracle_binding as ob
from kiraray_b200 import scenes
from __graft_entry__ import relmse

def both(desc, cam, w, h, spp, md, **kw):
    gpu = krr.Wfpt(params=dict(spp=spp, max_depth=md, **kw)); gpu.set_scene(desc); gpu.resize(w, h); gpu.begin_frame(1, cam)
    film = gpu.render_to_host()
    orc = ob.Oracle(desc, "reference"); ref = orc.render(cam, w, h, frame_index=1, spp=spp, max_depth=md, use_bvh=True); orc.close()
    return gpu, film, ref

rng = np.random.Generator(np.random.PCG64(7272))
sph = scenes.displaced_sphere(24, 16, rng, amplitude=0.05)
mats = {"diffuse": dict(diffuse=(0.7, 0.4, 0.3), bsdf_type=1), "dielectric": dict(diffuse=(1, 1, 1), roughness=0.0, bsdf_type=2, ior=1.5),
        "conductor": dict(diffuse=(0.9, 0.7, 0.3), roughness=0.3, bsdf_type=3, ior=0.4), "disney": dict(diffuse=(0.3, 0.5, 0.8), roughness=0.5, bsdf_type=4)}
for lname in ("infinite",):
    for mname, mk in mats.items():
        b = scenes.SceneBuilder()
        m = b.add_material(**mk)
        b.add_instance(b.add_mesh(sph[0], sph[2], sph[1], m), scenes.translation((0, 0, 0), 1.0))
        p, n, idx = scenes.quad((-5, -1.2, -5), (0, 0, 10), (10, 0, 0))
        b.add_instance(b.add_mesh(p, idx, n, b.add_material(diffuse=(0.5, 0.5, 0.5), bsdf_type=1)))
        if lname == "point": b.add_light(0, color=(1, 0.9, 0.8), scale=40.0, transform=scenes.translation((0, 4, 2)))
        elif lname == "infinite": b.add_light(4, color=(0.4, 0.5, 0.7), scale=1.0, scene_radius=12.0)
        else:
            p, n, idx = scenes.quad((-1, 3.0, -1), (2, 0, 0), (0, 0, 2))
            b.add_instance(b.add_mesh(p, idx, n, b.add_material(diffuse=(0, 0, 0), emissive=(17, 12, 4))))
        desc = b.build()
        cam = scenes.look_at_camera((0, 1.0, 4.5), (0, 0, 0), 1.0)
        gpu, film, ref = both(desc, cam, 64, 64, 4, 5)
        st, rs = gpu.stats(), ref["stats"]
        print(f"{lname:9s} {mname:11s} relmse {relmse(film, ref['film']):8.4f} mean {film[...,:3].mean():.4f}/{ref['film'][...,:3].mean():.4f} closest {st['closest_rays']}/{rs['closest_rays']} shadow {st['shadow_rays']}/{rs['shadow_rays']} miss {st['miss_items']}/{rs['miss_items']}")

def blocks(img, k=8):
    h, w = img.shape[:2]
    return img[..., :3].reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))

for cfg in ("cbox_mist.json", "cbox_smoke.json"):
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", cfg), asset_root=ROOT); app.set_resolution(96, 96); app.set_wfpt_params(spp=1, max_depth=8)
    cam = app.camera()
    gpu, film, ref = both(app.scene_desc(), cam, 96, 96, 1, 8)
    inst, prim = gpu.first_hits()
    bad = np.where((inst != ref["first_hits"][:, 0]) | (prim != ref["first_hits"][:, 1]))[0]
    print(cfg, "first-hit mismatches (1 spp)", len(bad), [(int(i), int(inst[i]), int(prim[i]), ref["first_hits"][i].tolist()) for i in bad[:6]])
    for spp in (32,):
        gpu, film, ref = both(app.scene_desc(), cam, 96, 96, spp, 8)
        st, rs = gpu.stats(), ref["stats"]
        print(spp, {k: round(st[k] / rs[k], 4) for k in ("closest_rays", "shadow_rays", "scatter_items", "miss_items", "medium_sample_items", "medium_scatter_items", "hit_light_items")})
        print("   closest_by_depth gpu", st["closest_by_depth"][:10]); print("   closest_by_depth ref", rs["closest_by_depth"][:10])
        print("   shadow_by_depth  gpu", st["shadow_by_depth"][:10]); print("   shadow_by_depth  ref", rs["shadow_by_depth"][:10])
        gb, rb = blocks(film), blocks(ref["film"])
        print("   relmse", round(relmse(film, ref["film"]), 4), "block-relmse", round(relmse(gb, rb), 5), "mean", film[..., :3].mean(), ref["film"][..., :3].mean(),
              "ratio", film[..., :3].mean() / ref["film"][..., :3].mean())
