#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE's own KRR_CALLABLE code (oracle/_ref, built from
/root/reference/src by oracle/build_oracle.py).  Run in the dev container only (needs the reference
tree to build oracle/_ref); the vectors are committed so that the GPU box -- which has no
/root/reference -- and the port backend can be checked against them.

  python tools/make_golden.py            # rewrites tests/golden/leaf_vectors.npz, cbox_48.npz

Everything is seeded (numpy PCG64, seed 7272 = KRR_DEFAULT_RND_SEED, reference core/config.in.h:19).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from leaf_cases import LeafCases  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(GOLD, exist_ok=True)
    lib = ob.load("reference")
    assert lib.ol_backend_name() == b"reference"
    cases = LeafCases(seed=7272)
    out = cases.evaluate(lib)
    np.savez_compressed(os.path.join(GOLD, "leaf_vectors.npz"), **out)
    print("leaf_vectors.npz:", {k: v.shape for k, v in out.items()})

    import kiraray_b200 as krr
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", "cbox.json"), asset_root=ROOT)
    w = h = 48
    app.set_resolution(w, h)
    app.set_wfpt_params(spp=2, max_depth=5)
    cam = app.camera()
    orc = ob.Oracle(app.scene_desc(), "reference")
    ref = orc.render(cam, w, h, frame_index=1, spp=2, max_depth=5, use_bvh=False, capture=(1, 1))
    st = ref["stats"]
    np.savez_compressed(
        os.path.join(GOLD, "cbox_48.npz"), film=ref["film"], first_hits=ref["first_hits"], sampler=ref["sampler"],
        lambda_=ref["lambda"], camera_sample=ref["camera_sample"],
        closest_by_depth=np.array(st["closest_by_depth"], np.int64), shadow_by_depth=np.array(st["shadow_by_depth"], np.int64),
        totals=np.array([st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["scatter_items"], st["hit_light_items"], st["miss_items"]], np.int64),
        **{f"queue{q}": ref["queues"][q] for q in range(6)})
    print("cbox_48.npz: rays", st["closest_rays"] + st["shadow_rays"])


if __name__ == "__main__":
    main()
