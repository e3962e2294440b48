"""Seeded inputs for the leaf functions of the hot path (sampler, wavelengths, colour, camera, BSDFs,
lights, phase function) and their evaluation through any backend that exports oracle/oracle_leaf.h.

Used three ways:
  * tools/make_golden.py evaluates them with the REFERENCE's own code (oracle/_ref) -> tests/golden/
  * tests/test_oracle_golden.py re-evaluates with whatever oracle backend is present and compares
  * tests/test_gpu_leaf_parity.py sends the same inputs through the CUDA device functions (debug
    entry points of the C ABI) and compares with the golden outputs
TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C

import numpy as np

import oracle_binding as ob

F = C.c_float


class OlLight(C.Structure):
    _fields_ = [("type", C.c_int), ("color", F * 3), ("scale", F), ("position", F * 3), ("rotation", F * 9), ("sceneRadius", F),
                ("cosInner", F), ("cosOuter", F), ("xform", F * 12), ("xformInv", F * 12)]


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _rot(rng):
    q = _unit(rng.normal(size=4))
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class LeafCases:
    N_BSDF = 96  # per material type

    def __init__(self, seed=7272):
        r = np.random.Generator(np.random.PCG64(seed))
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        self.pcg = np.stack([r.integers(0, 1920, 32), r.integers(0, 1080, 32), r.integers(0, 1 << 20, 32)], 1).astype(np.int64)
        self.wl_u = f32(r.random(64))
        self.rgb = f32(np.concatenate([r.random((48, 3)), r.random((16, 3)) * 20.0]))
        self.rgb_u = f32(r.random(64))
        self.spec = f32(r.random((64, 4)) * 3.0)
        self.spec_u = f32(r.random(64))
        # cameras: [film w, film h, focal length, focal distance, lens radius, aspect, shutter open, shutter time] + 3x4
        cams = []
        for k in range(2):
            R, t = _rot(r), r.normal(size=3) * 3
            cams.append(np.concatenate([[36.0, 24.0, 21.0 + 10 * k, 5.0, 0.0 if k == 0 else 0.05, 1.5, 0.5 * k, 0.05 * k],
                                        np.concatenate([R, t[:, None]], 1).ravel()]))
        self.cams = f32(cams)
        self.cam_px = np.stack([r.integers(0, 640, 32), r.integers(0, 480, 32)], 1).astype(np.int32)
        self.cam_cs = f32(r.random((32, 5)))
        # BSDF cases
        n = self.N_BSDF * 5
        bt = np.repeat(np.arange(5), self.N_BSDF)
        rough = r.random(n)
        rough[r.random(n) < 0.12] = 5e-4  # delta lobes
        metallic = np.where(r.random(n) < 0.3, 0.0, np.where(r.random(n) < 0.3, 1.0, r.random(n)))
        strans = np.where(r.random(n) < 0.5, 0.0, np.where(r.random(n) < 0.3, 1.0, r.random(n)))
        self.bsdf = dict(
            type=bt.astype(np.int32), ior=f32(1.1 + r.random(n)), diffuse=f32(r.random((n, 4))), specular=f32(r.random((n, 4)) * (r.random((n, 1)) < 0.7)),
            strans=f32(strans), rough=f32(rough), metallic=f32(metallic), aniso=f32(np.where(r.random(n) < 0.5, 0.0, r.random(n))),
            eta_kind=(r.random(n) < 0.5).astype(np.int32), eta=f32(0.2 + 2 * r.random(n)), k_kind=(r.random(n) < 0.5).astype(np.int32), k=f32(3 * r.random(n)),
            wl_u=f32(r.random(n)), seed=r.integers(0, 1 << 30, (n, 2)).astype(np.int64))
        wo = _unit(r.normal(size=(n, 3)))
        upper = ~np.isin(bt, [2, 4]) | (r.random(n) < 0.7)  # transmissive materials also see wo from below
        wo[:, 2] = np.where(upper, np.abs(wo[:, 2]), wo[:, 2])
        self.bsdf["wo"] = f32(wo)
        self.bsdf["wi"] = f32(_unit(r.normal(size=(n, 3))))
        # emissive triangles
        m = 64
        P = r.normal(size=(m, 3, 3))
        fn = _unit(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]))
        xf = []
        for _ in range(m):
            R = _rot(r) * (0.5 + 2 * r.random())
            xf.append(np.concatenate([R, (r.normal(size=3) * 2)[:, None]], 1).ravel())
        self.tri = dict(p=f32(P), n=f32(np.repeat(fn[:, None, :], 3, 1)), xform=f32(xf), Le=f32(r.random((m, 3))), scale=f32(1 + 20 * r.random(m)),
                        two_sided=(r.random(m) < 0.3).astype(np.int32), u=f32(r.random((m, 2))), ctx_p=f32(r.normal(size=(m, 3)) * 4),
                        ctx_n=f32(_unit(r.normal(size=(m, 3)))), wl_u=f32(r.random(m)))
        # analytic lights
        m = 48
        rots, xfs, invs = [], [], []
        for _ in range(m):
            R, t = _rot(r), r.normal(size=3) * 3
            rots.append(R.ravel())
            xfs.append(np.concatenate([R, t[:, None]], 1).ravel())
            invs.append(np.concatenate([R.T, (-R.T @ t)[:, None]], 1).ravel())
        inner = 5 + 30 * r.random(m)
        self.light = dict(type=np.tile(np.array([0, 1, 2, 4], np.int32), m // 4), color=f32(r.random((m, 3))), scale=f32(1 + 5 * r.random(m)),
                          rotation=f32(rots), xform=f32(xfs), xform_inv=f32(invs), radius=f32(5 + 10 * r.random(m)),
                          cos_inner=f32(np.cos(np.radians(inner))), cos_outer=f32(np.cos(np.radians(inner + 5 + 20 * r.random(m)))),
                          u=f32(r.random((m, 2))), ctx_p=f32(r.normal(size=(m, 3)) * 2), wi=f32(_unit(r.normal(size=(m, 3)))), wl_u=f32(r.random(m)))
        self.hg = dict(g=f32(r.uniform(-0.9, 0.9, 32)), wo=f32(_unit(r.normal(size=(32, 3)))), wi=f32(_unit(r.normal(size=(32, 3)))), u=f32(r.random((32, 2))))
        self.metal = dict(d=f32(r.random((32, 3))), s=f32(r.random((32, 3)) * (r.random((32, 1)) < 0.8)))

    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def wavelengths(lib, u):
        lam, pdf = (F * 4)(), (F * 4)()
        lib.ol_sample_wavelengths(float(u), lam, pdf)
        return lam, pdf

    def shading(self, lib, i):
        b = self.bsdf
        lam, pdf = self.wavelengths(lib, b["wl_u"][i])
        sd = ob.OlShading(IoR=b["ior"][i], diffuse=(F * 4)(*b["diffuse"][i]), specular=(F * 4)(*b["specular"][i]),
                          specularTransmission=b["strans"][i], roughness=b["rough"][i], metallic=b["metallic"][i], anisotropic=b["aniso"][i],
                          bsdfType=int(b["type"][i]), woWorld=(F * 3)(*b["wo"][i]), lambda_=lam, pdf=pdf,
                          etaKind=int(b["eta_kind"][i]), etaValue=(F * 4)(*([b["eta"][i]] * 4)), kKind=int(b["k_kind"][i]), kValue=(F * 4)(*([b["k"][i]] * 4)))
        return sd

    def evaluate(self, lib):
        """Runs every case through `lib` (an oracle_leaf.h backend) -> dict of numpy arrays."""
        fa = ob.fa
        out = {}
        # PCG
        st, fl = [], []
        for px, py, idx in self.pcg:
            s = ob.OlSampler()
            lib.ol_pcg_set_pixel_sample(C.byref(s), int(px), int(py), int(idx))
            lib.ol_pcg_advance(C.byref(s), 256 * (int(py) * 1920 + int(px)))
            fl.append([lib.ol_pcg_get1d(C.byref(s)) for _ in range(8)])
            st.append([s.state, s.inc])
        out["pcg_state"], out["pcg_floats"] = np.array(st, np.uint64), np.array(fl, np.float32)
        # wavelengths / colour
        lam = []
        for u in self.wl_u:
            l, p = self.wavelengths(lib, u)
            lam.append(list(l) + list(p))
        out["wavelengths"] = np.array(lam, np.float32)
        res = []
        for rgb, u in zip(self.rgb, self.rgb_u):
            l, _ = self.wavelengths(lib, u)
            row = []
            for t in range(3):
                o = (F * 4)()
                lib.ol_from_rgb(fa(*rgb), t, l, o)
                row += list(o)
            res.append(row)
        out["from_rgb"] = np.array(res, np.float32)
        res = []
        for s, u in zip(self.spec, self.spec_u):
            l, p = self.wavelengths(lib, u)
            rgb = (F * 3)()
            lib.ol_to_rgb(fa(*s), l, p, rgb)
            res.append(list(rgb) + [lib.ol_lum(fa(*s), l, p)])
        out["to_rgb_lum"] = np.array(res, np.float32)
        # camera
        res = []
        for k, (px, cs) in enumerate(zip(self.cam_px, self.cam_cs)):
            c = self.cams[k % 2]
            cam = ob.OlCamera(filmSize=(F * 2)(c[0], c[1]), focalLength=c[2], focalDistance=c[3], lensRadius=c[4], aspectRatio=c[5],
                              shutterOpen=c[6], shutterTime=c[7], transform=(F * 12)(*c[8:20]))
            o, d, t = (F * 3)(), (F * 3)(), F()
            lib.ol_camera_ray(C.byref(cam), int(px[0]), int(px[1]), 640, 480, fa(*cs), o, d, C.byref(t))
            res.append(list(o) + list(d) + [t.value])
        out["camera"] = np.array(res, np.float32)
        # BSDFs
        b = self.bsdf
        ev, sm, ty = [], [], []
        for i in range(len(b["type"])):
            sd = self.shading(lib, i)
            ty.append(lib.ol_bsdf_type(C.byref(sd)))
            f, pdf = (F * 4)(), F()
            lib.ol_bsdf_f_pdf(C.byref(sd), fa(*b["wo"][i]), fa(*b["wi"][i]), f, C.byref(pdf))
            ev.append(list(f) + [pdf.value])
            s = ob.OlSampler()
            lib.ol_pcg_set_pixel_sample(C.byref(s), int(b["seed"][i][0]) & 0xffff, int(b["seed"][i][0]) >> 16, int(b["seed"][i][1]))
            f2, wi, pdf2, fl2 = (F * 4)(), (F * 3)(), F(), C.c_int()
            lib.ol_bsdf_sample(C.byref(sd), fa(*b["wo"][i]), C.byref(s), f2, wi, C.byref(pdf2), C.byref(fl2))
            sm.append(list(f2) + list(wi) + [pdf2.value, float(fl2.value)])
        out["bsdf_type"], out["bsdf_eval"], out["bsdf_sample"] = np.array(ty, np.int32), np.array(ev, np.float32), np.array(sm, np.float32)
        # area lights
        t = self.tri
        res = []
        for i in range(len(t["scale"])):
            tl = ob.OlTriLight(xform=(F * 12)(*t["xform"][i]), Le=(F * 3)(*t["Le"][i]), scale=t["scale"][i], twoSided=int(t["two_sided"][i]))
            for c in range(3):
                for k in range(3):
                    tl.p[c][k], tl.n[c][k] = t["p"][i][c][k], t["n"][i][c][k]
            l, _ = self.wavelengths(lib, t["wl_u"][i])
            p, n, L, pdf = (F * 3)(), (F * 3)(), (F * 4)(), F()
            lib.ol_arealight_sample_li(C.byref(tl), fa(*t["u"][i]), fa(*t["ctx_p"][i]), fa(*t["ctx_n"][i]), l, p, n, L, C.byref(pdf))
            w = np.array(t["ctx_p"][i]) - np.array(list(p))
            w = (w / max(np.linalg.norm(w), 1e-20)).astype(np.float32)
            L2 = (F * 4)()
            lib.ol_arealight_L(C.byref(tl), p, n, fa(*w), l, L2)
            pdf2 = lib.ol_arealight_pdf_li(C.byref(tl), p, n, fa(*t["ctx_p"][i]), fa(*t["ctx_n"][i]))
            res.append(list(p) + list(n) + list(L) + [pdf.value] + list(L2) + [pdf2])
        out["arealight"] = np.array(res, np.float32)
        # analytic lights
        g = self.light
        res = []
        for i in range(len(g["type"])):
            ol = OlLight(type=int(g["type"][i]), color=(F * 3)(*g["color"][i]), scale=g["scale"][i], position=(F * 3)(*g["xform"][i][[3, 7, 11]]),
                         rotation=(F * 9)(*g["rotation"][i]), sceneRadius=g["radius"][i], cosInner=g["cos_inner"][i], cosOuter=g["cos_outer"][i],
                         xform=(F * 12)(*g["xform"][i]), xformInv=(F * 12)(*g["xform_inv"][i]))
            l, _ = self.wavelengths(lib, g["wl_u"][i])
            p, L, pdf, Li = (F * 3)(), (F * 4)(), F(), (F * 4)()
            lib.ol_light_sample_li(C.byref(ol), fa(*g["u"][i]), fa(*g["ctx_p"][i]), l, p, L, C.byref(pdf))
            if g["type"][i] == 4:
                lib.ol_inflight_Li(C.byref(ol), fa(*g["wi"][i]), l, Li)
            res.append(list(p) + list(L) + [pdf.value] + list(Li))
        out["light"] = np.array(res, np.float32)
        # HG phase function, getMetallic
        res = []
        h = self.hg
        lib.ol_hg_sample.argtypes = [F, C.POINTER(F), C.POINTER(F), C.POINTER(F), C.POINTER(F), C.POINTER(F)]
        for i in range(len(h["g"])):
            wi, p, pdf = (F * 3)(), F(), F()
            lib.ol_hg_sample(float(h["g"][i]), fa(*h["wo"][i]), fa(*h["u"][i]), wi, C.byref(p), C.byref(pdf))
            res.append([lib.ol_hg_p(float(h["g"][i]), fa(*h["wo"][i]), fa(*h["wi"][i]))] + list(wi) + [p.value, pdf.value])
        out["hg"] = np.array(res, np.float32)
        out["metallic"] = np.array([lib.ol_get_metallic(fa(*d), fa(*s)) for d, s in zip(self.metal["d"], self.metal["s"])], np.float32)
        return out
