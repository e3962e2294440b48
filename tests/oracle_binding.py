"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this)."""
import ctypes as C
import os

import numpy as np

from kiraray_b200.binding import KrrCameraData, KrrSceneDesc, KrrStats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F, I32, U64, P = C.c_float, C.c_int32, C.c_uint64, C.c_void_p


class OrcParams(C.Structure):
    _fields_ = [("nee", I32), ("enable_medium", I32), ("max_depth", I32), ("enable_clamp", I32), ("spp", I32),
                ("rr", F), ("clamp_max", F), ("use_bvh", I32), ("threads", I32), ("row_begin", I32), ("row_end", I32)]


class OlSampler(C.Structure):
    _fields_ = [("state", U64), ("inc", U64)]


class OlCamera(C.Structure):
    _fields_ = [("filmSize", F * 2), ("focalLength", F), ("focalDistance", F), ("lensRadius", F), ("aspectRatio", F),
                ("shutterOpen", F), ("shutterTime", F), ("transform", F * 12)]


class OlShading(C.Structure):
    _fields_ = [("IoR", F), ("diffuse", F * 4), ("specular", F * 4), ("specularTransmission", F), ("roughness", F),
                ("metallic", F), ("anisotropic", F), ("bsdfType", I32), ("woWorld", F * 3), ("lambda_", F * 4), ("pdf", F * 4),
                ("etaKind", I32), ("etaValue", F * 4), ("kKind", I32), ("kValue", F * 4)]


class OlTriLight(C.Structure):
    _fields_ = [("p", (F * 3) * 3), ("n", (F * 3) * 3), ("xform", F * 12), ("Le", F * 3), ("scale", F), ("twoSided", I32)]


def paths():
    return {"reference": os.path.join(ROOT, "oracle/_ref/libkrr_oracle_ref.so")}


_libs = {}


def available(kind):
    return os.path.exists(paths()[kind])


def load(kind="reference"):
    """kind: 'reference' -- the reference's own classes compiled host-side (oracle/_ref) under the restated stage driver."""
    if kind in _libs:
        return _libs[kind]
    path = paths()[kind]
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: build with `python oracle/build_oracle.py ref`")
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    lib.ol_backend_name.restype = C.c_char_p
    lib.ol_pcg_get1d.restype = F
    lib.ol_pcg_get1d.argtypes = [C.POINTER(OlSampler)]
    lib.ol_pcg_set_pixel_sample.argtypes = [C.POINTER(OlSampler), C.c_uint32, C.c_uint32, C.c_uint32]
    lib.ol_pcg_advance.argtypes = [C.POINTER(OlSampler), C.c_int64]
    lib.ol_sample_wavelengths.argtypes = [F, C.POINTER(F), C.POINTER(F)]
    lib.ol_from_rgb.argtypes = [C.POINTER(F), C.c_int, C.POINTER(F), C.POINTER(F)]
    lib.ol_to_rgb.argtypes = [C.POINTER(F)] * 4
    lib.ol_lum.argtypes = [C.POINTER(F)] * 3
    lib.ol_lum.restype = F
    lib.ol_camera_ray.argtypes = [C.POINTER(OlCamera), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(F), C.POINTER(F), C.POINTER(F), C.POINTER(F)]
    lib.ol_bsdf_type.argtypes = [C.POINTER(OlShading)]
    lib.ol_bsdf_f_pdf.argtypes = [C.POINTER(OlShading), C.POINTER(F), C.POINTER(F), C.POINTER(F), C.POINTER(F)]
    lib.ol_bsdf_sample.argtypes = [C.POINTER(OlShading), C.POINTER(F), C.POINTER(OlSampler), C.POINTER(F), C.POINTER(F), C.POINTER(F), C.POINTER(C.c_int)]
    lib.ol_arealight_sample_li.argtypes = [C.POINTER(OlTriLight)] + [C.POINTER(F)] * 8
    lib.ol_arealight_L.argtypes = [C.POINTER(OlTriLight)] + [C.POINTER(F)] * 5
    lib.ol_arealight_pdf_li.argtypes = [C.POINTER(OlTriLight)] + [C.POINTER(F)] * 4
    lib.ol_arealight_pdf_li.restype = F
    lib.ol_get_metallic.argtypes = [C.POINTER(F), C.POINTER(F)]
    lib.ol_get_metallic.restype = F
    lib.ol_hg_p.argtypes = [F, C.POINTER(F), C.POINTER(F)]
    lib.ol_hg_p.restype = F
    lib.orc_scene_create.argtypes = [C.POINTER(KrrSceneDesc)]
    lib.orc_scene_create.restype = P
    lib.orc_scene_destroy.argtypes = [P]
    lib.orc_scene_num_lights.argtypes = [P]
    lib.orc_render.restype = C.c_double
    lib.orc_render.argtypes = [P, C.POINTER(OrcParams), C.POINTER(KrrCameraData), I32, I32, U64, P, P, P, P, P,
                               C.POINTER(KrrStats), I32, I32, C.POINTER(P), C.POINTER(I32)]
    lib.orc_instance_xf.argtypes = [P, I32, F, C.POINTER(F), C.POINTER(F)]
    lib.orc_render_megakernel.restype = C.c_double
    lib.orc_render_megakernel.argtypes = [P, C.POINTER(OrcParams), C.POINTER(KrrCameraData), I32, I32, U64, P]
    lib.orc_instance_xf_div.argtypes = [P, I32, F, C.POINTER(F), C.POINTER(F)]
    lib.orc_intersect_triangle.argtypes = [C.POINTER(F)] * 5 + [F] + [C.POINTER(F)] * 3
    lib.ol_init()
    _libs[kind] = lib
    return lib


def fa(*v):
    return (F * len(v))(*v)


class Oracle:
    def __init__(self, scene_desc_ptr, kind="reference"):
        self.lib = load(kind)
        self.kind = kind
        self.scene = self.lib.orc_scene_create(scene_desc_ptr)

    def close(self):
        if self.scene:
            self.lib.orc_scene_destroy(self.scene)
            self.scene = None

    def instance_xf(self, inst, time, division_form=False):
        """object->world / world->object of an instance at a ray time; division_form: the original division-based
        SRT evaluation (an independent second statement of the same transform)"""
        m, inv = (F * 12)(), (F * 12)()
        (self.lib.orc_instance_xf_div if division_form else self.lib.orc_instance_xf)(self.scene, inst, time, m, inv)
        return np.array(m, np.float32), np.array(inv, np.float32)

    def render(self, cam, w, h, frame_index=1, spp=1, max_depth=10, rr=0.8, nee=True, use_bvh=True, threads=0,
               capture=None, rows=None, enable_clamp=False, clamp_max=1e3, enable_medium=True):
        p = OrcParams(nee=int(nee), enable_medium=int(enable_medium), max_depth=max_depth, enable_clamp=int(enable_clamp), spp=spp, rr=rr,
                      clamp_max=clamp_max, use_bvh=int(use_bvh), threads=threads,
                      row_begin=rows[0] if rows else 0, row_end=rows[1] if rows else 0)
        n = w * h
        out = {"film": np.zeros((h, w, 4), np.float32), "first_hits": np.full((n, 2), -2, np.int32),
               "sampler": np.zeros((n, 2), np.uint64), "lambda": np.zeros((n, 4), np.float32),
               "camera_sample": np.zeros((n, 5), np.float32)}
        stats = KrrStats()
        caps = [np.zeros((n, 4), np.int32) for _ in range(6)]
        cap_ptrs = (P * 6)(*[c.ctypes.data for c in caps])
        cap_counts = (I32 * 6)()
        cs, cd = capture if capture else (-1, -1)
        secs = self.lib.orc_render(self.scene, C.byref(p), C.byref(cam), w, h, frame_index,
                                   out["film"].ctypes.data, out["first_hits"].ctypes.data, out["sampler"].ctypes.data,
                                   out["lambda"].ctypes.data, out["camera_sample"].ctypes.data, C.byref(stats), cs, cd,
                                   cap_ptrs, cap_counts)
        out["seconds"] = secs
        out["stats"] = stats.as_dict()
        if capture:
            out["queues"] = [caps[q][: cap_counts[q]].copy() for q in range(6)]
        return out

    def render_megakernel(self, cam, w, h, frame_id=1, spp=1, max_depth=10, rr=0.8, nee=True, use_bvh=True, threads=0):
        """The reference's MegakernelPathTracer estimator (src/render/megakernel/device.cu:148-195) restated on the CPU;
        the film holds the SUM over the samples, as the reference writes it."""
        p = OrcParams(nee=int(nee), enable_medium=0, max_depth=max_depth, enable_clamp=0, spp=spp, rr=rr, clamp_max=1e3, use_bvh=int(use_bvh),
                      threads=threads, row_begin=0, row_end=0)
        film = np.zeros((h, w, 4), np.float32)
        secs = self.lib.orc_render_megakernel(self.scene, C.byref(p), C.byref(cam), w, h, frame_id, film.ctypes.data)
        return {"film": film, "seconds": secs}
