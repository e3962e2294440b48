"""ctypes mirror of include/krr_wfpt.h and include/krr_host_c.h (no logic, no fallback)."""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
KRR_MAX_DEPTH_STATS = 64
KRR_TEX_COUNT = 5


class NativeLibraryMissing(RuntimeError):
    pass


def lib_dir():
    return os.path.join(_HERE, "lib")


def data_dir():
    return os.path.join(_HERE, "data")


F, I32, U64, P = C.c_float, C.c_int32, C.c_uint64, C.c_void_p


class KrrTextureDesc(C.Structure):
    _fields_ = [("valid", I32), ("value", F * 4), ("image", C.POINTER(F)), ("width", I32), ("height", I32)]


class KrrSpectrumDesc(C.Structure):
    _fields_ = [("kind", I32), ("a", F * 3), ("b", F * 3), ("lambdas", C.POINTER(F)), ("values", C.POINTER(F)), ("n", I32)]


class KrrMaterialDesc(C.Structure):
    _fields_ = [("diffuse", F * 4), ("specular", F * 4), ("specular_transmission", F), ("anisotropic", F), ("ior", F),
                ("spectral_eta", KrrSpectrumDesc), ("spectral_k", KrrSpectrumDesc), ("textures", KrrTextureDesc * KRR_TEX_COUNT),
                ("bsdf_type", I32), ("shading_model", I32), ("color_space", I32)]


class KrrMeshDesc(C.Structure):
    _fields_ = [("positions", C.POINTER(F)), ("normals", C.POINTER(F)), ("texcoords", C.POINTER(F)), ("tangents", C.POINTER(F)),
                ("indices", C.POINTER(I32)), ("n_vertices", I32), ("n_triangles", I32), ("material", I32),
                ("medium_inside", I32), ("medium_outside", I32), ("Le", F * 3)]


class KrrSRT(C.Structure):
    _fields_ = [("s", F * 3), ("q", F * 4), ("t", F * 3)]


class KrrTransformNodeDesc(C.Structure):
    _fields_ = [("parent", I32), ("transform", F * 12), ("n_motion_keys", I32), ("motion_keys", C.POINTER(KrrSRT)),
                ("time_begin", F), ("time_end", F)]


class KrrInstanceDesc(C.Structure):
    _fields_ = [("mesh", I32), ("transform", F * 12), ("n_motion_keys", I32), ("motion_keys", C.POINTER(KrrSRT)),
                ("transform_node", I32)]


class KrrLightDesc(C.Structure):
    _fields_ = [("type", I32), ("color", F * 3), ("scale", F), ("transform", F * 12), ("inner_cone_deg", F),
                ("outer_cone_deg", F), ("scene_radius", F), ("texture", KrrTextureDesc)]


class KrrMediumDesc(C.Structure):
    _fields_ = [("type", I32), ("sigma_t", F * 3), ("albedo", F * 3), ("Le", F * 3), ("g", F), ("transform", F * 12),
                ("bounds_min", F * 3), ("bounds_max", F * 3), ("res", I32 * 3), ("density", C.POINTER(F)), ("scale", F),
                ("albedo_grid", C.POINTER(F))]


class KrrSceneOptions(C.Structure):
    _fields_ = [("animated", I32), ("multilevel", I32), ("motionblur", I32), ("starttime", F), ("endtime", F)]


class KrrSceneDesc(C.Structure):
    _fields_ = [("meshes", C.POINTER(KrrMeshDesc)), ("n_meshes", I32), ("instances", C.POINTER(KrrInstanceDesc)), ("n_instances", I32),
                ("materials", C.POINTER(KrrMaterialDesc)), ("n_materials", I32), ("lights", C.POINTER(KrrLightDesc)), ("n_lights", I32),
                ("media", C.POINTER(KrrMediumDesc)), ("n_media", I32), ("options", KrrSceneOptions),
                ("transform_nodes", C.POINTER(KrrTransformNodeDesc)), ("n_transform_nodes", I32)]


class KrrCameraData(C.Structure):
    _fields_ = [("film_size", F * 2), ("focal_length", F), ("focal_distance", F), ("lens_radius", F), ("aspect_ratio", F),
                ("shutter_open", F), ("shutter_time", F), ("transform", F * 12), ("medium", I32)]


class KrrColorSpaceData(C.Structure):
    _fields_ = [("cie_x", C.POINTER(F)), ("cie_y", C.POINTER(F)), ("cie_z", C.POINTER(F)), ("illuminant", C.POINTER(F)),
                ("xyz_from_rgb", F * 9), ("rgb_from_xyz", F * 9), ("z_nodes", C.POINTER(F)), ("coeffs", C.POINTER(F))]


class KrrStats(C.Structure):
    _fields_ = [("camera_rays", U64), ("closest_rays", U64), ("shadow_rays", U64), ("scatter_items", U64), ("hit_light_items", U64),
                ("miss_items", U64), ("medium_sample_items", U64), ("medium_scatter_items", U64),
                ("closest_by_depth", U64 * KRR_MAX_DEPTH_STATS), ("shadow_by_depth", U64 * KRR_MAX_DEPTH_STATS),
                ("kernel_launches", U64), ("bvh_nodes", U64), ("bvh_triangles", U64), ("tlas_nodes", U64)]

    def as_dict(self):
        d = {k: int(getattr(self, k)) for k, _ in self._fields_ if not k.endswith("_by_depth")}
        d["closest_by_depth"] = [int(v) for v in self.closest_by_depth]
        d["shadow_by_depth"] = [int(v) for v in self.shadow_by_depth]
        return d


class KrrLeafBsdfQuery(C.Structure):
    _fields_ = [("ior", F), ("diffuse", F * 4), ("specular", F * 4), ("specular_transmission", F), ("roughness", F), ("metallic", F),
                ("anisotropic", F), ("bsdf_type", I32), ("wo", F * 3), ("wi", F * 3), ("wavelength_u", F),
                ("seed_px", C.c_uint32), ("seed_py", C.c_uint32), ("seed_index", C.c_uint32), ("eta_kind", I32), ("eta", F)]


class KrrLeafBsdfResult(C.Structure):
    _fields_ = [("type_flags", I32), ("f", F * 4), ("pdf", F), ("s_f", F * 4), ("s_wi", F * 3), ("s_pdf", F), ("s_flags", I32)]


class KrrLeafLightQuery(C.Structure):
    _fields_ = [("type", I32), ("p", (F * 3) * 3), ("n", (F * 3) * 3), ("transform", F * 12), ("color", F * 3), ("scale", F),
                ("two_sided", I32), ("scene_radius", F), ("cos_inner", F), ("cos_outer", F), ("u", F * 2), ("ctx_p", F * 3),
                ("ctx_n", F * 3), ("wi", F * 3), ("wavelength_u", F)]


class KrrLeafLightResult(C.Structure):
    _fields_ = [("p", F * 3), ("n", F * 3), ("L", F * 4), ("pdf", F), ("L_eval", F * 4), ("pdf_li", F)]


_wfpt = None
_host = None


def _load(name):
    path = os.path.join(lib_dir(), name)
    if not os.path.exists(path):
        raise NativeLibraryMissing(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(the product has no CPU fallback)")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def load_wfpt():
    """libkrr_wfpt.so with argtypes set for every symbol include/krr_wfpt.h declares."""
    global _wfpt
    if _wfpt is not None:
        return _wfpt
    override = os.environ.get("KRR_WFPT_LIB")  # tuning aid: a variant built by build.py build_variant()
    lib = C.CDLL(override, mode=C.RTLD_GLOBAL) if override else _load("libkrr_wfpt.so")
    lib.krr_wfpt_last_error.restype = C.c_char_p
    sig = {
        "krr_wfpt_create": [C.c_char_p, C.POINTER(P)],
        "krr_wfpt_destroy": [P],
        "krr_wfpt_set_params": [P, C.c_char_p],
        "krr_wfpt_set_color_space": [P, C.POINTER(KrrColorSpaceData)],
        "krr_wfpt_set_scene": [P, C.POINTER(KrrSceneDesc)],
        "krr_wfpt_resize": [P, I32, I32],
        "krr_wfpt_update_instances": [P, C.POINTER(I32), C.POINTER(F), I32, P],
        "krr_wfpt_begin_frame": [P, U64, C.POINTER(KrrCameraData), P],
        "krr_wfpt_render": [P, P, P],
        "krr_wfpt_render_to_host": [P, P, P],
        "krr_wfpt_render_to_host_async": [P, P, P],
        "krr_wfpt_wait_host": [P],
        "krr_wfpt_comm_unique_id": [P],
        "krr_wfpt_comm_init_rank": [P, P, I32, I32],
        "krr_wfpt_comm_init_all": [C.POINTER(P), I32],
        "krr_wfpt_comm_destroy": [P],
        "krr_wfpt_reduce_film": [P, P, I32, F, P],
        "krr_wfpt_render_reduce_to_host_async": [P, P, I32, F, P],
        "krr_wfpt_render_megakernel": [P, U64, C.POINTER(KrrCameraData), P, P],
        "krr_wfpt_set_partition": [P, I32, I32],
        "krr_wfpt_get_stats": [P, C.POINTER(KrrStats)],
        "krr_wfpt_set_profiling": [P, I32],
        "krr_wfpt_get_stage_times": [P, C.POINTER(C.c_double), C.POINTER(I32), I32],
        "krr_wfpt_get_launch_times": [P, C.POINTER(I32), C.POINTER(F), I32],
        "krr_wfpt_debug_first_hits": [P, P, P],
        "krr_wfpt_debug_pixel_state": [P, P, P, P],
        "krr_wfpt_debug_capture": [P, I32, I32],
        "krr_wfpt_debug_queue": [P, I32, P, I32],
        "krr_accumulate_f32": [P, P, C.c_int64, U64, U64, I32, P],
        "krr_accumulate_f64": [P, P, C.c_int64, U64, U64, I32, P],
        "krr_accumulate_read_average": [P, I32, U64, C.c_int64, P, P],
        "krr_error_metric_f32": [P, P, C.c_int64, I32, C.POINTER(C.c_double), P],
        "krr_tonemap_f32": [P, C.c_int64, I32, F, I32, P],
        "krr_wfpt_debug_eval_bsdf": [P, C.POINTER(KrrLeafBsdfQuery), I32, C.POINTER(KrrLeafBsdfResult)],
        "krr_wfpt_debug_eval_light": [P, C.POINTER(KrrLeafLightQuery), I32, C.POINTER(KrrLeafLightResult)],
        "krr_wfpt_debug_eval_color": [P, C.POINTER(F), I32, C.POINTER(F)],
        "krr_wfpt_debug_camera_rays": [P, C.POINTER(KrrCameraData), I32, I32, C.POINTER(F), I32, C.POINTER(F)],
        "krr_wfpt_debug_instance_xf": [P, C.POINTER(I32), C.POINTER(F), I32, C.POINTER(F)],
        "krr_wfpt_abi_version": [],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = None if name == "krr_wfpt_destroy" else C.c_int
    _wfpt = lib
    return lib


def load_host():
    global _host
    if _host is not None:
        return _host
    load_wfpt()
    lib = _load("libkrr_host.so")
    lib.krr_host_last_error.restype = C.c_char_p
    lib.krr_host_app_create.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(P)]
    lib.krr_host_app_destroy.argtypes = [P]
    lib.krr_host_app_destroy.restype = None
    lib.krr_host_app_get_resolution.argtypes = [P, C.POINTER(I32), C.POINTER(I32)]
    lib.krr_host_app_set_resolution.argtypes = [P, I32, I32]
    lib.krr_host_app_scene_desc.argtypes = [P]
    lib.krr_host_app_scene_desc.restype = C.POINTER(KrrSceneDesc)
    lib.krr_host_app_get_camera.argtypes = [P, C.c_double, C.POINTER(KrrCameraData)]
    lib.krr_host_app_get_wfpt_params.argtypes = [P, C.c_char_p, I32]
    lib.krr_host_app_set_wfpt_params.argtypes = [P, C.c_char_p]
    lib.krr_host_app_render_frames.argtypes = [P, I32, P]
    lib.krr_host_app_wfpt_handle.argtypes = [P]
    lib.krr_host_app_wfpt_handle.restype = P
    lib.krr_host_app_frame_index.argtypes = [P]
    lib.krr_host_app_frame_index.restype = U64
    lib.krr_host_app_run.argtypes = [P, I32, I32]
    lib.krr_host_app_set_output_dir.argtypes = [P, C.c_char_p]
    lib.krr_host_app_get_pass_json.argtypes = [P, C.c_char_p, C.c_char_p, I32]
    lib.krr_host_app_accum_count.argtypes = [P]
    lib.krr_host_app_accum_count.restype = C.c_int64
    lib.krr_host_app_read_accumulated.argtypes = [P, P]
    lib.krr_host_app_set_reference.argtypes = [P, P, I32, I32]
    lib.krr_host_app_evaluate_next_frame.argtypes = [P]
    lib.krr_host_app_last_error_metric.argtypes = [P, C.POINTER(C.c_double), C.POINTER(I32)]
    lib.krr_host_image_load.argtypes = [C.c_char_p, I32, C.POINTER(I32), C.POINTER(I32), P]
    lib.krr_host_image_save.argtypes = [C.c_char_p, P, I32, I32, I32, I32]
    lib.krr_host_image_save_exr.argtypes = [C.c_char_p, P, I32, I32, I32, I32]
    lib.krr_host_set_data_dir.argtypes = [C.c_char_p]
    lib.krr_multi_create.argtypes = [C.POINTER(KrrSceneDesc), C.c_char_p, I32, I32, C.POINTER(I32), I32, I32, C.POINTER(P)]
    lib.krr_multi_destroy.argtypes = [P]
    lib.krr_multi_destroy.restype = None
    lib.krr_multi_uses_nccl.argtypes = [P]
    lib.krr_multi_render.argtypes = [P, C.POINTER(KrrCameraData), U64, I32, P, C.POINTER(C.c_double), C.POINTER(U64)]
    lib.krr_multi_last_error.restype = C.c_char_p
    lib.krr_host_set_data_dir(data_dir().encode())
    _host = lib
    return lib


_cs_cache = None


def color_space():
    """sRGB colour-space tables (kiraray_b200/data/spectral_srgb.bin) as a KrrColorSpaceData."""
    global _cs_cache
    if _cs_cache is not None:
        return _cs_cache[0]
    path = os.path.join(data_dir(), "spectral_srgb.bin")
    if not os.path.exists(path):
        raise NativeLibraryMissing(f"{path} is missing (python oracle/build_oracle.py ref spectral)")
    raw = np.fromfile(path, dtype=np.uint32, count=4)
    assert raw[0] == 0x4B525253 and raw[2] == 471 and raw[3] == 64, "bad spectral_srgb.bin"
    blob = np.fromfile(path, dtype=np.float32, offset=16)
    cs = KrrColorSpaceData()
    fp = lambda a: a.ctypes.data_as(C.POINTER(F))
    parts = {"cie_x": blob[0:471], "cie_y": blob[471:942], "cie_z": blob[942:1413], "illuminant": blob[1413:1884]}
    for k, v in parts.items():
        setattr(cs, k, fp(v))
    cs.xyz_from_rgb = (F * 9)(*blob[1884:1893])
    cs.rgb_from_xyz = (F * 9)(*blob[1893:1902])
    z = blob[1902:1966]
    co = blob[1966:]
    cs.z_nodes, cs.coeffs = fp(z), fp(co)
    _cs_cache = (cs, blob)
    return cs


class HostApp:
    """RenderApp of the C++ host layer (headless).  Loading a config touches no GPU."""

    def __init__(self, config, asset_root=None):
        self.lib = load_host()
        self.h = P()
        if isinstance(config, dict):
            text, is_path = json.dumps(config).encode(), 0
        else:
            text, is_path = str(config).encode(), 1
        root = asset_root.encode() if asset_root else (None if is_path else b".")
        rc = self.lib.krr_host_app_create(text, is_path, root, C.byref(self.h))
        if rc != 0:
            raise RuntimeError("krr_host_app_create: " + self.lib.krr_host_last_error().decode())

    def close(self):
        if self.h:
            self.lib.krr_host_app_destroy(self.h)
            self.h = P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc < 0:
            raise RuntimeError(f"{what}: {self.lib.krr_host_last_error().decode()}")
        return rc

    @property
    def resolution(self):
        w, h = I32(), I32()
        self.lib.krr_host_app_get_resolution(self.h, C.byref(w), C.byref(h))
        return w.value, h.value

    def set_resolution(self, w, h):
        self._ck(self.lib.krr_host_app_set_resolution(self.h, w, h), "set_resolution")

    def scene_desc(self):
        p = self.lib.krr_host_app_scene_desc(self.h)
        if not p:
            raise RuntimeError(self.lib.krr_host_last_error().decode())
        return p

    def camera(self, time=0.0):
        cam = KrrCameraData()
        self._ck(self.lib.krr_host_app_get_camera(self.h, time, C.byref(cam)), "get_camera")
        return cam

    def wfpt_params(self):
        buf = C.create_string_buffer(1024)
        self._ck(self.lib.krr_host_app_get_wfpt_params(self.h, buf, 1024), "get_wfpt_params")
        return json.loads(buf.value.decode())

    def set_wfpt_params(self, **kw):
        self._ck(self.lib.krr_host_app_set_wfpt_params(self.h, json.dumps(kw).encode()), "set_wfpt_params")

    def render_frames(self, n=1):
        w, h = self.resolution
        film = np.empty((h, w, 4), dtype=np.float32)
        self._ck(self.lib.krr_host_app_render_frames(self.h, n, film.ctypes.data_as(P)), "render_frames")
        return film

    def run(self, max_frames=0, finalize=True):
        """RenderApp main loop: frames until a pass requests the exit or max_frames; returns frames rendered."""
        n = self.lib.krr_host_app_run(self.h, int(max_frames), int(finalize))
        if n < 0:
            raise RuntimeError("krr_host_app_run: " + self.lib.krr_host_last_error().decode())
        return n

    def set_output_dir(self, d):
        self._ck(self.lib.krr_host_app_set_output_dir(self.h, str(d).encode()), "set_output_dir")

    def pass_json(self, name):
        buf = C.create_string_buffer(4096)
        n = self.lib.krr_host_app_get_pass_json(self.h, name.encode(), buf, 4096)
        if n < 0:
            raise RuntimeError("krr_host_app_get_pass_json: " + self.lib.krr_host_last_error().decode())
        import json as _json
        return _json.loads(buf.value.decode())

    def accum_count(self):
        return int(self.lib.krr_host_app_accum_count(self.h))

    def read_accumulated(self):
        w, h = self.resolution
        out = np.empty((h, w, 4), np.float32)
        self._ck(self.lib.krr_host_app_read_accumulated(self.h, out.ctypes.data_as(P)), "read_accumulated")
        return out

    def set_reference(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.float32)
        self._ck(self.lib.krr_host_app_set_reference(self.h, rgba.ctypes.data_as(P), rgba.shape[1], rgba.shape[0]), "set_reference")

    def evaluate_next_frame(self):
        self._ck(self.lib.krr_host_app_evaluate_next_frame(self.h), "evaluate_next_frame")

    def last_error_metric(self):
        v, n = C.c_double(), I32()
        self._ck(self.lib.krr_host_app_last_error_metric(self.h, C.byref(v), C.byref(n)), "last_error_metric")
        return v.value, n.value

    def wfpt_handle(self):
        return self.lib.krr_host_app_wfpt_handle(self.h)

    @property
    def frame_index(self):
        return int(self.lib.krr_host_app_frame_index(self.h))


class Wfpt:
    """Thin wrapper over the C ABI handle (include/krr_wfpt.h)."""

    def __init__(self, params=None, handle=None):
        self.lib = load_wfpt()
        self.owned = handle is None
        self.h = P(handle) if handle is not None else P()
        if handle is None:
            # the parity taps (first_hits / pixel_state) need the pass to record them: the test binding asks
            # for that unless told otherwise; the C ABI default (and bench.py) is off
            text = json.dumps(dict({"debug_taps": True}, **(params or {}))).encode()
            self._ck(self.lib.krr_wfpt_create(text, C.byref(self.h)), "create")
            self._ck(self.lib.krr_wfpt_set_color_space(self.h, C.byref(color_space())), "set_color_space")
        self.size = None

    def _ck(self, rc, what):
        if rc < 0:
            raise RuntimeError(f"krr_wfpt_{what}: {self.lib.krr_wfpt_last_error().decode()}")
        return rc

    def close(self):
        if self.owned and self.h:
            self.lib.krr_wfpt_destroy(self.h)
        self.h = P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **kw):
        self._ck(self.lib.krr_wfpt_set_params(self.h, json.dumps(kw).encode()), "set_params")

    def set_scene(self, desc_ptr):
        self._ck(self.lib.krr_wfpt_set_scene(self.h, desc_ptr), "set_scene")

    def resize(self, w, h):
        self._ck(self.lib.krr_wfpt_resize(self.h, w, h), "resize")
        self.size = (w, h)
        self.rows = (0, h)

    def set_partition(self, r0, r1):
        self._ck(self.lib.krr_wfpt_set_partition(self.h, r0, r1), "set_partition")
        self.rows = (r0, r1)

    def begin_frame(self, frame_index, cam, stream=None):
        self._ck(self.lib.krr_wfpt_begin_frame(self.h, frame_index, C.byref(cam), P(stream or 0)), "begin_frame")

    def render(self, film_device_ptr, stream=None):
        self._ck(self.lib.krr_wfpt_render(self.h, P(film_device_ptr), P(stream or 0)), "render")

    def render_megakernel(self, frame_index, cam, film_device_ptr, stream=None):
        """The reference's MegakernelPathTracer on this handle's scene (cross-validation estimator)."""
        self._ck(self.lib.krr_wfpt_render_megakernel(self.h, frame_index, C.byref(cam), P(film_device_ptr), P(stream or 0)), "render_megakernel")

    def render_to_host(self, film=None, stream=None):
        w, h = self.size
        if film is None:
            film = np.empty((h, w, 4), dtype=np.float32)
        self._ck(self.lib.krr_wfpt_render_to_host(self.h, film.ctypes.data_as(P), P(stream or 0)), "render_to_host")
        return film

    def render_to_host_async(self, film, stream=None):
        """Pipelined read-back: `film` (ideally pinned) holds the frame after wait_host()."""
        self._ck(self.lib.krr_wfpt_render_to_host_async(self.h, film.ctypes.data_as(P), P(stream or 0)), "render_to_host_async")

    def wait_host(self):
        self._ck(self.lib.krr_wfpt_wait_host(self.h), "wait_host")

    # ---- multi-GPU film reduction (NCCL inside the product library) ----
    def comm_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._ck(self.lib.krr_wfpt_comm_unique_id(buf), "comm_unique_id")
        return bytes(buf)

    def comm_init_rank(self, uid, world, rank):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self.lib.krr_wfpt_comm_init_rank(self.h, buf, world, rank), "comm_init_rank")

    def comm_destroy(self):
        self._ck(self.lib.krr_wfpt_comm_destroy(self.h), "comm_destroy")

    def reduce_film(self, film_device_ptr, root=0, scale=1.0, stream=None):
        self._ck(self.lib.krr_wfpt_reduce_film(self.h, P(film_device_ptr), root, scale, P(stream or 0)), "reduce_film")

    def render_reduce_to_host_async(self, film, root=0, scale=1.0, stream=None):
        ptr = film.ctypes.data_as(P) if film is not None else P(0)
        self._ck(self.lib.krr_wfpt_render_reduce_to_host_async(self.h, ptr, root, scale, P(stream or 0)), "render_reduce_to_host_async")

    def instance_xf(self, ids, times):
        """(n, 2, 12): object->world and world->object of each instance at each ray time, from the device"""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        t = np.ascontiguousarray(times, dtype=np.float32)
        out = np.empty((len(ids), 2, 12), np.float32)
        self._ck(self.lib.krr_wfpt_debug_instance_xf(self.h, ids.ctypes.data_as(C.POINTER(I32)), t.ctypes.data_as(C.POINTER(F)), len(ids),
                                                     out.ctypes.data_as(C.POINTER(F))), "debug_instance_xf")
        return out

    def update_instances(self, ids, transforms, stream=None):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        xf = np.ascontiguousarray(transforms, dtype=np.float32).reshape(-1, 12)
        self._ck(self.lib.krr_wfpt_update_instances(self.h, ids.ctypes.data_as(C.POINTER(I32)), xf.ctypes.data_as(C.POINTER(F)),
                                                    len(ids), P(stream or 0)), "update_instances")

    def stats(self):
        s = KrrStats()
        self._ck(self.lib.krr_wfpt_get_stats(self.h, C.byref(s)), "get_stats")
        return s.as_dict()

    STAGES = ["camera", "closest", "hit_miss", "scatter", "shadow", "resolve", "medium", "trace", "tail"]

    def set_profiling(self, on):
        self._ck(self.lib.krr_wfpt_set_profiling(self.h, int(on)), "set_profiling")

    def stage_times(self, reset=True):
        ms, n = (C.c_double * len(self.STAGES))(), (I32 * len(self.STAGES))()
        self._ck(self.lib.krr_wfpt_get_stage_times(self.h, ms, n, int(reset)), "get_stage_times")
        return {k: {"ms": ms[i], "launches": n[i]} for i, k in enumerate(self.STAGES)}

    def launch_times(self, capacity=8192):
        st, ms = (I32 * capacity)(), (F * capacity)()
        n = self._ck(self.lib.krr_wfpt_get_launch_times(self.h, st, ms, capacity), "get_launch_times")
        return [(self.STAGES[st[i]], ms[i]) for i in range(min(n, capacity))]

    def _npix(self):
        return (self.rows[1] - self.rows[0]) * self.size[0]

    def first_hits(self):
        n = self._npix()
        inst, prim = np.empty(n, np.int32), np.empty(n, np.int32)
        self._ck(self.lib.krr_wfpt_debug_first_hits(self.h, inst.ctypes.data_as(P), prim.ctypes.data_as(P)), "debug_first_hits")
        return inst, prim

    def pixel_state(self):
        n = self._npix()
        s, l, c = np.empty((n, 2), np.uint64), np.empty((n, 4), np.float32), np.empty((n, 5), np.float32)
        self._ck(self.lib.krr_wfpt_debug_pixel_state(self.h, s.ctypes.data_as(P), l.ctypes.data_as(P), c.ctypes.data_as(P)), "debug_pixel_state")
        return s, l, c

    # ---- leaf-function taps (parity tests) ----
    def eval_bsdf(self, queries):
        n = len(queries)
        q, r = (KrrLeafBsdfQuery * n)(*queries), (KrrLeafBsdfResult * n)()
        self._ck(self.lib.krr_wfpt_debug_eval_bsdf(self.h, q, n, r), "debug_eval_bsdf")
        return list(r)

    def eval_light(self, queries):
        n = len(queries)
        q, r = (KrrLeafLightQuery * n)(*queries), (KrrLeafLightResult * n)()
        self._ck(self.lib.krr_wfpt_debug_eval_light(self.h, q, n, r), "debug_eval_light")
        return list(r)

    def eval_color(self, in8):
        a = np.ascontiguousarray(in8, dtype=np.float32).reshape(-1, 8)
        out = np.empty((len(a), 20), np.float32)
        self._ck(self.lib.krr_wfpt_debug_eval_color(self.h, a.ctypes.data_as(C.POINTER(F)), len(a), out.ctypes.data_as(C.POINTER(F))), "debug_eval_color")
        return out

    def camera_rays(self, cam, w, h, in7):
        a = np.ascontiguousarray(in7, dtype=np.float32).reshape(-1, 7)
        out = np.empty((len(a), 7), np.float32)
        self._ck(self.lib.krr_wfpt_debug_camera_rays(self.h, C.byref(cam), w, h, a.ctypes.data_as(C.POINTER(F)), len(a),
                                                     out.ctypes.data_as(C.POINTER(F))), "debug_camera_rays")
        return out

    def capture(self, sample_id, depth):
        self._ck(self.lib.krr_wfpt_debug_capture(self.h, sample_id, depth), "debug_capture")

    def queue(self, q):
        n = self._npix()
        items = np.empty((n, 4), np.int32)
        cnt = self._ck(self.lib.krr_wfpt_debug_queue(self.h, q, items.ctypes.data_as(P), n), "debug_queue")
        return items[:cnt].copy()


# ---- HDR image files (host layer, no GPU): reference Image::loadImage / saveImage ----
def load_image(path, flip=False):
    lib = load_host()
    w, h = I32(), I32()
    if lib.krr_host_image_load(str(path).encode(), int(flip), C.byref(w), C.byref(h), None) != 0:
        raise RuntimeError("krr_host_image_load: " + lib.krr_host_last_error().decode())
    out = np.empty((h.value, w.value, 4), np.float32)
    if lib.krr_host_image_load(str(path).encode(), int(flip), C.byref(w), C.byref(h), out.ctypes.data_as(P)) != 0:
        raise RuntimeError("krr_host_image_load: " + lib.krr_host_last_error().decode())
    return out


def save_image(path, rgba, flip=False, reference_channel_order=True):
    lib = load_host()
    rgba = np.ascontiguousarray(rgba, np.float32)
    if lib.krr_host_image_save(str(path).encode(), rgba.ctypes.data_as(P), rgba.shape[1], rgba.shape[0], int(flip), int(reference_channel_order)) != 0:
        raise RuntimeError("krr_host_image_save: " + lib.krr_host_last_error().decode())


def save_exr(path, rgba, half=True, zip=False):
    lib = load_host()
    rgba = np.ascontiguousarray(rgba, np.float32)
    if lib.krr_host_image_save_exr(str(path).encode(), rgba.ctypes.data_as(P), rgba.shape[1], rgba.shape[0], int(half), int(zip)) != 0:
        raise RuntimeError("krr_host_image_save_exr: " + lib.krr_host_last_error().decode())


# ---- device kernels of the passes that follow the path tracer (C ABI, device pointers) ----
def error_metric(film_ptr, ref_ptr, n_pixels, metric=3, stream=None):
    lib = load_wfpt()
    v = C.c_double()
    if lib.krr_error_metric_f32(P(film_ptr), P(ref_ptr), n_pixels, metric, C.byref(v), P(stream or 0)) != 0:
        raise RuntimeError("krr_error_metric_f32: " + lib.krr_wfpt_last_error().decode())
    return v.value


def tonemap(film_ptr, n_pixels, op=0, exposure=1.0, gamma=True, stream=None):
    lib = load_wfpt()
    if lib.krr_tonemap_f32(P(film_ptr), n_pixels, op, exposure, int(gamma), P(stream or 0)) != 0:
        raise RuntimeError("krr_tonemap_f32: " + lib.krr_wfpt_last_error().decode())


def accumulate_f64(accum_ptr, film_ptr, n_pixels, accum_count, max_accum=0, moving_average=False, stream=None):
    lib = load_wfpt()
    if lib.krr_accumulate_f64(P(accum_ptr), P(film_ptr), n_pixels, accum_count, max_accum, int(moving_average), P(stream or 0)) != 0:
        raise RuntimeError("krr_accumulate_f64: " + lib.krr_wfpt_last_error().decode())


class MultiDeviceApp:
    """ctypes view of the host layer's MultiDeviceRenderApp (kiraray_b200/host/multi_device.cpp): one process, one
    pass handle + one host thread per device, NCCL film reduce inside the product library."""

    def __init__(self, desc_ptr, params, w, h, devices, tiles=1):
        self.lib = load_host()
        self.w, self.h = w, h
        self.p = P()
        dev = (I32 * len(devices))(*devices)
        rc = self.lib.krr_multi_create(desc_ptr, json.dumps(params).encode(), w, h, dev, len(devices), tiles, C.byref(self.p))
        if rc != 0:
            raise RuntimeError("krr_multi_create: " + self.lib.krr_multi_last_error().decode())

    @property
    def uses_nccl(self):
        return bool(self.lib.krr_multi_uses_nccl(self.p))

    def render(self, cam, first_frame, steps=1):
        """-> (film (H, W, 4) of the last step, wall ms of all steps, rays of the last step)"""
        import numpy as np
        film = np.empty((self.h, self.w, 4), np.float32)
        ms, rays = C.c_double(), U64()
        rc = self.lib.krr_multi_render(self.p, C.byref(cam), first_frame, steps, film.ctypes.data_as(P), C.byref(ms), C.byref(rays))
        if rc != 0:
            raise RuntimeError("krr_multi_render: " + self.lib.krr_multi_last_error().decode())
        return film, ms.value, rays.value

    def close(self):
        if self.p:
            self.lib.krr_multi_destroy(self.p)
            self.p = P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
