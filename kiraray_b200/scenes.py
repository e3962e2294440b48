"""Synthetic scene descriptors (KrrSceneDesc) for the parity tests and the workload runs of
BASELINE.json configs 3 and 5.  Input generation only: everything here produces
the flat host arrays `krr_wfpt_set_scene` (and the CPU oracle) consume; no rendering logic.

All randomness comes from numpy PCG64 seeded with KRR_DEFAULT_RND_SEED = 7272
(reference src/core/config.in.h:19).
"""
import ctypes as C

import numpy as np

from .binding import (F, I32, KrrInstanceDesc, KrrLightDesc, KrrMaterialDesc, KrrMediumDesc, KrrMeshDesc, KrrSceneDesc,
                      KrrSRT, KrrTransformNodeDesc)

SEED = 7272
IDENTITY = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(F))


class SceneBuilder:
    """Accumulates meshes / instances / materials / lights and builds a KrrSceneDesc whose pointers
    stay valid as long as this object lives."""

    def __init__(self):
        self.meshes, self.instances, self.materials, self.lights, self.media, self.nodes = [], [], [], [], [], []
        self.options = dict(animated=0, multilevel=0, motionblur=0, starttime=0.0, endtime=1.0)
        self._keep = []

    def add_material(self, diffuse=(0.7, 0.7, 0.7), specular=(0.0, 0.0, 0.0), roughness=1.0, bsdf_type=4, specular_transmission=0.0,
                     ior=1.5, emissive=None, anisotropic=0.0, shading_model=1, specular4=None, images=None):
        """Disney by default, SpecularGlossiness shading model like an OBJ import (specular.a = 1 - roughness).
        shading_model 0 = MetallicRoughness (specular = (occlusion, roughness, metallic, -)); specular4 overrides the
        four specular channels; images: {texture slot (KRR_TEX_*): (H, W, 4) float32 array} -- RGBA32F image textures
        (diffuse 0, specular 1, emissive 2, normal 3, transmission / alpha 4)."""
        m = KrrMaterialDesc()
        m.diffuse = (F * 4)(*diffuse, 1.0)
        m.specular = (F * 4)(*(specular4 if specular4 is not None else (*specular, 1.0 - roughness)))
        m.specular_transmission, m.anisotropic, m.ior = specular_transmission, anisotropic, ior
        m.bsdf_type, m.shading_model, m.color_space = bsdf_type, shading_model, 0
        if emissive is not None:
            m.textures[2].valid = 1
            m.textures[2].value = (F * 4)(*emissive, 1.0)
        for slot, img in (images or {}).items():
            a = np.ascontiguousarray(img, np.float32)
            assert a.ndim == 3 and a.shape[2] == 4
            self._keep.append(a)
            t = m.textures[slot]
            t.valid, t.image, t.width, t.height = 1, _fp(a), a.shape[1], a.shape[0]
            t.value = (F * 4)(*[float(x) for x in a.reshape(-1, 4).mean(0)])
        self.materials.append(m)
        return len(self.materials) - 1

    def add_mesh(self, positions, indices, normals=None, material=0, medium_inside=-1, medium_outside=-1, texcoords=None, tangents=None):
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(indices, np.int32).reshape(-1, 3)
        n = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        uv = None if texcoords is None else np.ascontiguousarray(texcoords, np.float32).reshape(-1, 2)
        tg = None if tangents is None else np.ascontiguousarray(tangents, np.float32).reshape(-1, 3)
        self._keep += [p, i, n, uv, tg]
        m = KrrMeshDesc()
        m.positions, m.indices = _fp(p), i.ctypes.data_as(C.POINTER(I32))
        if n is not None:
            m.normals = _fp(n)
        if uv is not None:
            m.texcoords = _fp(uv)
        if tg is not None:
            m.tangents = _fp(tg)
        m.n_vertices, m.n_triangles, m.material = len(p), len(i), material
        m.medium_inside, m.medium_outside = medium_inside, medium_outside
        self.meshes.append(m)
        return len(self.meshes) - 1

    def add_transform_node(self, parent=-1, transform=IDENTITY, motion_keys=None, time_begin=0.0, time_end=1.0):
        """One scene-graph node above mesh instances: a static local transform, or (motion_keys given, (K, 10)
        SRT keys regularly spaced over [time_begin, time_end]) an SRT motion transform replacing it."""
        n = KrrTransformNodeDesc()
        n.parent = parent
        n.transform = (F * 12)(*np.asarray(transform, np.float32).ravel())
        n.time_begin, n.time_end = time_begin, time_end
        if motion_keys is not None:
            k = np.ascontiguousarray(motion_keys, np.float32).reshape(-1, 10)
            self._keep.append(k)
            n.n_motion_keys = len(k)
            n.motion_keys = k.ctypes.data_as(C.POINTER(KrrSRT))
        self.nodes.append(n)
        return len(self.nodes) - 1

    def add_instance(self, mesh, transform=IDENTITY, motion_keys=None, transform_node=-1):
        """motion_keys: (K, 10) array of SRT keys (scale xyz, quaternion xyzw, translation xyz); transform_node:
        the node (add_transform_node) holding this instance -- `transform` must then be the chain's product at
        the current animation time."""
        inst = KrrInstanceDesc()
        inst.mesh = mesh
        inst.transform_node = transform_node
        inst.transform = (F * 12)(*np.asarray(transform, np.float32).ravel())
        if motion_keys is not None:
            k = np.ascontiguousarray(motion_keys, np.float32).reshape(-1, 10)
            self._keep.append(k)
            inst.n_motion_keys = len(k)
            inst.motion_keys = k.ctypes.data_as(C.POINTER(KrrSRT))
        self.instances.append(inst)
        return len(self.instances) - 1

    def add_light(self, type_, color=(1, 1, 1), scale=1.0, transform=IDENTITY, scene_radius=10.0, inner=30.0, outer=45.0):
        l = KrrLightDesc()
        l.type, l.scale, l.scene_radius, l.inner_cone_deg, l.outer_cone_deg = type_, scale, scene_radius, inner, outer
        l.color = (F * 3)(*color)
        l.transform = (F * 12)(*np.asarray(transform, np.float32).ravel())
        self.lights.append(l)
        return len(self.lights) - 1

    def add_medium(self, type_=0, sigma_t=(1, 1, 1), albedo=(0.8, 0.8, 0.8), Le=(0, 0, 0), g=0.0, transform=IDENTITY,
                   bounds=((0, 0, 0), (1, 1, 1)), density=None, scale=1.0, albedo_grid=None):
        m = KrrMediumDesc()
        m.type, m.g, m.scale = type_, g, scale
        m.sigma_t, m.albedo, m.Le = (F * 3)(*sigma_t), (F * 3)(*albedo), (F * 3)(*Le)
        m.transform = (F * 12)(*np.asarray(transform, np.float32).ravel())
        m.bounds_min, m.bounds_max = (F * 3)(*bounds[0]), (F * 3)(*bounds[1])
        if density is not None:
            d = np.ascontiguousarray(density, np.float32)  # indexed [z][y][x]
            self._keep.append(d)
            m.res = (I32 * 3)(d.shape[2], d.shape[1], d.shape[0])
            m.density = _fp(d)
            if albedo_grid is not None:
                a = np.ascontiguousarray(albedo_grid, np.float32)  # indexed [z][y][x][rgb], same lattice as the density
                assert a.shape == d.shape + (3,)
                self._keep.append(a)
                m.albedo_grid = _fp(a)
        self.media.append(m)
        return len(self.media) - 1

    def build(self):
        d = KrrSceneDesc()

        def arr(items, typ):
            a = (typ * max(len(items), 1))(*items)
            self._keep.append(a)
            return a

        d.meshes, d.n_meshes = arr(self.meshes, KrrMeshDesc), len(self.meshes)
        d.instances, d.n_instances = arr(self.instances, KrrInstanceDesc), len(self.instances)
        d.materials, d.n_materials = arr(self.materials, KrrMaterialDesc), len(self.materials)
        d.lights, d.n_lights = arr(self.lights, KrrLightDesc), len(self.lights)
        d.media, d.n_media = arr(self.media, KrrMediumDesc), len(self.media)
        d.transform_nodes, d.n_transform_nodes = arr(self.nodes, KrrTransformNodeDesc), len(self.nodes)
        for k, v in self.options.items():
            setattr(d.options, k, v)
        self.desc = d
        return C.pointer(d)

    def triangle_count(self):
        return sum(self.meshes[i.mesh].n_triangles for i in self.instances)


# ---------------------------------------------------------------------------------------------------
def translation(t, s=1.0):
    m = IDENTITY.copy()
    m[[0, 5, 10]] = s
    m[[3, 7, 11]] = t
    return m


def quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def srt_to_mat(k):
    """T * R * S of one SRT key (s[3], q[4] = xyzw, t[3]) as 12 floats -- the pass interpolates keys
    the same way (scale and translation linearly, quaternion nlerp)."""
    k = np.asarray(k, np.float64)
    m = quat_to_mat(k[3:7] / np.linalg.norm(k[3:7])) * k[0:3][None, :]
    return np.concatenate([m, k[7:10][:, None]], 1).astype(np.float32).ravel()


def displaced_sphere(nu, nv, rng, amplitude=0.15):
    """Closed lat-long sphere of 2*nu*(nv-1) triangles with smooth radial noise (a few random
    low-frequency harmonics); returns positions, normals (from the displaced surface), indices."""
    u = np.arange(nu) / nu * 2 * np.pi
    v = (np.arange(1, nv)) / nv * np.pi
    uu, vv = np.meshgrid(u, v, indexing="xy")  # (nv-1, nu)
    d = np.stack([np.sin(vv) * np.cos(uu), np.cos(vv), np.sin(vv) * np.sin(uu)], -1).reshape(-1, 3)
    d = np.concatenate([d, [[0, 1, 0], [0, -1, 0]]])
    r = np.ones(len(d))
    for _ in range(6):
        f = rng.normal(size=3) * 3
        r += amplitude / 6 * np.sin(d @ f + rng.uniform(0, 6.28)) * rng.uniform(0.5, 1.5)
    p = d * r[:, None]
    top, bot = len(d) - 2, len(d) - 1
    idx = []
    rows = nv - 1
    i0 = (np.arange(rows - 1)[:, None] * nu + np.arange(nu)[None, :])
    i1 = (np.arange(rows - 1)[:, None] * nu + (np.arange(nu)[None, :] + 1) % nu)
    idx.append(np.stack([i0, i1, i0 + nu], -1).reshape(-1, 3))
    idx.append(np.stack([i1, i1 + nu, i0 + nu], -1).reshape(-1, 3))
    a = np.arange(nu)
    idx.append(np.stack([np.full(nu, top), (a + 1) % nu, a], -1))
    last = (rows - 1) * nu
    idx.append(np.stack([np.full(nu, bot), last + a, last + (a + 1) % nu], -1))
    idx = np.concatenate(idx).astype(np.int32)
    # vertex normals: area-weighted face normals
    fn = np.cross(p[idx[:, 1]] - p[idx[:, 0]], p[idx[:, 2]] - p[idx[:, 0]])
    n = np.zeros_like(p)
    for c in range(3):
        np.add.at(n, idx[:, c], fn)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
    flip = np.einsum("ij,ij->i", n, d) < 0
    n[flip] *= -1
    return p.astype(np.float32), n.astype(np.float32), idx


def quad(p0, e1, e2):
    p0, e1, e2 = (np.asarray(a, np.float32) for a in (p0, e1, e2))
    p = np.stack([p0, p0 + e1, p0 + e1 + e2, p0 + e2])
    n = np.cross(e1, e2)
    n = np.tile(n / np.linalg.norm(n), (4, 1))
    return p, n.astype(np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.int32)


def tessellated_scene(n_objects=200, tris_per_object=100_000, n_emissive=1000, seed=SEED):
    """BASELINE.json config 3: a unit-cube arrangement of tessellated displaced spheres with Disney
    materials (roughness ~ U(0.05,1), metallic p=.2, specular transmission p=.1, base colour U(.1,.9)^3),
    a floor, and `n_emissive` emissive triangles (Le = 17,12,4) on a ceiling grid."""
    rng = np.random.Generator(np.random.PCG64(seed))
    b = SceneBuilder()
    nv = max(4, int(round(np.sqrt(tris_per_object / 4))))
    nu = max(3, int(round(tris_per_object / (2 * (nv - 1)))))
    grid = int(np.ceil(n_objects ** (1 / 3)))
    cell = 2.0 / grid
    for o in range(n_objects):
        metallic, trans = rng.random() < 0.2, rng.random() < 0.1
        base = rng.uniform(0.1, 0.9, 3)
        # SpecularGlossiness: a metal is a coloured specular with black diffuse (getMetallic, shading.h:17-30)
        mat = b.add_material(diffuse=(0.02, 0.02, 0.02) if metallic else base, specular=base if metallic else (0.04, 0.04, 0.04),
                             roughness=rng.uniform(0.05, 1.0), specular_transmission=1.0 if trans else 0.0)
        p, n, idx = displaced_sphere(nu, nv, rng)
        mesh = b.add_mesh(p, idx, n, mat)
        c = np.array([o % grid, (o // grid) % grid, o // (grid * grid)]) * cell - 1 + cell / 2 + rng.uniform(-0.1, 0.1, 3) * cell
        b.add_instance(mesh, translation(c, 0.42 * cell))
    floor = b.add_material(diffuse=(0.6, 0.6, 0.6), roughness=0.9)
    p, n, idx = quad((-3, -1.05, -3), (0, 0, 6), (6, 0, 0))
    b.add_instance(b.add_mesh(p, idx, n, floor))
    light = b.add_material(diffuse=(0, 0, 0), emissive=(17, 12, 4))
    k = int(np.ceil(np.sqrt(n_emissive / 2)))
    ps, ns, ids = [], [], []
    for q in range((n_emissive + 1) // 2):
        x, z = (q % k) / k * 4 - 2, (q // k) / k * 4 - 2
        p, n, idx = quad((x, 1.6, z), (0.5 * 4 / k, 0, 0), (0, 0, 0.5 * 4 / k))  # normal points down (-y)
        ids.append(idx + 4 * q), ps.append(p), ns.append(n)
    b.add_instance(b.add_mesh(np.concatenate(ps), np.concatenate(ids)[:n_emissive], np.concatenate(ns), light))
    return b


def srt_lerp(keys, time, t0=0.0, t1=1.0):
    """OptiX SRT key interpolation (float64 restatement of kiraray_b200/csrc/motion.cuh srtNodeXf): time clamped,
    components linear, quaternion normalised afterwards."""
    keys = np.asarray(keys, np.float64)
    n = len(keys)
    u = min(max((time - t0) / (t1 - t0), 0.0), 1.0) * (n - 1)
    k = min(int(u), n - 2)
    v = keys[k] + (u - k) * (keys[k + 1] - keys[k])
    v[3:7] /= np.linalg.norm(v[3:7])
    return v


def mat_mul(a, b):
    """product of two 3x4 row-major affine transforms given as 12 floats"""
    A, B = np.vstack([np.reshape(a, (3, 4)), [0, 0, 0, 1]]), np.vstack([np.reshape(b, (3, 4)), [0, 0, 0, 1]])
    return (A @ B)[:3].astype(np.float32).ravel()


def instanced_scene(n_blas=16, tris_per_blas=20_000, n_groups=100, per_group=100, motion=True, seed=SEED, n_keys=2, time=0.0, spin_scale=1.0, drift_scale=1.0):
    """BASELINE.json config 5: `n_groups * per_group` instances of `n_blas` BLASes arranged as a two-level
    graph: every GROUP node and every INSTANCE node under it carries its own `n_keys`-key SRT animation over
    [0, 1] (the reference wraps each animated node in an SRT motion transform, optix.cpp:400-563), plus a
    static floor and an emissive ceiling quad.  With motion=True the chains are passed as transform nodes
    and motion blur is on; the instances' `transform` is always the chain evaluated at `time` (what the
    scene-graph update of the reference would have produced).  Returns (builder, info) with
    info["group_keys"][g] / info["inst_keys"][i] the (K, 10) SRT keys and info["world"](i, t) the 12-float
    object->world transform of instance i at time t."""
    rng = np.random.Generator(np.random.PCG64(seed))
    b = SceneBuilder()
    nv = max(4, int(round(np.sqrt(tris_per_blas / 4))))
    nu = max(3, int(round(tris_per_blas / (2 * (nv - 1)))))
    meshes = []
    for m in range(n_blas):
        mat = b.add_material(diffuse=rng.uniform(0.1, 0.9, 3), roughness=rng.uniform(0.2, 1.0))
        p, n, idx = displaced_sphere(nu, nv, rng, amplitude=0.3)
        meshes.append(b.add_mesh(p, idx, n, mat))
    g = int(np.ceil(np.sqrt(n_groups)))
    s = int(np.ceil(per_group ** (1 / 3)))
    cell = 8 / g

    def animated_keys(scale, center, spin, drift):
        q0 = rng.normal(size=4)
        q0 /= np.linalg.norm(q0)
        dq, vel = rng.normal(size=4) * spin * spin_scale, rng.normal(size=3) * drift * drift_scale
        ks = []
        for k in range(n_keys):
            a = k / max(n_keys - 1, 1)
            q = q0 + a * dq
            ks.append(np.concatenate([[scale] * 3, q / np.linalg.norm(q), center + a * vel]))
        return np.array(ks, np.float32)

    group_keys, inst_keys, inst_group = [], [], []
    for gi in range(n_groups):
        gc = np.array([(gi % g) / g * 8 - 4 + cell / 2, 0.0, (gi // g) / g * 8 - 4 + cell / 2])
        gk = animated_keys(1.0, gc, 0.08, 0.15)
        group_keys.append(gk)
        gnode = b.add_transform_node(-1, srt_to_mat(srt_lerp(gk, time)), gk if motion else None) if motion else -1
        for ii in range(per_group):
            lc = (np.array([ii % s, (ii // s) % s, ii // (s * s)]) / s - 0.5 + 0.5 / s) * cell * 0.9
            ik = animated_keys(cell / s * 0.35 * rng.uniform(0.7, 1.0), lc, 0.15, 0.05)
            inst_keys.append(ik)
            inst_group.append(gi)
            world = mat_mul(srt_to_mat(srt_lerp(gk, time)), srt_to_mat(srt_lerp(ik, time)))
            node = b.add_transform_node(gnode, srt_to_mat(srt_lerp(ik, time)), ik) if motion else -1
            b.add_instance(meshes[(gi * per_group + ii) % n_blas], world, transform_node=node)
    floor = b.add_material(diffuse=(0.6, 0.6, 0.6), roughness=0.9)
    p, n, idx = quad((-6, -0.9, -6), (0, 0, 12), (12, 0, 0))
    b.add_instance(b.add_mesh(p, idx, n, floor))
    light = b.add_material(diffuse=(0, 0, 0), emissive=(17, 12, 4))
    p, n, idx = quad((-3, 4.0, -3), (6, 0, 0), (0, 0, 6))
    b.add_instance(b.add_mesh(p, idx, n, light))
    if motion:
        b.options.update(motionblur=1, multilevel=1, starttime=0.0, endtime=1.0)

    def world(i, t):
        return mat_mul(srt_to_mat(srt_lerp(group_keys[inst_group[i]], t)), srt_to_mat(srt_lerp(inst_keys[i], t)))

    return b, dict(group_keys=group_keys, inst_keys=inst_keys, inst_group=inst_group, world=world, n_moving=len(inst_keys))


def look_at_camera(eye, target, aspect, focal_length=21.0, up=(0, 1, 0), shutter_open=0.0, shutter_time=0.0, lens_radius=0.0, focal_distance=10.0):
    """rt::CameraData (reference src/core/camera.h:20-30): 24 mm film height, camera looks down -z."""
    from .binding import KrrCameraData
    eye, target, up = (np.asarray(a, np.float64) for a in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, up)
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    m = np.concatenate([np.stack([r, u, -f], 1), eye[:, None]], 1).astype(np.float32).ravel()
    cam = KrrCameraData()
    cam.film_size = (F * 2)(24.0 * aspect, 24.0)
    cam.focal_length, cam.focal_distance, cam.lens_radius, cam.aspect_ratio = focal_length, focal_distance, lens_radius, aspect
    cam.shutter_open, cam.shutter_time, cam.medium = shutter_open, shutter_time, -1
    cam.transform = (F * 12)(*m)
    return cam
