"""GPU leaf-function parity: the DEVICE BSDFs / lights / colour / camera routines, called through the
C ABI's debug taps, against the golden vectors the REFERENCE's own host-compiled code produced
(tests/golden/leaf_vectors.npz, tools/make_golden.py) on the same seeded inputs (tests/leaf_cases.py).

Tolerances (floating point): the device code uses CUDA's libm (sinf/cosf/powf/atan2f differ from
glibc in the last ulp) and lets nvcc contract a*b+c into FMA, so values agree to a few ulp, not bit
for bit: |gpu - ref| <= 2e-4 * |ref| + 1e-6 for every value of a case, required of >= 97 % of the
cases of each family, and <= 1e-2 * |ref| + 1e-5 of ALL cases (1e-1 for near-specular GGX lobes, alpha < 0.02,
which amplify a 1-ulp change of the half vector by ~1/alpha^2); discrete outputs (BSDF type flags, sampled lobe flags) must agree wherever the
continuous outputs do.  Integer / RNG-derived values (wavelengths, camera rays) are exact."""
import ctypes as C
import os

import numpy as np
import pytest

import kiraray_b200 as krr
from leaf_cases import LeafCases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F = C.c_float


@pytest.fixture(scope="module")
def env():
    return dict(gpu=krr.Wfpt(params={}), cases=LeafCases(seed=7272), gold=dict(np.load(os.path.join(GOLD, "leaf_vectors.npz"))))


def close_rows(got, ref, rtol=2e-4, atol=1e-6):
    return np.all(np.abs(got - ref) <= rtol * np.abs(ref) + atol, axis=1)


def test_bsdf_f_pdf_sample_all_material_types(env):
    c, g = env["cases"], env["gold"]
    b = c.bsdf
    n = len(b["type"])
    qs = []
    for i in range(n):
        qs.append(krr.KrrLeafBsdfQuery(
            ior=b["ior"][i], diffuse=(F * 4)(*b["diffuse"][i]), specular=(F * 4)(*b["specular"][i]), specular_transmission=b["strans"][i],
            roughness=b["rough"][i], metallic=b["metallic"][i], anisotropic=b["aniso"][i], bsdf_type=int(b["type"][i]),
            wo=(F * 3)(*b["wo"][i]), wi=(F * 3)(*b["wi"][i]), wavelength_u=b["wl_u"][i], seed_px=int(b["seed"][i][0]) & 0xffff,
            seed_py=int(b["seed"][i][0]) >> 16, seed_index=int(b["seed"][i][1]), eta_kind=int(b["eta_kind"][i]), eta=b["eta"][i]))
    res = env["gpu"].eval_bsdf(qs)
    types = np.array([r.type_flags for r in res], np.int32)
    assert np.array_equal(types, g["bsdf_type"]), "BSDFData::getBsdfType differs"
    ev = np.array([list(r.f) + [r.pdf] for r in res], np.float32)
    sm = np.array([list(r.s_f) + list(r.s_wi) + [r.s_pdf, float(r.s_flags)] for r in res], np.float32)
    names = ["null", "diffuse", "dielectric", "conductor", "disney"]
    for t in range(5):
        sl = slice(t * c.N_BSDF, (t + 1) * c.N_BSDF)
        ok_e = close_rows(ev[sl], g["bsdf_eval"][sl])
        ok_s = close_rows(sm[sl], g["bsdf_sample"][sl])
        broad = b["rough"][sl] ** 2 >= 0.02  # the tight tolerance is asked of the well-conditioned lobes (see below)
        assert ok_e[broad].mean() >= 0.97, (names[t], "f/pdf", np.where(~ok_e)[0][:5], ev[sl][~ok_e][:3], g["bsdf_eval"][sl][~ok_e][:3])
        assert ok_s[broad].mean() >= 0.97, (names[t], "sample", np.where(~ok_s)[0][:5], sm[sl][~ok_s][:3], g["bsdf_sample"][sl][~ok_s][:3])
        # every case: 1e-2, or 1e-1 for near-specular lobes (alpha = roughness^2 < 0.02: D(wh) ~ 1/alpha^2 turns a
        # 2-ulp change of wh -- the device uses the 2-ulp fast division -- into percents)
        sharp = (b["rough"][sl] ** 2 < 0.02)[:, None]
        tol = np.where(sharp, 1e-1, 1e-2)
        all_e = np.all(np.abs(ev[sl] - g["bsdf_eval"][sl]) <= tol * np.abs(g["bsdf_eval"][sl]) + 1e-5, axis=1)
        all_s = np.all(np.abs(sm[sl] - g["bsdf_sample"][sl]) <= tol * np.abs(g["bsdf_sample"][sl]) + 1e-5, axis=1)
        assert all_e.all() and all_s.all(), (names[t], np.where(~all_e)[0], np.where(~all_s)[0])
        assert np.array_equal(sm[sl][:, 8], g["bsdf_sample"][sl][:, 8]), (names[t], "sampled lobe flags differ")


def test_area_and_analytic_lights(env):
    c, g = env["cases"], env["gold"]
    t = c.tri
    qs = []
    for i in range(len(t["scale"])):
        q = krr.KrrLeafLightQuery(type=3, transform=(F * 12)(*t["xform"][i]), color=(F * 3)(*t["Le"][i]), scale=t["scale"][i], two_sided=int(t["two_sided"][i]),
                                  u=(F * 2)(*t["u"][i]), ctx_p=(F * 3)(*t["ctx_p"][i]), ctx_n=(F * 3)(*t["ctx_n"][i]), wavelength_u=t["wl_u"][i])
        for a in range(3):
            for k in range(3):
                q.p[a][k], q.n[a][k] = t["p"][i][a][k], t["n"][i][a][k]
        qs.append(q)
    res = env["gpu"].eval_light(qs)
    got = np.array([list(r.p) + list(r.n) + list(r.L) + [r.pdf] + list(r.L_eval) + [r.pdf_li] for r in res], np.float32)
    ok = close_rows(got, g["arealight"], rtol=5e-4, atol=1e-5)
    assert ok.mean() >= 0.99, (np.where(~ok)[0], got[~ok][:3], g["arealight"][~ok][:3])
    l = c.light
    qs = []
    for i in range(len(l["type"])):
        qs.append(krr.KrrLeafLightQuery(type=int(l["type"][i]), transform=(F * 12)(*l["xform"][i]), color=(F * 3)(*l["color"][i]), scale=l["scale"][i],
                                        scene_radius=l["radius"][i], cos_inner=l["cos_inner"][i], cos_outer=l["cos_outer"][i], u=(F * 2)(*l["u"][i]),
                                        ctx_p=(F * 3)(*l["ctx_p"][i]), wi=(F * 3)(*l["wi"][i]), wavelength_u=l["wl_u"][i]))
    res = env["gpu"].eval_light(qs)
    got = np.array([list(r.p) + list(r.L) + [r.pdf] + list(r.L_eval) for r in res], np.float32)
    ok = close_rows(got, g["light"], rtol=5e-4, atol=1e-5)
    assert ok.mean() >= 0.97, (np.where(~ok)[0], got[~ok][:3], g["light"][~ok][:3])


def test_colour_conversions(env):
    c, g = env["cases"], env["gold"]
    # fromRGB: wavelengths of rgb_u; toRGB / lum: wavelengths of spec_u  -> two passes
    a = env["gpu"].eval_color(np.concatenate([c.rgb, c.rgb_u[:, None], np.zeros((64, 4), np.float32)], 1))
    # RGBBounded is only defined on [0,1]^3 (rows 0..47); rows 48.. are HDR colours for the unbounded types
    np.testing.assert_allclose(a[:48, :4], g["from_rgb"][:48, :4], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(a[:, 4:12], g["from_rgb"][:, 4:12], rtol=2e-4, atol=1e-6)
    b = env["gpu"].eval_color(np.concatenate([np.zeros((64, 3), np.float32), c.spec_u[:, None], c.spec], 1))
    np.testing.assert_allclose(b[:, 12:16], g["to_rgb_lum"], rtol=2e-4, atol=2e-5)
    w = env["gpu"].eval_color(np.concatenate([np.zeros((64, 3), np.float32), c.wl_u[:, None], np.zeros((64, 4), np.float32)], 1))
    assert np.array_equal(w[:, 16:20].view(np.uint32), g["wavelengths"][:, :4].view(np.uint32)), "sampled wavelengths must be bit-exact"


def test_camera_rays_bit_exact_pinhole_and_close_thin_lens(env):
    c, g = env["cases"], env["gold"]
    for k in range(2):
        cc = c.cams[k]
        cam = krr.KrrCameraData(film_size=(F * 2)(cc[0], cc[1]), focal_length=cc[2], focal_distance=cc[3], lens_radius=cc[4], aspect_ratio=cc[5],
                                shutter_open=cc[6], shutter_time=cc[7], transform=(F * 12)(*cc[8:20]), medium=-1)
        idx = np.arange(k, 32, 2)
        inp = np.concatenate([c.cam_px[idx].astype(np.float32), c.cam_cs[idx]], 1)
        out = env["gpu"].camera_rays(cam, 640, 480, inp)
        if k == 0:
            assert np.array_equal(out.view(np.uint32), g["camera"][idx].view(np.uint32)), "pinhole camera rays must be bit-exact"
        else:
            np.testing.assert_allclose(out, g["camera"][idx], rtol=1e-5, atol=1e-6)
