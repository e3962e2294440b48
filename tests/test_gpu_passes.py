"""GPU tests of the passes that follow the path tracer on the reference's example configs:
AccumulatePass (float / double, budgets, save_on_finish), ErrorMeasurePass and ToneMappingPass, each
against a numpy restatement of the reference kernel it replaces."""
import json
import os

import numpy as np
import pytest
import torch

import kiraray_b200 as krr
from __graft_entry__ import relmse

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cbox_config(passes, w=48, h=48, **extra):
    cfg = json.load(open(os.path.join(ROOT, "assets", "configs", "cbox.json")))
    cfg["resolution"] = [w, h]
    cfg["passes"] = passes
    cfg.update(extra)
    return cfg


WFPT = {"enable": True, "name": "WavefrontPathTracer", "params": {"nee": True, "rr": 0.8, "max_depth": 5}}


def metric_numpy(y, ref, metric):
    """metrics.cu:64-128 in float64: per-pixel error, mean over RGB, clamp at 100, mean over pixels."""
    y, ref = y[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
    d = np.abs(y - ref)
    with np.errstate(divide="ignore", invalid="ignore"):
        e = {0: d ** 2, 1: d / ref, 2: d / (ref + y), 3: np.where(ref == 0, 0.0, (d / ref) ** 2)}[metric]
    pp = e.mean(axis=-1)
    pp = np.where(np.isnan(pp), 100.0, np.minimum(pp, 100.0))  # fminf(NaN, 100) = 100
    bad = ~np.isfinite(ref).all(axis=-1)
    pp[bad] = 0
    return float(pp.mean())


@pytest.mark.parametrize("metric", [0, 1, 2, 3])
def test_error_metric_kernel(metric):
    rng = np.random.Generator(np.random.PCG64(7272))
    h, w = 120, 200
    ref = rng.uniform(0.01, 2, (h, w, 4)).astype(np.float32)
    y = (ref * rng.uniform(0.5, 1.5, (h, w, 4))).astype(np.float32)
    ref[0, 0, :3] = 0            # rel_mse: ref == 0 -> 0; mape: division by zero -> clamp
    ref[1, 1, 0] = np.inf        # invalid reference pixel counts as 0
    y[2, 2, :3] = 1e9            # clamps at 100
    dy, dr = torch.from_numpy(y).cuda(), torch.from_numpy(ref).cuda()
    got = krr.error_metric(dy.data_ptr(), dr.data_ptr(), h * w, metric)
    want = metric_numpy(y, ref, metric)
    assert got == pytest.approx(want, rel=2e-5), (metric, got, want)


def tonemap_numpy(img, op, exposure, gamma):
    c = img[..., :3].astype(np.float32) * np.float32(exposure)
    if op == 1:
        lum = c @ np.array([0.299, 0.587, 0.114], np.float32)
        r = lum / (lum + 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            c = np.where(lum[..., None] == 0, 0, c * r[..., None] / lum[..., None])
    elif op == 2:
        c = c * 0.6
        c = np.clip((c * (2.51 * c + 0.03)) / (c * (2.43 * c + 0.59) + 0.14), 0, 1)
    elif op == 3:
        A, B, C_, D, E, F = 0.22, 0.3, 0.1, 0.2, 0.01, 0.3
        c = ((c * (A * c + C_ * B) + D * E) / (c * (A * c + B) + D * F)) - (E / F)
    elif op == 4:
        c = np.maximum(c - 0.004, 0)
        c = ((c * (6.2 * c + 0.5)) / (c * (6.2 * c + 1.7) + 0.06)) ** 2.2
    if gamma:
        c = np.power(c.astype(np.float64), 0.45454545)
    out = np.ones(img.shape, np.float32)
    out[..., :3] = c
    return out


@pytest.mark.parametrize("op", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("gamma", [False, True])
def test_tonemap_kernel(op, gamma):
    rng = np.random.Generator(np.random.PCG64(7272 + op))
    img = rng.uniform(0, 3, (90, 160, 4)).astype(np.float32)
    if op != 3:  # Uncharted2 at 0 is D*E/(D*F) - E/F: pure cancellation, amplified by the gamma curve
        img[0, 0, :3] = 0
    d = torch.from_numpy(img).cuda()
    krr.tonemap(d.data_ptr(), 90 * 160, op, 1.3, gamma)
    got = d.cpu().numpy()
    want = tonemap_numpy(img, op, 1.3, gamma)
    assert np.all(got[..., 3] == 1)
    assert np.allclose(got[..., :3], want[..., :3], rtol=2e-5, atol=2e-6, equal_nan=True)


def test_accumulate_double_matches_float64_mean():
    rng = np.random.Generator(np.random.PCG64(7272))
    n = 64 * 64
    frames = [rng.uniform(0, 5, (n, 4)).astype(np.float32) for _ in range(9)]
    acc = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
    for k, f in enumerate(frames):
        d = torch.from_numpy(f).cuda()
        krr.accumulate_f64(acc.data_ptr(), d.data_ptr(), n, k)
        want = (np.sum([x.astype(np.float64) for x in frames[:k + 1]], axis=0) / (k + 1)).astype(np.float32)
        assert np.array_equal(d.cpu().numpy(), want)


def test_config_with_all_passes_runs_to_its_budget_and_saves(tmp_path):
    """A reference-style config: path tracer -> accumulate (double, spp budget, save + exit on finish) ->
    error measure (continuous, saved log) -> tone mapping; driven by the headless main loop."""
    ref_app = krr.HostApp(cbox_config([WFPT, {"enable": True, "name": "AccumulatePass", "params": {"spp": 0}}]), asset_root=ROOT)
    ref_app.render_frames(24)
    reference = ref_app.read_accumulated()
    ref_app.close()

    passes = [WFPT,
              {"enable": True, "name": "AccumulatePass", "params": {"spp": 6, "precision": "double", "save_on_finish": True, "exit_on_finish": True,
                                                                  "task": {"type": "spp", "value": 6}}},
              {"enable": True, "name": "ErrorMeasurePass", "params": {"metric": "rel_mse", "continuous": True, "interval": 2, "save": True}},
              {"enable": True, "name": "ToneMappingPass", "params": {"operator": "aces", "exposure": 1.5, "gamma": True}}]
    app = krr.HostApp(cbox_config(passes, name="cbox_test"), asset_root=ROOT)
    app.set_output_dir(tmp_path)
    assert app.pass_json("AccumulatePass")["precision"] == "double" and app.pass_json("ToneMappingPass")["operator"] == "aces"
    app.render_frames(1)  # initialises the passes; the reference image can only be set on a live pass
    app.set_reference(reference)
    n = 1 + app.run(max_frames=50)
    assert n == 6 and app.accum_count() == 6, "the spp budget ends the loop"
    value, n_eval = app.last_error_metric()
    # setting the reference resets the pass's frame counter (errormeasure.cpp:86-91, 110): the 5 frames after it
    # are numbered 1..5 and interval 2 evaluates #2 and #4
    assert n_eval == 2 and value > 0
    # save_on_finish: <output>/<name>.exr holds the average, written like AccumulatePass::saveImage
    avg = app.read_accumulated()
    saved = krr.load_image(tmp_path / "cbox_test.exr", flip=True)[..., [3, 0, 1, 2]]
    assert np.allclose(saved, avg, rtol=1e-3, atol=1e-4)  # half precision
    log = json.load(open(tmp_path / "error" / "cbox_test.json"))
    assert log["timesteps"] == [2, 4] and len(log["data"]) == 2 and log["data"][-1]["RelMSE"] == pytest.approx(value)
    # the "Evaluate" button, one more frame.  The budget is spent, so nothing is added any more, and the film the
    # pass hands on is sum * 1 / (count + 1) exactly as in the reference (accumulate.cu:35, 46): 6/7 of the average
    app.evaluate_next_frame()
    app.render_frames(1)
    value, _ = app.last_error_metric()
    assert app.accum_count() == 6
    assert value == pytest.approx(relmse(avg * np.float32(6 / 7), reference), rel=1e-3)


def test_gltf_scene_renders_and_animates(tmp_path):
    """The glTF fixture of tests/test_gltf.py (textured emissive quads, one keyframe-animated) through the
    headless app: frames at different times differ (instance update -> TLAS refit), films are finite and lit."""
    from test_gltf import make_fixture
    make_fixture(str(tmp_path))
    cfg = {"resolution": [64, 64],
           "passes": [{"enable": True, "name": "WavefrontPathTracer", "params": {"max_depth": 3, "spp": 4}}],
           "scene": {"model": [{"model": "quad.gltf"}],
                     "camera": {"mData": {"focalLength": 21.0}},
                     "cameraController": {"mData": {"target": [4.0, 4.0, 10.0], "radius": 30.0, "pitch": 0.0, "yaw": 0.0}}}}
    app = krr.HostApp(cfg, asset_root=str(tmp_path))
    f1 = app.render_frames(1).copy()
    assert np.isfinite(f1).all() and f1[..., :3].max() > 0.1, "the emissive quads are visible"
    lib = app.lib
    # same frame index semantics, later time: the animated quad has moved
    app2 = krr.HostApp(cfg, asset_root=str(tmp_path))
    app2.camera(2.0)
    x = np.array(list(app2.scene_desc().contents.instances[0].transform)).reshape(3, 4)
    assert not np.allclose(x[:, 3], [18, 18, 28]), "animation sampled"


def test_megakernel_and_wavefront_estimators_agree(cbox_app):
    """The reference has two independent estimators of the same integral: the wavefront pass (spectral MIS
    with path pdfs pu / pl, integrator.cpp) and the megakernel (power-heuristic MIS on bsdf / light pdfs,
    megakernel/device.cu).  Both are built here from the same device routines but share no control flow,
    so their agreement cross-validates queues, routing, MIS weights and Russian roulette.  256 frames of
    1 spp at 64x64 each; compared on 8x8-pixel block means (noise of a block mean ~1.5 %)."""
    w = h = 64
    app = cbox_app(w, h, spp=1, max_depth=6)
    cam = app.camera()
    gpu = krr.Wfpt(params=dict(app.wfpt_params()))
    gpu.set_scene(app.scene_desc())
    gpu.resize(w, h)
    film = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    acc_w = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda")
    acc_m = torch.zeros_like(acc_w)
    n = 256
    for f in range(1, n + 1):
        gpu.begin_frame(f, cam)
        gpu.render(film.data_ptr())
        acc_w += film
        gpu.render_megakernel(f, cam, film.data_ptr())
        acc_m += film
    torch.cuda.synchronize()
    a = (acc_w / n).cpu().numpy()[..., :3]
    b = (acc_m / n).cpu().numpy()[..., :3]
    assert np.isfinite(a).all() and np.isfinite(b).all()
    blocks = lambda x: x.reshape(8, 8, 8, 8, 3).mean(axis=(1, 3))
    ba, bb = blocks(a), blocks(b)
    lit = ba.mean(axis=-1) > 0.02 * ba.mean()
    rel = np.abs(ba - bb)[lit] / np.maximum(ba[lit], 1e-3)
    assert a.mean() == pytest.approx(b.mean(), rel=0.02), (a.mean(), b.mean())
    assert np.median(rel) < 0.03 and np.percentile(rel, 95) < 0.15, (np.median(rel), np.percentile(rel, 95))


def test_megakernel_pass_from_config():
    passes = [{"enable": True, "name": "MegakernelPathTracer", "params": {"nee": True, "max_depth": 4, "rr": 0.8}},
              {"enable": True, "name": "AccumulatePass", "params": {"spp": 0}}]
    app = krr.HostApp(cbox_config(passes, 48, 48), asset_root=ROOT)
    film = app.render_frames(8)
    assert app.pass_json("MegakernelPathTracer") == {"nee": True, "max_depth": 4, "rr": pytest.approx(0.8), "spp": 1}
    assert np.isfinite(film).all() and film[..., :3].mean() > 0.05 and app.accum_count() == 8
