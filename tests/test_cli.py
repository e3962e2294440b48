"""krr_render (kiraray_b200/host/krr_render.cpp): the headless counterpart of src/main/kiraray.cpp:5-32."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "kiraray_b200", "lib", "krr_render")
ENV = dict(os.environ, KRR_DATA_DIR=os.path.join(ROOT, "kiraray_b200", "data"), KRR_ASSET_ROOT=ROOT)


def test_cli_is_built_and_reports_errors():
    assert os.path.exists(CLI), "python -m kiraray_b200.build"
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode != 0 and "usage" in r.stderr
    r = subprocess.run([CLI, "/nonexistent/config.json"], capture_output=True, text=True, env=ENV)
    assert r.returncode != 0 and "krr_render" in r.stderr


def test_data_manifest_matches_the_tables():
    from kiraray_b200 import build
    assert build.check_data()


@pytest.mark.gpu
def test_cli_renders_the_cornell_box(tmp_path):
    import kiraray_b200 as krr
    out = str(tmp_path / "film.pfm")
    r = subprocess.run([CLI, os.path.join(ROOT, "assets", "configs", "cbox.json"), "3", out], capture_output=True, text=True, env=ENV, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert "3 frames rendered" in r.stderr
    img = krr.load_image(out)
    assert np.isfinite(img).all() and img[..., :3].mean() > 0.01
