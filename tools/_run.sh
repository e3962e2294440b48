KRR_DATA_DIR=$PWD/kiraray_b200/data ./kiraray_b200/lib/krr_render assets/configs/cbox.json 3 /tmp/film.pfm; echo rc=$?; ls -la /tmp/film.pfm
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_cli.py::test_cli_renders_the_cornell_box 2>&1 | tail -3
