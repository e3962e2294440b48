timeout 900 python -m pytest tests/test_gpu_media.py tests/test_gpu_motion.py tests/test_gpu_sort_rays.py tests/test_cli.py -q -s 2>&1 | grep -E "albedo grid:|passed|failed|Error|assert" | head -20
PARAMS='{}' bash tools/abv.sh tess20m base sv1
PARAMS='{}' bash tools/abv.sh inst10k base sv1
for P in '{"l2_persist_mb":32}' '{"l2_persist_mb":96}'; do PARAMS=$P bash tools/abv.sh tess20m base; done
PARAMS='{"l2_persist_mb":8}' bash tools/abv.sh inst10k base
