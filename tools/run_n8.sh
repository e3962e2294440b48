#!/bin/bash
# GPU box with 8 GPUs (gpurun --gpus 8): the tree workloads with the hybrid tile x spp split (BASELINE configs[2] / [4] are stated for
# 8 x B200) and the headline with a fixed total number of samples (strong scaling).  usage: tools/run_n8.sh <tag> [N]
TAG=${1:-r2j}; N=${2:-8}
mkdir -p gpurun_out
run() { # name, bench args...
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps 5 --warmup 3 --min-seconds 2 "$@" > gpurun_out/bench_${name}_n${N}_$TAG.json 2> gpurun_out/bench_${name}_n${N}_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${name}_n${N}_$TAG.json"))
    print("$name N=$N", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), d["scaling"], d["config"]["parallelism"], d["clocks"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/bench_${name}_n${N}_$TAG.err").read()[-1500:])
PY
}
run tess20m_hybrid --workload tess20m --partition hybrid
run inst10k_hybrid --workload inst10k --partition hybrid
run cbox_strong_hybrid --workload cbox --partition hybrid --strong
