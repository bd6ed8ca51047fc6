#!/bin/bash
# short C2 bench lines for tuning A/B on the GPU box: tools/gpu_quick_bench.sh "<tag>:<VAR=VAL ...>" ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  env $envs python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/qb_$tag.json 2> gpurun_out/qb_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/qb_%s.json" % tag).read().strip().splitlines()[-1])
    s = d["stage_ms_per_step"]
    print("%-10s step %.1f ms  prefilter kernel %.1f  sw %.1f  traceback %.1f  e2e %.1f ms" % (tag, d["ms_per_step"], s["prefilter_kernel_ms"], s["sw_kernel_ms"], s["traceback_ms"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print(tag, "failed", e, open("gpurun_out/qb_%s.err" % tag).read()[-400:])
PY
done
