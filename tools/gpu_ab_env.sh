#!/bin/bash
# A/B of environment switches on the C2 prefilter (run on the GPU box): tools/gpu_ab_env.sh "<tag>:<VAR=VAL ...>" ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  echo "=== $tag [$envs] streaming"; env $envs python tools/gpu_pf_only.py 20 3000 3 2>&1 | grep -v "^\[sd\]" | tail -2
  echo "=== $tag [$envs] binned"; env $envs SD_PF_STREAM_MAX=0 python tools/gpu_pf_only.py 20 3000 2 2>&1 | grep -v "^\[sd\]" | tail -1
  echo "=== $tag ncu traffic (streaming)"
  env $envs ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
     --clock-control none -k regex:prefilter -c 1 --csv --log-file gpurun_out/ab_${tag}_traffic.csv python tools/gpu_pf_only.py 20 3000 1 > /dev/null 2>&1
  grep -E "prefilter" gpurun_out/ab_${tag}_traffic.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
