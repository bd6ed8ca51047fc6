#!/bin/bash
# A/B of library variants on the C2 prefilter (run on the GPU box): tools/gpu_ab_prefilter.sh <tag> <lib> [<tag> <lib> ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  tag=$1; lib=$2; shift 2
  echo "=== $tag streaming"; SD_LIB_PATH=$lib python tools/gpu_pf_only.py 20 3000 3 2>&1 | grep -v "^\[sd\]"
  echo "=== $tag binned"; SD_PF_STREAM_MAX=0 SD_LIB_PATH=$lib python tools/gpu_pf_only.py 20 3000 2 2>&1 | grep -v "^\[sd\]"
  echo "=== $tag ncu traffic (streaming)"
  SD_LIB_PATH=$lib ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct \
     --clock-control none -k regex:prefilter -c 1 --csv --log-file gpurun_out/ab_${tag}_traffic.csv python tools/gpu_pf_only.py 20 3000 1 > /dev/null 2>&1
  grep -E "prefilter" gpurun_out/ab_${tag}_traffic.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
