#!/bin/bash
# ncu launch lists (gpu__time_duration per launch) of a short state-machine pass under environment switches.
#   gpurun -- 'bash tools/gpu_session_launches.sh TAG cfg1 cfg2 ...'    (cfg = base | VAR=v[,VAR=v])
TAG=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  envs=""; [[ "$cfg" != base ]] && envs=$(echo $cfg | tr ',' ' ')
  out=gpurun_out/${TAG}_launches_${cfg//[,=]/_}.csv
  env $envs timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out python tools/dev_vo_pass.py 60 1 > /dev/null 2>&1
  echo "== $cfg"; python tools/launch_summary.py $out | head -30
done
