#!/bin/bash
# A/B timings of one instrumented pass under environment switches: each argument after the tag is "VAR=value[,VAR=value]" (or "base")
TAG=${1:-r2ab}; shift
mkdir -p gpurun_out
S=gpurun_out/${TAG}
for cfg in "$@"; do
  envs=""; [[ "$cfg" != base ]] && envs=$(echo $cfg | tr ',' ' ')
  echo "== $cfg" >> ${S}_ab.log
  env $envs MVO_VO_DEBUG=1 timeout 300 python tools/dev_vo_pass.py 150 3 2>&1 | grep -E "^pass|tracked|init " >> ${S}_ab.log
done
cat ${S}_ab.log
