#!/bin/bash
# A/B of one state-machine pass under environment switches + a subset of the GPU suite.
#   gpurun --timeout 900 -- 'bash tools/gpu_session_ab2.sh TAG "test files" cfg1 cfg2 ...'   (cfg = base | VAR=v[,VAR=v])
TAG=$1; TESTS=$2; shift 2
mkdir -p gpurun_out
S=gpurun_out/${TAG}
rm -f ${S}_ab.log
if [ -n "$TESTS" ]; then timeout 600 python -m pytest $TESTS -m gpu -q -x > ${S}_pytest.log 2>&1; echo "pytest rc=$?" >> ${S}_ab.log; tail -3 ${S}_pytest.log >> ${S}_ab.log; fi
for cfg in "$@"; do
  envs=""; [[ "$cfg" != base ]] && envs=$(echo $cfg | tr ',' ' ')
  echo "== $cfg" >> ${S}_ab.log
  env $envs MVO_VO_DEBUG=1 timeout 300 python tools/dev_vo_pass.py 150 3 > ${S}_pass_${cfg//[,=]/_}.log 2>&1
  grep -E "^pass|tracked|checksum" ${S}_pass_${cfg//[,=]/_}.log >> ${S}_ab.log
  grep -E "^keyframe" ${S}_pass_${cfg//[,=]/_}.log | tail -3 >> ${S}_ab.log
done
cat ${S}_ab.log
