#!/bin/bash
# ncu --set full captures (one launch each, source imported) of the kernels named on the command line after the tag;
# "name:skip" picks the launch after skipping `skip` earlier ones
TAG=${1:-r2p}; shift
mkdir -p gpurun_out
S=gpurun_out/${TAG}
for spec in "$@"; do
  k=${spec%%:*}; skip=0; [[ "$spec" == *:* ]] && skip=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o ${S}_prof_$k -f python tools/dev_vo_pass.py 40 1 > ${S}_ncu_$k.log 2>&1; echo "ncu $k rc=$?" >> ${S}_session.log
done
cat ${S}_session.log
