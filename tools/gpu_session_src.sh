#!/bin/bash
# ncu --set full source-level captures (dense sampling) of the kernels named after the tag: "name:skip"; exports raw + source CSV pages
TAG=$1; shift
mkdir -p gpurun_out
S=gpurun_out/${TAG}
cp monocular-visual-odometry_b200/libmvo.so ${S}_libmvo.so 2>/dev/null
for spec in "$@"; do
  k=${spec%%:*}; skip=20; [[ "$spec" == *:* ]] && skip=${spec##*:}
  timeout 300 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:"^$k" -s $skip -c 1 -o ${S}_prof_$k -f python tools/dev_vo_pass.py 40 1 > ${S}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
  ncu -i ${S}_prof_$k.ncu-rep --page raw --csv > ${S}_raw_$k.csv 2>/dev/null
  ncu -i ${S}_prof_$k.ncu-rep --page source --csv > ${S}_src_$k.csv 2>/dev/null
  rm -f ${S}_prof_$k.ncu-rep
done
rm -f ${S}_libmvo.so
