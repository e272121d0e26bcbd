#!/bin/bash
# Refresh of the end-of-round evidence after the last kernel change: bench line, launch list, ncu capture of the pose LM.
mkdir -p gpurun_out
S=gpurun_out/r2z
timeout 500 python bench.py > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc=$?"
MVO_VO_DEBUG=1 timeout 200 python tools/dev_vo_pass.py 150 3 > ${S}_vo_debug.log 2>&1; echo "vo debug rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${S}_launches.csv python tools/dev_vo_pass.py 150 1 > ${S}_launches.log 2>&1; echo "launch list rc=$?"
k=k_ba_pose
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 40 -c 2 -o ${S}_prof_$k -f python tools/dev_vo_pass.py 60 1 > ${S}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
ncu -i ${S}_prof_$k.ncu-rep --page raw --csv > ${S}_raw_$k.csv 2>/dev/null
ncu -i ${S}_prof_$k.ncu-rep --page source --csv > ${S}_src_$k.csv 2>/dev/null
head -c 600 ${S}_bench.json
