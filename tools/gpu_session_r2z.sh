#!/bin/bash
# Round-2 FINAL evidence session (one gpurun call): suite, bench line, stage breakdown, ncu launch list, ncu --set full captures.
#   /usr/local/graft/bin/gpurun --timeout 2100 -- 'bash tools/gpu_session_r2z.sh'
mkdir -p gpurun_out
S=gpurun_out/r2z
rm -f ${S}_session.log
python -c "import __graft_entry__ as g; g.smoke()" > ${S}_smoke.log 2>&1; echo "smoke rc=$?" >> ${S}_session.log
timeout 900 python -m pytest tests -m gpu -q -x > ${S}_pytest.log 2>&1; echo "pytest rc=$?" >> ${S}_session.log
timeout 500 python bench.py > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc=$?" >> ${S}_session.log
MVO_VO_DEBUG=1 timeout 200 python tools/dev_vo_pass.py 150 3 > ${S}_vo_debug.log 2>&1; echo "vo debug rc=$?" >> ${S}_session.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${S}_launches.csv python tools/dev_vo_pass.py 150 1 > ${S}_launches.log 2>&1; echo "launch list rc=$?" >> ${S}_session.log
for spec in k_ba_pose:40:2 k_match_filter:20:1 k_fast:20:1 k_describe:20:1 k_retain:20:1 k_blur2:20:1 match_kernel:20:1 k_select_kept:20:1 \
            k_harris_all:20:1 k_pyramid:20:1 k_select:20:1 k_pnp_score:20:1 k_pnp_hypotheses:20:1 k_pnp_finish:20:1 k_epi_finish:1:1 k_track_glue:20:1; do
  IFS=: read k skip cnt <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s $skip -c $cnt -o ${S}_prof_$k -f python tools/dev_vo_pass.py 60 1 > ${S}_ncu_$k.log 2>&1; echo "ncu $k rc=$?" >> ${S}_session.log
  ncu -i ${S}_prof_$k.ncu-rep --page raw --csv > ${S}_raw_$k.csv 2>/dev/null
  ncu -i ${S}_prof_$k.ncu-rep --page source --csv > ${S}_src_$k.csv 2>/dev/null
  case $k in k_ba_pose|k_match_filter|k_fast|k_describe|k_retain) ;; *) rm -f ${S}_prof_$k.ncu-rep ;; esac
done
du -sh gpurun_out >> ${S}_session.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> ${S}_session.log
cat ${S}_session.log
tail -3 ${S}_pytest.log
head -c 1500 ${S}_bench.json
