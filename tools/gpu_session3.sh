#!/bin/bash
TAG=${1:-r2s3}
mkdir -p gpurun_out
S=gpurun_out/${TAG}
timeout 600 python -m pytest tests/test_vo_pipeline_gpu.py tests/test_tracker_gpu.py tests/test_pnp_gpu.py -m gpu -q > ${S}_pytest_vo.log 2>&1; echo "pytest vo rc=$?" >> ${S}_session.log
MVO_VO_DEBUG=1 timeout 300 python tools/dev_vo_pass.py 150 2 > ${S}_vo_pass.log 2>&1; echo "vo pass rc=$?" >> ${S}_session.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file ${S}_launches.csv python tools/dev_vo_pass.py 40 1 > ${S}_ncu_list.log 2>&1; echo "ncu list rc=$?" >> ${S}_session.log
for k in k_epi_finish k_homo_finish k_epi_hypotheses; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o ${S}_prof_$k -f python tools/dev_vo_pass.py 30 1 > ${S}_ncu_$k.log 2>&1; echo "ncu $k rc=$?" >> ${S}_session.log
done
timeout 900 python bench.py --steps 5 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc=$?" >> ${S}_session.log
cat ${S}_session.log; tail -5 ${S}_pytest_vo.log; tail -3 ${S}_vo_pass.log; tail -3 ${S}_bench.err
