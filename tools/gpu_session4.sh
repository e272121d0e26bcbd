#!/bin/bash
# short hardware session: the state-machine / two-view tests, one instrumented pass, the launch list, the bench line
TAG=${1:-r2s9}
mkdir -p gpurun_out
S=gpurun_out/${TAG}
timeout 600 python -m pytest tests/test_vo_pipeline_gpu.py tests/test_tracker_gpu.py tests/test_homography_gpu.py tests/test_epipolar_gpu.py -m gpu -q > ${S}_pytest_vo.log 2>&1; echo "pytest vo rc=$?" >> ${S}_session.log
MVO_VO_DEBUG=1 timeout 300 python tools/dev_vo_pass.py 150 2 > ${S}_vo_pass.log 2>&1; echo "vo pass rc=$?" >> ${S}_session.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file ${S}_launches.csv python tools/dev_vo_pass.py 40 1 > ${S}_ncu_list.log 2>&1; echo "ncu list rc=$?" >> ${S}_session.log
timeout 900 python bench.py --steps 5 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc=$?" >> ${S}_session.log
cat ${S}_session.log; tail -5 ${S}_pytest_vo.log; tail -3 ${S}_vo_pass.log; tail -3 ${S}_bench.err
