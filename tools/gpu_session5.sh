#!/bin/bash
# tests of the tracking / two-view path + instrumented passes (stage clocks, tracker debug)
TAG=${1:-r2s13}
mkdir -p gpurun_out
S=gpurun_out/${TAG}
timeout 900 python -m pytest tests/test_track_filter_gpu.py tests/test_vo_pipeline_gpu.py tests/test_tracker_gpu.py tests/test_orb_gpu.py -m gpu -q > ${S}_pytest.log 2>&1; echo "pytest rc=$?" >> ${S}_session.log
MVO_VO_DEBUG=1 timeout 300 python tools/dev_vo_pass.py 150 3 > ${S}_vo_pass.log 2>&1; echo "vo pass rc=$?" >> ${S}_session.log
MVO_TRACK_DEBUG=1 timeout 300 python tools/dev_vo_pass.py 150 2 > ${S}_trk_dbg.log 2>&1; echo "trk dbg rc=$?" >> ${S}_session.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file ${S}_launches.csv python tools/dev_vo_pass.py 40 1 > ${S}_ncu_list.log 2>&1; echo "ncu list rc=$?" >> ${S}_session.log
cat ${S}_session.log; tail -5 ${S}_pytest.log; grep -E "^pass|^  init" ${S}_vo_pass.log | tail -4; grep -E "tracker\(device\)|tracker: waited" ${S}_trk_dbg.log | tail -6
