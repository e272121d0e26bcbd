#!/bin/bash
# One GPU session: bash tools/gpu_session.sh <tag> [pytest-selection...]   (outputs under gpurun_out/<tag>_*)
TAG=${1:-s}; shift
mkdir -p gpurun_out
S=gpurun_out/${TAG}
timeout 600 python -m pytest tests/test_vo_pipeline_gpu.py tests/test_tracker_gpu.py -m gpu -q -x > ${S}_pytest_vo.log 2>&1; echo "pytest vo rc=$?" >> ${S}_session.log
timeout 900 python -m pytest tests -m gpu -q > ${S}_pytest.log 2>&1; echo "pytest rc=$?" >> ${S}_session.log
timeout 900 python bench.py --steps 5 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc=$?" >> ${S}_session.log
cat ${S}_session.log; tail -5 ${S}_pytest_vo.log; tail -3 ${S}_pytest.log; tail -5 ${S}_bench.err
