#!/bin/bash
# First GPU call of the next round, in one gpurun invocation (about 12 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session_r2.sh'
# Everything lands under gpurun_out/.  Order: the validated suite first, then the two paths that have not run on hardware
# (homography kernels, assembled VO pipeline) in their own processes with their own timeouts, then the bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/session.log
timeout 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/session.log
MVO_EPI_DEBUG=1 timeout 300 python -m pytest tests/test_homography_gpu.py -m gpu -q -rxX --runxfail > gpurun_out/homography.log 2>&1; echo "homography rc=$?" >> gpurun_out/session.log
timeout 400 python -m pytest tests/test_vo_pipeline_gpu.py -m gpu -q -rxX --runxfail > gpurun_out/vo_pipeline.log 2>&1; echo "vo pipeline rc=$?" >> gpurun_out/session.log
timeout 300 python tests/dev/run_vo_synth.py 40 0 > gpurun_out/run_vo_synth_e.json 2> gpurun_out/run_vo_synth_e.err; echo "run_vo_synth(E) rc=$?" >> gpurun_out/session.log
timeout 300 python tests/dev/run_vo_synth.py 40 1 > gpurun_out/run_vo_synth_eh.json 2> gpurun_out/run_vo_synth_eh.err; echo "run_vo_synth(E+H) rc=$?" >> gpurun_out/session.log
timeout 300 python tools/dev_orb_variants.py > gpurun_out/orb_variants.jsonl 2> gpurun_out/orb_variants.err; echo "orb variants rc=$?" >> gpurun_out/session.log
timeout 400 python tools/multi_sequence_bench.py 1 2 4 8 > gpurun_out/multi_sequence.jsonl 2> gpurun_out/multi_sequence.err; echo "multi sequence rc=$?" >> gpurun_out/session.log
timeout 600 python bench.py --steps 300 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/session.log
cat gpurun_out/session.log
