#!/bin/bash
# Round-2 GPU session 1 (one gpurun invocation, ~15 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session_r2.sh'
# Regression run of the suite, then the experiments VERDICT r1 asked for: concurrent sequences per GPU, the kernel
# variants' timing, the whole state machine on the 150-frame loop next to the oracle pipeline, the bench line.
mkdir -p gpurun_out
S=gpurun_out/r2s1
python -c "import __graft_entry__ as g; g.smoke()" > ${S}_smoke.log 2>&1; echo "smoke rc=$?" >> ${S}_session.log
timeout 900 python -m pytest tests -m gpu -q -x > ${S}_pytest.log 2>&1; echo "pytest rc=$?" >> ${S}_session.log
timeout 400 python tests/dev/run_vo_synth.py 150 1 loop > ${S}_run_vo_loop.json 2> ${S}_run_vo_loop.err; echo "run_vo_synth(loop) rc=$?" >> ${S}_session.log
timeout 300 python tools/dev_orb_variants.py > ${S}_orb_variants.jsonl 2> ${S}_orb_variants.err; echo "orb variants rc=$?" >> ${S}_session.log
timeout 500 python tools/multi_sequence_bench.py 1 2 4 8 16 > ${S}_multi_sequence.jsonl 2> ${S}_multi_sequence.err; echo "multi sequence rc=$?" >> ${S}_session.log
timeout 600 python bench.py --steps 300 --warmup 10 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc=$?" >> ${S}_session.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> ${S}_session.log
cat ${S}_session.log
