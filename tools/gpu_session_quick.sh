#!/bin/bash
# tests (files in $2) + one bench line; summary on stdout.   gpurun -- 'bash tools/gpu_session_quick.sh TAG "tests..." [bench args]'
TAG=$1; TESTS=$2; shift 2
mkdir -p gpurun_out
S=gpurun_out/${TAG}
if [ -n "$TESTS" ]; then timeout 900 python -m pytest $TESTS -m gpu -q -x > ${S}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${S}_pytest.log; fi
timeout 400 python bench.py "$@" > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("${S}_bench.json"))
print("fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e_pageable"]["value"]), "cpu", round(d.get("cpu_baseline", {}).get("value", 0), 1))
print({k: round(v["ms_mean"], 3) for k, v in d["detail"]["ms_per_frame_by_kind"].items()})
print("orb_batch", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["extra"]["orb_batch"].items() if k != "note"})
print("stages", {k: round(v["us_per_frame"], 1) for k, v in sorted(d["stages"].items(), key=lambda x: -x[1]["us_per_frame"])})
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["us_per_launch"], 1), d["roofline"]["frac"], d["roofline"]["issue_frac"])
PY
