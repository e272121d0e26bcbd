"""The COMPLETE library built for the host (tests/emu_build.py: every translation unit of csrc/, kernels through
tests/cpp/cuda_emu.h — the cluster LM kernels of ba.cu included: the blocks of a thread-block cluster run concurrently, with a
cluster barrier and distributed-shared-memory address translation — and the CUDA runtime through tests/emu/cuda_runtime_emu.cpp)
and driven through the ordinary ctypes wrapper (MVO_LIB points it at the emulated library) in a child process.

* always: bundle adjustment (both kernels) and the whole solvePnPRansac replacement (hypotheses, scoring, consensus, LM refit)
  against the oracles;
* MVO_SLOW_TESTS=1: `__graft_entry__.smoke()` — every stage plus three frames through the device-resident tracker (worker thread,
  pre-match, fused match-list filter, PnP, BA from the frame ring) against the CPU tracker — about five minutes."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    import emu_build
    import mvo_b200
    so = emu_build.build(tmp_path_factory.mktemp("fullemu"), emu_build.ALL_UNITS)
    import ctypes as C
    lib = C.CDLL(str(so))
    missing = [n for n in mvo_b200.SIGNATURES if not hasattr(lib, n)]
    assert not missing, missing                                     # the emulated build exports the whole C ABI
    return so


def _child(emu_lib, code, timeout):
    env = dict(os.environ)
    env["MVO_LIB"] = str(emu_lib)
    r = subprocess.run([sys.executable, "-c", code.format(root=str(ROOT))], capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and "child ok" in r.stdout, (r.stdout[-1500:], r.stderr[-2500:])
    return r.stdout


BA_PNP = r'''
import sys
import numpy as np
sys.path.insert(0, r"{root}"); sys.path.insert(0, r"{root}/monocular-visual-odometry_b200/python")
import mvo_b200, mvo_synth
from oracle import oracle_lib, pnp_oracle as po
ctx = mvo_b200.Context(0, max_keypoints=500, pnp_hypotheses=512)
pb = mvo_synth.ba_problem(0, n_frames=3, n_points=120)
ctx.set_params(ba_iterations=5)
for fix in (True, False):
    gp, gx, gs = ctx.bundle_adjustment(pb["T_w_c"], pb["points"], pb["edge_frame"], pb["edge_point"], pb["obs"], pb["K"], fix_points=fix, update_points=not fix)
    op, ox, os_ = oracle_lib.bundle_adjustment(pb["T_w_c"], pb["points"], pb["edge_frame"], pb["edge_point"], pb["obs"], pb["K"], fix_points=fix,
                                               update_points=not fix, iterations=5)
    assert np.abs(gp - op).max() < 1e-8 and abs(gs[1] - os_[1]) < 1e-9 * os_[1] and gs[2] == os_[2], (fix, np.abs(gp - op).max())
    if not fix:
        assert np.abs(gx - ox).max() < 1e-6
P, uv, rvec_t, tvec_t, _ = mvo_synth.pnp_problem(0, n=300)
rvec, tvec, inl = ctx.solve_pnp_ransac(P, uv, mvo_synth.K_DEFAULT)
assert np.abs(rvec - rvec_t).max() < 5e-3 and np.abs(tvec - tvec_t).max() < 2e-2 and len(inl) > 150 and np.all(np.diff(inl) > 0)
ro, to = po.refine(P[inl], uv[inl], mvo_synth.K_DEFAULT, rvec, tvec)           # the refit converged: the least-squares optimum of the consensus set
assert np.abs(rvec - ro).max() < 1e-6 and np.abs(tvec - to).max() < 1e-6
print("child ok")
'''


def test_emulated_ba_and_pnp_match_the_oracles(emu_lib):
    _child(emu_lib, BA_PNP, 600)


SMOKE = r'''
import sys
sys.path.insert(0, r"{root}")
import __graft_entry__ as g
g.smoke()
print("child ok")
'''


@pytest.mark.skipif(os.environ.get("MVO_SLOW_TESTS", "0") == "0", reason="about five minutes: set MVO_SLOW_TESTS=1")
def test_smoke_against_the_emulated_library(emu_lib):
    out = _child(emu_lib, SMOKE, 1800)
    assert "tracked frames ok" in out
