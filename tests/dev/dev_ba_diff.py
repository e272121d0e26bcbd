import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python")); sys.path.insert(0, str(ROOT))
import numpy as np, mvo_b200, mvo_synth
from oracle import oracle_lib
ctx = mvo_b200.Context(0)
for fix in (True, False):
  for F,P,iters,ff in [(5,2000,10,0),(5,300,10,0),(1,100,10,0),(3,37,5,0),(8,500,10,0),(16,200,4,0),(5,400,50,1),(5,2000,50,0)]:
    pb = mvo_synth.ba_problem(F*7+P, n_frames=F, n_points=P, visibility=1.0 if P!=300 else 0.7)
    ctx.set_params(ba_iterations=iters, ba_fix_first_pose=ff)
    gp,gx,gs = ctx.bundle_adjustment(pb["T_w_c"], pb["points"], pb["edge_frame"], pb["edge_point"], pb["obs"], pb["K"], fix_points=fix, update_points=not fix)
    op,ox,os_ = oracle_lib.bundle_adjustment(pb["T_w_c"], pb["points"], pb["edge_frame"], pb["edge_point"], pb["obs"], pb["K"], fix_points=fix, update_points=not fix, iterations=iters, fix_first_pose=ff)
    print(f"fix={fix} F={F} P={P} it={iters} ff={ff}: iters {gs[2]:.0f}/{os_[2]:.0f} chi {gs[1]:.9g}/{os_[1]:.9g} rel {abs(gs[1]-os_[1])/max(os_[1],1e-300):.2e} lam {gs[3]:.4g}/{os_[3]:.4g} dT {np.abs(gp-op).max():.2e} dX {np.abs(gx-ox).max():.2e}")
