"""Generates the committed golden vectors from the third-party code the reference calls
(OpenCV, via the cv2 4.13 wheel in this container).  Run from the repo root:
    python tests/golden/make_golden.py [match|orb|pnp|all]
The reference ships no fixtures of its own; these pin the oracle."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python"))
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent


def make_match():
    import cv2
    import mvo_synth
    d1 = mvo_synth.random_descriptors(101, 211, dup_every=9)
    d2 = mvo_synth.random_descriptors(102, 197, dup_every=6)
    d1[::4] = d2[(np.arange(0, 211, 4) * 5) % 197]
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    nn = bf.match(d1, d2)
    kn = bf.knnMatch(d1, d2, k=2)
    np.savez_compressed(OUT / "match_golden.npz", d1=d1, d2=d2,
                        nn_train=np.array([m.trainIdx for m in nn], np.int32),
                        nn_dist=np.array([m.distance for m in nn], np.float32),
                        knn_train=np.array([[m.trainIdx for m in mm] for mm in kn], np.int32),
                        knn_dist=np.array([[m.distance for m in mm] for mm in kn], np.float32),
                        cv_version=cv2.__version__)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("match", "all"):
        make_match()
    if what in ("orb", "all") and "make_orb" in globals():
        make_orb()
    if what in ("pnp", "all") and "make_pnp" in globals():
        make_pnp()
