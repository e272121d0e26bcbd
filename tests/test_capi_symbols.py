"""CPU-side checks of the drop-in boundary: libmvo.so loads and exports every symbol that
include/mvo.h declares; host-only entry points behave like the reference."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_exported(built):
    import mvo_b200
    lib = mvo_b200.load_library()
    header = (ROOT / "include" / "mvo.h").read_text()
    declared = set(re.findall(r"\b(mvo_[a-z0-9_]+)\s*\(", header))
    declared -= {"mvo_status"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mvo.h but not exported"
    assert declared == set(mvo_b200.SIGNATURES), declared ^ set(mvo_b200.SIGNATURES)


def test_default_params_match_config_yaml(built):
    import mvo_b200
    p = mvo_b200.default_params()
    # reference config/config.yaml:65-69,66,84-85,94-95; g2o_ba.cpp:275; vo.cpp:316
    assert (p.orb_nfeatures, p.orb_nlevels, p.orb_fast_threshold) == (8000, 4, 20)
    assert abs(p.orb_scale_factor - 1.2) < 1e-6
    assert (p.max_keypoints, p.grid_size, p.max_pts_per_grid) == (1500, 16, 8)
    assert (p.xiang_gao_ratio, p.lowe_ratio) == (2.0, 1.0)
    assert p.ba_iterations == 50 and p.pnp_reproj_error == 2.0


def test_no_cpu_fallback(built):
    """Without a GPU mvo_create must fail loudly (no CPU path exists in the product)."""
    import ctypes as C
    import mvo_b200
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = mvo_b200.load_library()
    h = C.c_void_p()
    assert lib.mvo_create(C.byref(h), 0, None) == -2   # MVO_ERR_NO_DEVICE
    with pytest.raises(mvo_b200.MvoError):
        mvo_b200.Context(0)


def test_remove_duplicated_matches_host(built):
    import mvo_b200
    from oracle import oracle_lib
    rng = np.random.default_rng(3)
    for n in (0, 1, 5, 16, 17, 200, 2000):
        m = np.zeros(n, mvo_b200.DMATCH_DTYPE)
        m["query_idx"] = np.arange(n)
        m["train_idx"] = rng.integers(0, max(1, n // 2), n)
        m["distance"] = rng.integers(0, 100, n)
        got = mvo_b200.remove_duplicated_matches(m)
        ref = oracle_lib.remove_duplicated_matches(m)
        assert got.tobytes() == ref.tobytes()
        assert len(np.unique(got["train_idx"])) == len(got)
        assert np.all(np.diff(got["train_idx"]) > 0)


def test_dedup_proxy_matches_dmatch_sort(built):
    """The device-resident tracker removes duplicated matches on 8-byte (train, map) records; libstdc++'s
    std::sort must apply the same permutation to them as to the reference's 16-byte cv::DMatch array
    (feature_match.cpp:241-260: unstable sort on trainIdx, first of each run survives)."""
    import ctypes as C
    import mvo_b200
    lib = mvo_b200.load_library()
    lib.mvo_test_dedup_pairs.restype = C.c_int
    lib.mvo_test_dedup_pairs.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    rng = np.random.default_rng(11)
    for n in (0, 1, 2, 15, 16, 17, 33, 300, 1300, 2001, 5000):
        for spread in (2, 3, 10):
            m = np.zeros(n, mvo_b200.DMATCH_DTYPE)
            m["query_idx"] = np.arange(n)                      # = position in the match list before the sort
            m["train_idx"] = rng.integers(0, max(1, n // spread), n)
            ref = mvo_b200.remove_duplicated_matches(m)
            train = np.ascontiguousarray(m["train_idx"], np.int32).copy()
            tag = np.ascontiguousarray(m["query_idx"], np.int32).copy()
            cnt = C.c_int(n)
            assert lib.mvo_test_dedup_pairs(train.ctypes.data, tag.ctypes.data, C.byref(cnt)) == 0
            assert cnt.value == len(ref)
            assert np.array_equal(train[: cnt.value], ref["train_idx"]) and np.array_equal(tag[: cnt.value], ref["query_idx"])


def test_partition_restatement_equals_libstdcxx(tmp_path):
    """The formula k_match_filter uses for one introsort step against libstdc++'s own __unguarded_partition_pivot
    (tests/cpp/introsort_partition_check.cpp): identical arrays and cuts on 4000 random / adversarial sequences."""
    import subprocess
    exe = tmp_path / "introsort_check"
    subprocess.run(["g++", "-O2", "-std=c++17", str(ROOT / "tests" / "cpp" / "introsort_partition_check.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "mismatches 0" in r.stdout
    # organ-pipe inputs do leave the quicksort phase (heapsort in libstdc++): the case the kernel hands back to the host
    assert int(r.stdout.split("beyond_depth_limit")[1]) > 0
