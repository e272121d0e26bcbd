"""GPU parity: all-pairs matching through the C ABI vs the oracle (bit-exact), plus
size-independent properties at the BASELINE size (2001 x 2001)."""
import numpy as np
import pytest

import mvo_synth
from oracle import oracle_lib

pytestmark = pytest.mark.gpu


def _eq(a, b):
    assert a.shape == b.shape
    for f in ("query_idx", "train_idx", "img_idx", "distance"):
        assert np.array_equal(a[f], b[f]), f


@pytest.mark.parametrize("n1,n2,dup", [(1, 2, 0), (1, 1, 0), (37, 53, 0), (128, 128, 4), (129, 31, 0),
                                        (300, 257, 5), (1000, 3000, 9), (2001, 2001, 11)])
def test_hamming_nn_and_knn2(ctx, n1, n2, dup):
    d1 = mvo_synth.random_descriptors(10 + n1, n1, dup_every=dup)
    d2 = mvo_synth.random_descriptors(20 + n2, n2, dup_every=dup)
    if dup:
        d1[::3] = d2[(np.arange(0, n1, 3) * 7) % n2]
    _eq(ctx.match_hamming_nn(d1, d2), oracle_lib.hamming_nn(d1, d2))
    if n2 >= 2:
        _eq(ctx.match_hamming_knn2(d1, d2), oracle_lib.hamming_knn2(d1, d2))


def test_extreme_descriptors(ctx):
    d1 = np.zeros((5, 32), np.uint8)
    d2 = np.full((7, 32), 255, np.uint8)
    d2[3] = 0
    r = ctx.match_hamming_nn(d1, d2)
    assert np.all(r["train_idx"] == 3) and np.all(r["distance"] == 0)
    k = ctx.match_hamming_knn2(d1, d2)
    assert np.all(k[:, 1]["train_idx"] == 0) and np.all(k[:, 1]["distance"] == 256)


@pytest.mark.parametrize("n1,n2,radius", [(50, 60, 50.0), (400, 380, 50.0), (2001, 2001, 50.0), (2001, 1999, 100.0), (100, 100, 0.0)])
def test_radius_sad(ctx, n1, n2, radius):
    rng = np.random.default_rng(n1 + n2)
    d1, d2 = mvo_synth.random_descriptors(1, n1, dup_every=6), mvo_synth.random_descriptors(2, n2, dup_every=4)
    # keypoint-like coordinates (multiples of the level scales), some exactly on the radius
    xy1 = (rng.integers(31, 600, (n1, 2)) * np.float32(1.2)).astype(np.float32)
    xy2 = (rng.integers(31, 440, (n2, 2)) * np.float32(1.0)).astype(np.float32)
    xy2[: min(n1, n2) // 4] = xy1[: min(n1, n2) // 4] + np.float32([radius, 0])
    _eq(ctx.match_radius_sad(d1, xy1, d2, xy2, radius), oracle_lib.match_radius_bf(xy1, xy2, d1, d2, radius))


@pytest.mark.parametrize("method", [1, 2, 3])
@pytest.mark.parametrize("n1,n2", [(0, 10), (10, 0), (3, 2), (700, 650), (2001, 2001)])
def test_match_features(ctx, method, n1, n2):
    rng = np.random.default_rng(method * 100 + n1)
    d2 = mvo_synth.random_descriptors(2, n2)
    d1 = mvo_synth.random_descriptors(1, n1)
    if n1 and n2:
        # make half the queries noisy copies of train rows so thresholds keep something
        src = rng.integers(0, n2, n1)
        noisy = d2[src] ^ (rng.random((n1, 32)) < 0.03).astype(np.uint8) * rng.integers(1, 255, (n1, 32), dtype=np.uint8)
        half = rng.random(n1) < 0.5
        d1[half] = noisy[half]
    xy1 = rng.uniform(31, 600, (n1, 2)).astype(np.float32)
    xy2 = rng.uniform(31, 600, (n2, 2)).astype(np.float32)
    if method == 2 and n2 < 2:
        import mvo_b200
        with pytest.raises(mvo_b200.MvoError):
            ctx.match_features(d1, d2, method, xy1, xy2, 100.0)
        return
    got = ctx.match_features(d1, d2, method, xy1, xy2, 100.0)
    ref = oracle_lib.match_features(d1, d2, method, xy1, xy2, 100.0)
    _eq(got, ref)


def test_wrong_method_index(ctx):
    import mvo_b200
    d = mvo_synth.random_descriptors(1, 8)
    with pytest.raises(mvo_b200.MvoError) as e:
        ctx.match_features(d, d, 4)
    assert e.value.code == -1 and "wrong method index" in str(e.value)


def test_properties_full_size(ctx):
    """Self-match: every descriptor's nearest neighbour in its own set is itself at distance 0;
    knn2's first column equals match(); swapping bytes consistently leaves distances unchanged."""
    d = mvo_synth.random_descriptors(7, 2001)
    nn = ctx.match_hamming_nn(d, d)
    assert np.array_equal(nn["train_idx"], np.arange(2001)) and np.all(nn["distance"] == 0)
    d2 = mvo_synth.random_descriptors(8, 2001)
    nn = ctx.match_hamming_nn(d, d2)
    kn = ctx.match_hamming_knn2(d, d2)
    assert kn[:, 0].tobytes() == nn.tobytes()
    assert np.all(kn[:, 0]["distance"] <= kn[:, 1]["distance"])
    perm = np.random.default_rng(0).permutation(32)
    nn_p = ctx.match_hamming_nn(d[:, perm], d2[:, perm])
    assert np.array_equal(nn_p["distance"], nn["distance"]) and np.array_equal(nn_p["train_idx"], nn["train_idx"])
