"""Pin the oracle's PRNG on public JAX known-answer values and pin the product's
host-side key derivation on the oracle."""
import numpy as np
import pytest

from oracle import prng
from iactrace_b200 import random as R


def test_threefry_block_kats():
    # Random123 / JAX test-suite vectors
    assert [int(v) for v in prng.threefry2x32(0, 0, 0, 0)] == [0x6B200159, 0x99BA4EFE]
    assert [int(v) for v in prng.threefry2x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)] == [0x1CB996FC, 0xBB002BE7]
    assert [int(v) for v in prng.threefry2x32(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3)] == [0xC4923A9C, 0x483DF7A0]


def test_split_kats_both_modes():
    assert prng.split(prng.key(0), 2, prng.LEGACY).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert prng.split(prng.key(0), 2, prng.PARTITIONABLE).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


def test_normal_uniform_kats():
    assert prng.normal(prng.key(0), 1, prng.LEGACY)[0] == np.float32(-0.20584226)
    assert prng.normal(prng.key(42), 1, prng.LEGACY)[0] == np.float32(-0.18471177)
    assert prng.normal(prng.key(42), 1, prng.PARTITIONABLE)[0] == np.float32(-0.028304616)
    assert prng.normal(prng.key(0), 1, prng.PARTITIONABLE)[0] == np.float32(1.6226422)
    assert prng.uniform(prng.key(0), 1, mode=prng.LEGACY)[0] == np.float32(0.41845703)
    assert prng.uniform(prng.key(0), 1, mode=prng.PARTITIONABLE)[0] == np.float32(0.947667)


def test_key_and_uniform_properties():
    assert prng.key(42).tolist() == [0, 42]
    assert prng.key((7 << 32) | 9).tolist() == [7, 9]
    for mode in (prng.LEGACY, prng.PARTITIONABLE):
        u = prng.uniform(prng.key(3), 10001, mode=mode)
        assert u.dtype == np.float32 and 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.02
        n = prng.normal(prng.key(3), 20001, mode)
        assert abs(n.mean()) < 0.03 and abs(n.std() - 1) < 0.03
        # legacy odd-length padding: a prefix relation does NOT hold, but lengths do
        assert len(prng.bits(prng.key(1), 7, mode)) == 7
        c = prng.choice_p(prng.key(5), np.array([0.25, 0.25, 0.25, 0.25], np.float32), 4000, mode)
        assert set(np.unique(c)) == {0, 1, 2, 3} and abs((c == 0).mean() - 0.25) < 0.04


@pytest.mark.parametrize("mode", [R.PARTITIONABLE, R.LEGACY])
def test_product_key_derivation_matches_oracle(mode):
    for seed in (0, 42, 4242, (3 << 32) | 1):
        assert R.key(seed).tolist() == prng.key(seed).tolist()
        for n in (1, 2, 3, 380, 877):
            assert np.array_equal(R.split(R.key(seed), n, mode), prng.split(prng.key(seed), n, mode))
    assert R.as_key(None).tolist() == [0, 0]
    assert R.as_key(7).tolist() == [0, 7]
    assert R.as_key([1, 2]).tolist() == [1, 2]
    with pytest.raises(ValueError):
        R.as_key([1, 2, 3])


def test_choice_with_uneven_probabilities_is_distributed_as_p():
    """``jax.random.choice(key, n, shape, p=p)`` (sample_polygon, utils/sampling.py:53-58) is restated from the
    upstream source as recalled (cumsum, r = cum[-1] * (1 - U), searchsorted side='left'); no public known-answer
    vector exists for it offline, so the exact stream stays UNPINNED.  What can be pinned: with
    very uneven triangle areas the restatement draws each triangle with its probability (4-sigma binomial bound),
    in both threefry modes, and the resulting points are uniform over the polygon (first moments = centroid)."""
    from oracle import sample as osample
    verts = np.float32([[-0.45, -0.30], [0.10, -0.42], [0.48, 0.05], [0.05, 0.40], [-0.35, 0.22]])
    v0 = verts[0]
    tri_area = np.array([0.5 * abs((verts[i][0] - v0[0]) * (verts[i + 1][1] - v0[1]) - (verts[i + 1][0] - v0[0]) * (verts[i][1] - v0[1]))
                         for i in range(1, 4)])
    p = tri_area / tri_area.sum()
    assert p.max() / p.min() > 2.0
    n = 40000
    for mode in (prng.PARTITIONABLE, prng.LEGACY):
        idx = prng.choice_p(prng.key(5), p.astype(np.float32), n, mode)
        freq = np.bincount(idx, minlength=3) / n
        assert np.all(np.abs(freq - p) < 4.0 * np.sqrt(p * (1 - p) / n)), (freq, p)
        pts = osample.sample_polygon(prng.key(6), verts, n, mode)
        # polygon centroid = area-weighted mean of the fan triangles' centroids
        cen = sum(a * (v0 + verts[i] + verts[i + 1]) / 3.0 for a, i in zip(tri_area, range(1, 4))) / tri_area.sum()
        assert np.abs(pts.mean(0) - cen).max() < 4.0 * 0.3 / np.sqrt(n)
    # edge of the convention: U = 0 -> r = cum[-1] -> the last index whose cumulative sum reaches it (never out of range)
    cum = np.cumsum(p.astype(np.float32), dtype=np.float32)
    assert np.searchsorted(cum, cum[-1], side="left") == 2
