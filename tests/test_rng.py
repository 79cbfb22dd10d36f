"""Pins for ``qml_essentials_b200.rng`` (the ``jax.random`` restatement; SURVEY section 8(c),
VERDICT round 1 weak #1).  jax cannot be installed here, so these are the published
vectors that reach this path:

* the three Threefry-2x32 (20 rounds) known-answer vectors of Random123
  (``kat_vectors``; JAX's own ``random_test.py::testThreefry2x32`` asserts the same three);
* ``jax.random.split(PRNGKey(0))`` as printed in the JAX documentation for the legacy
  (``jax_threefry_partitionable=False``) layout - ``[[4146024105 967050713] [2718843009
  1272950319]]`` - which checks the block function on the counter pairing JAX uses;
* the partitionable layout (default of the pinned jax 0.9.0.1): child ``i`` of ``split`` is
  ``threefry(key, (0, i))``, so child 0 of ``split(key(0))`` is the first Random123 vector.
"""

import numpy as np

from qml_essentials_b200 import rng

KAT = [  # counter, key, expected (Random123 kat_vectors, threefry2x32 20)
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
    ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
    ((0x243F6A88, 0x85A308D3), (0x13198A2E, 0x03707344), (0xC4923A9C, 0x483DF7A0)),
]


def test_threefry2x32_random123_vectors():
    for ctr, key, want in KAT:
        assert rng._threefry_scalar(key[0], key[1], ctr[0], ctr[1]) == want
        a, b = rng.threefry2x32(np.uint32(key[0]), np.uint32(key[1]),
                                np.uint32([ctr[0]]), np.uint32([ctr[1]]))
        assert (int(a[0]), int(b[0])) == want


def test_vectorised_and_scalar_block_functions_agree():
    g = np.random.default_rng(0)
    k = g.integers(0, 2**32, (50, 2), dtype=np.uint64).astype(np.uint32)
    x = g.integers(0, 2**32, (50, 2), dtype=np.uint64).astype(np.uint32)
    a, b = rng.threefry2x32(k[:, 0], k[:, 1], x[:, 0], x[:, 1])
    for i in range(50):
        assert rng._threefry_scalar(int(k[i, 0]), int(k[i, 1]), int(x[i, 0]), int(x[i, 1])) == (
            int(a[i]), int(b[i]))


def test_legacy_split_of_key0_matches_the_jax_documentation():
    # legacy layout: counters arange(2 * num) are paired (i, i + num); outputs concatenated
    a0, b0 = rng._threefry_scalar(0, 0, 0, 2)
    a1, b1 = rng._threefry_scalar(0, 0, 1, 3)
    assert [[a0, a1], [b0, b1]] == [[4146024105, 967050713], [2718843009, 1272950319]]


def test_key_and_partitionable_split_layout():
    assert rng.key(42).data.tolist() == [0, 42]
    assert rng.key((7 << 32) | 5).data.tolist() == [7, 5]
    kids = rng.split(rng.key(0), 3).data
    assert kids.shape == (3, 2)
    assert kids[0].tolist() == [0x6B200159, 0x99BA4EFE]  # threefry(key, (0, 0)): KAT 1
    for i in range(3):
        assert tuple(kids[i].tolist()) == rng._threefry_scalar(0, 0, 0, i)
    # batched keys split element-wise, children on the leading axis like jax.vmap(split)
    many = rng.split(rng.key(1), 5)
    grand = rng.split(many, 2).data
    assert grand.shape == (2, 5, 2)
    for b in range(5):
        assert grand[:, b].tolist() == rng.split(many[b], 2).data.tolist()


def test_uniform_and_normal_derivations():
    k = rng.key(123)
    u = rng.uniform(k, (1000,))
    assert u.dtype == np.float64 and (u >= 0).all() and (u < 1).all()
    # mantissa fill: 52 random bits of word i -> u_i = bits / 2^52
    w = rng._bits64(k.data, 4)
    assert np.array_equal(rng.uniform(k, (4,)), (w >> np.uint64(12)).astype(np.float64) / 2.0**52)
    assert abs(u.mean() - 0.5) < 0.03
    z = rng.normal(k, (4000,))
    assert abs(z.mean()) < 0.06 and abs(z.std() - 1.0) < 0.05
    assert np.array_equal(rng.uniform(k, (7,), 2.0, 5.0), 2.0 + 3.0 * rng.uniform(k, (7,)))
    assert rng.choice_uniforms(rng.split(k, 3), 11).shape == (3, 11)
