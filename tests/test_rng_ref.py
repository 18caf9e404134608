"""The numpy restatement of Philox4x32-10 (oracle/philox_ref.py) against the Random123 known-answer
vectors, and basic properties of the uniform / normal mappings.  CPU only; the CUDA generator is
compared with this restatement in tests/test_gpu_rng.py."""
import numpy as np

from oracle import philox_ref as P


def test_philox_known_answers():
    for ctr, key, want in P.KNOWN_ANSWERS:
        got = P.philox4x32_10([np.uint32(c) for c in ctr], key)
        assert tuple(int(x) for x in got) == want, (ctr, key, [hex(int(x)) for x in got])


def test_streams_rows_and_groups_are_distinct_and_shift_invariant():
    a = P.uniforms(7, 1, 0, 64, 128)
    assert a.shape == (64, 128) and a.dtype == np.float32 and a.min() >= 0. and a.max() < 1.
    assert np.array_equal(P.uniforms(7, 1, 10, 20, 128), a[10:30]), "rows are keyed by the global ray index"
    assert not np.array_equal(P.uniforms(7, 0, 0, 64, 128), a) and not np.array_equal(P.uniforms(8, 1, 0, 64, 128), a)
    assert np.array_equal(P.uniforms(7, 1, 0, 64, 7), a[:, :7]), "a narrower row is a prefix"
    assert len(np.unique(a)) > 0.99 * a.size
    assert abs(a.mean() - .5) < 0.01 and abs(a.var() - 1 / 12) < 0.005


def test_normals_are_standard():
    z = P.normals(3, 2, 0, 4096, 64).astype(np.float64)
    assert np.isfinite(z).all()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.) < 0.01
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3.) < 0.1
