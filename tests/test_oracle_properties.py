"""CPU: hypothesis-driven properties of the oracle itself -- cheap sanity that the checker
checks the right thing: translation invariance, permutation equivariance, adjacency structure."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_oracle


def _state(n, seed):
    rs = np.random.RandomState(seed)
    x = np.zeros((n, 4))
    side = int(np.ceil(np.sqrt(n)))
    g = np.stack(np.meshgrid(np.arange(side), np.arange(side)), -1).reshape(-1, 2)[:n] * 0.6
    x[:, :2] = g + rs.uniform(-0.1, 0.1, size=(n, 2))
    x[:, 2:] = rs.uniform(-2, 2, size=(n, 2))
    return x


@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 40), seed=st.integers(0, 10 ** 6), dx=st.integers(-64, 64), dy=st.integers(-64, 64))
def test_translation_invariance_and_structure(n, seed, dx, dy):
    x = _state(n, seed)
    h = c_oracle.helpers(x)
    y = x.copy(); y[:, 0] += dx; y[:, 1] += dy                      # integers: small exact shifts
    h2 = c_oracle.helpers(y)
    assert np.array_equal(h["adj"], h2["adj"]) or np.abs(h["values"] - h2["values"]).max() < 1e-6
    assert np.allclose(h["values"], h2["values"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(h["adj"], h["adj"].T) and not np.diag(h["adj"]).any()
    rows = h["adj_mean"].sum(1)
    assert np.all((np.abs(rows - 1) < 1e-12) | (rows == 0))
    u = c_oracle.controller(x)
    assert np.all(np.abs(u) <= 1.0)
    assert c_oracle.reward(x) <= 0


@settings(max_examples=25, deadline=None)
@given(n=st.integers(8, 40), seed=st.integers(0, 10 ** 6))
def test_permutation_equivariance(n, seed):
    x = _state(n, seed)
    perm = np.random.RandomState(seed + 1).permutation(n)
    h, hp = c_oracle.helpers(x), c_oracle.helpers(x[perm])
    assert np.array_equal(hp["adj"], h["adj"][perm][:, perm])
    assert np.allclose(hp["values"], h["values"][perm], rtol=1e-9, atol=1e-9)
    assert np.allclose(c_oracle.controller(x[perm]), c_oracle.controller(x)[perm], rtol=0, atol=1e-9)
    assert np.allclose(c_oracle.knn_obs(x[perm]), c_oracle.knn_obs(x)[perm], rtol=0, atol=1e-12)
