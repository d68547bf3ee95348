"""Shared test helpers (test infrastructure; may import the oracle)."""
import numpy as np

from athena_b200.graph import PackedGraphs
from oracle.oracle import Batch, LayerSpec

# Tolerances of BASELINE.json north_star:
RTOL_ACT = 1e-5     # forward activations and gradients, fp32, relative
RTOL_PARAM = 1e-4   # parameters after N training steps, relative


def rel_err(a, b) -> float:
    """max |a-b| relative to the largest reference magnitude (norm-wise relative error)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max() / scale)


def to_oracle_batch(p: PackedGraphs) -> Batch:
    return Batch(nv=p.nv, ne=p.ne, ia=p.ia, ja=p.ja, x=p.x, e=p.e)


def kipf_spec(nvf, T, act="none") -> LayerSpec:
    return LayerSpec("kipf", list(nvf), T, activation=act)


def duvenaud_spec(nvf, nef, T, min_deg, max_deg, n_out, act="sigmoid", ract="softmax") -> LayerSpec:
    return LayerSpec("duvenaud", list(nvf), T, nef, min_deg, max_deg, n_out, act, ract)


def random_params(n, rng, scale=0.5):
    return (rng.standard_normal(n) * scale).astype(np.float32)
