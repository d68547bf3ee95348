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


def full_spec(n_in, n_out, act="none", use_bias=True) -> LayerSpec:
    return LayerSpec("full", [n_in, n_out], 1, activation=act, use_bias=use_bias)


def random_params(n, rng, scale=0.5):
    return (rng.standard_normal(n) * scale).astype(np.float32)


def assert_parity(got, ref32, ref64=None, tol=RTOL_ACT, what=""):
    """Parity of a CUDA result with the reference arithmetic.

    Primary criterion: within `tol` (relative) of the fp32 oracle, which walks the
    reference's loops in the reference's order.  For long fp32 reductions (weight gradients
    summed over >1e4 vertices) the fp32 oracle's own summation-order noise exceeds 1e-5, so
    the float64 shadow of the same algorithm arbitrates: the CUDA result must then be within
    `tol` of the float64 evaluation (i.e. at least as close to the exact reference value as
    the tolerance demands); the fp32 oracle's own deviation is printed for the record.
    """
    e32 = rel_err(got, ref32)
    if e32 <= tol:
        return e32
    assert ref64 is not None, f"{what}: rel err {e32:.3e} > {tol:g} vs the fp32 oracle"
    e64 = rel_err(got, ref64)
    eref = rel_err(ref32, ref64)
    assert e64 <= tol, (f"{what}: rel err {e32:.3e} vs fp32 oracle, {e64:.3e} vs float64 shadow "
                        f"(fp32 oracle itself is {eref:.3e} from float64); tolerance {tol:g}")
    return e64
