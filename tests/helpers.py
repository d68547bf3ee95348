"""Shared test helpers (test infrastructure; may import the oracle)."""
import numpy as np

from athena_b200.graph import PackedGraphs
from oracle.oracle import Batch, LayerSpec

# Tolerances of BASELINE.json north_star:
RTOL_ACT = 1e-5     # forward activations and gradients, fp32, relative
RTOL_PARAM = 1e-4   # parameters after N training steps, relative


# Element-wise companion of the norm-wise metric: every element is compared relative to ITS OWN
# reference magnitude, with an absolute floor of ELEM_FLOOR x the largest reference magnitude
# (an element that is the cancelled sum of much larger terms carries the rounding noise of those
# terms in the reference's fp32 arithmetic too, so below the floor a relative error is
# meaningless).  Asserted at ELEM_TOL wherever rel_err is asserted at <= 1e-5.
ELEM_FLOOR = 1e-3
ELEM_TOL = 1e-4
ELEM_LOG = []      # (norm-wise, element-wise) of every comparison, dumped by conftest.py


def elem_rel_err(a, b, floor: float = ELEM_FLOOR) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    fin = np.isfinite(b)
    if not fin.all():
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[np.isinf(b)], b[np.isinf(b)])
        a, b = a[fin], b[fin]
        if a.size == 0:
            return 0.0
    scale = max(float(np.abs(b).max()), 1e-30)
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor * scale)).max())


def rel_err(a, b) -> float:
    """max |a-b| relative to the largest reference magnitude (norm-wise relative error).
    Every call also records the element-wise metric (elem_rel_err) and fails when a comparison
    that is within 1e-5 norm-wise is worse than ELEM_TOL element-wise."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = max(float(np.abs(b).max()), 1e-30)
    err = float(np.abs(a - b).max() / scale)
    if np.isfinite(err):
        ew = elem_rel_err(a, b)
        ELEM_LOG.append((err, ew))
        if err <= RTOL_ACT:
            assert ew <= ELEM_TOL, (f"element-wise relative error {ew:.3e} > {ELEM_TOL:g} "
                                    f"(norm-wise {err:.3e})")
    return err


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
