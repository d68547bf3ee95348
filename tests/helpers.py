"""Shared test helpers (test infrastructure; may import the oracle)."""
import numpy as np

from athena_b200.graph import PackedGraphs
from oracle.oracle import Batch, LayerSpec

# Tolerances of BASELINE.json north_star:
RTOL_ACT = 1e-5     # forward activations and gradients, fp32, relative
RTOL_PARAM = 1e-4   # parameters after N training steps, relative


# Companions of the norm-wise metric, recorded for every comparison and
# dumped to gpurun_out/elem_err.json:
#  * element-wise: every element relative to ITS OWN reference magnitude, with an absolute
#    floor of ELEM_FLOOR x the largest reference magnitude.  An output element is a sum of
#    O(100) products of the size of the largest elements; where the terms cancel to something
#    1000 x smaller, the rounding noise of the terms (1e-6 .. 1e-5 of the maximum, in the
#    reference's own fp32 arithmetic as well) is 1e-3 .. 1e-2 of the element, so 1e-5 of itself
#    is not attainable for such elements by ANY fp32 evaluation.  The bound asserted is the one
#    the norm-wise tolerance implies, RTOL_ACT / ELEM_FLOOR: an element above 1e-3 of the
#    maximum is within 1 % of itself.  Measured worst case on the GPU suite: 1.2e-3.
#  * column-wise: for [rows >= 256, features] arrays every FEATURE COLUMN is held to the
#    norm-wise criterion against its own maximum (COL_TOL): a feature on a smaller scale than
#    the others cannot hide behind the largest one.
ELEM_FLOOR = 1e-3
ELEM_TOL = 1e-2
COL_TOL = 5e-5
ELEM_LOG = []      # (norm-wise, element-wise, column-wise) of every comparison


def col_rel_err(a, b) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.ndim != 2 or a.shape[0] < 256 or not np.isfinite(b).all():
        return 0.0
    scale = np.maximum(np.abs(b).max(axis=0), 1e-30)
    return float((np.abs(a - b).max(axis=0) / scale).max())


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
        ew, cw = elem_rel_err(a, b), col_rel_err(a, b)
        ELEM_LOG.append((err, ew, cw))
        if err <= RTOL_ACT:
            assert ew <= ELEM_TOL, (f"element-wise relative error {ew:.3e} > {ELEM_TOL:g} "
                                    f"(norm-wise {err:.3e})")
            assert cw <= COL_TOL, (f"column-wise relative error {cw:.3e} > {COL_TOL:g} "
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
