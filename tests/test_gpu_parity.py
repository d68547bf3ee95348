"""GPU parity tests proper: the CUDA path, called through the C ABI, against
the CPU oracle on the same seeded inputs.

  - integer structures (CSR / CSC / degrees / buckets / permutation): bit-exact
  - forward activations and gradients (fp32): <= 1e-5 relative
  - parameters after N training steps: <= 1e-4 relative
(BASELINE.json north_star).  Relative = max|a-b| / max|ref| (see helpers.rel_err).
"""
import os
import zlib

import numpy as np
import pytest

import athena_b200 as ab
from athena_b200 import synth
from helpers import (RTOL_ACT, RTOL_PARAM, assert_parity, duvenaud_spec, full_spec, kipf_spec,
                     random_params, rel_err, to_oracle_batch)
from oracle.oracle import Batch, OptimSpec

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------------------
# K1: batch build, bit-exact
# ---------------------------------------------------------------------------
def _check_batch(cuda, oracle32, p, buckets=((1, 10), (2, 5), (1, 1))):
    ref = oracle32.batch_build(p.nv, p.ne, p.ia, p.ja)
    gb = ab.GraphBatch(p)
    for k in ("row_ptr", "col", "eid", "deg", "vgraph", "csc_ptr", "csc_src", "csc_ent"):
        got = gb.export(k)
        assert np.array_equal(got, ref[k]), k
    # coefficient: float regime (powf vs 1/sqrtf), 1e-6 relative
    v_of = np.repeat(np.arange(p.V), ref["deg"])
    prod = ref["deg"][v_of].astype(np.float64) * ref["deg"][ref["col"]]
    coef = gb.export("coef")
    ok = prod > 0                       # 0 ** -0.5 = Infinity in the reference too
    assert np.all(np.isinf(coef[~ok]))
    assert rel_err(coef[ok], 1.0 / np.sqrt(prod[ok])) <= 1e-6
    for mn, mx in buckets:
        gb.bucketize(mn, mx)
        bkt, perm, ptr = oracle32.bucketize(ref["deg"], mn, mx)
        assert np.array_equal(gb.export("bucket"), bkt)
        assert np.array_equal(gb.export("perm"), perm)
        assert np.array_equal(gb.export("bucket_ptr"), ptr)
    gb.destroy()


def test_batch_build_reference_fixture_graphs(cuda, oracle32):
    # test/test_kipf_msgpass_layer.f90:83-90 (6 v / 8 e) and test_msgpass_network.f90:268-273 (5 v / 6 e)
    g1 = ab.graph_type(); g1.set_num_vertices(6, 3); g1.set_num_edges(8, 1)
    g1.generate_adjacency([[1, 2], [1, 3], [2, 3], [2, 4], [3, 5], [4, 5], [4, 6], [5, 6]])
    g2 = ab.graph_type(); g2.set_num_vertices(5, 3); g2.set_num_edges(6, 1)
    g2.generate_adjacency([[1, 2], [1, 3], [2, 3], [2, 4], [3, 5], [4, 5]])
    g2.add_self_loops()
    _check_batch(cuda, oracle32, ab.pack_graphs([g1, g2, g1]))


def test_batch_build_ragged_and_empty_graphs(cuda, oracle32):
    rng = np.random.default_rng(1)
    p = synth.molecular_batch(300, 4, 2, rng, nv_range=(1, 40))
    _check_batch(cuda, oracle32, p)
    # graphs with zero vertices at the ends and in the middle
    nv = np.array([0, 3, 0, 2, 0], np.int32)
    ne = np.array([0, 2, 0, 1, 0], np.int32)
    nz = np.array([0, 4, 0, 2, 0], np.int32)
    # graph 1: rows {2,3},{1},{1}; graph 3: rows {2},{1}; empty graphs contribute adj_ia = [1]
    ia = np.array([1, 1, 3, 4, 5, 1, 1, 2, 3, 1], np.int32)
    ja = np.array([[2, 1], [3, 2], [1, 1], [1, 2], [2, 1], [1, 1]], np.int32)
    _check_batch(cuda, oracle32, ab.PackedGraphs(nv, ne, nz, ia, ja))


def test_batch_build_directed_and_multi_edges(cuda, oracle32):
    rng = np.random.default_rng(2)
    nv = np.array([50, 70], np.int64)
    src = np.concatenate([rng.integers(0, 50, 400), 50 + rng.integers(0, 70, 900)])
    dst = np.concatenate([rng.integers(0, 50, 400), 50 + rng.integers(0, 70, 900)])
    p, _ = synth.packed_from_edges(nv, src, dst, directed=True, add_self_loops=False)
    _check_batch(cuda, oracle32, p)


def test_batch_build_power_law_long_columns(cuda, oracle32):
    rng = np.random.default_rng(3)
    p = synth.powerlaw_batch(3, 20000, 4, rng, max_degree=10000)
    deg = oracle32.batch_build(p.nv, p.ne, p.ia, p.ja)["deg"]
    assert deg.max() > 1000          # exercises the block-sort path for long CSC columns
    _check_batch(cuda, oracle32, p, buckets=((1, 10), (3, 64)))


def test_batch_build_one_huge_column(cuda, oracle32):
    # star graph: hub column longer than the shared-memory sort capacity (16384)
    n = 40000
    p, _ = synth.packed_from_edges(np.array([n]), np.zeros(n - 1, np.int64), np.arange(1, n))
    _check_batch(cuda, oracle32, p, buckets=((1, 4),))


def test_batch_build_rejects_bad_adjacency(cuda):
    nv = np.array([2], np.int32); ne = np.array([0], np.int32); nz = np.array([2], np.int32)
    ia = np.array([1, 2, 3], np.int32)
    ja = np.array([[1, 0], [3, 0]], np.int32)        # neighbour 3 > num_vertices
    with pytest.raises(ab.AthenaCudaError) as ei:
        ab.GraphBatch(ab.PackedGraphs(nv, ne, nz, ia, ja))
    assert ei.value.code == -4 and "greater than the number of vertices" in str(ei.value)


# ---------------------------------------------------------------------------
# Kipf layer: forward / backward
# ---------------------------------------------------------------------------
def test_kipf_identity_graph_known_answer(cuda):
    """test/test_diffstruc_extd_kipf.f90:23-45 through the layer with W = I."""
    g = ab.graph_type(); g.set_num_vertices(2, 2); g.set_num_edges(0, 0)
    g.adj_ia = np.array([1, 2, 3], np.int32); g.adj_ja = np.array([[1, 0], [2, 0]], np.int32)
    g.vertex_features = np.array([[1, 2], [3, 4]], np.float32)
    L = ab.kipf_msgpass_layer_type([2, 2], 1)
    L.set_params(np.eye(2, dtype=np.float32).ravel())
    L.set_graph([g])
    out = L.forward()
    assert np.abs(out - g.vertex_features).max() <= 1e-6
    gin = L.backward(np.ones((2, 2), np.float32), want_input_grad=True)
    assert np.abs(gin - 1.0).max() <= 1e-6


KIPF_CASES = [
    # nvf, T, act, generator
    ([5, 5], 1, "none", "mol"),
    ([3, 6, 5], 2, "relu", "mol"),
    ([7, 7, 7, 7], 3, "sigmoid", "mol"),
    ([8, 4], 1, "softmax", "mol"),
    ([16, 32, 8], 2, "tanh", "reg"),
    ([64, 64, 64], 2, "leaky_relu", "reg"),
    ([33, 65], 1, "relu", "reg"),
    ([128, 128], 1, "none", "reg"),
    ([130, 20], 1, "sigmoid", "mol"),
    # tensor-core (tcgen05) shapes: 32/64 wide, ragged vertex counts
    ([64, 64], 1, "relu", "mol"),
    ([64, 32, 64], 2, "sigmoid", "mol"),
    ([32, 32, 32], 2, "tanh", "mol"),
    ([32, 64], 1, "none", "reg"),
]


def _make(gen, F, rng, Fe=0):
    if gen == "mol":
        return synth.molecular_batch(37, F, Fe, rng, nv_range=(2, 30))
    return synth.regular_batch(24, 64, 6, F, rng, Fe=Fe, self_loop_features=bool(Fe))


@pytest.mark.parametrize("nvf,T,act,gen", KIPF_CASES)
def test_kipf_layer_forward_backward_parity(cuda, oracle32, oracle64, nvf, T, act, gen):
    rng = np.random.default_rng(zlib.crc32(repr((nvf, T, act)).encode()))
    p = _make(gen, nvf[0], rng)
    spec = kipf_spec(nvf, T, act)
    n = oracle32.num_params([spec])
    params = random_params(n, rng, 0.4)
    g_out = rng.standard_normal((p.V, nvf[-1])).astype(np.float32)
    ob = to_oracle_batch(p)
    out_ref, dp_ref, dx_ref = oracle32.layer_fwd_bwd(spec, params, ob, g_out, want_dx=True)
    out64, dp64, dx64 = oracle64.layer_fwd_bwd(spec, params, ob, g_out, want_dx=True)

    L = ab.kipf_msgpass_layer_type(nvf, T, activation=act)
    assert L.num_params == n
    L.set_params(params)
    L.set_graph(p)
    out = L.forward()
    assert rel_err(out, out_ref) <= RTOL_ACT
    L.zero_gradients()
    dx = L.backward(g_out, want_input_grad=True)
    dp = L.get_gradients()
    assert rel_err(dp, dp_ref) <= RTOL_ACT
    assert rel_err(dx, dx_ref) <= RTOL_ACT
    # the CUDA path is as close to exact arithmetic as the fp32 oracle is
    assert rel_err(out, out64) <= 4 * max(rel_err(out_ref, out64), 5e-7)
    assert rel_err(dp, dp64) <= 4 * max(rel_err(dp_ref, dp64), 5e-7)
    # gradients accumulate (params(t)%grad +=): second backward doubles them
    L.backward(g_out)
    assert rel_err(L.get_gradients(), 2 * dp) <= 1e-6
    # get/set round trip (test/test_kipf_msgpass_layer.f90:115-146)
    L.set_gradients(np.full(n, 0.1, np.float32))
    assert np.allclose(L.get_gradients(), 0.1)
    assert np.array_equal(L.get_params(), params)


def test_kipf_tensor_core_path_many_tiles(cuda, oracle32, oracle64):
    """More 128-row tiles than resident CTAs: exercises the persistent loops, the mbarrier
    phase flips and the double-buffered dW kernel of the tcgen05 path.  tanh (smooth) so that
    the comparison is not dominated by relu'(0) sign flips; dW is a 64000-term fp32 reduction,
    so the float64 shadow arbitrates (helpers.assert_parity)."""
    rng = np.random.default_rng(77)
    p = synth.regular_batch(1000, 64, 6, 64, rng)          # V = 64000 = 500 tiles of 128
    spec = kipf_spec([64, 64, 64], 2, "tanh")
    params = random_params(oracle32.num_params([spec]), rng, 0.2)
    g_out = rng.standard_normal((p.V, 64)).astype(np.float32)
    ob = to_oracle_batch(p)
    out_ref, dp_ref, dx_ref = oracle32.layer_fwd_bwd(spec, params, ob, g_out, want_dx=True)
    out64, dp64, dx64 = oracle64.layer_fwd_bwd(spec, params, ob, g_out, want_dx=True)
    L = ab.kipf_msgpass_layer_type([64, 64, 64], 2, activation="tanh")
    L.set_params(params)
    L.set_graph(p)
    out = L.forward()
    assert_parity(out, out_ref, out64, what="out")
    L.zero_gradients()
    dx = L.backward(g_out, want_input_grad=True)
    dp = L.get_gradients()
    assert_parity(dp, dp_ref, dp64, what="dW")
    assert_parity(dx, dx_ref, dx64, what="dx")
    # bitwise run-to-run determinism (no float atomics anywhere)
    out2 = L.forward()
    L.zero_gradients()
    dx2 = L.backward(g_out, want_input_grad=True)
    assert np.array_equal(out, out2) and np.array_equal(dx, dx2)
    assert np.array_equal(dp, L.get_gradients())


def test_kipf_forward_shape_matches_reference_test(cuda):
    """test/test_kipf_msgpass_layer.f90:148-165: 6 vertices, [5 -> 5], params = 1."""
    g = ab.graph_type(); g.set_num_vertices(6, 5); g.set_num_edges(8, 1)
    g.vertex_features[:] = 1.0
    g.generate_adjacency([[1, 2], [1, 3], [2, 3], [2, 4], [3, 5], [4, 5], [4, 6], [5, 6]])
    L = ab.kipf_msgpass_layer_type([5, 5], 1)
    L.set_params(np.ones(L.num_params, np.float32))
    L.set_graph([g])
    assert L.forward().shape == (6, 5)


# ---------------------------------------------------------------------------
# Duvenaud layer
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["chem", "mindeg2"])
def test_duvenaud_matches_reference_python_golden(cuda, name):
    d = np.load(os.path.join(GOLD, f"duvenaud_ref_{name}.npz"))
    fv, fe, T, no, mn, mx = [int(v) for v in d["hyper"]]
    nv, ne = d["nv"], d["ne"]
    ia, ja = d["ia"], d["ja"].reshape(-1, 2)
    nz = np.array([ia[o + n] - 1 for o, n in zip(np.cumsum(np.r_[0, nv[:-1] + 1]), nv)], np.int32)
    p = ab.PackedGraphs(nv, ne, nz, ia, ja, d["x"], d["e"])
    L = ab.duvenaud_msgpass_layer_type([fv], [fe], T, mx, no, min_vertex_degree=mn)
    L.set_params(d["params"])
    L.set_graph(p)
    out = L.forward()
    assert rel_err(out, d["out"]) <= RTOL_ACT
    L.zero_gradients()
    dx = L.backward(d["g_out"], want_input_grad=True)
    assert rel_err(L.get_gradients(), d["dparams"]) <= RTOL_ACT
    assert rel_err(dx, d["dx"]) <= RTOL_ACT


DUV_CASES = [
    # nvf, nef, T, min, max, n_out, act, ract, gen
    ([6], 1, 4, 1, 10, 10, "sigmoid", "softmax", "mol"),
    ([4, 8, 4], 2, 2, 2, 4, 3, "sigmoid", "softmax", "mol"),
    ([8], 1, 2, 1, 4, 3, "sigmoid", "linear", "mol"),       # test_msgpass_network.f90:97-107
    ([32], 4, 2, 1, 6, 32, "relu", "softmax", "mol"),        # cfg4 dims
    ([64], 3, 1, 10, 14, 16, "tanh", "softmax", "reg"),      # all vertices in bucket 13-10
    ([5], 0, 2, 1, 3, 4, "sigmoid", "softmax", "mol"),       # no edge features
]


@pytest.mark.parametrize("nvf,nef,T,mn,mx,no,act,ract,gen", DUV_CASES)
def test_duvenaud_layer_forward_backward_parity(cuda, oracle32, nvf, nef, T, mn, mx, no, act, ract,
                                                gen):
    rng = np.random.default_rng(zlib.crc32(repr((nvf, nef, T, mn, mx, no)).encode()))
    p = _make(gen, nvf[0], rng, Fe=nef)
    full = nvf * (T + 1) if len(nvf) == 1 else nvf
    spec = duvenaud_spec(full, nef, T, mn, mx, no, act, ract)
    n = oracle32.num_params([spec])
    params = random_params(n, rng, 0.3)
    g_out = rng.standard_normal((p.B, no)).astype(np.float32)
    out_ref, dp_ref, dx_ref = oracle32.layer_fwd_bwd(spec, params, to_oracle_batch(p), g_out,
                                                     want_dx=True)
    L = ab.duvenaud_msgpass_layer_type(nvf, [nef], T, mx, no, min_vertex_degree=mn,
                                       message_activation=act, readout_activation=ract)
    assert L.num_params == n
    L.set_params(params)
    L.set_graph(p)
    out = L.forward()
    assert out.shape == (p.B, no)
    assert rel_err(out, out_ref) <= RTOL_ACT
    L.zero_gradients()
    dx = L.backward(g_out, want_input_grad=True)
    assert rel_err(L.get_gradients(), dp_ref) <= RTOL_ACT
    assert rel_err(dx, dx_ref) <= RTOL_ACT


def test_duvenaud_self_loop_without_edge_feature(cuda, oracle32):
    """adj_ja(2,w) = 0 (add_self_loops marker) contributes a zero edge feature."""
    rng = np.random.default_rng(11)
    p = synth.chemical_batch(8, rng)
    assert (p.ja[:, 1] == 0).sum() == p.V
    spec = duvenaud_spec([6] * 5, 1, 4, 1, 10, 10)
    params = random_params(oracle32.num_params([spec]), rng, 0.3)
    out_ref, _, _ = oracle32.layer_fwd_bwd(spec, params, to_oracle_batch(p))
    L = ab.duvenaud_msgpass_layer_type([6], [1], 4, 10, 10)
    L.set_params(params)
    L.set_graph(p)
    assert rel_err(L.forward(), out_ref) <= RTOL_ACT


# ---------------------------------------------------------------------------
# network: loss, gradients, optimiser steps
# ---------------------------------------------------------------------------
def _train_compare(cuda, oracle32, specs, layers, p, target, optim: OptimSpec, ab_opt, steps=5,
                   input_lists=None, oracle64=None):
    net = ab.network_type()
    for k, L in enumerate(layers):
        if input_lists is not None and input_lists[k] is not None:
            net.add(L, input_list=input_lists[k], operator="concatenate")
        else:
            net.add(L)
    net.compile(ab_opt, loss_method="mse", batch_size=p.B)
    n = oracle32.num_params(specs)
    assert net.num_params == n
    rng = np.random.default_rng(123)
    params = random_params(n, rng, 0.3)
    net.set_params(params)
    batch = ab.GraphBatch(p)
    # loss + gradients of the first step
    ob = to_oracle_batch(p)
    loss_ref, out_ref, g_ref = oracle32.stack_fwd_bwd(specs, params, ob, target)
    loss = net.loss_and_gradients(batch, target)
    assert abs(loss - loss_ref) <= 1e-5 * max(1.0, abs(loss_ref))
    if oracle64 is None:
        assert rel_err(net.get_gradients(), g_ref) <= RTOL_ACT
    else:   # long reductions: the float64 shadow arbitrates (helpers.assert_parity)
        _, _, g64 = oracle64.stack_fwd_bwd(specs, params, ob, target)
        assert_parity(net.get_gradients(), g_ref, g64, RTOL_ACT, "network gradients")
    assert rel_err(net.forward(batch), out_ref) <= RTOL_ACT
    net.update()                                   # consumes those gradients (step 1)
    ref = params.copy()
    s1 = np.zeros(n, np.float32); s2 = np.zeros(n, np.float32)
    l0, _ = oracle32.train_step(specs, ref, ob, target, optim, s1, s2, 1)
    assert rel_err(net.get_params(), ref) <= RTOL_PARAM
    assert np.all(net.get_gradients() == 0)        # reset_gradients
    losses = []
    for it in range(2, steps + 1):
        lr_, _ = oracle32.train_step(specs, ref, ob, target, optim, s1, s2, it)
        l = net.train_step(batch, target)
        losses.append((l, lr_))
    for l, lr_ in losses:
        assert abs(l - lr_) <= 1e-4 * max(1.0, abs(lr_))
    assert rel_err(net.get_params(), ref) <= RTOL_PARAM
    net.destroy()


def test_kipf_network_sgd_training_parity(cuda, oracle32):
    """Shape of test/test_msgpass_network.f90:40-88: Kipf net, SGD lr 0.01, graph target."""
    rng = np.random.default_rng(5)
    p = synth.molecular_batch(9, 8, 0, rng, nv_range=(3, 9), self_loop_features=False)
    spec = kipf_spec([8, 8, 8], 2)
    target = p.x.copy()                            # train(graph, graph): the target is the input graph
    _train_compare(cuda, oracle32, [spec], [ab.kipf_msgpass_layer_type([8, 8, 8], 2)], p, target,
                   OptimSpec("sgd", lr=0.01), ab.sgd_optimiser_type(learning_rate=0.01))


def test_kipf_two_layer_relu_momentum(cuda, oracle32):
    rng = np.random.default_rng(6)
    p = synth.regular_batch(16, 64, 6, 64, rng)
    specs = [kipf_spec([64, 64], 1, "relu"), kipf_spec([64, 64], 1, "none")]
    layers = [ab.kipf_msgpass_layer_type([64, 64], 1, "relu"),
              ab.kipf_msgpass_layer_type([64, 64], 1, "none")]
    target = rng.standard_normal((p.V, 64)).astype(np.float32)
    _train_compare(cuda, oracle32, specs, layers, p, target,
                   OptimSpec("sgd", lr=0.05, momentum=0.9, nesterov=True),
                   ab.sgd_optimiser_type(0.05, momentum=0.9, nesterov=True))


def test_duvenaud_network_adam_clip_training_parity(cuda, oracle32):
    """example/msgpass_chemical dims (main.f90:129-192): T=4, Fv=6, Fe=1, D=10, n_out=10,
    Adam lr 1e-2, clip_norm 0.1, batch 8."""
    rng = np.random.default_rng(7)
    p = synth.chemical_batch(8, rng)
    spec = duvenaud_spec([6] * 5, 1, 4, 1, 10, 10)
    target = rng.random((8, 10)).astype(np.float32)
    _train_compare(cuda, oracle32, [spec], [ab.duvenaud_msgpass_layer_type([6], [1], 4, 10, 10)],
                   p, target, OptimSpec("adam", lr=1e-2, clip_norm=0.1),
                   ab.adam_optimiser_type(1e-2, clip_dict=ab.clip_type(clip_norm=0.1)), steps=8)


def test_kipf_duvenaud_stack_training_parity(cuda, oracle32):
    """cfg4 wiring: Kipf(32->32) x2 -> Duvenaud(T=2, D=6, n_out=32), Adam, min/max clip."""
    rng = np.random.default_rng(8)
    p = synth.molecular_batch(40, 32, 4, rng)
    specs = [kipf_spec([32, 32], 1, "relu"), kipf_spec([32, 32], 1, "relu"),
             duvenaud_spec([32] * 3, 4, 2, 1, 6, 32)]
    layers = [ab.kipf_msgpass_layer_type([32, 32], 1, "relu"),
              ab.kipf_msgpass_layer_type([32, 32], 1, "relu"),
              ab.duvenaud_msgpass_layer_type([32], [4], 2, 6, 32)]
    target = rng.random((40, 32)).astype(np.float32)
    _train_compare(cuda, oracle32, specs, layers, p, target,
                   OptimSpec("adam", lr=5e-3, clip_min=-0.05, clip_max=0.05),
                   ab.adam_optimiser_type(5e-3, clip_dict=ab.clip_type(-0.05, 0.05)))


@pytest.mark.parametrize("hidden2", [64, 128])
def test_chemical_example_network_training_parity(cuda, oracle32, hidden2):
    """The whole network of example/msgpass_chemical (main.f90:129-192): Duvenaud(T=4, Fv=6,
    Fe=1, D=10, n_out=10) -> full 10->128 -> 64 -> 1, leaky_relu, Adam lr 1e-2, clip_norm 0.1,
    batch 8.  The full layers take num_inputs from the previous layer, as in the example.
    (11 649 parameters: clipping and the step ride on the finalize launch; hidden2 = 128 makes it
    19 969, past that launch's limit: clip partial sums there, the step in a launch of its own.)"""
    rng = np.random.default_rng(17)
    p = synth.chemical_batch(8, rng)
    specs = [duvenaud_spec([6] * 5, 1, 4, 1, 10, 10), full_spec(10, 128, "leaky_relu"),
             full_spec(128, hidden2, "leaky_relu"), full_spec(hidden2, 1, "leaky_relu")]
    layers = [ab.duvenaud_msgpass_layer_type([6], [1], 4, 10, 10),
              ab.full_layer_type(128, activation="leaky_relu"),
              ab.full_layer_type(hidden2, activation="leaky_relu"),
              ab.full_layer_type(1, activation="leaky_relu")]
    target = rng.random((8, 1)).astype(np.float32)
    _train_compare(cuda, oracle32, specs, layers, p, target,
                   OptimSpec("adam", lr=1e-2, clip_norm=0.1),
                   ab.adam_optimiser_type(1e-2, clip_dict=ab.clip_type(clip_norm=0.1)), steps=8)


@pytest.mark.parametrize("act,bias", [("none", True), ("tanh", False), ("softmax", True),
                                      ("sigmoid", True)])
def test_full_layer_forward_backward_parity(cuda, oracle32, act, bias):
    """full_layer_type forward / reverse sweep on its own (athena_full_layer.f90:839-874):
    y = act(W x + b), dW = gz x^T, db = sum gz, dx = W^T gz."""
    rng = np.random.default_rng(zlib.crc32(f"{act}{bias}".encode()))
    p = synth.chemical_batch(37, rng)
    spec = full_spec(19, 23, act, bias)
    L = ab.full_layer_type(23, 19, use_bias=bias, activation=act)
    n = oracle32.num_params([spec])
    assert L.num_params == n == 19 * 23 + (23 if bias else 0)
    params = random_params(n, rng, 0.4)
    L.set_params(params)
    L.set_graph(p)
    x = rng.standard_normal((37, 19)).astype(np.float32)
    g = rng.standard_normal((37, 23)).astype(np.float32)
    ob = to_oracle_batch(p)
    out_ref, grad_ref, dx_ref = oracle32.layer_fwd_bwd(spec, params, ob, g_out=g, want_dx=True,
                                                       x=x)
    assert rel_err(L.forward(x), out_ref) <= RTOL_ACT
    L.zero_gradients()
    dx = L.backward(g, want_input_grad=True)
    assert rel_err(L.get_gradients(), grad_ref) <= RTOL_ACT
    assert rel_err(dx, dx_ref) <= RTOL_ACT
    L.destroy()


def test_full_layer_placement_rules(cuda):
    """A full layer consumes a graph-level array: it may only follow a Duvenaud or full layer,
    and nothing but full layers may follow a graph-level output."""
    net = ab.network_type()
    with pytest.raises(ab.AthenaCudaError):
        net.add(ab.full_layer_type(4, 4))
    net.add(ab.kipf_msgpass_layer_type([4, 4], 1))
    with pytest.raises(ab.AthenaCudaError):
        net.add(ab.full_layer_type(4, 4))
    net.add(ab.duvenaud_msgpass_layer_type([4], [0], 1, 3, 5))
    with pytest.raises(ab.AthenaCudaError):
        net.add(ab.full_layer_type(4, 7))          # width mismatch
    with pytest.raises(ab.AthenaCudaError):
        net.add(ab.kipf_msgpass_layer_type([5, 5], 1))
    net.add(ab.full_layer_type(4))
    assert net.model[-1].num_inputs == 5
    net.destroy()


@pytest.mark.parametrize("kind", ["rmsprop", "rmsprop_beta", "adagrad"])
def test_rmsprop_adagrad_training_parity(cuda, oracle32, kind):
    """minimise_rmsprop / minimise_adagrad on the flat parameter vector
    (athena_optimiser.f90:771-803, 898-925), fused Kipf steps underneath, with clipping."""
    rng = np.random.default_rng(31)
    p = synth.regular_batch(12, 64, 4, 64, rng)
    specs = [kipf_spec([64, 64], 1, "relu"), kipf_spec([64, 64], 1, "none")]
    layers = [ab.kipf_msgpass_layer_type([64, 64], 1, "relu"),
              ab.kipf_msgpass_layer_type([64, 64], 1, "none")]
    target = rng.standard_normal((p.V, 64)).astype(np.float32)
    clip = ab.clip_type(clip_norm=5.0)
    if kind == "adagrad":
        spec_o, opt = OptimSpec("adagrad", lr=0.01, clip_norm=5.0), ab.adagrad_optimiser_type(0.01, clip_dict=clip)
    else:
        beta = 0.9 if kind == "rmsprop_beta" else 0.0
        spec_o = OptimSpec("rmsprop", lr=0.002, beta1=beta, clip_norm=5.0)
        opt = ab.rmsprop_optimiser_type(0.002, beta=beta, clip_dict=clip)
    _train_compare(cuda, oracle32, specs, layers, p, target, spec_o, opt)


@pytest.mark.parametrize("case", ["sgd_l1", "rmsprop_l1l2", "adamw", "adam_l2"])
def test_regulariser_training_parity(cuda, oracle32, case):
    """regulariser%regularise in front of the step and the AdamW / L2 branches of
    minimise_adam (athena_regulariser.f90:85-137, athena_optimiser.f90:1064-1078)."""
    rng = np.random.default_rng(41)
    p = synth.molecular_batch(24, 32, 0, rng)
    specs = [kipf_spec([32, 32], 1, "tanh"), kipf_spec([32, 32], 1, "none")]
    layers = [ab.kipf_msgpass_layer_type([32, 32], 1, "tanh"),
              ab.kipf_msgpass_layer_type([32, 32], 1, "none")]
    target = rng.standard_normal((p.V, 32)).astype(np.float32)
    if case == "sgd_l1":
        so = OptimSpec("sgd", lr=0.05, momentum=0.9, regulariser="l1", l1=0.02)
        opt = ab.sgd_optimiser_type(0.05, momentum=0.9, regulariser=ab.l1_regulariser_type(0.02))
    elif case == "rmsprop_l1l2":
        so = OptimSpec("rmsprop", lr=0.002, beta1=0.9, regulariser="l1l2", l1=0.01, l2=0.03)
        opt = ab.rmsprop_optimiser_type(0.002, beta=0.9,
                                        regulariser=ab.l1l2_regulariser_type(0.01, 0.03))
    else:
        dec = case == "adamw"
        so = OptimSpec("adam", lr=0.01, regulariser="l2", l2=0.05, l2_decoupled=dec)
        opt = ab.adam_optimiser_type(0.01, regulariser=ab.l2_regulariser_type(0.05, decoupled=dec))
    _train_compare(cuda, oracle32, specs, layers, p, target, so, opt)


def test_network_train_loop_runs_epochs(cuda):
    """network%train batch loop with a ragged last batch (athena_network_sub.f90:3575-3670)."""
    rng = np.random.default_rng(9)
    p = synth.chemical_batch(19, rng)
    net = ab.network_type()
    net.add(ab.duvenaud_msgpass_layer_type([6], [1], 2, 10, 4))
    net.compile(ab.adam_optimiser_type(1e-2), batch_size=8)
    target = rng.random((19, 4)).astype(np.float32)
    hist = net.train(p, target, num_epochs=6, shuffle_batches=False)
    assert len(hist) == 6 and hist[-1] < hist[0] and np.isfinite(hist).all()


def test_shard_sum_equals_full_batch_gradient(cuda, oracle32):
    """The N>1 decomposition on one GPU: gradients of contiguous graph shards (with the
    global-batch normalisation) sum to the full-batch gradient."""
    rng = np.random.default_rng(10)
    p = synth.molecular_batch(64, 32, 4, rng)
    target = rng.random((64, 8)).astype(np.float32)
    full = ab.GraphBatch(p)
    first = np.zeros(5, np.int32)
    nz64 = p.nz.astype(np.int64)  # named: the buffer must outlive the call
    ab.check(ab.lib().athena_cuda_shard_graphs(64, ab.ptr(nz64), 4, ab.ptr(first)))
    net = ab.network_type()
    for L in (ab.kipf_msgpass_layer_type([32, 32], 1, "relu"),
              ab.duvenaud_msgpass_layer_type([32], [4], 2, 6, 8)):
        net.add(L)
    net.compile(ab.sgd_optimiser_type(0.0), batch_size=64)   # lr = 0: update only zeroes grads
    params = random_params(net.num_params, rng, 0.3)
    net.set_params(params)
    g_sum = np.zeros(net.num_params, np.float32)
    loss_sum = 0.0
    loss_full = net.loss_and_gradients(full, target)
    g_full = net.get_gradients().copy()
    net.update()
    for r in range(4):
        sh = ab.GraphBatch(p.slice(first[r], first[r + 1]))
        loss_sum += net.loss_and_gradients(sh, target[first[r]:first[r + 1]], global_batch=64)
        g_sum += net.get_gradients()
        net.update()
    assert abs(loss_sum - loss_full) <= 1e-5 * abs(loss_full)
    assert rel_err(g_sum, g_full) <= RTOL_ACT


def test_kipf_fused_training_multi_edge_batch_uses_list_gather(cuda, oracle32):
    """Repeated (row, column) pairs cannot be adjacency BITS: such a batch must take the
    list-gather kernels (k_pipe_gather) and still match the reference, which sums a repeated
    neighbour as often as it is listed (athena_diffstruc_extd_sub_kipf.f90:36-44)."""
    rng = np.random.default_rng(64)
    nv = rng.integers(20, 70, 24).astype(np.int64)
    voff = np.concatenate([[0], np.cumsum(nv)])
    src, dst = [], []
    for g, n in enumerate(nv):
        m = 3 * int(n)
        a = voff[g] + rng.integers(0, n, m)
        b = voff[g] + rng.integers(0, n, m)
        src += [a, a[: m // 4]]                    # a quarter of the edges is listed twice
        dst += [b, b[: m // 4]]
    p, _ = synth.packed_from_edges(nv, np.concatenate(src), np.concatenate(dst))
    ref = oracle32.batch_build(p.nv, p.ne, p.ia, p.ja)
    rows = np.repeat(np.arange(p.V), ref["deg"])
    pairs = rows.astype(np.int64) * p.V + ref["col"]
    assert np.unique(pairs).size < pairs.size      # the batch really has multi-edges
    p.x = rng.standard_normal((p.V, 64)).astype(np.float32)
    specs = [kipf_spec([64, 64], 1, "relu"), kipf_spec([64, 64], 1, "none")]
    layers = [ab.kipf_msgpass_layer_type([64, 64], 1, "relu"),
              ab.kipf_msgpass_layer_type([64, 64], 1, "none")]
    target = rng.standard_normal((p.V, 64)).astype(np.float32)
    _train_compare(cuda, oracle32, specs, layers, p, target, OptimSpec("sgd", lr=0.05),
                   ab.sgd_optimiser_type(0.05))


def test_kipf_fused_directed_graph_with_sinks(cuda, oracle32):
    """Directed graphs with vertices that have no outgoing entry (degree 0, deg^-1/2 =
    Infinity as in the reference) but are somebody's neighbour: only the rows that list such a
    vertex may become non-finite; every other row must match.  (In a dense adjacency product
    0 * Infinity would poison the whole tile: these batches take the list kernels.)"""
    rng = np.random.default_rng(77)
    nv = np.array([60, 68, 50], np.int64)
    voff = np.concatenate([[0], np.cumsum(nv)])
    src, dst = [], []
    for g, n in enumerate(nv):
        n = int(n)
        s_ = np.repeat(np.arange(n - 2), 3)              # the last two vertices are sinks
        d_ = rng.integers(0, n - 2, s_.size)
        d_[rng.random(s_.size) < 0.02] = n - 1           # a few edges point at a sink
        keep = np.unique(s_ * n + d_, return_index=True)[1]
        src.append(voff[g] + s_[keep]); dst.append(voff[g] + d_[keep])
    p, _ = synth.packed_from_edges(nv, np.concatenate(src), np.concatenate(dst), directed=True,
                                   add_self_loops=False)
    p.x = rng.standard_normal((p.V, 64)).astype(np.float32)
    spec = kipf_spec([64, 64], 1, "none")
    params = random_params(oracle32.num_params([spec]), rng, 0.3)
    with np.errstate(all="ignore"):
        out_ref, _, _ = oracle32.layer_fwd_bwd(spec, params, to_oracle_batch(p))
    L = ab.kipf_msgpass_layer_type([64, 64], 1)
    L.set_params(params)
    L.set_graph(p)
    out = L.forward()
    finite = np.isfinite(out_ref).all(axis=1)
    assert 0 < (~finite).sum() < p.V // 2                # some rows list a sink, most do not
    assert np.isfinite(out[finite]).all()
    assert rel_err(out[finite], out_ref[finite]) <= RTOL_ACT
    assert not np.isfinite(out[~finite]).all(axis=1).any()
    L.destroy()


@pytest.mark.parametrize("act", ["none", "relu", "tanh"])
def test_kipf_large_graph_width_128_fused_path(cuda, oracle32, oracle64, act):
    """One graph far larger than a 128-row tile, F = 128: the fused SpMM + tcgen05 kernel of
    agg_tc.cu (forward, aggregate saved) followed by the generic reverse sweep; and the
    inference path of network%predict, which does not save the aggregate."""
    rng = np.random.default_rng(128)
    p = synth.random_graph(3001, 5, 128, rng)            # last tile is partial (3001 % 128 = 57)
    spec = kipf_spec([128, 128, 128], 2, act)
    params = random_params(oracle32.num_params([spec]), rng, 0.1)
    g_out = rng.standard_normal((p.V, 128)).astype(np.float32)
    ob = to_oracle_batch(p)
    r32 = oracle32.layer_fwd_bwd(spec, params, ob, g_out, want_dx=True)
    r64 = oracle64.layer_fwd_bwd(spec, params, ob, g_out, want_dx=True)
    L = ab.kipf_msgpass_layer_type([128, 128, 128], 2, activation=act)
    L.set_params(params)
    L.set_graph(p)
    out = L.forward()
    assert_parity(out, r32[0], r64[0], what="out")
    L.zero_gradients()
    dx = L.backward(g_out, want_input_grad=True)
    assert_parity(L.get_gradients(), r32[1], r64[1], what="dW")
    assert_parity(dx, r32[2], r64[2], what="dx")
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([128, 128, 128], 2, activation=act))
    net.compile(ab.sgd_optimiser_type(0.01), batch_size=1)
    net.set_params(params)
    assert_parity(net.predict(p), r32[0], r64[0], what="predict")


@pytest.mark.parametrize("act2,optim", [("none", "sgd"), ("tanh", "adam_clip"), ("leaky_relu", "sgd")])
def test_kipf_fused_training_ragged_graphs(cuda, oracle32, act2, optim):
    """The fused cfg2 path end to end on RAGGED graphs: partial 128-row tiles, the MSE fused
    into the last layer's epilogue (with and without an activation to fold), sign-bit and
    saved-activation variants of the backward epilogue, and both the one-launch finalise+step
    (SGD) and the clip + step path (Adam with clip_norm)."""
    rng = np.random.default_rng(31)
    p = synth.molecular_batch(300, 64, 0, rng, nv_range=(2, 50))
    specs = [kipf_spec([64, 64], 1, "relu"), kipf_spec([64, 64, 64], 2, act2)]
    layers = [ab.kipf_msgpass_layer_type([64, 64], 1, "relu"),
              ab.kipf_msgpass_layer_type([64, 64, 64], 2, act2)]
    target = rng.standard_normal((p.V, 64)).astype(np.float32) * 0.5
    if optim == "sgd":
        o, a = OptimSpec("sgd", lr=0.02, momentum=0.5), ab.sgd_optimiser_type(0.02, momentum=0.5)
    else:
        o = OptimSpec("adam", lr=2e-3, clip_norm=0.5)
        a = ab.adam_optimiser_type(2e-3, clip_dict=ab.clip_type(clip_norm=0.5))
    _train_compare(cuda, oracle32, specs, layers, p, target, o, a, steps=4)


# ---------------------------------------------------------------------------
# f2: skip connections (network%add(..., input_list, operator='concatenate')) and swish
# ---------------------------------------------------------------------------
def _euler_graphs():
    """The bump-channel mesh of example/msgpass_euler as mod_read_euler.f90:14-53 builds it:
    generate_adjacency(index_list) and NO self loops (tests/golden/make_euler_fixture.py)."""
    d = np.load(os.path.join(GOLD, "euler_bump.npz"))
    graphs, targets = [], []
    for s in (1, 2):
        g = ab.graph_type()
        g.set_num_vertices(d[f"in_{s}"].shape[0], d[f"in_{s}"].shape[1])
        g.vertex_features[:] = d[f"in_{s}"]
        g.set_num_edges(d["index_list"].shape[0])
        g.generate_adjacency(d["index_list"])
        graphs.append(g)
        targets.append(d[f"out_{s}"])
    return graphs, np.concatenate(targets)


EULER_WIDTHS = [(3, 6), (9, 14), (17, 32), (35, 64), (67, 32), (35, 14), (17, 7)]
EULER_ACTS = ["softmax"] * 6 + ["swish"]


def test_euler_skip_network_training_parity(cuda, oracle32, oracle64):
    """example/msgpass_euler/src/main.f90:182-276 on its real 12 800-vertex mesh: seven Kipf
    layers, each after the first reading [input | previous] (input_list = [0, -1]), softmax
    message activations and a swish head, Adam lr 2e-2 with clip(-1, 1), batch of 2 graphs."""
    from oracle.oracle import LayerSpec
    graphs, target = _euler_graphs()
    p = ab.pack_graphs(graphs)
    assert p.V == 25600 and p.Z == 2 * 2 * 37681
    specs, layers, lists = [], [], []
    for k, ((fi, fo), act) in enumerate(zip(EULER_WIDTHS, EULER_ACTS)):
        specs.append(LayerSpec("kipf", [fi, fo], 1, activation=act,
                               inputs=None if k == 0 else [-1, k - 1]))
        layers.append(ab.kipf_msgpass_layer_type([fi, fo], 1, act))
        lists.append(None if k == 0 else [0, -1])
    _train_compare(cuda, oracle32, specs, layers, p, target,
                   OptimSpec("adam", lr=2e-2, clip_min=-1.0, clip_max=1.0),
                   ab.adam_optimiser_type(2e-2, clip_dict=ab.clip_type(-1.0, 1.0)), steps=4,
                   input_lists=lists, oracle64=oracle64)


@pytest.mark.parametrize("shape", ["tiles64", "ragged"])
def test_skip_network_two_consumers_parity(cuda, oracle32, oracle64, shape):
    """A layer read by two later layers (gradient = sum of both consumers' slices), ids given
    absolutely (k), relatively (-k) and as 0 = the network input; on a tileable batch the
    fused tile kernels run underneath."""
    from oracle.oracle import LayerSpec
    rng = np.random.default_rng(77)
    if shape == "tiles64":
        p = synth.regular_batch(10, 64, 4, 64, rng)
        F0, w = 64, [64, 64, 32, 16]
    else:
        p = synth.molecular_batch(23, 5, 0, rng, nv_range=(2, 40), self_loop_features=False)
        F0, w = 5, [7, 6, 9, 4]
    # L1: x -> w0 (relu) ; L2: L1 -> w1 (tanh) ; L3: [L1 | L2] -> w2 (sigmoid) ;
    # L4: [x | L3 | L1] -> w3 (swish)
    specs = [LayerSpec("kipf", [F0, w[0]], 1, activation="relu"),
             LayerSpec("kipf", [w[0], w[1]], 1, activation="tanh"),
             LayerSpec("kipf", [w[0] + w[1], w[2]], 1, activation="sigmoid", inputs=[0, 1]),
             LayerSpec("kipf", [F0 + w[2] + w[0], w[3]], 1, activation="swish", inputs=[-1, 2, 0])]
    layers = [ab.kipf_msgpass_layer_type([F0, w[0]], 1, "relu"),
              ab.kipf_msgpass_layer_type([w[0], w[1]], 1, "tanh"),
              ab.kipf_msgpass_layer_type([w[0] + w[1], w[2]], 1, "sigmoid"),
              ab.kipf_msgpass_layer_type([F0 + w[2] + w[0], w[3]], 1, "swish")]
    lists = [None, None, [1, -1], [0, 3, -3]]
    target = rng.standard_normal((p.V, w[3])).astype(np.float32)
    _train_compare(cuda, oracle32, specs, layers, p, target, OptimSpec("sgd", lr=0.05),
                   ab.sgd_optimiser_type(0.05), input_lists=lists, oracle64=oracle64)


def test_network_add_input_list_rules(cuda):
    """Errors of network%add(layer, input_list, operator): ids out of range
    (athena_network_sub.f90:835-847), invalid operator (:820-823), widths that do not add up."""
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([3, 6], 1, "softmax"))
    with pytest.raises(ab.AthenaCudaError):
        net.add(ab.kipf_msgpass_layer_type([9, 4], 1), input_list=[0, 2])       # layer 2 is itself
    with pytest.raises(ab.AthenaCudaError):
        net.add(ab.kipf_msgpass_layer_type([9, 4], 1), input_list=[0, -2])      # before the input
    with pytest.raises(ab.AthenaCudaError):
        net.add(ab.kipf_msgpass_layer_type([8, 4], 1), input_list=[0, -1])      # 3 + 6 != 8
    with pytest.raises(ab.AthenaCudaError):
        net.add(ab.kipf_msgpass_layer_type([9, 4], 1), input_list=[0, -1], operator="*")
    with pytest.raises(ab.AthenaCudaError):
        net.add(ab.duvenaud_msgpass_layer_type([9], [0], 1, 3, 5), input_list=[0, -1])
    net.add(ab.kipf_msgpass_layer_type([9, 4], 1), input_list=[0, -1], operator="concatenate")
    net.add(ab.kipf_msgpass_layer_type([13, 2], 1), input_list=[1, 0, 2], operator="||")
    net.destroy()


@pytest.mark.parametrize("kind", ["kipf_tiles", "kipf_T2_ragged", "duvenaud", "full"])
def test_swish_layer_parity(cuda, oracle32, kind):
    """swish (athena_activation_swish.f90:29-34) as message / dense activation: forward and the
    derivative on the saved pre-activation (get_partial_swish_val)."""
    rng = np.random.default_rng(zlib.crc32(kind.encode()))
    x = None
    if kind == "kipf_tiles":
        p = synth.regular_batch(9, 64, 5, 64, rng)
        spec, L = kipf_spec([64, 64], 1, "swish"), ab.kipf_msgpass_layer_type([64, 64], 1, "swish")
    elif kind == "kipf_T2_ragged":
        p = synth.molecular_batch(15, 6, 0, rng, nv_range=(1, 30), self_loop_features=False)
        spec, L = kipf_spec([6, 9, 5], 2, "swish"), ab.kipf_msgpass_layer_type([6, 9, 5], 2, "swish")
    elif kind == "duvenaud":
        p = synth.chemical_batch(6, rng)
        spec = duvenaud_spec([6] * 3, 1, 2, 1, 10, 5, "swish", "softmax")
        L = ab.duvenaud_msgpass_layer_type([6], [1], 2, 10, 5, message_activation="swish")
    else:
        p = synth.chemical_batch(11, rng)
        spec, L = full_spec(7, 5, "swish"), ab.full_layer_type(5, 7, activation="swish")
        x = rng.standard_normal((11, 7)).astype(np.float32)
    n = oracle32.num_params([spec])
    params = random_params(n, rng, 0.4)
    L.set_params(params)
    L.set_graph(p)
    ob = to_oracle_batch(p)
    out_shape = oracle32.out_shape([spec], ob)
    g = rng.standard_normal(out_shape).astype(np.float32)
    out_ref, grad_ref, dx_ref = oracle32.layer_fwd_bwd(spec, params, ob, g_out=g, want_dx=True, x=x)
    out = L.forward(x) if x is not None else L.forward()
    assert rel_err(out, out_ref) <= RTOL_ACT
    L.zero_gradients()
    dx = L.backward(g, want_input_grad=True)
    assert rel_err(L.get_gradients(), grad_ref) <= RTOL_ACT
    assert rel_err(dx, dx_ref) <= RTOL_ACT
    L.destroy()


# ---------------------------------------------------------------------------
# f3 / f4: graph construction on the device, real chemical data, ONNX graph inputs
# ---------------------------------------------------------------------------
BATCH_INTS = ("row_ptr", "col", "eid", "deg", "vgraph", "csc_ptr", "csc_src", "csc_ent")


def _host_graphs(nvs, index_lists, loops):
    graphs = []
    for nv, il in zip(nvs, index_lists):
        g = ab.graph_type()
        g.set_num_vertices(int(nv), 1)
        g.set_num_edges(len(il), 0)
        g.generate_adjacency([tuple(int(t) for t in e) for e in il])
        if loops:
            g.add_self_loops()
        graphs.append(g)
    return graphs


def _random_edge_lists(rng, B, nv_max, e_max, self_edges=True):
    nvs, ils = [], []
    for _ in range(B):
        nv = int(rng.integers(0, nv_max + 1))
        ne = int(rng.integers(0, e_max + 1)) if nv > 0 else 0
        il = rng.integers(1, max(nv, 1) + 1, (ne, 2)).astype(np.int32)
        if not self_edges and ne:
            il = il[il[:, 0] != il[:, 1]]
        nvs.append(nv)
        ils.append(il)
    return nvs, ils


@pytest.mark.parametrize("loops", [False, True])
@pytest.mark.parametrize("case", ["fixtures", "random", "star", "dup_hub"])
def test_generate_adjacency_on_device_is_bit_exact(cuda, case, loops):
    """athena_cuda_batch_create_from_edges against the host restatement of
    generate_adjacency / add_self_loops fed through the CSR route: every integer structure of
    the batch identical (empty graphs, isolated vertices, self edges, repeated edges, hubs)."""
    rng = np.random.default_rng(zlib.crc32(f"{case}{loops}".encode()))
    if case == "fixtures":     # test_kipf_msgpass_layer.f90:83-90, test_msgpass_network.f90:264-271
        nvs = [6, 5, 6]
        ils = [np.array([[1, 2], [1, 3], [2, 3], [2, 4], [3, 5], [4, 5], [4, 6], [5, 6]]),
               np.array([[1, 2], [1, 3], [2, 3], [2, 4], [3, 5], [4, 5]])]
        ils.append(ils[0])
    elif case == "random":
        nvs, ils = _random_edge_lists(rng, 57, 40, 150)
        nvs[3], ils[3] = 0, np.zeros((0, 2), np.int32)        # an empty graph
        nvs[9], ils[9] = 7, np.zeros((0, 2), np.int32)        # seven isolated vertices
    elif case == "star":        # one 40 000-leaf hub: the long-row sort
        n = 40001
        il = np.stack([np.ones(n - 1, np.int32), rng.permutation(np.arange(2, n + 1))], 1)
        nvs, ils = [5, n, 3], [np.array([[1, 2], [5, 4]]), il.astype(np.int32), np.array([[3, 3]])]
    else:                       # repeated edges onto a few hubs (rows of 33..2000 entries)
        nvs, ils = [300], [np.stack([rng.integers(1, 4, 5000), rng.integers(1, 301, 5000)], 1)]
    graphs = _host_graphs(nvs, ils, loops)
    p = ab.pack_graphs(graphs, with_features=False)
    ref = ab.GraphBatch(p)
    dev = ab.GraphBatch.from_edges(nvs, ils, add_self_loops=loops)
    assert (dev.B, dev.V, dev.Z, dev.E) == (ref.B, ref.V, ref.Z, ref.E)
    for k in BATCH_INTS:
        assert np.array_equal(dev.export(k), ref.export(k)), k
    assert np.array_equal(dev.export("coef"), ref.export("coef"))
    # with the entry counts known to the caller the build does not synchronise; a wrong count
    # is caught by the row-pointer check of the CSR build
    hinted = ab.GraphBatch.from_edges(nvs, ils, add_self_loops=loops, num_entries=p.nz)
    for k in BATCH_INTS:
        assert np.array_equal(hinted.export(k), ref.export(k)), k
    hinted.destroy()
    if p.Z > 0:
        wrong = p.nz.copy()
        wrong[int(np.argmax(wrong))] += 1
        with pytest.raises(ab.AthenaCudaError):
            ab.GraphBatch.from_edges(nvs, ils, add_self_loops=loops, num_entries=wrong)
    ref.destroy()
    dev.destroy()


def test_generate_adjacency_on_device_rejects_bad_index(cuda):
    with pytest.raises(ab.AthenaCudaError) as ei:
        ab.GraphBatch.from_edges([4, 3], [np.array([[1, 2]]), np.array([[1, 2], [2, 4]])])
    assert ei.value.code == -4 and "sample 2" in str(ei.value)
    with pytest.raises(ab.AthenaCudaError):
        ab.GraphBatch.from_edges([4], [np.array([[0, 2]])], add_self_loops=True)


def _edge_index_of(g):
    """build_edge_index of example/msgpass_chemical/validate_onnx.py:38-57."""
    ncsr = g.adj_ja.shape[0]
    ei = np.zeros((3, ncsr), np.int64)
    k = 0
    for v in range(g.num_vertices):
        for j in range(g.adj_ia[v] - 1, g.adj_ia[v + 1] - 1):
            ei[0, k] = g.adj_ja[j, 0] - 1
            ei[1, k] = g.adj_ja[j, 1] - 1
            ei[2, k] = v
            k += 1
    return ei, np.diff(g.adj_ia).astype(np.int64)


def test_onnx_edge_index_ingestion_is_bit_exact(cuda):
    """[3, ncsr] edge_index + degree (athena_onnx_msgpass_utils.f90:53-92) -> the same batch as
    the CSR route, self loops (edge-feature index -1) included."""
    rng = np.random.default_rng(8)
    nvs, ils = _random_edge_lists(rng, 23, 30, 90)
    graphs = _host_graphs(nvs, ils, True)
    for g, il in zip(graphs, ils):
        g.num_edges = len(il)
    p = ab.pack_graphs(graphs, with_features=False)
    ref = ab.GraphBatch(p)
    pairs = [_edge_index_of(g) for g in graphs]
    deg = np.concatenate([d for _, d in pairs])
    dev = ab.GraphBatch.from_edge_index(p.nv, p.ne, [e for e, _ in pairs], deg)
    for k in BATCH_INTS:
        assert np.array_equal(dev.export(k), ref.export(k)), k
    dev.destroy()
    # degrees that do not add up / an entry outside its target's row / a bad source
    bad = deg.copy(); bad[np.nonzero(deg)[0][0]] += 1
    with pytest.raises(ab.AthenaCudaError):
        ab.GraphBatch.from_edge_index(p.nv, p.ne, [e for e, _ in pairs], bad)
    k = next(i for i, (e, _) in enumerate(pairs) if e.shape[1] > 3)
    e_bad = [e.copy() for e, _ in pairs]
    e_bad[k][2, 0], e_bad[k][2, -1] = e_bad[k][2, -1], e_bad[k][2, 0]
    with pytest.raises(ab.AthenaCudaError):
        ab.GraphBatch.from_edge_index(p.nv, p.ne, e_bad, deg)
    e_bad = [e.copy() for e, _ in pairs]
    e_bad[k][0, 1] = nvs[k]
    with pytest.raises(ab.AthenaCudaError):
        ab.GraphBatch.from_edge_index(p.nv, p.ne, e_bad, deg)
    ref.destroy()


def _chemical_dataset():
    from athena_b200.read_chemical_graphs import get_graph_from_basis
    d = np.load(os.path.join(GOLD, "chemical_database.npz"))
    graphs = []
    for s in range(d["energy"].size):
        g = get_graph_from_basis(d["lattice"][s], ["C"] * 8, d["positions"][s], d["forces"][s])
        graphs.append(g)
    e = d["energy"].astype(np.float32)
    labels = (e - e.min()) / (e.max() - e.min())      # main.f90:246-251
    return graphs, labels


def test_chemical_database_epoch_training_parity(cuda, oracle32):
    """BASELINE configs[0] on its REAL data: the 198 cells of example/msgpass_chemical/
    database.xyz read as mod_read_chemical_graphs.f90:196-278 reads them, add_self_loops,
    Duvenaud(T=4, D=10, 10 outputs) -> full 128 -> 64 -> 1, Adam lr 1e-2 + clip_norm 0.1, batches
    of 8 (the last one of 6), one whole epoch of network%train against the oracle; the graphs of
    every batch are built on the device from the edge lists."""
    graphs, labels = _chemical_dataset()
    assert len(graphs) == 198
    index_lists = []
    for g in graphs:     # recover the undirected edge list: entry (nb, k) with nb >= row lists edge k
        il = np.zeros((g.num_edges, 2), np.int32)
        for v in range(g.num_vertices):
            for nb, k in g.adj_ja[g.adj_ia[v] - 1:g.adj_ia[v + 1] - 1]:
                if nb >= v + 1:
                    il[k - 1] = (v + 1, nb)
        index_lists.append(il)
        g.add_self_loops()                            # main.f90:107-110
    specs = [duvenaud_spec([6] * 5, 1, 4, 1, 10, 10), full_spec(10, 128, "leaky_relu"),
             full_spec(128, 64, "leaky_relu"), full_spec(64, 1, "leaky_relu")]
    net = ab.network_type()
    net.add(ab.duvenaud_msgpass_layer_type([6], [1], 4, 10, 10))
    for w in (128, 64, 1):
        net.add(ab.full_layer_type(w, activation="leaky_relu"))
    net.compile(ab.adam_optimiser_type(1e-2, clip_dict=ab.clip_type(clip_norm=0.1)), batch_size=8)
    n = oracle32.num_params(specs)
    assert net.num_params == n
    rng = np.random.default_rng(2)
    params = random_params(n, rng, 0.3)
    net.set_params(params)
    ref = params.copy()
    s1 = np.zeros(n, np.float32); s2 = np.zeros(n, np.float32)
    optim = OptimSpec("adam", lr=1e-2, clip_norm=0.1)
    it = 0
    for s0 in range(0, 198, 8):
        gs = graphs[s0:s0 + 8]
        p = ab.pack_graphs(gs)
        tgt = labels[s0:s0 + 8].reshape(-1, 1)
        dev = ab.GraphBatch.from_edges(p.nv, index_lists[s0:s0 + 8], add_self_loops=True)
        dev.packed = p                                # features travel with the call
        if s0 == 0:
            host = ab.GraphBatch(p)
            for k in BATCH_INTS:
                assert np.array_equal(dev.export(k), host.export(k)), k
            host.destroy()
        it += 1
        loss_ref, _ = oracle32.train_step(specs, ref, to_oracle_batch(p), tgt, optim, s1, s2, it)
        loss = net.train_step(dev, tgt)
        assert abs(loss - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref)), (it, loss, loss_ref)
        dev.destroy()
    assert it == 25
    assert rel_err(net.get_params(), ref) <= RTOL_PARAM
    net.destroy()


@pytest.mark.parametrize("entry", ["layer", "network"])
def test_nonfinite_features_stay_confined_to_neighbours(cuda, oracle32, entry):
    """A NaN / Inf feature reaches exactly the rows that list its vertex, as in the reference's
    entry-by-entry gather (_sub_kipf.f90:36-45).  The tensor-core gather would spread it over
    the vertex's whole 128-row tile (0 * Inf in the dense adjacency product): the forward
    kernel detects it and the pass is repeated with the list gather."""
    rng = np.random.default_rng(99)
    p = synth.regular_batch(8, 64, 3, 64, rng)       # tileable, F = 64: the tcgen05 path
    x = p.x.copy()
    x[5, 7] = np.nan          # graph 0
    x[64 + 9, 0] = np.inf     # graph 1 (same 128-row tile as graph 0)
    x[300, 63] = -np.inf      # graph 4
    spec = kipf_spec([64, 64], 1, "tanh")
    params = random_params(64 * 64, rng, 0.2)
    ob = Batch(p.nv, p.ne, p.ia, p.ja, x, None)
    ref, _, _ = oracle32.layer_fwd_bwd(spec, params, ob)
    if entry == "layer":
        L = ab.kipf_msgpass_layer_type([64, 64], 1, "tanh")
        L.set_params(params)
        L.set_graph(p)
        out = L.forward(x)
        clean = L.forward(p.x)                        # the next call is back on the tensor core
        L.destroy()
    else:
        net = ab.network_type()
        net.add(ab.kipf_msgpass_layer_type([64, 64], 1, "tanh"))
        net.compile(ab.sgd_optimiser_type(0.1), batch_size=p.B)
        net.set_params(params)
        gb = ab.GraphBatch(p)
        out = net.forward(gb, vertex_features=x)
        clean = net.forward(gb, vertex_features=p.x)
        net.destroy()
    bad = ~np.isfinite(ref)
    assert bad.any() and bad.sum() < 40 * 64          # only neighbours of the three vertices
    assert np.array_equal(~np.isfinite(out), bad)
    assert np.abs(out[~bad] - ref[~bad]).max() <= RTOL_ACT * np.abs(ref[~bad]).max()
    ref_clean, _, _ = oracle32.layer_fwd_bwd(spec, params, to_oracle_batch(p))
    assert rel_err(clean, ref_clean) <= RTOL_ACT


@pytest.mark.parametrize("kind", ["kipf", "duvenaud"])
def test_dataset_resident_training_equals_streaming(cuda, kind):
    """network%train with the data set kept on the device (the reference keeps it in memory for
    the whole call, athena_network_sub.f90:3564-3565) is the same computation as re-sending
    every mini-batch: identical losses and bitwise identical parameters."""
    rng = np.random.default_rng(41)

    def make():
        net = ab.network_type()
        if kind == "kipf":
            net.add(ab.kipf_msgpass_layer_type([8, 16], 1, "tanh"))
            net.add(ab.kipf_msgpass_layer_type([16, 4], 1, "none"))
        else:
            net.add(ab.duvenaud_msgpass_layer_type([6], [1], 2, 10, 4))
        net.compile(ab.adam_optimiser_type(5e-3), batch_size=8)
        return net
    if kind == "kipf":
        p = synth.molecular_batch(37, 8, 0, rng, nv_range=(3, 20), self_loop_features=False)
        target = rng.standard_normal((p.V, 4)).astype(np.float32)
    else:
        p = synth.chemical_batch(37, rng)
        target = rng.random((37, 4)).astype(np.float32)
    a, b = make(), make()
    params = random_params(a.num_params, rng, 0.3)
    a.set_params(params)
    b.set_params(params)
    ha = a.train(p, target, num_epochs=3, shuffle_batches=True, seed=5)
    hb = b.train(p, target, num_epochs=3, shuffle_batches=True, seed=5, resident=True)
    assert ha == hb
    assert np.array_equal(a.get_params(), b.get_params())
    a.destroy()
    b.destroy()


@pytest.mark.parametrize("decay", ["exp", "step"])
def test_lr_decay_training_parity(cuda, oracle32, decay):
    """network%train with a decaying learning rate (athena_optimiser.f90:414 get_lr(lr, iter)):
    exp decay per iteration (example/msgpass_euler: exp_lr_decay_type(1e-3)); step decay advances
    the optimiser's counter once per EPOCH, and Adam's bias correction reads that same counter
    (athena_network_sub.f90:2834-2841, athena_optimiser.f90:1058-1059)."""
    rng = np.random.default_rng(13)
    p = synth.molecular_batch(24, 6, 0, rng, nv_range=(3, 12), self_loop_features=False)
    specs = [kipf_spec([6, 10], 1, "tanh"), kipf_spec([10, 3], 1, "none")]
    lr0 = 0.02
    dec = ab.exp_lr_decay_type(0.05) if decay == "exp" else ab.step_lr_decay_type(0.5, 2)
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([6, 10], 1, "tanh"))
    net.add(ab.kipf_msgpass_layer_type([10, 3], 1, "none"))
    net.compile(ab.adam_optimiser_type(lr0, lr_decay=dec), batch_size=8)
    n = oracle32.num_params(specs)
    params = random_params(n, rng, 0.4)
    net.set_params(params)
    target = rng.standard_normal((p.V, 3)).astype(np.float32)
    hist = net.train(p, target, num_epochs=4, shuffle_batches=False)
    ref = params.copy()
    s1 = np.zeros(n, np.float32); s2 = np.zeros(n, np.float32)
    voff = np.concatenate([[0], np.cumsum(p.nv)])
    it, hist_ref = 0, []
    for epoch in range(1, 5):
        if decay == "step":
            it = epoch                                  # the counter advances once per epoch
        tot = 0.0
        for s0 in range(0, 24, 8):
            if decay == "exp":
                it += 1
            q = p.slice(s0, s0 + 8)
            o = OptimSpec("adam", lr=dec.get_lr(lr0, it))
            l, _ = oracle32.train_step(specs, ref, to_oracle_batch(q),
                                       target[voff[s0]:voff[s0 + 8]], o, s1, s2, it)
            tot += l
        hist_ref.append(tot / 3)
    assert np.allclose(hist, hist_ref, rtol=1e-4)
    assert rel_err(net.get_params(), ref) <= RTOL_PARAM
    net.destroy()


def test_generate_adjacency_from_device_edge_list(cuda):
    """athena_cuda_batch_create_from_edges with the index list already in device memory
    (ATHENA_MEM_DEVICE), e.g. produced by an earlier kernel: same batch as from host memory."""
    import ctypes as C
    rng = np.random.default_rng(3)
    nvs, ils = _random_edge_lists(rng, 31, 25, 70)
    nv = np.asarray(nvs, np.int32)
    ne = np.asarray([len(il) for il in ils], np.int32)
    il = np.ascontiguousarray(np.concatenate(ils), np.int32)
    ref = ab.GraphBatch.from_edges(nvs, ils, add_self_loops=True)
    d_il = ab.DeviceArray.from_host(il)
    h = C.c_int64()
    ab.check(ab.lib().athena_cuda_batch_create_from_edges(
        C.byref(h), nv.size, ab.ptr(nv), ab.ptr(ne), C.c_void_p(d_il.addr), None, 1,
        ab.MEM_DEVICE, 1))
    dev = ab.GraphBatch._adopt(h.value)
    assert (dev.V, dev.Z) == (ref.V, ref.Z)
    for k in BATCH_INTS:
        assert np.array_equal(dev.export(k), ref.export(k)), k
    dev.destroy()
    ref.destroy()
    d_il.free()


def test_onnx_edge_index_single_large_graph_and_empty_graphs(cuda):
    """One graph that is not tileable (5 000 vertices) and a batch with empty graphs through the
    edge_index route."""
    rng = np.random.default_rng(4)
    nvs, ils = [5000, 0, 3, 0], [rng.integers(1, 5001, (20000, 2)).astype(np.int32),
                                 np.zeros((0, 2), np.int32), np.array([[1, 2], [2, 3]], np.int32),
                                 np.zeros((0, 2), np.int32)]
    graphs = _host_graphs(nvs, ils, True)
    p = ab.pack_graphs(graphs, with_features=False)
    ref = ab.GraphBatch(p)
    pairs = [_edge_index_of(g) for g in graphs]
    dev = ab.GraphBatch.from_edge_index(p.nv, p.ne, [e for e, _ in pairs],
                                        np.concatenate([d for _, d in pairs]))
    for k in BATCH_INTS:
        assert np.array_equal(dev.export(k), ref.export(k)), k
    dev.destroy()
    ref.destroy()


def test_skip_network_multi_step_sources(cuda, oracle32, oracle64):
    """Sources with several time steps: the concatenation takes the LAST step's width
    (num_vertex_features(T)), the gradient re-enters the source layer at its last step."""
    from oracle.oracle import LayerSpec
    rng = np.random.default_rng(78)
    p = synth.molecular_batch(17, 4, 0, rng, nv_range=(2, 25), self_loop_features=False)
    specs = [LayerSpec("kipf", [4, 6, 5], 2, activation="tanh"),
             LayerSpec("kipf", [9, 7, 3], 2, activation="sigmoid", inputs=[0, -1]),
             LayerSpec("kipf", [12, 2], 1, activation="none", inputs=[-1, 0, 1])]
    layers = [ab.kipf_msgpass_layer_type([4, 6, 5], 2, "tanh"),
              ab.kipf_msgpass_layer_type([9, 7, 3], 2, "sigmoid"),
              ab.kipf_msgpass_layer_type([12, 2], 1, "none")]
    lists = [None, [1, 0], [0, 1, 2]]
    target = rng.standard_normal((p.V, 2)).astype(np.float32)
    _train_compare(cuda, oracle32, specs, layers, p, target, OptimSpec("adam", lr=0.01),
                   ab.adam_optimiser_type(0.01), input_lists=lists, oracle64=oracle64)
