"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/athena_cuda.h declares, fails loudly without a GPU, and its
pure-host helpers behave.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import athena_b200 as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "athena_cuda.h")).read()
    return sorted(set(re.findall(r"\b(athena_cuda_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(ab.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol():
    L = ab.lib()
    for n in declared_symbols():
        assert getattr(L, n).restype is not None


def test_optimiser_descriptor_layout_matches_the_header(tmp_path):
    """The ctypes mirror of struct athena_optimiser_desc (and the constants it carries) must be
    the struct the C compiler sees: size and the offset of every field."""
    import subprocess
    from athena_b200 import _lib
    fields = [f for f, _ in _lib.OptimiserDesc._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "athena_cuda.h"\n'
                   "int main(void) {\n"
                   '  printf("%zu", sizeof(athena_optimiser_desc));\n'
                   + "".join(f'  printf(" %zu", offsetof(athena_optimiser_desc, {f}));\n'
                             for f in fields)
                   + '  printf(" %d %d %d %d", ATHENA_OPT_SGD, ATHENA_OPT_ADAM, ATHENA_OPT_RMSPROP,'
                     " ATHENA_OPT_ADAGRAD);\n"
                   '  printf(" %d %d %d %d", ATHENA_REG_NONE, ATHENA_REG_L1, ATHENA_REG_L2,'
                     " ATHENA_REG_L1L2);\n  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.OptimiserDesc)] + [getattr(_lib.OptimiserDesc, f).offset for f in fields]
    want += [_lib.OPT_SGD, _lib.OPT_ADAM, _lib.OPT_RMSPROP, _lib.OPT_ADAGRAD,
             _lib.REG_NONE, _lib.REG_L1, _lib.REG_L2, _lib.REG_L1L2]
    assert got == want


def test_optimiser_mirror_fills_the_descriptor_like_the_reference_defaults():
    """Host logic of the Python mirror: constructor defaults are the reference's
    (athena_optimiser.f90:119-250, athena_regulariser.f90:40-76, athena_clipper.f90:14-30)."""
    from athena_b200 import _lib
    d = ab.sgd_optimiser_type().desc()
    assert (d.kind, d.momentum, d.nesterov, d.regulariser) == (_lib.OPT_SGD, 0.0, 0, _lib.REG_NONE)
    assert abs(d.learning_rate - 0.01) < 1e-7 and d.clip_min_max == 0 and d.clip_norm_on == 0
    d = ab.adam_optimiser_type(1e-3, regulariser=ab.l2_regulariser_type()).desc()
    assert d.kind == _lib.OPT_ADAM and abs(d.beta1 - 0.9) < 1e-7 and abs(d.beta2 - 0.999) < 1e-7
    assert abs(d.epsilon - 1e-8) < 1e-12
    assert (d.regulariser, d.l2_decoupled) == (_lib.REG_L2, 1) and abs(d.l2 - 0.01) < 1e-7
    d = ab.rmsprop_optimiser_type(clip_dict=ab.clip_type(clip_norm=0.5)).desc()
    assert d.kind == _lib.OPT_RMSPROP and d.beta1 == 0.0 and d.clip_norm_on == 1
    assert abs(d.clip_norm - 0.5) < 1e-7
    d = ab.adagrad_optimiser_type(regulariser=ab.l1l2_regulariser_type(0.02, 0.03)).desc()
    assert d.kind == _lib.OPT_ADAGRAD and d.regulariser == _lib.REG_L1L2
    assert abs(d.l1 - 0.02) < 1e-7 and abs(d.l2 - 0.03) < 1e-7
    d = ab.sgd_optimiser_type(clip_dict=ab.clip_type(-0.1, 0.2)).desc()
    assert d.clip_min_max == 1 and abs(d.clip_min + 0.1) < 1e-7 and abs(d.clip_max - 0.2) < 1e-7


def test_header_cites_the_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "athena_cuda.h")).read()
    for cite in ("athena_msgpass_layer_sub.f90:144-174", "athena_kipf_msgpass_layer.f90:915-959",
                 "athena_duvenaud_msgpass_layer.f90:755-859", "athena_network_sub.f90:3611-3670",
                 "athena_diffstruc_extd_sub_kipf.f90:85-111"):
        assert cite in hdr, cite


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on CPU-only hosts")
    rc = ab.lib().athena_cuda_init(-1)
    assert rc == -1
    assert b"no CPU fallback" in ab.lib().athena_cuda_last_error()
    with pytest.raises(ab.AthenaCudaError):
        ab.kipf_msgpass_layer_type([4, 4], 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "athena_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cc")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert not re.search(r"#include[^\n]*oracle", src) and "liboracle" not in src, f


def test_shard_graphs_balances_entries():
    w = np.array([10, 10, 10, 10, 100, 10, 10, 10], np.int64)
    first = np.zeros(3, np.int32)
    ab.check(ab.lib().athena_cuda_shard_graphs(w.size, ab.ptr(w), 2, ab.ptr(first)))
    assert first[0] == 0 and first[2] == 8 and 4 <= first[1] <= 5
    first = np.zeros(9, np.int32)
    ab.check(ab.lib().athena_cuda_shard_graphs(w.size, ab.ptr(w), 8, ab.ptr(first)))
    # as many graphs as ranks: every rank owns exactly one graph, however skewed the weights
    assert first.tolist() == list(range(9))
    # skewed batches never leave a rank empty while num_graphs >= world_size
    # (an empty rank cannot build a batch and its peers would wait in the gradient exchange)
    for w, world in ((np.array([1, 100], np.int64), 2),
                     (np.array([1000, 1, 1, 1, 1, 1, 1, 1, 1, 1], np.int64), 4),
                     (np.array([1, 1, 1, 1, 1, 1, 1, 1, 1, 1000], np.int64), 8)):
        first = np.zeros(world + 1, np.int32)
        ab.check(ab.lib().athena_cuda_shard_graphs(w.size, ab.ptr(w), world, ab.ptr(first)))
        assert first[0] == 0 and first[world] == w.size and np.all(np.diff(first) >= 1), first
    # fewer graphs than ranks: trailing ranks are empty by necessity, boundaries stay ordered
    w = np.array([5, 5], np.int64)
    first = np.zeros(5, np.int32)
    ab.check(ab.lib().athena_cuda_shard_graphs(w.size, ab.ptr(w), 4, ab.ptr(first)))
    assert first[0] == 0 and first[4] == 2 and np.all(np.diff(first) >= 0)
    # uniform weights -> equal contiguous shards
    w = np.full(4096, 832, np.int64)
    first = np.zeros(5, np.int32)
    ab.check(ab.lib().athena_cuda_shard_graphs(w.size, ab.ptr(w), 4, ab.ptr(first)))
    assert first.tolist() == [0, 1024, 2048, 3072, 4096]


def test_graph_type_helpers_and_packing():
    g = ab.graph_type()
    g.set_num_vertices(5, 8)
    g.set_num_edges(6, 2)
    # test/test_msgpass_network.f90:249-276
    g.generate_adjacency([[1, 2], [1, 3], [2, 3], [2, 4], [3, 5], [4, 5]])
    assert g.adj_ia.tolist() == [1, 3, 6, 9, 11, 13]
    assert g.adj_ja[:2].tolist() == [[2, 1], [3, 2]]
    g.add_self_loops()
    assert g.num_entries == 12 + 5
    assert g.adj_ja[g.adj_ia[0] - 1 + 2].tolist() == [1, 0]
    g2 = ab.graph_type()
    g2.set_num_vertices(2, 8)
    g2.set_num_edges(1, 2)
    g2.generate_adjacency([[1, 2]])
    p = ab.pack_graphs([g, g2])
    assert p.B == 2 and p.V == 7 and p.Z == 19 and p.E == 7
    assert p.ia.tolist() == g.adj_ia.tolist() + g2.adj_ia.tolist()
    s = p.slice(1, 2)
    assert s.ia.tolist() == [1, 2, 3] and s.ja.tolist() == [[2, 1], [1, 1]] and s.x.shape == (2, 8)


def test_synthetic_configs_have_the_documented_shapes():
    from athena_b200 import synth
    rng = np.random.default_rng(0)
    p = synth.regular_batch(32, 64, 6, 64, rng)           # cfg2 at 1/128 scale
    assert p.V == 32 * 64 and p.Z == 32 * 64 * 13 and p.x.shape == (2048, 64)
    assert np.all(np.diff(p.ia.reshape(32, 65), axis=1) == 13)
    p = synth.molecular_batch(64, 32, 4, rng)
    deg = np.concatenate([np.diff(p.ia[o:o + n + 1]) for o, n in
                          zip(np.cumsum(np.r_[0, p.nv[:-1] + 1]), p.nv)])
    assert deg.max() <= 5 and deg.min() >= 3             # degree <= 4 + self loop
    assert p.ja[:, 1].min() >= 1                           # self loops carry edge features


def test_extxyz_reader_follows_the_reference_reader():
    """athena_b200/read_chemical_graphs.py on the real data set (fixture of database.xyz):
    the sizes SURVEY.md section 8d re-derived from mod_read_chemical_graphs.f90:196-278
    (46 undirected edges per cell on average, 32..61; degree 6..17 before the self loop),
    the vertex-feature layout of :268-272 and edge features in (0.5, 3.0) / 3.0."""
    import os
    from athena_b200.read_chemical_graphs import get_graph_from_basis, parse_extxyz
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "chemical_database.npz"))
    assert d["energy"].size == 198
    ne, degs = [], []
    for s in range(0, 198, 9):
        g = get_graph_from_basis(d["lattice"][s], ["C"] * 8, d["positions"][s], d["forces"][s])
        assert g.num_vertices == 8 and g.num_vertex_features == 6 and g.num_edge_features == 1
        ne.append(g.num_edges)
        degs.extend(np.diff(g.adj_ia))
        assert g.adj_ja.shape[0] == 2 * g.num_edges          # no self images below 3 A here
        assert np.allclose(g.vertex_features[:, 3], 0.06) and \
            np.allclose(g.vertex_features[:, 4], 12.011 / 52.0)
        assert np.all(g.edge_features > 0.5 / 3.0) and np.all(g.edge_features < 1.0)
        # feature 6 = (edges found from this atom to itself or LATER atoms) / 6
        cnt = np.zeros(8)
        for v in range(8):
            for nb, k in g.adj_ja[g.adj_ia[v] - 1:g.adj_ia[v + 1] - 1]:
                if nb >= v + 1:
                    cnt[v] += 1
        assert np.allclose(g.vertex_features[:, 5], cnt / 6.0)
        assert np.allclose(g.vertex_features[:, :3], d["forces"][s], rtol=1e-6)
    assert 32 <= min(ne) and max(ne) <= 61 and 6 <= min(degs) and max(degs) <= 17
    txt = ('2\nLattice="4.0 0.0 0.0 0.0 4.0 0.0 0.0 0.0 4.0" Properties=species:S:1:pos:R:3:'
           'forces:R:3 energy=-1.5 free_energy=-9.0 pbc="T T T"\n'
           'C 0.0 0.0 0.0 0.1 0.2 0.3\nC 1.0 0.0 0.0 -0.1 -0.2 -0.3\n\n')
    fr = parse_extxyz(txt)
    assert len(fr) == 1 and fr[0]["energy"] == -1.5 and fr[0]["positions"].shape == (2, 3)
    g = get_graph_from_basis(fr[0]["lattice"], fr[0]["species"], fr[0]["positions"], fr[0]["forces"])
    # distances: 1.0 (the pair) and 3.0 (pair through the far face) -- 3.0 is NOT < cutoff_max
    assert g.num_edges == 1 and abs(float(g.edge_features[0, 0]) - 1.0 / 3.0) < 1e-6


def test_lr_decay_mirrors_follow_the_reference_formulas():
    """athena_lr_decay.f90:218-270 in real32: exp lr*exp(-it*rate), step lr*rate**(it/steps)
    with integer division and a per-epoch counter, inv lr*(1+rate*it)**(-power)."""
    from athena_b200.network import (base_lr_decay_type, exp_lr_decay_type, inv_lr_decay_type,
                                     step_lr_decay_type)
    f = np.float32
    assert base_lr_decay_type().get_lr(0.02, 7) == float(f(0.02))
    e = exp_lr_decay_type(1e-3)
    assert not e.iterate_per_epoch
    assert abs(e.get_lr(2e-2, 150) - float(f(2e-2) * np.exp(f(-150) * f(1e-3)))) <= 1e-9
    s = step_lr_decay_type(0.5, 5)
    assert s.iterate_per_epoch
    assert [s.get_lr(1.0, it) for it in (1, 4, 5, 9, 10)] == [1.0, 1.0, 0.5, 0.5, 0.25]
    assert exp_lr_decay_type().decay_rate == 0.9 and step_lr_decay_type().decay_steps == 100
    i = inv_lr_decay_type(0.01, 2.0)
    assert abs(i.get_lr(0.1, 10) - 0.1 / 1.1 ** 2) <= 1e-8


def test_tile_kernel_reciprocal_division_table_is_exact():
    """csrc/tile_fma.cu replaces floor(n / d) by (n * tf_inv20[d]) >> 20 for n < 4096, d <= 64:
    the table in the source must be ceil(2^20 / d) and the formula exact over the whole range
    (and the product must fit 32 bits)."""
    import re
    src = open(os.path.join(ROOT, "athena_b200", "csrc", "tile_fma.cu")).read()
    m = re.search(r"tf_inv20\[65\] = \{(.*?)\};", src, re.S)
    assert m is not None
    table = [int(x) for x in m.group(1).replace("\n", " ").split(",")]
    assert len(table) == 65
    n = np.arange(4096, dtype=np.uint64)
    for d in range(1, 65):
        assert table[d] == (2 ** 20 + d - 1) // d
        prod = n * np.uint64(table[d])
        assert int(prod.max()) < 2 ** 32
        assert np.array_equal(prod >> np.uint64(20), n // np.uint64(d))
