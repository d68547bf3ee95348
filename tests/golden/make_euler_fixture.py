"""Packs the mesh of the reference's example/msgpass_euler (the bump-channel Euler data set,
example/msgpass_euler/data/bump_*.txt) into a compressed fixture, read the way
example/msgpass_euler/src/mod_read_euler.f90:14-53 reads it: first line = counts, then the
(F, V) vertex features in Fortran order (one vertex per line) and the (2, E) 1-based edge list.
/root/reference does not exist on the GPU box, so the parity tests load the fixture.

    python tests/golden/make_euler_fixture.py      ->  tests/golden/euler_bump.npz
"""
import os

import numpy as np

REF = "/root/reference/example/msgpass_euler/data"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_vertices(path):
    with open(path) as f:
        nv, nf = (int(t) for t in f.readline().split())
        vals = np.array(f.read().split(), dtype=np.float64)
    assert vals.size == nv * nf
    # read(unit,*) graph%vertex_features fills the (F, V) array in column-major order:
    # feature index fastest = our [V, F] rows; stored as real32 like the reference holds them
    return vals.reshape(nv, nf).astype(np.float32)


def read_edges(path):
    with open(path) as f:
        ne = int(f.readline().split()[0])
        idx = np.array(f.read().split(), dtype=np.int64)
    assert idx.size == 2 * ne
    return idx.reshape(ne, 2).astype(np.int32)  # index_list(2, E), 1-based


def main():
    out = {"index_list": read_edges(os.path.join(REF, "bump_edgeData_1.txt"))}
    for s in (1, 2):
        out[f"in_{s}"] = read_vertices(os.path.join(REF, f"bump_nodeData_in_{s}.txt"))
        out[f"out_{s}"] = read_vertices(os.path.join(REF, f"bump_nodeData_out_{s}.txt"))
    path = os.path.join(HERE, "euler_bump.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
