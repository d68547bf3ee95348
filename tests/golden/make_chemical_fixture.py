"""Packs example/msgpass_chemical/database.xyz (198 periodic 8-atom carbon cells, the data set of
BASELINE configs[0]) into a compressed fixture: lattices, Cartesian positions, forces, energies
as the file states them (float64).  The graphs are then built by
athena_b200/read_chemical_graphs.py, which follows the reference's reader
(example/example_library/src/mod_read_chemical_graphs.f90:139-278).  /root/reference does not
exist on the GPU box, so the tests load the fixture.

    python tests/golden/make_chemical_fixture.py   ->  tests/golden/chemical_database.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from athena_b200.read_chemical_graphs import parse_extxyz  # noqa: E402

REF = "/root/reference/example/msgpass_chemical/database.xyz"


def main():
    with open(REF) as f:
        frames = parse_extxyz(f.read())
    assert all(fr["species"] == ["C"] * 8 for fr in frames)
    out = dict(lattice=np.stack([fr["lattice"] for fr in frames]),
               positions=np.stack([fr["positions"] for fr in frames]),
               forces=np.stack([fr["forces"] for fr in frames]),
               energy=np.array([fr["energy"] for fr in frames]))
    path = os.path.join(HERE, "chemical_database.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
