"""Reader of example/msgpass_chemical's data set: extended-xyz cells -> graph_type samples,
following example/example_library/src/mod_read_chemical_graphs.f90:139-278
(read_extxyz_db / get_graph_from_basis) loop by loop, in real32 like the reference.

What is NOT in the reference tree and therefore restated from its published behaviour:
  * atomstruc's geom_read (extxyz -> basis_type with FRACTIONAL coordinates, atoms grouped by
    species in order of first appearance) and set_element_properties_to_default (charge =
    atomic number, mass = standard atomic weight);
  * graphstruc's convert_to_sparse / generate_adjacency (see athena_b200/graph.py).
Data loading is host work in the reference too; the graph construction itself
(generate_adjacency + add_self_loops) can run on the device: GraphBatch.from_edges.
"""
from __future__ import annotations

import math
import re
from typing import List, Sequence, Tuple

import numpy as np

from .graph import graph_type

# charge (atomic number), mass (standard atomic weight) -- atomstruc defaults
ELEMENT_PROPERTIES = {
    "H": (1.0, 1.008), "He": (2.0, 4.0026), "Li": (3.0, 6.94), "Be": (4.0, 9.0122),
    "B": (5.0, 10.81), "C": (6.0, 12.011), "N": (7.0, 14.007), "O": (8.0, 15.999),
    "F": (9.0, 18.998), "Na": (11.0, 22.990), "Mg": (12.0, 24.305), "Al": (13.0, 26.982),
    "Si": (14.0, 28.085), "P": (15.0, 30.974), "S": (16.0, 32.06), "Cl": (17.0, 35.45),
}

CUTOFF_MIN = np.float32(0.5)   # mod_read_chemical_graphs.f90:229
CUTOFF_MAX = np.float32(3.0)   # :230


def parse_extxyz(text: str) -> List[dict]:
    """Frames of an extended-xyz file: natoms / comment line with Lattice="..." and energy=... /
    `species x y z fx fy fz` rows (Properties=species:S:1:pos:R:3:forces:R:3)."""
    lines = text.splitlines()
    frames, i = [], 0
    while i < len(lines):
        if not lines[i].strip():
            i += 1
            continue
        n = int(lines[i].split()[0])
        header = lines[i + 1]
        lat = re.search(r'Lattice="([^"]*)"', header)
        en = re.search(r'(?<![A-Za-z_])energy=([-+0-9.eEdD]+)', header)
        if lat is None or en is None:
            raise ValueError(f"frame at line {i + 1}: Lattice / energy missing")
        lattice = np.array(lat.group(1).split(), np.float64).reshape(3, 3)
        species, pos, force = [], [], []
        for row in lines[i + 2:i + 2 + n]:
            t = row.split()
            species.append(t[0])
            pos.append([float(v) for v in t[1:4]])
            force.append([float(v) for v in t[4:7]] if len(t) >= 7 else [0.0, 0.0, 0.0])
        frames.append(dict(lattice=lattice, species=species, positions=np.array(pos, np.float64),
                           forces=np.array(force, np.float64),
                           energy=float(en.group(1).replace("d", "e").replace("D", "e"))))
        i += 2 + n
    return frames


def _modu(v: np.ndarray) -> np.ndarray:
    """misc_linalg modu: sqrt(sum(v**2)) in real32, summed in index order."""
    v = v.astype(np.float32)
    acc = v[..., 0] * v[..., 0]
    acc = acc + v[..., 1] * v[..., 1]
    acc = acc + v[..., 2] * v[..., 2]
    return np.sqrt(acc, dtype=np.float32)


def basis_edges(lattice, species: Sequence[str], positions, forces
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """get_graph_from_basis (mod_read_chemical_graphs.f90:196-274) up to the edge list:
    -> (vertex_features [V, 6], index_list [E, 2] 1-based, edge_features [E, 1])."""
    lat = np.asarray(lattice, np.float32)
    # geom_read: atoms grouped by species (first appearance), fractional coordinates
    order, names = [], []
    for name in species:
        if name not in names:
            names.append(name)
    for name in names:
        order.extend(k for k, sp in enumerate(species) if sp == name)
    spec_of = [names.index(species[k]) for k in order]
    frac = (np.asarray(positions, np.float64)[order] @ np.linalg.inv(np.asarray(lattice, np.float64))
            ).astype(np.float32)
    frc = np.asarray(forces, np.float32)[order]
    V = len(order)
    amax = [int(math.ceil(float(CUTOFF_MAX) / float(_modu(lat[d])))) for d in range(3)]  # :231-233
    ii, jj, kk = np.meshgrid(np.arange(-amax[0], amax[0] + 2), np.arange(-amax[1], amax[1] + 2),
                             np.arange(-amax[2], amax[2] + 2), indexing="ij")   # do i / j / k
    shifts = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float32)
    first_of_spec = [spec_of.index(s) for s in range(len(names))]
    index_list, feats, degree = [], [], np.zeros(V, np.int64)
    for iatom in range(V):
        is_ = spec_of[iatom]
        # spec_loop2 starts at js = is with jatom RESET to 0 (:240-246): for several species
        # the second index is counted from the first atom of species `is`, as the reference does
        for pos_j in range(first_of_spec[is_], V):
            jatom = pos_j - first_of_spec[is_]            # 0-based value of the reference's jatom-1
            if spec_of[pos_j] == is_ and pos_j < iatom:   # is == js and ja < ia: cycle
                continue
            diff = frac[iatom] - frac[pos_j]
            diff = diff - np.ceil(diff - np.float32(0.5))                           # :248-249
            vt = diff[None, :] + shifts                                             # :251-255
            cart = (vt[:, 0:1] * lat[0][None, :] + vt[:, 1:2] * lat[1][None, :]) + \
                vt[:, 2:3] * lat[2][None, :]                                        # matmul(vtmp1, lat)
            r = _modu(cart)
            hit = np.nonzero((r > CUTOFF_MIN) & (r < CUTOFF_MAX))[0]               # :257-258
            degree[iatom] += hit.size
            for h in hit:
                index_list.append((iatom + 1, jatom + 1))
                feats.append(r[h] / CUTOFF_MAX)                                    # :261
    x = np.zeros((V, 6), np.float32)
    for v in range(V):
        charge, mass = ELEMENT_PROPERTIES[names[spec_of[v]]]
        x[v, 0:3] = frc[v]                                                         # :268-272
        x[v, 3] = np.float32(charge) / np.float32(100.0)
        x[v, 4] = np.float32(mass) / np.float32(52.0)
        x[v, 5] = np.float32(degree[v]) / np.float32(6.0)
    il = np.asarray(index_list, np.int32).reshape(-1, 2)
    ef = np.asarray(feats, np.float32).reshape(-1, 1)
    return x, il, ef


def get_graph_from_basis(lattice, species, positions, forces) -> graph_type:
    x, il, ef = basis_edges(lattice, species, positions, forces)
    g = graph_type()
    g.set_num_vertices(x.shape[0], x.shape[1])
    g.vertex_features[:] = x
    g.set_num_edges(il.shape[0], 1)
    g.edge_features[:] = ef
    g.generate_adjacency(il)     # convert_to_sparse + generate_adjacency (:275-276)
    return g


def read_extxyz_db(file: str):
    """-> (graphs, labels[num_samples]) like read_extxyz_db(file, graphs, labels) (:139-190)."""
    with open(file) as f:
        frames = parse_extxyz(f.read())
    graphs = [get_graph_from_basis(fr["lattice"], fr["species"], fr["positions"], fr["forces"])
              for fr in frames]
    labels = np.asarray([fr["energy"] for fr in frames], np.float32)
    return graphs, labels
