#!/usr/bin/env python
"""example/msgpass_chemical/src/main.f90 on the device: energies of 198 periodic 8-atom carbon
cells from a Duvenaud fingerprint (T = 4, degree buckets 1..10, 10 outputs) and a dense head
128 -> 64 -> 1 (leaky_relu), Adam lr 1e-2 with clip_norm 0.1, batches of 8.  The cells come from
tests/golden/chemical_database.npz (database.xyz of the reference), the graphs are built as
mod_read_chemical_graphs.f90 builds them, generate_adjacency + add_self_loops run on the device.

    python examples/msgpass_chemical.py [num_epochs]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athena_b200 as ab  # noqa: E402
from athena_b200.read_chemical_graphs import get_graph_from_basis  # noqa: E402


def main():
    num_epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ab.check(ab.lib().athena_cuda_init(-1))
    d = np.load(os.path.join(ROOT, "tests", "golden", "chemical_database.npz"))
    graphs = [get_graph_from_basis(d["lattice"][s], ["C"] * 8, d["positions"][s], d["forces"][s])
              for s in range(d["energy"].size)]
    for g in graphs:
        g.add_self_loops()
    energy = d["energy"].astype(np.float32)
    output = ((energy - energy.min()) / (energy.max() - energy.min())).reshape(-1, 1)

    network = ab.network_type()
    network.add(ab.duvenaud_msgpass_layer_type(num_time_steps=4, num_vertex_features=[6],
                                               num_edge_features=[1], num_outputs=10,
                                               readout_activation="softmax",
                                               min_vertex_degree=1, max_vertex_degree=10))
    network.add(ab.full_layer_type(num_inputs=10, num_outputs=128, activation="leaky_relu"))
    network.add(ab.full_layer_type(num_outputs=64, activation="leaky_relu"))
    network.add(ab.full_layer_type(num_outputs=1, activation="leaky_relu"))
    network.compile(optimiser=ab.adam_optimiser_type(clip_dict=ab.clip_type(clip_norm=0.1),
                                                     learning_rate=1e-2),
                    loss_method="mse", batch_size=8)
    rng = np.random.default_rng(1)
    network.set_params((rng.standard_normal(network.num_params) * 0.1).astype(np.float32))
    print("Number of layers:", network.num_layers, " Number of parameters:", network.num_params,
          " Number of samples:", len(graphs))
    history = network.train(graphs, output, num_epochs=num_epochs, batch_size=8,
                            shuffle_batches=True, resident=True)
    for epoch, loss in enumerate(history, 1):
        if epoch == 1 or epoch % 5 == 0 or epoch == len(history):
            print(f"epoch={epoch}, loss={loss:.6f}")
    pred = network.predict(graphs[:8])
    print("predicted", pred.ravel().round(3).tolist())
    print("expected ", output[:8].ravel().round(3).tolist())
    assert np.isfinite(history).all() and history[-1] < history[0]


if __name__ == "__main__":
    main()
