#!/usr/bin/env python
"""example/msgpass_euler/src/main.f90 on the device: the skip-connected 7-layer Kipf network
(every layer after the first reads [input | previous layer], softmax message activations, a
swish head) trained on the bump-channel Euler data set with Adam (lr 2e-2, exp decay 1e-3,
clip(-1, 1)), batch of 2 graphs.  Same call sequence as the Fortran program; the mesh comes
from tests/golden/euler_bump.npz (packed from the reference's data files).

    python examples/msgpass_euler.py [num_epochs]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athena_b200 as ab  # noqa: E402


def read_graphs():
    d = np.load(os.path.join(ROOT, "tests", "golden", "euler_bump.npz"))
    graphs_in, graphs_out = [], []
    for s in (1, 2):                       # read_graph(vertex_file, edge_file, graph)
        g = ab.graph_type()
        g.set_num_vertices(*d[f"in_{s}"].shape)
        g.vertex_features[:] = d[f"in_{s}"]
        g.set_num_edges(d["index_list"].shape[0])
        g.generate_adjacency(d["index_list"])
        graphs_in.append(g)
        graphs_out.append(d[f"out_{s}"])
    return graphs_in, np.concatenate(graphs_out)


def main():
    num_epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ab.check(ab.lib().athena_cuda_init(-1))
    graphs_in, target = read_graphs()
    network = ab.network_type()
    network.add(ab.kipf_msgpass_layer_type(num_time_steps=1, num_vertex_features=[3, 6],
                                           activation="softmax"))
    for nvf, act in (([9, 14], "softmax"), ([17, 32], "softmax"), ([35, 64], "softmax"),
                     ([67, 32], "softmax"), ([35, 14], "softmax"), ([17, 7], "swish")):
        network.add(ab.kipf_msgpass_layer_type(num_time_steps=1, num_vertex_features=nvf,
                                               activation=act),
                    input_list=[0, -1], operator="concatenate")
    network.compile(optimiser=ab.adam_optimiser_type(clip_dict=ab.clip_type(-1.0, 1.0),
                                                     learning_rate=2e-2,
                                                     lr_decay=ab.exp_lr_decay_type(1e-3)),
                    loss_method="mse", batch_size=2)
    rng = np.random.default_rng(1)
    params = np.concatenate([rng.standard_normal(fi * fo) * np.sqrt(2.0 / fi)      # he_normal
                             for fi, fo in ((3, 6), (9, 14), (17, 32), (35, 64), (67, 32),
                                            (35, 14), (17, 7))]).astype(np.float32)
    network.set_params(params)
    print("NUMBER OF LAYERS", network.num_layers, " Number of parameters", network.num_params)
    history = network.train(graphs_in, target, num_epochs=num_epochs, batch_size=2, resident=True)
    for epoch, loss in enumerate(history, 1):
        if epoch == 1 or epoch % 5 == 0 or epoch == len(history):
            print(f"epoch={epoch}, loss={loss:.6f}")
    pred = network.predict(graphs_in[:1])
    print("predicted", pred[:2].round(4).tolist())
    print("expected ", target[:2].round(4).tolist())
    assert np.isfinite(history).all() and history[-1] < history[0]


if __name__ == "__main__":
    main()
