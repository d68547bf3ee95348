"""Generate golden vectors from the REFERENCE's own Python code.

Run in the development container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/duvenaud_ref_*.npz, tests/golden/adam_ref.npz

1. Duvenaud forward + gradients: imports the reference's PyTorch restatement of
   its Duvenaud network, example/msgpass_chemical/pytorch_network.py
   (DuvenaudMPNN, DuvenaudLayer, csr_to_pyg), UNMODIFIED.  That file imports
   torch_geometric, which is not installed here, so a ~40-line stand-in with
   PyG's documented semantics is registered in sys.modules first:
     MessagePassing(aggr='add', flow='source_to_target').propagate(edge_index, x=, edge_attr=)
        -> update(aggregate(message(x_j = x[edge_index[0]], edge_attr), index = edge_index[1]))
     utils.scatter(src, index, dim=0, dim_size, reduce='sum') -> index_add
   Gradients come from torch autograd through the reference's code with
   loss = sum(out * G), i.e. upstream gradient G.

2. Adam: torch.optim.Adam on the two analytic problems of
   example/adam_benchmark/compare_fortran_pytorch.py:14-20,52-71 (the reference
   declares its Adam equivalent to torch's on these).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_pyg_stub():
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_utils = types.ModuleType("torch_geometric.utils")
    tg_loader = types.ModuleType("torch_geometric.loader")
    tg_data = types.ModuleType("torch_geometric.data")

    class MessagePassing(torch.nn.Module):
        def __init__(self, aggr="add", flow="source_to_target"):
            super().__init__()
            assert aggr == "add" and flow == "source_to_target"

        def propagate(self, edge_index, size=None, **kw):
            x = kw["x"]
            x_j = x[edge_index[0]]
            msg = self.message(x_j, kw["edge_attr"])
            out = self.aggregate(msg, edge_index[1], dim_size=x.shape[0])
            return self.update(out)

        def aggregate(self, inputs, index, ptr=None, dim_size=None):
            out = torch.zeros(dim_size, inputs.shape[1], dtype=inputs.dtype)
            return out.index_add(0, index, inputs)

    def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
        assert dim == 0 and reduce == "sum"
        out = torch.zeros(dim_size, src.shape[1], dtype=src.dtype)
        return out.index_add(0, index, src)

    class Data:  # only constructed by the loader helpers we do not call
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class DataLoader:
        pass

    tg_nn.MessagePassing = MessagePassing
    tg_utils.scatter = scatter
    tg_loader.DataLoader = DataLoader
    tg_data.Data = Data
    tg.nn, tg.utils, tg.loader, tg.data = tg_nn, tg_utils, tg_loader, tg_data
    for m in (tg, tg_nn, tg_utils, tg_loader, tg_data):
        sys.modules[m.__name__] = m


def load_reference_module():
    install_pyg_stub()
    path = os.path.join(REF, "example/msgpass_chemical/pytorch_network.py")
    spec = importlib.util.spec_from_file_location("athena_ref_pytorch_network", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def random_graph(rng, nv, p_edge, fv, fe):
    """Undirected graph in athena's CSR convention: every undirected edge gets
    one edge-feature column shared by both directions; one self-loop per vertex
    with its OWN edge-feature column (so no edge id is <= 0, which is undefined
    in the reference)."""
    pairs = [(i, j) for i in range(nv) for j in range(i + 1, nv) if rng.random() < p_edge]
    # keep it connected enough: chain
    for i in range(nv - 1):
        if (i, i + 1) not in pairs:
            pairs.append((i, i + 1))
    pairs.sort()
    ne = len(pairs) + nv
    rows = [[] for _ in range(nv)]
    for k, (i, j) in enumerate(pairs):
        rows[i].append((j + 1, k + 1))
        rows[j].append((i + 1, k + 1))
    for v in range(nv):
        rows[v].append((v + 1, len(pairs) + v + 1))
        rows[v].sort()
    ia = [1]
    ja = []
    for v in range(nv):
        ja.extend(rows[v])
        ia.append(ia[-1] + len(rows[v]))
    x = rng.random((nv, fv), dtype=np.float32)
    e = rng.random((ne, fe), dtype=np.float32)
    return dict(nv=nv, ne=ne, ia=np.array(ia, np.int32), ja=np.array(ja, np.int32), x=x, e=e)


def duvenaud_case(mod, name, seed, n_graphs, nv_range, p_edge, fv, fe, T, n_out, min_deg, max_deg):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    graphs = [random_graph(rng, int(rng.integers(nv_range[0], nv_range[1] + 1)), p_edge, fv, fe)
              for _ in range(n_graphs)]
    net = mod.DuvenaudMPNN(fv, fe, max_deg, min_degree=min_deg, num_timesteps=T, num_outputs=n_out)
    D = max_deg - min_deg + 1
    w = [rng.normal(0, 0.6, fv * (fv + fe) * D).astype(np.float32) for _ in range(T)]
    r = [rng.normal(0, 0.8, n_out * fv).astype(np.float32) for _ in range(T)]
    with torch.no_grad():
        for t in range(T):
            net.layers[t].weight.copy_(torch.from_numpy(w[t]))
            net.readout_weights[t].copy_(torch.from_numpy(r[t]))
    G = rng.normal(0, 1, (n_graphs, n_out)).astype(np.float32)
    outs, dxs = [], []
    for p in net.parameters():
        p.grad = None
    for s, g in enumerate(graphs):
        vf = torch.from_numpy(g["x"].T.copy())            # [fv, nv]  (Fortran val(F,V))
        ef = torch.from_numpy(g["e"].T.copy())            # [fe, ne]
        ia = torch.from_numpy(g["ia"].astype(np.int64))
        ja = torch.from_numpy(g["ja"].T.copy().astype(np.int64))  # [2, Z]
        x, ei, ea, nd = mod.csr_to_pyg(vf, ef, ia, ja)
        x = x.clone().requires_grad_(True)
        out = net(x, ei, ea, nd)                          # [1, n_out]
        (out * torch.from_numpy(G[s:s + 1])).sum().backward()   # grads accumulate over samples
        outs.append(out.detach().numpy()[0])
        dxs.append(x.grad.detach().numpy().copy())
    params = np.concatenate(w + r)
    dparams = np.concatenate([net.layers[t].weight.grad.numpy() for t in range(T)] +
                             [net.readout_weights[t].grad.numpy() for t in range(T)])
    np.savez(os.path.join(OUT, f"duvenaud_ref_{name}.npz"),
             nv=np.array([g["nv"] for g in graphs], np.int32),
             ne=np.array([g["ne"] for g in graphs], np.int32),
             ia=np.concatenate([g["ia"] for g in graphs]),
             ja=np.concatenate([g["ja"] for g in graphs]),
             x=np.concatenate([g["x"] for g in graphs]),
             e=np.concatenate([g["e"] for g in graphs]),
             params=params, g_out=G, out=np.stack(outs), dparams=dparams,
             dx=np.concatenate(dxs),
             hyper=np.array([fv, fe, T, n_out, min_deg, max_deg], np.int32))
    print(name, "V=", sum(g["nv"] for g in graphs), "params=", params.size)


def adam_case():
    lr, b1, b2, eps, n = 0.01, 0.9, 0.999, 1e-8, 20

    def run(x0, loss_fn):
        x = torch.tensor(np.array(x0, np.float32)).requires_grad_(True)
        opt = torch.optim.Adam([x], lr=lr, betas=(b1, b2), eps=eps)
        hp, hg = [], []
        for _ in range(n):
            opt.zero_grad()
            loss_fn(x).backward()
            hg.append(x.grad.numpy().copy())
            opt.step()
            hp.append(x.detach().numpy().copy())
        return np.stack(hp), np.stack(hg)

    p1, g1 = run([0.0], lambda x: (x[0] - 3.0) ** 2)
    p2, g2 = run([0.0, 2.0], lambda x: (x[0] - 3.0) ** 2 + (x[1] + 1.0) ** 2)
    np.savez(os.path.join(OUT, "adam_ref.npz"), scalar_params=p1, scalar_grads=g1,
             multi_params=p2, multi_grads=g2, hyper=np.array([lr, b1, b2, eps], np.float32))
    print("adam", p1[-1], p2[-1])


if __name__ == "__main__":
    torch.set_num_threads(1)
    mod = load_reference_module()
    # dims of example/msgpass_chemical (main.f90:129-138): Fv=6, Fe=1, T=4, D=10, n_out=10
    duvenaud_case(mod, "chem", 42, n_graphs=4, nv_range=(6, 9), p_edge=0.7, fv=6, fe=1, T=4,
                  n_out=10, min_deg=1, max_deg=10)
    # min_degree > 1 exercises the bucket-index divide (differs from degree divide)
    duvenaud_case(mod, "mindeg2", 7, n_graphs=3, nv_range=(5, 12), p_edge=0.35, fv=4, fe=2, T=2,
                  n_out=3, min_deg=2, max_deg=5)
    adam_case()
