#!/usr/bin/env python
"""Multi-GPU correctness check of the data-parallel path on real GPUs (not part of pytest:
the round-end GPU tests run on one GPU).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp.py

Every rank builds the same global batch, takes its contiguous graph shard
(athena_cuda_shard_graphs) and trains 4 steps three times from the same initial
parameters: (1) gradient exchange fused with the step over peer memory, (2) NCCL all-reduce,
(3) rank 0 alone on the full batch (no exchange).  (1) and (2) must agree to rounding, both
must match (3) within the parameter tolerance of the north star (1e-4 relative), and the
replicas must be bitwise identical across ranks.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athena_b200 as ab  # noqa: E402
from athena_b200 import synth  # noqa: E402

F = 64


def build_net(kind, params):
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([F, F], 1, "relu"))
    if kind == "kipf":
        net.add(ab.kipf_msgpass_layer_type([F, F], 1, "none"))
    else:
        net.add(ab.duvenaud_msgpass_layer_type([F, 32, 32], [0], 2, 8, 16))
    net.compile(ab.sgd_optimiser_type(0.05, momentum=0.9), batch_size=1)
    if params is not None:
        net.set_params(params)
    return net


def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    L = ab.lib()
    ab.check(L.athena_cuda_init(local))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def nccl_init():
        idbuf = [None]
        if rank == 0:
            raw = C.create_string_buffer(ab._lib.COMM_ID_BYTES)
            ab.check(L.athena_cuda_comm_unique_id(raw))
            idbuf = [raw.raw]
        dist.broadcast_object_list(idbuf, src=0)
        ab.check(L.athena_cuda_comm_init(world, rank, C.create_string_buffer(idbuf[0], 128)))

    def p2p_init():
        mine = C.create_string_buffer(ab._lib.P2P_HANDLE_BYTES)
        ab.check(L.athena_cuda_comm_p2p_export(mine))
        allh = [None] * world
        dist.all_gather_object(allh, mine.raw)
        ab.check(L.athena_cuda_comm_p2p_import(world, rank, C.create_string_buffer(b"".join(allh))))

    ok = True
    for kind in ("kipf", "kipf_duvenaud"):
        rng = np.random.default_rng(5)
        p = synth.regular_batch(64 * world, 64, 6, F, rng)
        first = np.zeros(world + 1, np.int32)
        nz64 = p.nz.astype(np.int64)  # named: the buffer must outlive the call
        ab.check(L.athena_cuda_shard_graphs(p.B, ab.ptr(nz64), world, ab.ptr(first)))
        g0, g1 = int(first[rank]), int(first[rank + 1])
        voff = np.concatenate([[0], np.cumsum(p.nv)])
        if kind == "kipf":
            target = rng.standard_normal((p.V, F)).astype(np.float32)
            tgt_local = target[voff[g0]:voff[g1]]
        else:
            target = rng.random((p.B, 16)).astype(np.float32)
            tgt_local = target[g0:g1]
        n = build_net(kind, None).num_params
        params0 = (rng.standard_normal(n) * 0.1).astype(np.float32)
        shard = p.slice(g0, g1)
        results = {}
        for mode in ("p2p", "nccl"):
            nccl_init()
            if mode == "p2p":
                p2p_init()
            net = build_net(kind, params0)
            batch = ab.GraphBatch(shard)
            losses = [net.train_step(batch, tgt_local, global_batch=p.B) for _ in range(4)]
            results[mode] = (net.get_params(), losses)
            ab.check(L.athena_cuda_synchronize())
            net.destroy()
            batch.destroy()
            ab.check(L.athena_cuda_comm_destroy())
        # replicas identical across ranks (bitwise)
        mine = torch.from_numpy(results["p2p"][0].copy()).cuda()
        allp = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        same = all(torch.equal(allp[0], t) for t in allp)
        if rank == 0:
            net = build_net(kind, params0)
            batch = ab.GraphBatch(p)
            ref_losses = [net.train_step(batch, target) for _ in range(4)]
            ref = net.get_params()
            e_p2p, e_nccl = rel(results["p2p"][0], ref), rel(results["nccl"][0], ref)
            e_pn = rel(results["p2p"][0], results["nccl"][0])
            e_loss = abs(results["p2p"][1][-1] - ref_losses[-1]) / abs(ref_losses[-1])
            good = same and e_p2p <= 1e-4 and e_nccl <= 1e-4 and e_pn <= 1e-5 and e_loss <= 1e-5
            ok = ok and good
            print(f"{kind}: world {world}: params rel err vs single-GPU full batch: p2p {e_p2p:.2e}, "
                  f"nccl {e_nccl:.2e}; p2p vs nccl {e_pn:.2e}; global loss rel err {e_loss:.2e}; "
                  f"replicas bitwise identical: {same} -> {'OK' if good else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
