"""Data-parallel decomposition on CPU, world_size 2 over gloo.

The N > 1 path is: contiguous graph shards (athena_cuda_shard_graphs, pure host code in
libathena_cuda) -> per-rank loss + gradients with the GLOBAL batch normalisation -> one
all-reduce(sum) of the flat [gradients | loss] buffer -> the identical clip + optimiser
step on every rank.  Here the per-rank arithmetic is the CPU oracle (this is a test: the
oracle is the checker) and the collective is gloo; the result must equal the single-process
full-batch step.  On the GPU the same decomposition is exercised by
test_gpu_parity.py::test_shard_sum_equals_full_batch_gradient and by bench.py --gpus N.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import athena_b200 as ab
from athena_b200 import synth
from helpers import RTOL_ACT, RTOL_PARAM, rel_err
from oracle.oracle import Batch, LayerSpec, OptimSpec, Oracle

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _case(name):
    rng = np.random.default_rng(123)
    if name == "kipf":
        p = synth.regular_batch(24, 16, 2, 8, rng)
        specs = [LayerSpec("kipf", [8, 8], 1, activation="relu"),
                 LayerSpec("kipf", [8, 8], 1, activation="none")]
        target = rng.standard_normal((p.V, 8)).astype(np.float32)
        opt = OptimSpec("sgd", lr=0.05, momentum=0.9)
    else:
        p = synth.molecular_batch(24, 8, 2, rng)
        specs = [LayerSpec("kipf", [8, 8], 1, activation="relu"),
                 LayerSpec("duvenaud", [8] * 3, 2, 2, 1, 4, 5, "sigmoid", "softmax")]
        target = rng.random((p.B, 5)).astype(np.float32)
        opt = OptimSpec("adam", lr=0.01, clip_norm=0.1)
    o = Oracle("f32")
    params = (rng.standard_normal(o.num_params(specs)) * 0.3).astype(np.float32)
    return p, specs, target, opt, params


def _shards(p):
    first = np.zeros(WORLD + 1, np.int32)
    nz64 = p.nz.astype(np.int64)  # named: the buffer must outlive the call
    ab.check(ab.lib().athena_cuda_shard_graphs(p.B, ab.ptr(nz64), WORLD, ab.ptr(first)))
    return first


def _worker(rank, port, name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        p, specs, target, opt, params = _case(name)
        o = Oracle("f32")
        first = _shards(p)
        g0, g1 = int(first[rank]), int(first[rank + 1])
        sp = p.slice(g0, g1)
        voff = np.concatenate([[0], np.cumsum(p.nv)])
        kipf_last = specs[-1].kind == "kipf"
        tgt = target[voff[g0]:voff[g1]] if kipf_last else target[g0:g1]
        s1 = np.zeros_like(params)
        s2 = np.zeros_like(params)
        for it in (1, 2, 3):
            loss, _, grads = o.stack_fwd_bwd(specs, params, Batch(sp.nv, sp.ne, sp.ia, sp.ja, sp.x, sp.e),
                                             tgt, global_B=p.B)
            flat = torch.from_numpy(np.concatenate([grads, [loss]]).astype(np.float32))
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)        # the ONE exchange of the path
            flat = flat.numpy()
            params, s1, s2 = o.update(params, flat[:-1], opt, s1, s2, it)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), params=params, loss=flat[-1])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["kipf", "kipf_duvenaud"])
def test_two_rank_gloo_step_equals_full_batch_step(name, tmp_path):
    mp.spawn(_worker, args=(_free_port(), name, str(tmp_path)), nprocs=WORLD, join=True)
    p, specs, target, opt, params = _case(name)
    o = Oracle("f32")
    s1 = np.zeros_like(params)
    s2 = np.zeros_like(params)
    for it in (1, 2, 3):
        loss, _, grads = o.stack_fwd_bwd(specs, params, Batch(p.nv, p.ne, p.ia, p.ja, p.x, p.e), target)
        params, s1, s2 = o.update(params, grads, opt, s1, s2, it)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["params"], r1["params"]), "replicas diverged"
    assert rel_err(r0["params"], params) <= RTOL_PARAM
    assert abs(float(r0["loss"]) - loss) <= RTOL_ACT * abs(loss) * 10


def test_shards_are_contiguous_and_cover_the_batch():
    p, *_ = _case("kipf_duvenaud")
    first = _shards(p)
    assert first[0] == 0 and first[-1] == p.B and np.all(np.diff(first) > 0)
    z = [int(p.nz[first[r]:first[r + 1]].sum()) for r in range(WORLD)]
    # a boundary overshoots the ideal cut by at most one graph
    assert max(z) - min(z) <= 2 * (int(p.nz.max()) + 1)
