#!/usr/bin/env python
"""Device-resident throughput of the BASELINE.json configs that are NOT the bench.py line
(cfg3 large-graph inference, cfg4 Kipf+Duvenaud molecular training, cfg5 power-law), with
the per-kernel CUDA-event breakdown.  One JSON line per config; results are quoted in
DESIGN.md.  Usage:  python tools/bench_configs.py [cfg3] [cfg4] [cfg5] [cfg1]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athena_b200 as ab  # noqa: E402
from athena_b200 import synth  # noqa: E402

L = ab.lib()
PEAK = 6558.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    ab.check(L.athena_cuda_synchronize())
    ms = C.c_float()
    ab.check(L.athena_cuda_timer_start(1))
    for _ in range(steps):
        fn()
    ab.check(L.athena_cuda_timer_stop(1, C.byref(ms)))
    return ms.value / steps


def kernel_profile(fn, steps=3):
    ab.check(L.athena_cuda_synchronize())
    ab.check(L.athena_cuda_profile_begin())
    for _ in range(steps):
        fn()
    n = C.c_int32()
    ab.check(L.athena_cuda_profile_end(C.byref(n)))
    out = {}
    name = C.create_string_buffer(96)
    cnt = C.c_int64()
    tms = C.c_float()
    for i in range(n.value):
        ab.check(L.athena_cuda_profile_get(i, name, 96, C.byref(cnt), C.byref(tms)))
        out[name.value.decode()] = {"launches_per_step": cnt.value / steps,
                                    "us_per_launch": round(tms.value / cnt.value * 1e3, 1)}
    return out


def cfg3():
    rng = np.random.default_rng(1)
    t0 = time.time()
    p = synth.random_graph(2_000_000, 8, 128, rng)
    gen_s = time.time() - t0
    F = 128
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([F, F], 1, "relu"))
    net.add(ab.kipf_msgpass_layer_type([F, F], 1, "none"))
    net.compile(ab.sgd_optimiser_type(0.01), batch_size=1)
    net.set_params((rng.standard_normal(net.num_params) / np.sqrt(F)).astype(np.float32))
    batch = ab.GraphBatch(p)
    x = ab.DeviceArray.from_host(p.x)
    out = ab.DeviceArray((p.V, F))

    def step():
        ab.check(L.athena_cuda_network_forward(net.handle, batch.handle, ab.ptr(x), None,
                                               ab.ptr(out), ab.MEM_DEVICE))
    ms = timed(step, 20, 5)
    V, Z = p.V, p.Z
    comp = 2 * (4 * (V + 1) + 4 * Z + 4 * V + 8 * V * F)   # SURVEY 8(d), per layer, inference
    return {"config": "cfg3: Kipf 2 x (128->128) inference, one graph, V=2e6", "V": V, "Z": Z,
            "ms_per_step": ms, "edges_per_s": Z / ms * 1e3,
            "compulsory_GBps": comp / ms / 1e6, "frac_of_hbm": comp / ms / 1e6 / PEAK,
            "kernels": kernel_profile(step), "host_generation_s": round(gen_s, 1)}


def cfg4(graphs=8192):
    rng = np.random.default_rng(2)
    p = synth.molecular_batch(graphs, 32, 4, rng)
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([32, 32], 1, "relu"))
    net.add(ab.kipf_msgpass_layer_type([32, 32], 1, "relu"))
    net.add(ab.duvenaud_msgpass_layer_type([32], [4], 2, 6, 32))
    net.compile(ab.adam_optimiser_type(0.001), batch_size=p.B)
    net.set_params((rng.standard_normal(net.num_params) * 0.1).astype(np.float32))
    batch = ab.GraphBatch(p)
    x = ab.DeviceArray.from_host(p.x)
    e = ab.DeviceArray.from_host(p.e)
    t = ab.DeviceArray.from_host(rng.random((p.B, 32)).astype(np.float32))

    def step():
        ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x), ab.ptr(e),
                                                  ab.ptr(t), ab.MEM_DEVICE, p.B, None))
    ms = timed(step, 30, 5)
    return {"config": "cfg4: Kipf(32->32) x 2 + Duvenaud(T=2, D=6, 32 outputs) train, "
                      f"{graphs} molecular graphs per GPU (V~U[10,50], degree<=4+self), Adam",
            "V": p.V, "Z": p.Z, "graphs": p.B, "ms_per_step": ms,
            "graphs_per_s": p.B / ms * 1e3, "edges_per_s": p.Z / ms * 1e3,
            "kernels": kernel_profile(step)}


def cfg5(graphs=64):
    rng = np.random.default_rng(3)
    p = synth.powerlaw_batch(graphs, 16384, 64, rng, max_degree=10000)
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([64, 64], 1, "relu"))
    net.add(ab.kipf_msgpass_layer_type([64, 64], 1, "none"))
    net.compile(ab.sgd_optimiser_type(0.01), batch_size=p.B)
    net.set_params((rng.standard_normal(net.num_params) / 8).astype(np.float32))
    batch = ab.GraphBatch(p)
    x = ab.DeviceArray.from_host(p.x)
    t = ab.DeviceArray.from_host(rng.standard_normal((p.V, 64)).astype(np.float32))

    def step():
        ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x), None,
                                                  ab.ptr(t), ab.MEM_DEVICE, p.B, None))
    ms = timed(step, 20, 5)
    deg = np.diff(np.concatenate([[0], np.cumsum(p.nz)]))  # per graph entries
    return {"config": f"cfg5: Kipf 2 x (64->64) train, {graphs} power-law graphs x 16384 vertices "
                      "(Zipf 2.1, max degree 10000)", "V": p.V, "Z": p.Z, "ms_per_step": ms,
            "edges_per_s": p.Z / ms * 1e3, "max_entries_per_graph": int(deg.max()),
            "kernels": kernel_profile(step)}


def cfg1():
    rng = np.random.default_rng(42)
    p = synth.chemical_batch(8, rng)
    net = ab.network_type()
    net.add(ab.duvenaud_msgpass_layer_type([6], [1], 4, 10, 10))
    net.compile(ab.adam_optimiser_type(0.01, clip_dict=ab.clip_type(clip_norm=0.1)), batch_size=8)
    net.set_params((rng.standard_normal(net.num_params) * 0.3).astype(np.float32))
    batch = ab.GraphBatch(p)
    x = ab.DeviceArray.from_host(p.x)
    e = ab.DeviceArray.from_host(p.e)
    t = ab.DeviceArray.from_host(rng.random((p.B, 10)).astype(np.float32))

    def step():
        ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x), ab.ptr(e),
                                                  ab.ptr(t), ab.MEM_DEVICE, p.B, None))
    ms = timed(step, 200, 20)
    return {"config": "cfg1: chemical Duvenaud (8 graphs x 8 atoms, T=4, D=10, 10 outputs), Adam + "
                      "clip_norm, one train step", "V": p.V, "Z": p.Z, "us_per_step": ms * 1e3,
            "graphs_per_s": p.B / ms * 1e3, "kernels": kernel_profile(step),
            "note": "latency-bound (tens of kB per step)"}


if __name__ == "__main__":
    ab.check(L.athena_cuda_init(0))
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5", "cfg1"]
    for w in which:
        print(json.dumps({"cfg1": cfg1, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}[w]()), flush=True)
