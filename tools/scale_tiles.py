#!/usr/bin/env python
"""How much of a fused cfg2 kernel is per-launch ramp / drain and how much is per tile:
the per-kernel CUDA-event times of the training step at 1x, 2x and 4x the cfg2 batch."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athena_b200 as ab  # noqa: E402
from athena_b200 import synth  # noqa: E402

L = ab.lib()
ab.check(L.athena_cuda_init(0))
F = 64
for graphs in (2048, 4096, 8192, 16384):
    rng = np.random.default_rng(1)
    p = synth.regular_batch(graphs, 64, 6, F, rng)
    target = rng.standard_normal((p.V, F), dtype=np.float32)
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([F, F], 1, "relu"))
    net.add(ab.kipf_msgpass_layer_type([F, F], 1, "none"))
    net.compile(ab.sgd_optimiser_type(0.01), batch_size=p.B)
    net.set_params((rng.standard_normal(net.num_params) / 8).astype(np.float32))
    x_d = ab.DeviceArray.from_host(p.x)
    t_d = ab.DeviceArray.from_host(target)
    batch = ab.GraphBatch(p)

    def step():
        ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x_d), None,
                                                  ab.ptr(t_d), ab.MEM_DEVICE, p.B, None))
    for _ in range(5):
        step()
    ab.check(L.athena_cuda_synchronize())
    ms = C.c_float()
    ab.check(L.athena_cuda_timer_start(1))
    for _ in range(20):
        step()
    ab.check(L.athena_cuda_timer_stop(1, C.byref(ms)))
    ab.check(L.athena_cuda_profile_begin())
    for _ in range(5):
        step()
    n = C.c_int32()
    ab.check(L.athena_cuda_profile_end(C.byref(n)))
    ks = {}
    name = C.create_string_buffer(96)
    cnt = C.c_int64()
    tms = C.c_float()
    for i in range(n.value):
        ab.check(L.athena_cuda_profile_get(i, name, 96, C.byref(cnt), C.byref(tms)))
        ks[name.value.decode()] = round(tms.value / cnt.value * 1e3, 1)
    print(json.dumps({"graphs": graphs, "tiles_per_cta": graphs / 2 / 148,
                      "us_per_step": round(ms.value / 20 * 1e3, 1), "kernels_us": ks}))
    net.destroy(); batch.destroy(); x_d.free(); t_d.free()
