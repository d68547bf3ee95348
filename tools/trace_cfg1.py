#!/usr/bin/env python
"""Phase trace of the Duvenaud tile kernels on the one-tile cfg1 batch (see tools/trace_cfg4.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import athena_b200 as ab  # noqa: E402
from athena_b200 import synth  # noqa: E402

L = ab.lib()
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
for i in range(3):
    ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x), ab.ptr(e), ab.ptr(t),
                                              ab.MEM_DEVICE, p.B, None))
ab.check(L.athena_cuda_synchronize())
