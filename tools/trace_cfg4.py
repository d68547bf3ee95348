#!/usr/bin/env python
"""Phase trace of the Duvenaud tile kernels (debugging build: make OUT=../lib_trace OBJ=_obj_trace
EXTRA=-DTF_TRACE in athena_b200/csrc, run with ATHENA_CUDA_LIB=athena_b200/lib_trace/libathena_cuda.so):
block 0 / thread 0 prints the clock64 distance between phase boundaries of its first tiles."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import athena_b200 as ab  # noqa: E402
from athena_b200 import synth  # noqa: E402

L = ab.lib()
rng = np.random.default_rng(2)
G = 8192
q = synth.molecular_batch(G, 32, 4, rng)
tgt = rng.random((G, 32)).astype(np.float32)
net = ab.network_type()
net.add(ab.kipf_msgpass_layer_type([32, 32], 1, "relu"))
net.add(ab.kipf_msgpass_layer_type([32, 32], 1, "relu"))
net.add(ab.duvenaud_msgpass_layer_type([32], [4], 2, 6, 32))
net.compile(ab.adam_optimiser_type(0.001), batch_size=q.B)
net.set_params((np.random.default_rng(7).standard_normal(net.num_params) * 0.1).astype(np.float32))
batch = ab.GraphBatch(q)
x = ab.DeviceArray.from_host(q.x)
e = ab.DeviceArray.from_host(q.e)
t = ab.DeviceArray.from_host(tgt)
for i in range(2):
    ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x), ab.ptr(e), ab.ptr(t),
                                              ab.MEM_DEVICE, G, None))
ab.check(L.athena_cuda_synchronize())
