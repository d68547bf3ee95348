"""athena_b200 -- B200 (sm_100a) implementation of athena's graph message-passing
hot path behind athena's own layer / network API.

Product code: the CUDA kernels and C ABI in csrc/ (libathena_cuda.so) and this
thin ctypes mirror of the reference's Fortran interface.  The CPU oracle under
/oracle is test infrastructure and is never imported from here.
"""
from ._lib import (ACT, AthenaCudaError, DeviceArray, LIB_PATH, MEM_DEVICE, MEM_HOST, build,
                   check, lib, pinned_empty, ptr)
from .graph import PackedGraphs, graph_type, pack_graphs
from .layers import (GraphBatch, duvenaud_msgpass_layer_type, full_layer_type,
                     kipf_msgpass_layer_type, msgpass_layer_type)
from .network import (adagrad_optimiser_type, adam_optimiser_type, base_lr_decay_type,
                      base_optimiser_type, clip_type, exp_lr_decay_type, inv_lr_decay_type,
                      l1_regulariser_type, l1l2_regulariser_type, l2_regulariser_type,
                      network_type, rmsprop_optimiser_type, sgd_optimiser_type,
                      step_lr_decay_type)

__all__ = [
    "ACT", "AthenaCudaError", "DeviceArray", "LIB_PATH", "MEM_DEVICE", "MEM_HOST", "build",
    "check", "lib", "pinned_empty", "ptr", "PackedGraphs", "graph_type", "pack_graphs",
    "GraphBatch", "duvenaud_msgpass_layer_type", "full_layer_type", "kipf_msgpass_layer_type",
    "msgpass_layer_type", "adagrad_optimiser_type", "adam_optimiser_type", "base_optimiser_type",
    "clip_type", "l1_regulariser_type", "l1l2_regulariser_type", "l2_regulariser_type",
    "network_type", "rmsprop_optimiser_type", "sgd_optimiser_type", "base_lr_decay_type",
    "exp_lr_decay_type", "step_lr_decay_type", "inv_lr_decay_type",
]
