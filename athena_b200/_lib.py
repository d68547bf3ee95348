"""ctypes binding of libathena_cuda (include/athena_cuda.h).

This is the only place the Python mirror touches native code.  There is no
CPU fallback: if the shared library is missing, or no sm_100 device is
usable, calls raise AthenaCudaError.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ATHENA_CUDA_LIB: load another build of the library (A/B experiments on the GPU box)
LIB_PATH = os.environ.get("ATHENA_CUDA_LIB") or os.path.join(_HERE, "lib", "libathena_cuda.so")
CSRC = os.path.join(_HERE, "csrc")

MEM_HOST, MEM_DEVICE = 0, 1
ACT = {"none": 0, "linear": 1, "relu": 2, "leaky_relu": 3, "sigmoid": 4, "tanh": 5, "softmax": 6,
       "swish": 7}
OPT_SGD, OPT_ADAM, OPT_RMSPROP, OPT_ADAGRAD = 0, 1, 2, 3
REG_NONE, REG_L1, REG_L2, REG_L1L2 = 0, 1, 2, 3
COMM_ID_BYTES = 128
P2P_HANDLE_BYTES = 128

BATCH_FIELDS = {"row_ptr": 0, "col": 1, "eid": 2, "deg": 3, "vgraph": 4, "csc_ptr": 5,
                "csc_src": 6, "csc_ent": 7, "bucket": 8, "perm": 9, "bucket_ptr": 10, "coef": 11}


class AthenaCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libathena_cuda error {code}: {msg}")
        self.code = code


class OptimiserDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("learning_rate", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("epsilon", C.c_float), ("momentum", C.c_float),
                ("nesterov", C.c_int32), ("clip_min_max", C.c_int32), ("clip_min", C.c_float),
                ("clip_max", C.c_float), ("clip_norm_on", C.c_int32), ("clip_norm", C.c_float),
                ("regulariser", C.c_int32), ("l1", C.c_float), ("l2", C.c_float),
                ("l2_decoupled", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile libathena_cuda.so for sm_100a with nvcc (in-tree)."""
    cmd = ["make", "-C", CSRC, "-s", "-j8"] + (["-B"] if force else [])
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AthenaCudaError(-1, f"{LIB_PATH} not built; run athena_b200.build() "
                                  "(python -c 'import __graft_entry__ as g; g.build()')")
    L = C.CDLL(LIB_PATH)
    H, I32, I64, F32, P = C.c_int64, C.c_int32, C.c_int64, C.c_float, C.c_void_p
    # every pointer parameter is declared void* so that numpy buffers, device addresses and
    # byref(scalar) are all accepted
    PH = PI32 = PI64 = PF = P
    sig = {
        "athena_cuda_init": [I32],
        "athena_cuda_shutdown": [],
        "athena_cuda_version": [PI32, PI32],
        "athena_cuda_device_info": [PI32, PI32, PI64],
        "athena_cuda_synchronize": [],
        "athena_cuda_malloc": [P, C.c_size_t],
        "athena_cuda_free": [P],
        "athena_cuda_host_alloc": [P, C.c_size_t],
        "athena_cuda_host_free": [P],
        "athena_cuda_memcpy_h2d": [P, P, C.c_size_t],
        "athena_cuda_memcpy_d2h": [P, P, C.c_size_t],
        "athena_cuda_memset": [P, C.c_int, C.c_size_t],
        "athena_cuda_timer_start": [I32],
        "athena_cuda_timer_stop": [I32, PF],
        "athena_cuda_launch_count": [PI64],
        "athena_cuda_flush_l2": [],
        "athena_cuda_profile_begin": [],
        "athena_cuda_profile_end": [PI32],
        "athena_cuda_profile_get": [I32, P, I32, PI64, PF],
        "athena_cuda_batch_create": [PH, I32, P, P, P, P, P, I32, I32],
        "athena_cuda_batch_create_from_edges": [PH, I32, P, P, P, P, I32, I32, I32],
        "athena_cuda_batch_create_from_edge_index": [PH, I32, P, P, P, P, P, I32, I32],
        "athena_cuda_batch_destroy": [H],
        "athena_cuda_batch_status": [H],
        "athena_cuda_batch_info": [H, PI32, PI64, PI64, PI64],
        "athena_cuda_batch_bucketize": [H, I32, I32],
        "athena_cuda_batch_export": [H, I32, P, I64],
        "athena_cuda_kipf_layer_create": [PH, I32, P, I32],
        "athena_cuda_duvenaud_layer_create": [PH, I32, P, I32, I32, I32, I32, I32, I32],
        "athena_cuda_full_layer_create": [PH, I32, I32, I32, I32],
        "athena_cuda_layer_destroy": [H],
        "athena_cuda_layer_num_params": [H, PI64],
        "athena_cuda_layer_set_params": [H, P, I64],
        "athena_cuda_layer_get_params": [H, P, I64],
        "athena_cuda_layer_set_gradients": [H, P, I64],
        "athena_cuda_layer_get_gradients": [H, P, I64],
        "athena_cuda_layer_zero_gradients": [H],
        "athena_cuda_layer_forward": [H, H, P, P, P, I32],
        "athena_cuda_layer_backward": [H, H, P, P, I32],
        "athena_cuda_layer_backward_stage": [H, H, I32, P, I64],
        "athena_cuda_layer_backward_flush": [H, H],
        "athena_cuda_network_create": [PH],
        "athena_cuda_network_destroy": [H],
        "athena_cuda_network_add": [H, H],
        "athena_cuda_network_add_inputs": [H, H, I32, P, I32],
        "athena_cuda_network_compile": [H, P],
        "athena_cuda_network_num_params": [H, PI64],
        "athena_cuda_network_set_params": [H, P, I64],
        "athena_cuda_network_get_params": [H, P, I64],
        "athena_cuda_network_get_gradients": [H, P, I64],
        "athena_cuda_network_set_learning_rate": [H, F32],
        "athena_cuda_network_set_iteration": [H, I64],
        "athena_cuda_network_forward": [H, H, P, P, P, I32],
        "athena_cuda_network_train_step": [H, H, P, P, P, I32, I32, PF],
        "athena_cuda_network_loss_and_gradients": [H, H, P, P, P, I32, I32, PF],
        "athena_cuda_network_update": [H],
        "athena_cuda_network_last_loss": [H, PF],
        "athena_cuda_comm_unique_id": [P],
        "athena_cuda_comm_init": [I32, I32, P],
        "athena_cuda_comm_p2p_export": [P],
        "athena_cuda_comm_p2p_import": [I32, I32, P],
        "athena_cuda_comm_destroy": [],
        "athena_cuda_comm_info": [PI32, PI32],
        "athena_cuda_shard_graphs": [I32, P, I32, P],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.athena_cuda_last_error.argtypes = []
    L.athena_cuda_last_error.restype = C.c_char_p
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise AthenaCudaError(rc, lib().athena_cuda_last_error().decode())


def ptr(a):
    """void* of a numpy array (must be C-contiguous), an int device address, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags.c_contiguous, "array must be C-contiguous"
        # data_as keeps a reference to the array, so a temporary passed inline stays alive
        # until the foreign call has returned
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, DeviceArray):
        return C.c_void_p(a.addr)
    return C.c_void_p(int(a))


class DeviceArray:
    """A device allocation owned by Python (tests / device-resident bench leg)."""

    def __init__(self, shape, dtype=np.float32):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib().athena_cuda_malloc(C.byref(p), max(self.nbytes, 1)))
        self.addr = p.value

    @classmethod
    def from_host(cls, a: np.ndarray) -> "DeviceArray":
        a = np.ascontiguousarray(a)
        d = cls(a.shape, a.dtype)
        check(lib().athena_cuda_memcpy_h2d(C.c_void_p(d.addr), ptr(a), a.nbytes))
        check(lib().athena_cuda_synchronize())
        return d

    def to_host(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        check(lib().athena_cuda_memcpy_d2h(ptr(out), C.c_void_p(self.addr), self.nbytes))
        return out

    def free(self):
        if self.addr:
            lib().athena_cuda_free(C.c_void_p(self.addr))
            self.addr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array backed by page-locked host memory (fast async H2D)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib().athena_cuda_host_alloc(C.byref(p), max(n, 1)))
    buf = (C.c_byte * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr
