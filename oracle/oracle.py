"""ctypes binding of the CPU oracle (oracle/athena_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(athena_b200) never imports this module.

All arrays use the Fortran memory order of the reference, i.e. a feature
matrix val(F, V) is passed as a C-contiguous numpy array of shape [V, F].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

ACT = {"none": 0, "linear": 1, "relu": 2, "leaky_relu": 3, "sigmoid": 4,
       "tanh": 5, "softmax": 6, "swish": 7}


def build(force: bool = False) -> None:
    """Compile the three oracle variants with gcc (seconds)."""
    need = force or not all(
        os.path.exists(os.path.join(_BUILD, f"liboracle_{k}.so")) for k in ("f32", "f64", "fast"))
    if not need:
        src = os.path.getmtime(os.path.join(_HERE, "athena_oracle.c"))
        need = any(os.path.getmtime(os.path.join(_BUILD, f"liboracle_{k}.so")) < src
                   for k in ("f32", "f64", "fast"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


class _LayerT(C.Structure):
    _fields_ = [("kind", C.c_int), ("T", C.c_int), ("nvf", C.c_int * 17),
                ("nef", C.c_int), ("min_deg", C.c_int), ("max_deg", C.c_int),
                ("n_out", C.c_int), ("act", C.c_int), ("ract", C.c_int), ("use_bias", C.c_int),
                ("n_in", C.c_int), ("inputs", C.c_int * 4)]


@dataclass
class LayerSpec:
    """Mirror of the constructor arguments of kipf_msgpass_layer_type
    (athena_kipf_msgpass_layer.f90:80-97) / duvenaud_msgpass_layer_type
    (athena_duvenaud_msgpass_layer.f90:88-120)."""
    kind: str                      # "kipf" | "duvenaud" | "full" (num_vertex_features = [n_in, n_out])
    num_vertex_features: Sequence[int]  # (0:T)
    num_time_steps: int
    num_edge_features: int = 0
    min_vertex_degree: int = 1
    max_vertex_degree: int = 1
    num_outputs: int = 0
    activation: str = "none"
    readout_activation: str = "softmax"
    use_bias: bool = True
    # network%add(layer, input_list, operator='concatenate'): sources whose vertex features are
    # concatenated (in order) to form this layer's input; -1 = the network input, k >= 0 = layer
    # k of the stack.  None: the previous layer.
    inputs: Optional[Sequence[int]] = None

    def to_c(self) -> _LayerT:
        L = _LayerT()
        L.kind = {"kipf": 0, "duvenaud": 1, "full": 2}[self.kind]
        L.use_bias = int(self.use_bias)
        L.T = self.num_time_steps
        nvf = list(self.num_vertex_features)
        if len(nvf) == 1:
            nvf = nvf * (self.num_time_steps + 1)
        assert len(nvf) == self.num_time_steps + 1 <= 17
        for i, f in enumerate(nvf):
            L.nvf[i] = f
        L.nef = self.num_edge_features
        L.min_deg = self.min_vertex_degree
        L.max_deg = self.max_vertex_degree
        L.n_out = self.num_outputs
        L.act = ACT[self.activation]
        L.ract = ACT[self.readout_activation]
        L.n_in = 0 if self.inputs is None else len(self.inputs)
        assert L.n_in <= 4
        for i, src in enumerate(self.inputs or []):
            L.inputs[i] = src
        return L


class _OptimT(C.Structure):
    pass


@dataclass
class OptimSpec:
    kind: str = "sgd"
    lr: float = 0.01
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    momentum: float = 0.0
    nesterov: bool = False
    clip_min: Optional[float] = None
    clip_max: Optional[float] = None
    clip_norm: Optional[float] = None
    regulariser: Optional[str] = None   # "l1" | "l2" | "l1l2" (athena_regulariser.f90)
    l1: float = 0.01
    l2: float = 0.01
    l2_decoupled: bool = True


@dataclass
class Batch:
    """A mini-batch of graph_type samples in the packed host layout that the
    C ABI takes (see include/athena_cuda.h, athena_cuda_batch_create)."""
    nv: np.ndarray      # [B] int32 num_vertices
    ne: np.ndarray      # [B] int32 num_edges (edge-feature columns)
    ia: np.ndarray      # concatenated adj_ia, 1-based, sum(nv+1)
    ja: np.ndarray      # concatenated adj_ja (2,Z) interleaved -> [Z,2], 1-based
    x: np.ndarray       # [V_tot, F0]
    e: Optional[np.ndarray] = None  # [E_tot, Fe]

    @property
    def B(self) -> int:
        return int(self.nv.shape[0])

    @property
    def V(self) -> int:
        return int(self.nv.sum())

    @property
    def Z(self) -> int:
        return int(self.ja.shape[0])

    @property
    def E(self) -> int:
        return int(self.ne.sum())


def _p(a, ty):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ty))


class Oracle:
    def __init__(self, precision: str = "f32"):
        build()
        self.precision = precision
        self.lib = C.CDLL(os.path.join(_BUILD, f"liboracle_{precision}.so"))
        self.dtype = np.float64 if precision == "f64" else np.float32
        self.creal = C.c_double if precision == "f64" else C.c_float
        assert self.lib.oracle_real_bytes() == np.dtype(self.dtype).itemsize
        _OptimT_fields = [("kind", C.c_int), ("lr", self.creal), ("beta1", self.creal),
                          ("beta2", self.creal), ("eps", self.creal),
                          ("momentum", self.creal), ("nesterov", C.c_int),
                          ("clip_flags", C.c_int), ("clip_min", self.creal),
                          ("clip_max", self.creal), ("clip_norm", self.creal),
                          ("reg", C.c_int), ("l1", self.creal), ("l2", self.creal),
                          ("l2_decoupled", C.c_int)]
        self._OptimT = type("_OptimT_" + precision, (C.Structure,), {"_fields_": _OptimT_fields})
        R, I, RP = self.creal, C.c_int, C.POINTER(self.creal)
        self.lib.oracle_mse_cell.restype = self.creal
        self.lib.oracle_clip.argtypes = [I, RP, I, R, R, R]
        self.lib.oracle_clip_bias.argtypes = [I, RP, I, RP, I, R, R, R]
        self.lib.oracle_sgd.argtypes = [I, RP, RP, RP, R, R, I]
        self.lib.oracle_adam.argtypes = [I, RP, RP, RP, RP, R, R, R, R, I]
        self.lib.oracle_stack_fwd_bwd.restype = self.creal
        self.lib.oracle_train_step.restype = self.creal

    # -- helpers ----------------------------------------------------------
    def r(self, a) -> np.ndarray:
        return np.ascontiguousarray(a, dtype=self.dtype)

    @staticmethod
    def i(a) -> np.ndarray:
        return np.ascontiguousarray(a, dtype=np.int32)

    def rp(self, a):
        return _p(a, self.creal)

    def optim_c(self, o: OptimSpec):
        s = self._OptimT()
        s.kind = {"sgd": 0, "adam": 1, "rmsprop": 2, "adagrad": 3}[o.kind]  # rmsprop: beta in beta1
        s.lr, s.beta1, s.beta2, s.eps, s.momentum = o.lr, o.beta1, o.beta2, o.eps, o.momentum
        s.nesterov = int(o.nesterov)
        flags = 0
        if o.clip_min is not None or o.clip_max is not None:
            flags |= 1
            s.clip_min = o.clip_min if o.clip_min is not None else -np.finfo(np.float32).max
            s.clip_max = o.clip_max if o.clip_max is not None else np.finfo(np.float32).max
        if o.clip_norm is not None:
            flags |= 2
            s.clip_norm = o.clip_norm
        s.clip_flags = flags
        s.reg = {None: 0, "l1": 1, "l2": 2, "l1l2": 3}[o.regulariser]
        s.l1, s.l2, s.l2_decoupled = o.l1, o.l2, int(o.l2_decoupled)
        return s

    # -- primitive ops (single graph) -------------------------------------
    def kipf_propagate(self, x, ia, ja):
        x = self.r(x); ia = self.i(ia); ja = self.i(ja)
        V, F = x.shape
        out = np.empty_like(x)
        self.lib.oracle_kipf_propagate(F, V, self.rp(x), _p(ia, C.c_int), _p(ja, C.c_int), self.rp(out))
        return out

    def kipf_propagate_bwd(self, g, ia, ja):
        g = self.r(g); ia = self.i(ia); ja = self.i(ja)
        V, F = g.shape
        out = np.empty_like(g)
        self.lib.oracle_kipf_propagate_bwd(F, V, self.rp(g), _p(ia, C.c_int), _p(ja, C.c_int), self.rp(out))
        return out

    def duvenaud_propagate(self, x, e, ia, ja):
        x = self.r(x); e = self.r(e); ia = self.i(ia); ja = self.i(ja)
        V, F = x.shape
        Fe = e.shape[1]
        out = np.empty((V, F + Fe), self.dtype)
        self.lib.oracle_duvenaud_propagate(F, Fe, V, self.rp(x), self.rp(e), _p(ia, C.c_int),
                                           _p(ja, C.c_int), self.rp(out))
        return out

    def duvenaud_update(self, a, w, ia, min_deg, max_deg, Fo):
        a = self.r(a); w = self.r(w); ia = self.i(ia)
        V, K = a.shape
        out = np.empty((V, Fo), self.dtype)
        self.lib.oracle_duvenaud_update(K, Fo, V, self.rp(a), self.rp(w), _p(ia, C.c_int),
                                        min_deg, max_deg, self.rp(out))
        return out

    def matmul(self, w, M, K, p):
        """Y = W.P with W flat column-major [M,K]; p is [N,K] -> returns [N,M]."""
        w = self.r(w); p = self.r(p)
        N = p.shape[0]
        out = np.empty((N, M), self.dtype)
        self.lib.oracle_matmul(M, K, N, self.rp(w), self.rp(p), self.rp(out))
        return out

    def activation(self, kind, x):
        x = self.r(x)
        V, F = x.shape
        out = np.empty_like(x)
        self.lib.oracle_activation(ACT[kind], F, V, self.rp(x), self.rp(out))
        return out

    def activation_bwd(self, kind, y, g):
        y = self.r(y); g = self.r(g)
        V, F = y.shape
        out = np.empty_like(y)
        self.lib.oracle_activation_bwd(ACT[kind], F, V, self.rp(y), self.rp(g), self.rp(out))
        return out

    def activation_bwd_x(self, kind, x, y, g):
        """Derivative with the pre-activation at hand (swish differentiates on it)."""
        x = self.r(x); y = self.r(y); g = self.r(g)
        V, F = y.shape
        out = np.empty_like(y)
        self.lib.oracle_activation_bwd_x(ACT[kind], F, V, self.rp(x), self.rp(y), self.rp(g),
                                         self.rp(out))
        return out

    def mse_cell(self, p, e):
        p = self.r(p).ravel(); e = self.r(e).ravel()
        return float(self.lib.oracle_mse_cell(C.c_size_t(p.size), self.rp(p), self.rp(e)))

    def clip(self, g, clip_min=None, clip_max=None, clip_norm=None, bias=None):
        g = self.r(g).copy()
        o = self.optim_c(OptimSpec(clip_min=clip_min, clip_max=clip_max, clip_norm=clip_norm))
        if bias is None:
            self.lib.oracle_clip(g.size, self.rp(g), o.clip_flags, o.clip_min, o.clip_max, o.clip_norm)
            return g
        b = self.r(bias).copy()
        self.lib.oracle_clip_bias(g.size, self.rp(g), b.size, self.rp(b), o.clip_flags,
                                  o.clip_min, o.clip_max, o.clip_norm)
        return g, b

    def sgd(self, p, g, vel, lr, momentum=0.0, nesterov=False):
        p = self.r(p).copy(); g = self.r(g).copy(); vel = self.r(vel).copy()
        self.lib.oracle_sgd(p.size, self.rp(p), self.rp(g), self.rp(vel), lr, momentum, int(nesterov))
        return p, vel

    def adam(self, p, g, m, v, lr, beta1, beta2, eps, it):
        p = self.r(p).copy(); g = self.r(g); m = self.r(m).copy(); v = self.r(v).copy()
        self.lib.oracle_adam(p.size, self.rp(p), self.rp(g), self.rp(m), self.rp(v),
                             lr, beta1, beta2, eps, int(it))
        return p, m, v

    # -- integer structures ----------------------------------------------
    def batch_build(self, nv, ne, ia, ja):
        nv = self.i(nv); ne = self.i(ne); ia = self.i(ia); ja = self.i(ja)
        B = nv.size
        V = int(nv.sum()); Z = int(ja.size // 2)
        o = dict(row_ptr=np.empty(V + 1, np.int32), col=np.empty(Z, np.int32),
                 eid=np.empty(Z, np.int32), deg=np.empty(V, np.int32),
                 vgraph=np.empty(V, np.int32), csc_ptr=np.empty(V + 1, np.int32),
                 csc_src=np.empty(Z, np.int32), csc_ent=np.empty(Z, np.int32))
        rc = self.lib.oracle_batch_build(B, _p(nv, C.c_int), _p(ne, C.c_int), _p(ia, C.c_int),
                                         _p(ja, C.c_int), *[_p(o[k], C.c_int) for k in
                                                            ("row_ptr", "col", "eid", "deg", "vgraph",
                                                             "csc_ptr", "csc_src", "csc_ent")])
        if rc != 0:
            raise ValueError(f"graph {-rc - 1}: adjacency index outside 1..num_vertices")
        return o

    def bucketize(self, deg, min_deg, max_deg):
        deg = self.i(deg)
        V = deg.size
        D = max_deg - min_deg + 1
        bkt = np.empty(V, np.int32); perm = np.empty(V, np.int32); ptr = np.empty(D + 1, np.int32)
        self.lib.oracle_bucketize(V, _p(deg, C.c_int), min_deg, max_deg, _p(bkt, C.c_int),
                                  _p(perm, C.c_int), _p(ptr, C.c_int))
        return bkt, perm, ptr

    # -- composite ---------------------------------------------------------
    def num_params(self, layers: List[LayerSpec]) -> int:
        return sum(self.lib.oracle_layer_num_params(C.byref(L.to_c())) for L in layers)

    def _layers_c(self, layers):
        arr = (_LayerT * len(layers))()
        for k, L in enumerate(layers):
            arr[k] = L.to_c()
        return arr

    def out_shape(self, layers, b: Batch):
        last = layers[-1]
        if last.kind == "kipf":
            nvf = list(last.num_vertex_features)
            return (b.V, nvf[-1])
        if last.kind == "full":
            return (b.B, list(last.num_vertex_features)[1])
        return (b.B, last.num_outputs)

    def stack_fwd_bwd(self, layers: List[LayerSpec], params, b: Batch, target=None,
                      global_B: Optional[int] = None, want_grads: bool = True):
        """-> (loss, out, dparams)"""
        params = self.r(params)
        x = self.r(b.x); e = self.r(b.e) if b.e is not None else None
        out = np.zeros(self.out_shape(layers, b), self.dtype)
        dparams = np.zeros(params.size, self.dtype) if (want_grads and target is not None) else None
        tgt = self.r(target) if target is not None else None
        loss = self.lib.oracle_stack_fwd_bwd(
            len(layers), self._layers_c(layers), self.rp(params), b.B, _p(self.i(b.nv), C.c_int),
            _p(self.i(b.ne), C.c_int), _p(self.i(b.ia), C.c_int), _p(self.i(b.ja), C.c_int),
            self.rp(x), self.rp(e), self.rp(tgt), int(global_B or b.B), self.rp(out), self.rp(dparams))
        return float(loss), out, dparams

    def layer_fwd_bwd(self, layer: LayerSpec, params, b: Batch, g_out=None, want_dx=False,
                      x=None):
        """-> (out, dparams, dx).  `x` overrides the batch's vertex features (a full layer
        takes the graph-level [B, num_inputs] array)."""
        params = self.r(params)
        x = self.r(b.x if x is None else x); e = self.r(b.e) if b.e is not None else None
        out = np.zeros(self.out_shape([layer], b), self.dtype)
        g = self.r(g_out) if g_out is not None else None
        dparams = np.zeros(params.size, self.dtype) if g is not None else None
        dx = np.zeros_like(x) if (want_dx and g is not None) else None
        Lc = layer.to_c()
        self.lib.oracle_layer_fwd_bwd(
            C.byref(Lc), self.rp(params), b.B, _p(self.i(b.nv), C.c_int), _p(self.i(b.ne), C.c_int),
            _p(self.i(b.ia), C.c_int), _p(self.i(b.ja), C.c_int), self.rp(x), self.rp(e),
            self.rp(g), self.rp(out), self.rp(dparams), self.rp(dx))
        return out, dparams, dx

    def update(self, params, grads, o: OptimSpec, state1, state2, it):
        params = self.r(params).copy(); grads = self.r(grads).copy()
        s1 = self.r(state1).copy(); s2 = self.r(state2).copy()
        oc = self.optim_c(o)
        self.lib.oracle_update(params.size, self.rp(params), self.rp(grads), C.byref(oc),
                               self.rp(s1), self.rp(s2), int(it))
        return params, s1, s2

    def train_step(self, layers, params, b: Batch, target, o: OptimSpec, state1, state2, it):
        """In-place on params/state1/state2 (must be contiguous arrays of self.dtype).
        -> (loss, out)"""
        assert params.dtype == self.dtype and params.flags.c_contiguous
        x = self.r(b.x); e = self.r(b.e) if b.e is not None else None
        out = np.zeros(self.out_shape(layers, b), self.dtype)
        grads = np.zeros(params.size, self.dtype)
        oc = self.optim_c(o)
        tgt = self.r(target)
        loss = self.lib.oracle_train_step(
            len(layers), self._layers_c(layers), self.rp(params), b.B, _p(self.i(b.nv), C.c_int),
            _p(self.i(b.ne), C.c_int), _p(self.i(b.ia), C.c_int), _p(self.i(b.ja), C.c_int),
            self.rp(x), self.rp(e), self.rp(tgt), self.rp(out), self.rp(grads), C.byref(oc),
            self.rp(state1), self.rp(state2), int(it))
        return float(loss), out
