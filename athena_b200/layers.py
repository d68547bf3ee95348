"""Python mirror of athena's message-passing layer API on top of the C ABI.

Same names, argument meaning and error behaviour as
  kipf_msgpass_layer_type      athena_kipf_msgpass_layer.f90:80-97,143-308
  duvenaud_msgpass_layer_type  athena_duvenaud_msgpass_layer.f90:88-120,256-505
  msgpass_layer_type%set_graph / forward   athena_msgpass_layer.f90:58-69
  learnable_layer_type get/set_params, get/set_gradients  athena_base_layer.f90:499-533
so that the parity tests read like the reference's own tests
(test/test_kipf_msgpass_layer.f90, test/test_duvenaud_msgpass_layer.f90).
All compute happens in libathena_cuda; nothing here touches the oracle.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import ACT, AthenaCudaError, check, lib, ptr
from .graph import PackedGraphs, graph_type, pack_graphs


class GraphBatch:
    """Device representation of graph(:), built once and shared by all layers
    (replaces the per-forward deep copies of set_graph_msgpass,
    athena_msgpass_layer_sub.f90:144-174)."""

    def __init__(self, graphs: Union[Sequence[graph_type], PackedGraphs], validate: bool = True,
                 mem: int = _lib.MEM_HOST, ia_dev=None, ja_dev=None):
        self.packed = graphs if isinstance(graphs, PackedGraphs) else pack_graphs(graphs)
        p = self.packed
        h = C.c_int64()
        ia = ia_dev if mem == _lib.MEM_DEVICE else p.ia
        ja = ja_dev if mem == _lib.MEM_DEVICE else p.ja
        check(lib().athena_cuda_batch_create(C.byref(h), p.B, ptr(p.nv), ptr(p.ne), ptr(p.nz),
                                             ptr(ia), ptr(ja), mem, int(validate)))
        self.handle = h.value
        self.B, self.V, self.Z, self.E = p.B, p.V, p.Z, p.E

    @classmethod
    def _adopt(cls, handle: int, packed: Optional[PackedGraphs] = None) -> "GraphBatch":
        self = cls.__new__(cls)
        self.handle = handle
        self.packed = packed
        b, v, z, e = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().athena_cuda_batch_info(handle, C.byref(b), C.byref(v), C.byref(z), C.byref(e)))
        self.B, self.V, self.Z, self.E = b.value, v.value, z.value, e.value
        return self

    @classmethod
    def from_edges(cls, num_vertices, index_lists: Sequence, add_self_loops: bool = False,
                   validate: bool = True, num_entries=None) -> "GraphBatch":
        """generate_adjacency(index_list) (+ add_self_loops) on the device
        (athena_cuda_batch_create_from_edges): index_lists[s] = [E_s, 2] 1-based pairs."""
        nv = np.ascontiguousarray(num_vertices, np.int32)
        ils = [np.ascontiguousarray(il, np.int32).reshape(-1, 2) for il in index_lists]
        ne = np.asarray([il.shape[0] for il in ils], np.int32)
        il = np.ascontiguousarray(np.concatenate(ils) if ils else np.zeros((0, 2)), np.int32)
        h = C.c_int64()
        nz = None if num_entries is None else np.ascontiguousarray(num_entries, np.int32)
        check(lib().athena_cuda_batch_create_from_edges(C.byref(h), nv.size, ptr(nv), ptr(ne),
                                                        ptr(il), ptr(nz), int(add_self_loops),
                                                        _lib.MEM_HOST, int(validate)))
        return cls._adopt(h.value)

    @classmethod
    def from_edge_index(cls, num_vertices, num_edges, edge_indices: Sequence, degree,
                        validate: bool = True) -> "GraphBatch":
        """The ONNX graph inputs of athena's message-passing export: edge_indices[s] = int64
        [3, ncsr_s] (source, edge-feature index, target; 0-based), degree int64 [V]
        (athena_onnx_msgpass_utils.f90:53-92)."""
        nv = np.ascontiguousarray(num_vertices, np.int32)
        ne = np.ascontiguousarray(num_edges, np.int32)
        eis = [np.ascontiguousarray(ei, np.int64).reshape(3, -1) for ei in edge_indices]
        nz = np.asarray([ei.shape[1] for ei in eis], np.int32)
        ei = np.ascontiguousarray(np.concatenate([e.ravel() for e in eis]) if eis
                                  else np.zeros(0), np.int64)
        deg = np.ascontiguousarray(degree, np.int64)
        h = C.c_int64()
        check(lib().athena_cuda_batch_create_from_edge_index(
            C.byref(h), nv.size, ptr(nv), ptr(ne), ptr(nz), ptr(ei), ptr(deg), _lib.MEM_HOST,
            int(validate)))
        return cls._adopt(h.value)

    def export(self, what: str) -> np.ndarray:
        n = {"row_ptr": self.V + 1, "col": self.Z, "eid": self.Z, "deg": self.V, "vgraph": self.V,
             "csc_ptr": self.V + 1, "csc_src": self.Z, "csc_ent": self.Z, "bucket": self.V,
             "perm": self.V, "coef": self.Z}.get(what)
        if what == "bucket_ptr":
            n = self._D + 1
        dtype = np.float32 if what == "coef" else np.int32
        out = np.empty(n, dtype)
        check(lib().athena_cuda_batch_export(self.handle, _lib.BATCH_FIELDS[what], ptr(out), n))
        return out

    def bucketize(self, min_degree: int, max_degree: int):
        check(lib().athena_cuda_batch_bucketize(self.handle, min_degree, max_degree))
        self._D = max_degree - min_degree + 1

    def status(self):
        check(lib().athena_cuda_batch_status(self.handle))

    def destroy(self):
        if getattr(self, "handle", 0):
            lib().athena_cuda_batch_destroy(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class msgpass_layer_type:
    """Common part of the two message-passing layers."""
    name = "msgpass"

    def __init__(self):
        self.handle = 0
        self.graph: Optional[GraphBatch] = None
        self.output: Optional[np.ndarray] = None
        self._owned = True

    # -- learnable_layer_type --------------------------------------------
    @property
    def num_params(self) -> int:
        n = C.c_int64()
        check(lib().athena_cuda_layer_num_params(self.handle, C.byref(n)))
        return n.value

    def get_num_params(self) -> int:
        return self.num_params

    def get_params(self) -> np.ndarray:
        out = np.empty(self.num_params, np.float32)
        check(lib().athena_cuda_layer_get_params(self.handle, ptr(out), out.size))
        return out

    def set_params(self, params):
        a = np.ascontiguousarray(params, np.float32)
        check(lib().athena_cuda_layer_set_params(self.handle, ptr(a), a.size))

    def get_gradients(self) -> np.ndarray:
        out = np.empty(self.num_params, np.float32)
        check(lib().athena_cuda_layer_get_gradients(self.handle, ptr(out), out.size))
        return out

    def set_gradients(self, gradients):
        a = np.ascontiguousarray(gradients, np.float32)
        if a.size == 1:  # set_gradients(scalar) broadcast, athena_base_layer_sub.f90:651-691
            a = np.full(self.num_params, a.ravel()[0], np.float32)
        check(lib().athena_cuda_layer_set_gradients(self.handle, ptr(a), a.size))

    def zero_gradients(self):
        check(lib().athena_cuda_layer_zero_gradients(self.handle))

    # -- msgpass_layer_type -------------------------------------------------
    def set_graph(self, graph: Union[Sequence[graph_type], PackedGraphs, GraphBatch]):
        """layer%set_graph(graph(:)) -- athena_msgpass_layer.f90:111-117."""
        self.graph = graph if isinstance(graph, GraphBatch) else GraphBatch(graph)

    def forward(self, vertex_features=None, edge_features=None) -> np.ndarray:
        """layer%forward(input): input(1,s) vertex features, input(2,s) edge
        features, concatenated over samples ([V_tot, F] / [E_tot, Fe]).  When
        omitted, the features stored with set_graph's graphs are used (what
        input_layer%set_input_graph feeds, athena_input_layer.f90:511-556)."""
        if self.graph is None:
            raise AthenaCudaError(-5, "forward: set_graph has not been called")
        p = self.graph.packed
        x = np.ascontiguousarray(p.x if vertex_features is None else vertex_features, np.float32)
        e = p.e if edge_features is None else edge_features
        e = None if e is None else np.ascontiguousarray(e, np.float32)
        out = np.empty(self._out_shape(), np.float32)
        check(lib().athena_cuda_layer_forward(self.handle, self.graph.handle, ptr(x), ptr(e),
                                              ptr(out), _lib.MEM_HOST))
        self.output = out
        return out

    def backward(self, grad_output, want_input_grad: bool = False):
        """Reverse sweep for an upstream gradient of the layer output;
        accumulates into the parameter gradients.  Returns d(input) or None."""
        g = np.ascontiguousarray(grad_output, np.float32)
        gin = np.empty((self.graph.V, self.num_vertex_features[0]), np.float32) \
            if want_input_grad else None
        check(lib().athena_cuda_layer_backward(self.handle, self.graph.handle, ptr(g), ptr(gin),
                                               _lib.MEM_HOST))
        return gin

    def destroy(self):
        if self.handle and self._owned:
            lib().athena_cuda_layer_destroy(self.handle)
        self.handle = 0

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def _act_id(a) -> int:
    name = "none" if a is None else str(a).strip().lower()
    if name not in ACT:
        raise AthenaCudaError(-2, f"unsupported activation '{name}' on the CUDA path "
                                  f"(supported: {sorted(ACT)})")
    return ACT[name]


class kipf_msgpass_layer_type(msgpass_layer_type):
    """kipf_msgpass_layer_type(num_vertex_features, num_time_steps, activation="none")
    -- docs/source/layers/msgpass/kipf_msgpass_layer.rst, athena_kipf_msgpass_layer.f90:143-308."""
    name = "kipf"

    def __init__(self, num_vertex_features: Sequence[int], num_time_steps: int,
                 activation="none", kernel_initialiser: Optional[str] = None, verbose: int = 0):
        super().__init__()
        nvf = [int(f) for f in np.atleast_1d(num_vertex_features)]
        if num_time_steps < 1:
            # athena_kipf_msgpass_layer.f90:271-274 -> stop_program
            raise AthenaCudaError(-2, "Number of time steps must be at least 1")
        if len(nvf) == 1:
            nvf = nvf * (num_time_steps + 1)
        elif len(nvf) != num_time_steps + 1:
            # athena_kipf_msgpass_layer.f90:279-283
            raise AthenaCudaError(-2, "Number of vertex features must be a scalar or a vector "
                                      "of length num_time_steps + 1")
        self.num_vertex_features = nvf
        self.num_time_steps = int(num_time_steps)
        self.activation = "none" if activation is None else str(activation)
        arr = np.asarray(nvf, np.int32)
        h = C.c_int64()
        check(lib().athena_cuda_kipf_layer_create(C.byref(h), self.num_time_steps, ptr(arr),
                                                  _act_id(self.activation)))
        self.handle = h.value
        _initialise(self, kernel_initialiser)

    def _out_shape(self):
        return (self.graph.V, self.num_vertex_features[-1])


class duvenaud_msgpass_layer_type(msgpass_layer_type):
    """duvenaud_msgpass_layer_type(num_vertex_features, num_edge_features, num_time_steps,
    max_vertex_degree, num_outputs, min_vertex_degree=1, message_activation="sigmoid",
    readout_activation="softmax") -- athena_duvenaud_msgpass_layer.f90:88-124,256-505."""
    name = "duvenaud"

    def __init__(self, num_vertex_features, num_edge_features, num_time_steps: int,
                 max_vertex_degree: int, num_outputs: int, min_vertex_degree: int = 1,
                 message_activation="sigmoid", readout_activation="softmax",
                 kernel_initialiser: Optional[str] = None, verbose: int = 0):
        super().__init__()
        nvf = [int(f) for f in np.atleast_1d(num_vertex_features)]
        nef = [int(f) for f in np.atleast_1d(num_edge_features)]
        if num_time_steps < 1:
            raise AthenaCudaError(-2, "Number of time steps must be at least 1")
        if len(nvf) == 1:
            nvf = nvf * (num_time_steps + 1)
        elif len(nvf) != num_time_steps + 1:
            raise AthenaCudaError(-2, "num_vertex_features must have 1 or num_time_steps+1 entries")
        if min_vertex_degree < 1 or max_vertex_degree < min_vertex_degree:
            raise AthenaCudaError(-2, "min_vertex_degree must be at least 1 and max_vertex_degree "
                                      "at least min_vertex_degree")
        self.num_vertex_features = nvf
        self.num_edge_features = nef[0]  # sized with num_edge_features(0): :552
        self.num_time_steps = int(num_time_steps)
        self.min_vertex_degree = int(min_vertex_degree)
        self.max_vertex_degree = int(max_vertex_degree)
        self.num_outputs = int(num_outputs)
        self.message_activation = str(message_activation)
        self.readout_activation = str(readout_activation)
        arr = np.asarray(nvf, np.int32)
        h = C.c_int64()
        check(lib().athena_cuda_duvenaud_layer_create(
            C.byref(h), self.num_time_steps, ptr(arr), self.num_edge_features,
            self.min_vertex_degree, self.max_vertex_degree, self.num_outputs,
            _act_id(self.message_activation), _act_id(self.readout_activation)))
        self.handle = h.value
        _initialise(self, kernel_initialiser)

    def _out_shape(self):
        return (self.graph.B, self.num_outputs)


class full_layer_type(msgpass_layer_type):
    """full_layer_type(num_outputs, num_inputs, use_bias=True, activation="none") -- the dense
    head that follows the Duvenaud readout in example/msgpass_chemical (main.f90:139-157);
    athena_full_layer.f90:147-160 (constructor), :839-874 (forward: act(matmul(W, x) + b)).
    Parameters: W [num_outputs, num_inputs] column-major, then the bias (:371-396).  It takes
    the graph-level [batch, num_inputs] array of the previous layer, not vertex features."""
    name = "full"

    def __init__(self, num_outputs: int, num_inputs: Optional[int] = None, use_bias: bool = True,
                 activation="none", kernel_initialiser: Optional[str] = None,
                 bias_initialiser: Optional[str] = None, verbose: int = 0):
        super().__init__()
        self.num_outputs = int(num_outputs)
        self.use_bias = bool(use_bias)
        self.activation = "none" if activation is None else str(activation)
        self._init = (kernel_initialiser, bias_initialiser)
        _act_id(self.activation)
        # without num_inputs the layer is initialised when it is added to a network, from the
        # previous layer's output shape (athena_full_layer.f90:212-215, network%compile)
        if num_inputs is not None:
            self._create(num_inputs)

    def _create(self, num_inputs: int):
        kernel_initialiser, bias_initialiser = self._init
        if num_inputs < 1 or self.num_outputs < 1:
            raise AthenaCudaError(-2, "full_layer: num_inputs and num_outputs must be positive")
        self.num_inputs = int(num_inputs)
        self.num_vertex_features = [self.num_inputs, self.num_outputs]
        h = C.c_int64()
        check(lib().athena_cuda_full_layer_create(C.byref(h), self.num_inputs, self.num_outputs,
                                                  _act_id(self.activation), int(self.use_bias)))
        self.handle = h.value
        _initialise(self, kernel_initialiser)
        if self.use_bias:  # bias_initialiser defaults to zeros (athena_full_layer.f90:281-284)
            prm = self.get_params()
            prm[self.num_inputs * self.num_outputs:] = 1.0 if bias_initialiser == "ones" else 0.0
            self.set_params(prm)

    def _out_shape(self):
        return (self.graph.B, self.num_outputs)

    def forward(self, input=None, edge_features=None) -> np.ndarray:
        """layer%forward(input) with input [batch, num_inputs] (Fortran val(num_inputs, batch))."""
        if self.graph is None:
            raise AthenaCudaError(-5, "forward: set_graph has not been called")
        x = np.ascontiguousarray(input, np.float32)
        assert x.shape == (self.graph.B, self.num_inputs), "full_layer: input shape mismatch"
        out = np.empty(self._out_shape(), np.float32)
        check(lib().athena_cuda_layer_forward(self.handle, self.graph.handle, ptr(x), None,
                                              ptr(out), _lib.MEM_HOST))
        self.output = out
        return out

    def backward(self, grad_output, want_input_grad: bool = False):
        g = np.ascontiguousarray(grad_output, np.float32)
        gin = np.empty((self.graph.B, self.num_inputs), np.float32) if want_input_grad else None
        check(lib().athena_cuda_layer_backward(self.handle, self.graph.handle, ptr(g), ptr(gin),
                                               _lib.MEM_HOST))
        return gin


def _initialise(layer: msgpass_layer_type, kernel_initialiser: Optional[str]):
    """Host-side initialisers stay host-side (athena_initialiser*.f90 use the
    compiler RNG and are out of scope; parity tests inject parameters with
    set_params).  'ones'/'zeros' are exact; anything else gets a seeded
    glorot-uniform draw so a fresh layer is trainable."""
    n = layer.num_params
    kind = (kernel_initialiser or "glorot_uniform").lower()
    if kind == "ones":
        layer.set_params(np.ones(n, np.float32))
    elif kind == "zeros":
        layer.set_params(np.zeros(n, np.float32))
    else:
        nvf = layer.num_vertex_features
        fan = float(nvf[0] + nvf[-1])
        lim = np.sqrt(6.0 / fan)
        rng = np.random.default_rng(0)
        layer.set_params(rng.uniform(-lim, lim, n).astype(np.float32))
