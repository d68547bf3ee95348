"""Python mirror of the slice of network_type that drives the message-passing
layers: add / compile / train / forward / predict / update and the flat
parameter accessors (athena_network.f90:142-223; bodies
athena_network_sub.f90:2639-2929, 3387-3905, 4226-4303).  Optimisers and the
clipper are plain descriptors (athena_optimiser.f90, athena_clipper.f90); the
arithmetic runs in libathena_cuda.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import AthenaCudaError, OptimiserDesc, check, lib, ptr
from .graph import PackedGraphs, graph_type, pack_graphs
from .layers import GraphBatch, msgpass_layer_type


class clip_type:
    """clip_type(clip_min, clip_max, clip_norm) -- athena_clipper.f90:30-120."""

    def __init__(self, clip_min: Optional[float] = None, clip_max: Optional[float] = None,
                 clip_norm: Optional[float] = None):
        self.l_min_max = clip_min is not None or clip_max is not None
        big = float(np.finfo(np.float32).max)
        self.min = -big if clip_min is None else float(clip_min)
        self.max = big if clip_max is None else float(clip_max)
        self.l_norm = clip_norm is not None
        self.norm = 0.0 if clip_norm is None else float(clip_norm)


class l1_regulariser_type:
    """l1_regulariser_type(l1=0.01) -- athena_regulariser.f90:40-49, 85-101."""
    kind = _lib.REG_L1

    def __init__(self, l1: float = 0.01):
        self.l1, self.l2, self.decoupled = float(l1), 0.0, True


class l2_regulariser_type:
    """l2_regulariser_type(l2=0.01, decoupled=.true.) -- athena_regulariser.f90:51-66, 103-119;
    `decoupled` selects AdamW in minimise_adam (athena_optimiser.f90:1064-1084)."""
    kind = _lib.REG_L2

    def __init__(self, l2: float = 0.01, decoupled: bool = True):
        self.l1, self.l2, self.decoupled = 0.0, float(l2), bool(decoupled)


class l1l2_regulariser_type:
    """l1l2_regulariser_type(l1=0.01, l2=0.01) -- athena_regulariser.f90:121-137."""
    kind = _lib.REG_L1L2

    def __init__(self, l1: float = 0.01, l2: float = 0.01):
        self.l1, self.l2, self.decoupled = float(l1), float(l2), True


class base_lr_decay_type:
    """base_lr_decay_type: no decay (athena_lr_decay.f90:200-214)."""
    iterate_per_epoch = False

    def get_lr(self, learning_rate: float, iteration: int) -> float:
        return float(np.float32(learning_rate))


class exp_lr_decay_type(base_lr_decay_type):
    """exp_lr_decay_type(decay_rate = 0.9): lr * exp(-iteration * decay_rate)
    (athena_lr_decay.f90:127-143, 218-233)."""

    def __init__(self, decay_rate: float = 0.9):
        self.decay_rate = float(decay_rate)

    def get_lr(self, learning_rate, iteration):
        return float(np.float32(learning_rate) *
                     np.exp(-np.float32(iteration) * np.float32(self.decay_rate), dtype=np.float32))


class step_lr_decay_type(base_lr_decay_type):
    """step_lr_decay_type(decay_rate = 0.1, decay_steps = 100): lr * decay_rate ** (iteration /
    decay_steps), integer division, the counter advancing once per EPOCH
    (athena_lr_decay.f90:146-170, 236-251)."""
    iterate_per_epoch = True

    def __init__(self, decay_rate: float = 0.1, decay_steps: int = 100):
        self.decay_rate, self.decay_steps = float(decay_rate), int(decay_steps)

    def get_lr(self, learning_rate, iteration):
        return float(np.float32(learning_rate) *
                     np.float32(self.decay_rate) ** (int(iteration) // self.decay_steps))


class inv_lr_decay_type(base_lr_decay_type):
    """inv_lr_decay_type(decay_rate = 0.001, decay_power = 1): lr * (1 + decay_rate * iteration)
    ** (-decay_power) (athena_lr_decay.f90:173-195, 254-270)."""

    def __init__(self, decay_rate: float = 0.001, decay_power: float = 1.0):
        self.decay_rate, self.decay_power = float(decay_rate), float(decay_power)

    def get_lr(self, learning_rate, iteration):
        base = np.float32(1.0) + np.float32(self.decay_rate) * np.float32(iteration)
        return float(np.float32(learning_rate) * base ** np.float32(-self.decay_power))


class base_optimiser_type:
    kind = _lib.OPT_SGD

    def __init__(self, learning_rate: float = 0.01, clip_dict: Optional[clip_type] = None,
                 regulariser=None, lr_decay: Optional[base_lr_decay_type] = None):
        self.learning_rate = float(learning_rate)
        self.clip_dict = clip_dict or clip_type()
        self.regulariser = regulariser
        self.lr_decay = lr_decay or base_lr_decay_type()
        self.iter = 0
        self.epoch = 0

    def desc(self) -> OptimiserDesc:
        d = OptimiserDesc()
        d.kind = self.kind
        d.learning_rate = self.learning_rate
        d.beta1, d.beta2, d.epsilon = 0.9, 0.999, 1e-8
        d.momentum, d.nesterov = 0.0, 0
        c = self.clip_dict
        d.clip_min_max, d.clip_min, d.clip_max = int(c.l_min_max), c.min, c.max
        d.clip_norm_on, d.clip_norm = int(c.l_norm), c.norm
        r = self.regulariser
        d.regulariser = _lib.REG_NONE if r is None else r.kind
        d.l1, d.l2 = (0.0, 0.0) if r is None else (r.l1, r.l2)
        d.l2_decoupled = 1 if r is None else int(r.decoupled)
        return d


class sgd_optimiser_type(base_optimiser_type):
    """sgd_optimiser_type(learning_rate, momentum, nesterov) -- athena_optimiser.f90:560-673."""
    kind = _lib.OPT_SGD

    def __init__(self, learning_rate: float = 0.01, momentum: float = 0.0, nesterov: bool = False,
                 clip_dict: Optional[clip_type] = None, regulariser=None,
                 lr_decay: Optional[base_lr_decay_type] = None):
        super().__init__(learning_rate, clip_dict, regulariser, lr_decay)
        self.momentum, self.nesterov = float(momentum), bool(nesterov)

    def desc(self):
        d = super().desc()
        d.momentum, d.nesterov = self.momentum, int(self.nesterov)
        return d


class adam_optimiser_type(base_optimiser_type):
    """adam_optimiser_type(learning_rate, beta1, beta2, epsilon) -- athena_optimiser.f90:940-1091."""
    kind = _lib.OPT_ADAM

    def __init__(self, learning_rate: float = 0.01, beta1: float = 0.9, beta2: float = 0.999,
                 epsilon: float = 1e-8, clip_dict: Optional[clip_type] = None, regulariser=None,
                 lr_decay: Optional[base_lr_decay_type] = None):
        super().__init__(learning_rate, clip_dict, regulariser, lr_decay)
        self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)

    def desc(self):
        d = super().desc()
        d.beta1, d.beta2, d.epsilon = self.beta1, self.beta2, self.epsilon
        return d


class rmsprop_optimiser_type(base_optimiser_type):
    """rmsprop_optimiser_type(learning_rate, beta, epsilon) -- athena_optimiser.f90:160-166
    (defaults beta = 0, epsilon = 1e-8), minimise_rmsprop :771-803."""
    kind = _lib.OPT_RMSPROP

    def __init__(self, learning_rate: float = 0.01, beta: float = 0.0, epsilon: float = 1e-8,
                 clip_dict: Optional[clip_type] = None, regulariser=None,
                 lr_decay: Optional[base_lr_decay_type] = None):
        super().__init__(learning_rate, clip_dict, regulariser, lr_decay)
        self.beta, self.epsilon = float(beta), float(epsilon)

    def desc(self):
        d = super().desc()
        d.beta1, d.epsilon = self.beta, self.epsilon
        return d


class adagrad_optimiser_type(base_optimiser_type):
    """adagrad_optimiser_type(learning_rate, epsilon) -- athena_optimiser.f90:199-203,
    minimise_adagrad :898-925."""
    kind = _lib.OPT_ADAGRAD

    def __init__(self, learning_rate: float = 0.01, epsilon: float = 1e-8,
                 clip_dict: Optional[clip_type] = None, regulariser=None,
                 lr_decay: Optional[base_lr_decay_type] = None):
        super().__init__(learning_rate, clip_dict, regulariser, lr_decay)
        self.epsilon = float(epsilon)

    def desc(self):
        d = super().desc()
        d.epsilon = self.epsilon
        return d


class network_type:
    def __init__(self):
        h = C.c_int64()
        check(lib().athena_cuda_network_create(C.byref(h)))
        self.handle = h.value
        self.model: List[msgpass_layer_type] = []
        self.batch_size = 0
        self.loss_val = 0.0
        self.epoch = 0
        self.optimiser: Optional[base_optimiser_type] = None
        self.compiled = False

    # -- construction -------------------------------------------------------
    _OPERATORS = {"||": 1, "concat": 1, "concatenate": 1, "append": 1, "+": 2, "add": 2}

    def add(self, layer: msgpass_layer_type, input_list: Optional[Sequence[int]] = None,
            output_list=None, operator=None):
        """network%add(layer, input_list, output_list, operator) (athena_network_sub.f90:764-830).
        input_list ids: 0 = the input layer, k > 0 = the k-th added layer, k < 0 = counted back
        from this layer (-1 = the previously added one); operator 'concatenate' (default) joins
        the sources' vertex features in list order (example/msgpass_euler/src/main.f90:192-255)."""
        if not layer.handle and hasattr(layer, "_create"):
            if not self.model:
                raise AthenaCudaError(-2, "network_add: the first layer must be a message-passing layer")
            layer._create(self.model[-1].num_outputs if self.model[-1].name != "kipf"
                          else self.model[-1].num_vertex_features[-1])
        if output_list is not None:
            raise AthenaCudaError(-2, "network_add: output_list is outside the CUDA path")
        if input_list is None:
            check(lib().athena_cuda_network_add(self.handle, layer.handle))
        else:
            op = 1
            if operator is not None:
                op = operator if isinstance(operator, int) else \
                    self._OPERATORS.get(str(operator).strip().lower(), 0)
            if op < 1 or op > 2:
                raise AthenaCudaError(-2, "invalid operator")  # stop_program("invalid operator")
            ids = np.ascontiguousarray(input_list, np.int32)
            check(lib().athena_cuda_network_add_inputs(self.handle, layer.handle, ids.size,
                                                       ptr(ids), op))
        layer._owned = False  # the network owns the device object now
        self.model.append(layer)

    def compile(self, optimiser: base_optimiser_type, loss_method: str = "mse",
                accuracy_method: str = "mse", metrics=None, batch_size: int = 1, verbose: int = 0):
        if loss_method != "mse":
            raise AthenaCudaError(-2, f"loss_method '{loss_method}' is outside the CUDA path (mse only)")
        d = optimiser.desc()
        check(lib().athena_cuda_network_compile(self.handle, C.byref(d)))
        self.optimiser = optimiser
        self.batch_size = int(batch_size)
        self.compiled = True

    @property
    def num_layers(self) -> int:
        return len(self.model) + 1  # + the implicit input layer, as the reference counts

    @property
    def num_params(self) -> int:
        n = C.c_int64()
        check(lib().athena_cuda_network_num_params(self.handle, C.byref(n)))
        return n.value

    def get_num_params(self) -> int:
        return self.num_params

    def get_params(self) -> np.ndarray:
        out = np.empty(self.num_params, np.float32)
        check(lib().athena_cuda_network_get_params(self.handle, ptr(out), out.size))
        return out

    def set_params(self, params):
        a = np.ascontiguousarray(params, np.float32)
        check(lib().athena_cuda_network_set_params(self.handle, ptr(a), a.size))

    def get_gradients(self) -> np.ndarray:
        out = np.empty(self.num_params, np.float32)
        check(lib().athena_cuda_network_get_gradients(self.handle, ptr(out), out.size))
        return out

    def set_learning_rate(self, lr: float):
        check(lib().athena_cuda_network_set_learning_rate(self.handle, float(lr)))

    # -- one iteration of the batch loop ------------------------------------
    def _out_shape(self, batch: GraphBatch):
        last = self.model[-1]
        if last.name == "kipf":
            return (batch.V, last.num_vertex_features[-1])
        return (batch.B, last.num_outputs)  # duvenaud / full: graph-level output

    def forward(self, graphs: Union[Sequence[graph_type], PackedGraphs, GraphBatch],
                vertex_features=None, edge_features=None) -> np.ndarray:
        batch = graphs if isinstance(graphs, GraphBatch) else GraphBatch(graphs)
        p = batch.packed
        x = np.ascontiguousarray(p.x if vertex_features is None else vertex_features, np.float32)
        e = p.e if edge_features is None else edge_features
        e = None if e is None else np.ascontiguousarray(e, np.float32)
        out = np.empty(self._out_shape(batch), np.float32)
        check(lib().athena_cuda_network_forward(self.handle, batch.handle, ptr(x), ptr(e), ptr(out),
                                                _lib.MEM_HOST))
        return out

    predict = forward

    def train_step(self, batch: GraphBatch, target, global_batch: int = 0,
                   vertex_features=None, edge_features=None, want_loss: bool = True) -> float:
        p = batch.packed
        x = np.ascontiguousarray(p.x if vertex_features is None else vertex_features, np.float32)
        e = p.e if edge_features is None else edge_features
        e = None if e is None else np.ascontiguousarray(e, np.float32)
        t = np.ascontiguousarray(target, np.float32)
        assert t.size == int(np.prod(self._out_shape(batch))), "target shape mismatch"
        loss = C.c_float()
        check(lib().athena_cuda_network_train_step(
            self.handle, batch.handle, ptr(x), ptr(e), ptr(t), _lib.MEM_HOST, int(global_batch),
            C.byref(loss) if want_loss else None))
        self.loss_val = float(loss.value)
        return self.loss_val

    def loss_and_gradients(self, batch: GraphBatch, target, global_batch: int = 0) -> float:
        p = batch.packed
        t = np.ascontiguousarray(target, np.float32)
        loss = C.c_float()
        check(lib().athena_cuda_network_loss_and_gradients(
            self.handle, batch.handle, ptr(p.x), ptr(p.e), ptr(t), _lib.MEM_HOST,
            int(global_batch), C.byref(loss)))
        return float(loss.value)

    def update(self):
        check(lib().athena_cuda_network_update(self.handle))

    def _advance_optimiser(self):
        """The head of network%update (athena_network_sub.f90:2834-2841, athena_optimiser.f90:414):
        advance the optimiser's iteration counter -- once per epoch when the decay iterates per
        epoch -- and hand the decayed learning rate (and, for per-epoch counters, the iteration
        Adam's bias correction must use) to the device."""
        o = self.optimiser
        if o is None or type(o.lr_decay) is base_lr_decay_type:
            return
        if o.lr_decay.iterate_per_epoch:
            if self.epoch > o.epoch:
                o.epoch = self.epoch
                o.iter += 1
            check(lib().athena_cuda_network_set_iteration(self.handle, max(o.iter, 1)))
        else:
            o.iter += 1
        self.set_learning_rate(o.lr_decay.get_lr(o.learning_rate, o.iter))

    # -- network%train -------------------------------------------------------
    def train(self, input: Union[Sequence[graph_type], PackedGraphs], output, num_epochs: int = 1,
              batch_size: Optional[int] = None, shuffle_batches: bool = True, verbose: int = 0,
              seed: int = 0, resident: bool = False) -> List[float]:
        """Batch loop of athena_network_sub.f90:3575-3670.  `output` is the
        target: for a Kipf-last network the per-vertex target [V_tot, F_T] (the
        reference passes graph_type targets); for a Duvenaud-last network
        [num_samples, num_outputs].

        resident=True keeps the data set on the device, the way the reference keeps it in
        memory for the whole call (save_input / save_output, :3564-3565): features, edge
        features and targets are uploaded once, every mini-batch's device graph is built once,
        and a step then moves nothing over PCIe but the loss.  Same arithmetic, same results."""
        if not self.compiled:
            raise AthenaCudaError(-5, "network is not compiled")
        packed = input if isinstance(input, PackedGraphs) else pack_graphs(input)
        bs = int(batch_size or self.batch_size or packed.B)
        num_samples = packed.B
        num_batches = (num_samples + bs - 1) // bs
        order = np.arange(num_batches)
        rng = np.random.default_rng(seed)
        target = np.ascontiguousarray(output, np.float32)
        kipf_last = self.model[-1].name == "kipf"
        voff = np.concatenate([[0], np.cumsum(packed.nv, dtype=np.int64)])
        history = []
        if resident:
            eoff = np.concatenate([[0], np.cumsum(packed.ne, dtype=np.int64)])
            x_d = _lib.DeviceArray.from_host(np.ascontiguousarray(packed.x, np.float32))
            e_d = None if packed.e is None else \
                _lib.DeviceArray.from_host(np.ascontiguousarray(packed.e, np.float32))
            t_d = _lib.DeviceArray.from_host(target)
            fx = packed.x.shape[1]
            fe = 0 if packed.e is None else packed.e.shape[1]
            ft = int(target.size // (voff[-1] if kipf_last else num_samples))
            batches = {}
        for epoch in range(1, num_epochs + 1):
            self.epoch = epoch
            if shuffle_batches:
                rng.shuffle(order)
            avg = 0.0
            for b in order:
                self._advance_optimiser()
                s0, s1 = int(b) * bs, min((int(b) + 1) * bs, num_samples)
                if resident:
                    if int(b) not in batches:
                        batches[int(b)] = GraphBatch(packed.slice(s0, s1))
                    batch = batches[int(b)]
                    t0 = voff[s0] if kipf_last else s0
                    loss = C.c_float()
                    check(lib().athena_cuda_network_train_step(
                        self.handle, batch.handle, C.c_void_p(x_d.addr + 4 * fx * int(voff[s0])),
                        None if e_d is None else C.c_void_p(e_d.addr + 4 * fe * int(eoff[s0])),
                        C.c_void_p(t_d.addr + 4 * ft * int(t0)), _lib.MEM_DEVICE, 0,
                        C.byref(loss)))
                    avg += float(loss.value)
                    continue
                batch = GraphBatch(packed.slice(s0, s1))
                tgt = target[voff[s0]:voff[s1]] if kipf_last else target[s0:s1]
                avg += self.train_step(batch, tgt)
                batch.destroy()
            self.loss_val = avg / num_batches
            history.append(self.loss_val)
            if verbose:
                print(f"epoch {epoch}: loss {self.loss_val:.6e}")
        if resident:
            for batch in batches.values():
                batch.destroy()
            x_d.free()
            t_d.free()
            if e_d is not None:
                e_d.free()
        return history

    def destroy(self):
        if self.handle:
            lib().athena_cuda_network_destroy(self.handle)
            self.handle = 0
            for layer in self.model:
                layer.handle = 0

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
