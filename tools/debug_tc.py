"""Scratch diagnostics for the tcgen05 kernels (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import athena_b200 as ab
from athena_b200 import synth
from oracle.oracle import Oracle, Batch, LayerSpec

ab.check(ab.lib().athena_cuda_init(0))
o = Oracle("f32")
o64 = Oracle("f64")
def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))
for graphs in (24, 1000, 4096):
    for T, act in ((1, "none"), (2, "tanh"), (2, "relu")):
        rng = np.random.default_rng(graphs)
        p = synth.regular_batch(graphs, 64, 6, 64, rng)
        nvf = [64] * (T + 1)
        spec = LayerSpec("kipf", nvf, T, activation=act)
        params = (rng.standard_normal(o.num_params([spec])) * 0.2).astype(np.float32)
        g = rng.standard_normal((p.V, 64)).astype(np.float32)
        out_r, dp_r, dx_r = o.layer_fwd_bwd(spec, params, Batch(p.nv, p.ne, p.ia, p.ja, p.x, None), g, want_dx=True)
        L = ab.kipf_msgpass_layer_type(nvf, T, activation=act)
        L.set_params(params); L.set_graph(p)
        out = L.forward(); L.zero_gradients(); dx = L.backward(g, want_input_grad=True); dp = L.get_gradients()
        n1 = 64 * 64
        print(f"graphs={graphs:5d} V={p.V:7d} T={T} act={act:5s} out={rel(out,out_r):.2e} "
              f"dW1={rel(dp[:n1],dp_r[:n1]):.2e} " + (f"dW2={rel(dp[n1:],dp_r[n1:]):.2e} " if T == 2 else "") +
              f"dx={rel(dx,dx_r):.2e}", flush=True)
        _, dp64, dx64 = o64.layer_fwd_bwd(spec, params, Batch(p.nv, p.ne, p.ia, p.ja, p.x, None), g, want_dx=True)
        print(f"      vs f64: gpu dW={rel(dp,dp64):.2e} oracle32 dW={rel(dp_r,dp64):.2e} | gpu dx={rel(dx,dx64):.2e} oracle32 dx={rel(dx_r,dx64):.2e}", flush=True)
        L.destroy()
