#!/usr/bin/env python
"""Bitwise run-to-run determinism of a training step (needs a B200).

    python tools/check_determinism.py [repeats [steps [cfg2|cfg4|cfg1]]]

Trains `steps` (3) steps from the same parameters `repeats` times and compares parameters and losses
bit for bit with the first run.  cfg2: the Kipf step of the bench line (SGD); cfg4: Kipf x 2 +
Duvenaud on molecular graphs (Adam; the FP32 tile kernels with their two groups per SM, the
batched K = 32 dW product); cfg1: the chemical Duvenaud layer with Adam + norm clipping (the step
inside the finalize launch).  Useful with the A/B switches (ATHENA_CUDA_DISABLE_TCG,
ATHENA_CUDA_LIB) to locate a race.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athena_b200 as ab  # noqa: E402
from athena_b200 import synth  # noqa: E402


def build(cfg, rng):
    """-> (packed batch, target, make_net) ; make_net() builds a fresh network (optimiser state)"""
    if cfg == "cfg2":
        p = synth.regular_batch(4096, 64, 6, 64, rng)
        target = rng.standard_normal((p.V, 64)).astype(np.float32)

        def make_net():
            net = ab.network_type()
            net.add(ab.kipf_msgpass_layer_type([64, 64], 1, "relu"))
            net.add(ab.kipf_msgpass_layer_type([64, 64], 1, "none"))
            net.compile(ab.sgd_optimiser_type(0.01), batch_size=p.B)
            return net
        scale = 1.0 / 8
    elif cfg == "cfg4":
        p = synth.molecular_batch(8192, 32, 4, rng)
        target = rng.random((p.B, 32)).astype(np.float32)

        def make_net():
            net = ab.network_type()
            net.add(ab.kipf_msgpass_layer_type([32, 32], 1, "relu"))
            net.add(ab.kipf_msgpass_layer_type([32, 32], 1, "relu"))
            net.add(ab.duvenaud_msgpass_layer_type([32], [4], 2, 6, 32))
            net.compile(ab.adam_optimiser_type(0.001), batch_size=p.B)
            return net
        scale = 0.1
    else:
        p = synth.chemical_batch(8, rng)
        target = rng.random((p.B, 10)).astype(np.float32)

        def make_net():
            net = ab.network_type()
            net.add(ab.duvenaud_msgpass_layer_type([6], [1], 4, 10, 10))
            net.compile(ab.adam_optimiser_type(0.01, clip_dict=ab.clip_type(clip_norm=0.1)),
                        batch_size=p.B)
            return net
        scale = 0.3
    return p, target, make_net, scale


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
    rng = np.random.default_rng(2024)
    p, target, make_net, scale = build(cfg, rng)
    params0 = None
    batch = ab.GraphBatch(p)
    first = None
    bad = 0
    for r in range(reps):
        net = make_net()
        if params0 is None:
            params0 = (rng.standard_normal(net.num_params) * scale).astype(np.float32)
        net.set_params(params0)
        losses = [net.train_step(batch, target) for _ in range(steps)]
        prm = net.get_params()
        if first is None:
            first = (losses, prm)
        elif losses != first[0] or not np.array_equal(prm, first[1]):
            bad += 1
            d = np.abs(prm - first[1])
            print(f"run {r}: differs: losses {losses} vs {first[0]}; params max abs diff "
                  f"{d.max():.3e} at {int(d.argmax())} ({int((d > 0).sum())} elements differ)")
    print(f"{cfg}: {bad} of {reps - 1} repeats differ from the first run")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
