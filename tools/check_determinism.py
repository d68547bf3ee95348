#!/usr/bin/env python
"""Bitwise run-to-run determinism of the cfg2 training step (needs a B200).

    python tools/check_determinism.py [repeats [steps]]

Trains `steps` (3) steps from the same parameters `repeats` times and compares parameters and losses
bit for bit with the first run.  Useful with the A/B switches (ATHENA_CUDA_DISABLE_TCG,
ATHENA_CUDA_LIB) to locate a race.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athena_b200 as ab  # noqa: E402
from athena_b200 import synth  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rng = np.random.default_rng(2024)
    p = synth.regular_batch(4096, 64, 6, 64, rng)
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([64, 64], 1, "relu"))
    net.add(ab.kipf_msgpass_layer_type([64, 64], 1, "none"))
    net.compile(ab.sgd_optimiser_type(0.01), batch_size=p.B)
    params0 = (rng.standard_normal(net.num_params) / 8).astype(np.float32)
    target = rng.standard_normal((p.V, 64)).astype(np.float32)
    batch = ab.GraphBatch(p)
    first = None
    bad = 0
    for r in range(reps):
        net.set_params(params0)
        losses = [net.train_step(batch, target) for _ in range(steps)]
        prm = net.get_params()
        if first is None:
            first = (losses, prm)
        elif losses != first[0] or not np.array_equal(prm, first[1]):
            bad += 1
            d = np.abs(prm - first[1])
            print(f"run {r}: differs: losses {losses} vs {first[0]}; params max abs diff "
                  f"{d.max():.3e} at {int(d.argmax())} ({int((d > 0).sum())} elements differ; "
                  f"first layer {int((d[:4096] > 0).sum())}, second layer {int((d[4096:] > 0).sum())})")
    print(f"{bad} of {reps - 1} repeats differ from the first run")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
