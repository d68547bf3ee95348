import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu")


@pytest.fixture(scope="session")
def oracle32():
    from oracle.oracle import Oracle
    return Oracle("f32")


@pytest.fixture(scope="session")
def oracle64():
    from oracle.oracle import Oracle
    return Oracle("f64")


@pytest.fixture(scope="session")
def cuda():
    """Initialised libathena_cuda; the GPU tests FAIL (not skip) without a device:
    there is no CPU fallback to hide behind."""
    import athena_b200 as ab
    ab.check(ab.lib().athena_cuda_init(-1))
    return ab


def pytest_sessionfinish(session, exitstatus):
    """Dump the norm-wise / element-wise error pairs the parity comparisons saw."""
    try:
        import json
        import helpers
        if not helpers.ELEM_LOG:
            return
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        log = sorted(helpers.ELEM_LOG, key=lambda t: -t[1])
        with open(os.path.join(out, "elem_err.json"), "w") as f:
            json.dump({"comparisons": len(log), "floor": helpers.ELEM_FLOOR,
                       "worst_elementwise (norm, elem, col)": log[:10],
                       "max_normwise": max(t[0] for t in log),
                       "max_columnwise": max(t[2] for t in log)}, f, indent=1)
    except Exception:
        pass
