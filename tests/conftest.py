import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("QVC_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _gpu_unavailable_reason():
    """Why the `gpu` tests cannot run here, or None on a B200 box.  A missing libqvc_b200.so is NOT a reason to skip:
    with a B200 present the tests run and fail loudly in capi.load()."""
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device"
        if torch.cuda.get_device_capability(0) != (10, 0):
            return "device 0 is not sm_100 (B200)"
    except Exception as e:      # noqa: BLE001
        return f"{type(e).__name__}: {e}"
    return None


def pytest_collection_modifyitems(config, items):
    # a plain `pytest` on a CPU-only host skips the GPU tests instead of failing in torch.cuda init; on the GPU box
    # nothing is skipped, so a missing library or device there still fails loudly in the tests themselves
    if not any("gpu" in item.keywords for item in items):
        return
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=f"needs a B200: {reason}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def shapes():
    import json
    with open(os.path.join(GOLDEN, "state_dict_shapes.json")) as f:
        return {k: tuple(v) for k, v in json.load(f).items()}


@pytest.fixture(scope="session")
def sd(shapes):
    import synth
    return synth.synthetic_state_dict(shapes, 0)


@pytest.fixture(scope="session")
def model_cfg():
    import json
    with open(os.path.join(GOLDEN, "quickvc_model_config.json")) as f:
        return json.load(f)


def load_golden(name):
    import numpy as np
    import torch
    with np.load(os.path.join(GOLDEN, f"infer_{name}.npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


GOLDEN_CASES = {
    "small": (2, 24, 1, 200),
    "shortmel": (2, 16, 2, 100),
    "cfg1": (1, 250, 1, 250),
    "chunk": (1, 25, 1, 250),
}
