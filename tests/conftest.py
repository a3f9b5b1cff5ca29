import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import llama3_np_b200  # noqa: E402,F401  (import shim for the dotted package directory)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _probe_gpu():
    """(has_gpu, reason).  A library that is missing or fails to bind is NOT the same as "no device": on a box
    with an NVIDIA device node it is an error (a GPU run that silently skipped everything would go green with no
    coverage), on a CPU-only box the gpu tests are skipped with the real reason."""
    import ctypes as C
    try:
        from llama3_np_b200 import _cabi
        lib = _cabi.lib()
    except Exception as e:  # missing .so, unresolved symbol, ABI drift
        return False, f"libllama3_b200.so did not load: {e!r}"
    n = C.c_int()
    if lib.l3_device_count(C.byref(n)) != 0 or n.value <= 0:
        return False, "no CUDA device in this container"
    return True, ""


HAS_GPU, NO_GPU_REASON = _probe_gpu()
DEVICE_NODE = os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    if DEVICE_NODE and "did not load" in NO_GPU_REASON:
        raise pytest.UsageError(f"GPU present but {NO_GPU_REASON} - run `python __graft_entry__.py` to build it")
    skip = pytest.mark.skip(reason=NO_GPU_REASON)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_model(name):
    """(args, hidden, weights, fixture) of a golden model case: weights regenerated from the
    recorded seed and checked against the recorded digest."""
    import hashlib
    from llama3_np_b200.config import ModelArgs
    from llama3_np_b200.synth import make_weights
    g = load_golden(name)
    if name == "stories15m_c1":
        fields = dict(dim=288, n_layers=6, n_heads=6, n_kv_heads=None, vocab_size=32000,
                      max_seq_len=256, max_batch_size=1)
    else:
        fields = {}
        for k in g.files:
            if k.startswith("cfg_"):
                v = int(g[k])
                fields[k[4:]] = None if v == -1 else v
    args = ModelArgs(**fields)
    hidden = int(g["hidden"])
    w = make_weights(args, hidden, int(g["seed"]))
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k]).tobytes())
    assert h.hexdigest() == str(g["digest"]), "synthetic weights drifted from the golden fixture"
    return args, hidden, w, g


MODEL_CASES = ["tiny_mha", "tiny_gqa", "hd48_gqa", "hd128_gqa"]
