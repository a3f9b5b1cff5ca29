"""CPU suite: the C-ABI library loads, exports every symbol include/llama3_b200.h declares,
and the host side fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import HAS_GPU, ROOT
from llama3_np_b200 import _cabi
from llama3_np_b200.config import ModelArgs, named_config
from llama3_np_b200.synth import make_weights, param_count, weight_shapes


def _declared():
    text = open(os.path.join(ROOT, "include", "llama3_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(l3_\w+)\s*\(", text, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_cabi.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert b"sm_100a" in lib.l3_version()


def test_config_struct_matches_header():
    text = open(os.path.join(ROOT, "include", "llama3_b200.h")).read()
    body = re.search(r"typedef struct L3Config \{(.*?)\} L3Config;", text, flags=re.S).group(1)
    fields = re.findall(r"^\s*(?:int32_t|float)\s+(\w+);", body, flags=re.M)
    assert fields == [f[0] for f in _cabi.L3Config._fields_]


def test_modelargs_mirrors_reference_defaults():
    a = ModelArgs()
    assert (a.dim, a.n_layers, a.n_heads, a.n_kv_heads, a.vocab_size, a.max_seq_len, a.max_new_tokens,
            a.rope_theta, a.norm_eps, a.max_batch_size, a.dtype) == (
        288, 6, 6, None, 32000, 256, 150, 10000.0, 1e-6, 1, "float32")


def test_synth_layout_and_param_counts():
    args, hidden = named_config("stories15M")
    shapes = dict((k, s) for k, s, _ in weight_shapes(args, hidden))
    assert shapes["model.layers.0.mlp.up_proj.weight"] == (768, 288)
    assert shapes["lm_head.weight"] == (32000, 288)
    assert len(shapes) == 3 + 9 * 6
    # params read per decoded token = everything but the embedding table + one row
    assert param_count(args, hidden) - 32000 * 288 + 288 == 15_192_000
    a8, h8 = named_config("llama3-8b")
    assert param_count(a8, h8) == 8_030_261_248
    w = make_weights(ModelArgs(dim=64, n_layers=1, n_heads=4, vocab_size=32), 96, seed=3)
    assert all(v.dtype == np.float32 and v.flags.c_contiguous for v in w.values())


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from llama3_np_b200 import Llama
    with pytest.raises(RuntimeError):
        Llama(None, ModelArgs(), hidden_dim=768, random_seed=0)
    out = np.zeros((1, 8), np.float32)
    rc = _cabi.lib().l3_op_rmsnorm(0, _cabi.f32p(out), _cabi.f32p(out), 1e-6, 1, 8, _cabi.f32p(out))
    assert rc != 0


def test_bad_config_is_value_error():
    from llama3_np_b200 import Llama
    with pytest.raises(ValueError):
        Llama(None, ModelArgs(dim=100, n_heads=6), hidden_dim=64, random_seed=0)  # dim % n_heads
    with pytest.raises(ValueError):
        Llama(None, ModelArgs(dtype="float16"), hidden_dim=768, random_seed=0)


def test_every_runtime_switch_is_documented():
    """INTEGRATION.md's switch table names every L3_* environment variable the library or its Python host reads: a
    bench record lists the active switches by these names (bench.py ENV_SWITCHES), so an undocumented one would make a
    number unexplainable."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for path in glob.glob(os.path.join(root, "llama3.np_b200", "csrc", "*")):
        with open(path) as f:
            names |= set(re.findall(r'getenv\("(L3_[A-Z0-9_]+)"\)', f.read()))
    for path in glob.glob(os.path.join(root, "llama3.np_b200", "*.py")):
        with open(path) as f:
            names |= set(re.findall(r'environ\.get\("(L3_[A-Z0-9_]+)"', f.read()))
    assert len(names) > 20
    with open(os.path.join(root, "INTEGRATION.md")) as f:
        doc = f.read()
    missing = sorted(n for n in names if n not in doc)
    assert not missing, f"switches read by the code but absent from INTEGRATION.md: {missing}"
