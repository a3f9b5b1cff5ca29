"""Model hyper-parameters: the config surface of the reference, kept field for field.

Mirrors `/root/reference/config.py:5-19` (`ModelArgs`): same field names, order and
defaults, so code written against the reference (`ModelArgs(); args.dim = ...`) runs
unchanged.  As in the reference there is no `hidden_dim` field: the FFN width is
inferred from the `up_proj` weight shape at load time.

`rope_theta` and `dtype` are dead fields for the reference's `llama3.py` (the RoPE base
is the hard-coded default 10000 at `llama3.py:31`; activations are float64).  Here
`rope_theta` is likewise ignored (oracle parity), while `dtype` selects the device
arithmetic mode: "float32" (token-identical mode) or "bfloat16" (tensor-core mode).
"""
from dataclasses import dataclass
from typing import Optional


@dataclass
class ModelArgs:
    dim: int = 288  # D
    n_layers: int = 6
    n_heads: int = 6  # HN; HD = dim // n_heads = 48
    n_kv_heads: Optional[int] = None  # KVHN; None -> n_heads
    vocab_size: int = 32000  # VS
    max_seq_len: int = 256  # M
    max_new_tokens: int = 150
    rope_theta: float = 10000.0
    norm_eps: float = 1e-6
    max_batch_size: int = 1
    dtype: str = "float32"


# Named shapes of BASELINE.json's configs (SURVEY.md §8 legend).  `hidden_dim` is not a
# ModelArgs field (see above); it is returned beside the args.
def named_config(name: str, **overrides):
    table = {
        "stories15M": (dict(dim=288, n_layers=6, n_heads=6, n_kv_heads=None,
                            vocab_size=32000, max_seq_len=256), 768),
        "llama3.2-1b": (dict(dim=2048, n_layers=16, n_heads=32, n_kv_heads=8,
                             vocab_size=128256, max_seq_len=2305), 8192),
        "llama3-8b": (dict(dim=4096, n_layers=32, n_heads=32, n_kv_heads=8,
                           vocab_size=128256, max_seq_len=512), 14336),
    }
    fields, hidden = table[name]
    fields = dict(fields)
    hidden = overrides.pop("hidden_dim", hidden)
    fields.update(overrides)
    return ModelArgs(**fields), hidden
