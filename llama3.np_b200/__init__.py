"""llama3.np_b200 - a B200-native (sm_100a) implementation of llama3.np's Llama-3 forward
pass and greedy generate loop, behind the reference's own Python surface.

Import as `llama3_np_b200` (see the shim `llama3_np_b200.py` at the repo root):

    from llama3_np_b200 import Llama, ModelArgs

Everything numerical runs in hand-written CUDA kernels reached through the C-ABI in
`include/llama3_b200.h`; there is no CPU fallback.
"""
from .config import ModelArgs, named_config  # noqa: F401
from .utils import load_parameters  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require the built library
    if name in ("Llama", "compute_cos_sin_cache"):
        from . import llama3 as _l
        return getattr(_l, name)
    if name == "Tokenizer":
        from .tokenizer import Tokenizer
        return Tokenizer
    raise AttributeError(name)
