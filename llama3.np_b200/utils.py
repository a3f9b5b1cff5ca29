"""Weight IO - the reference's `utils.py:4-5` surface (`load_parameters = np.load`)."""
import numpy as np


def load_parameters(model_path):
    """Return a mapping key -> ndarray.  Accepts a path to an `.npz` (as the reference
    does) or an already-loaded mapping (extension: 8B-shaped synthetic weights are
    impractical as a file)."""
    if isinstance(model_path, (str, bytes)) or hasattr(model_path, "__fspath__"):
        return np.load(model_path)
    return model_path


def checkpoint_digest(model_path, chunk: int = 1 << 24) -> str:
    """sha256 (hex) of the checkpoint FILE's bytes: the key of the packed device-layout cache
    (`Llama(..., cache_dir=...)`), so an edited or replaced `.npz` can never be served from a stale pack."""
    import hashlib
    h = hashlib.sha256()
    with open(model_path, "rb") as f:
        while True:
            b = f.read(chunk)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def packed_cache_path(cache_dir, digest: str, dtype: str, tp_rank: int = 0, tp_world: int = 1) -> str:
    """Where the pack of one (checkpoint, dtype, tensor-parallel placement) lives inside `cache_dir`."""
    import os
    kind = "bf16" if dtype in ("bfloat16", "bf16") else "f32"
    return os.path.join(os.fspath(cache_dir), f"{digest[:40]}.{kind}.tp{tp_rank}of{tp_world}.l3pack")
