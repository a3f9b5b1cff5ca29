"""Weight IO - the reference's `utils.py:4-5` surface (`load_parameters = np.load`)."""
import numpy as np


def load_parameters(model_path):
    """Return a mapping key -> ndarray.  Accepts a path to an `.npz` (as the reference
    does) or an already-loaded mapping (extension: 8B-shaped synthetic weights are
    impractical as a file)."""
    if isinstance(model_path, (str, bytes)) or hasattr(model_path, "__fspath__"):
        return np.load(model_path)
    return model_path
