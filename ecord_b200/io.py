"""Result logs on disk (the step after the path).

Mirrors `experiments/utils/system.py:26-106` as used by `experiments/validate.py:244-250` (`--dump_logs`):
`save_to_disk(fname, log, compressed=True)` / `load_from_disk(fname, compressed=True)`.  The reference compresses
with python-snappy; that package is optional here -- when it is missing the same call writes a bz2 pickle
(`.pbz2`, the reference's other codec, system.py:54-72), and `load_from_disk` finds whichever file exists.
Tensors in the log are moved to the CPU first (a rollout log holds CPU tensors already; best_state etc. stay as is).
"""
import bz2
import os
import pickle

import torch

try:
    import snappy as _snappy
except ImportError:                      # not in this image; the reference pins python-snappy in requirements.txt
    _snappy = None


def _with_ext(fname, ext):
    return fname if os.path.splitext(fname)[-1] == ext else fname + ext


def _to_cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


def save_pickle(fname, obj, verbose=False):
    fname = _with_ext(fname, ".pkl")
    with open(fname, "wb") as f:
        pickle.dump(_to_cpu(obj), f)
    return fname


def load_pickle(fname, verbose=False):
    with open(_with_ext(fname, ".pkl"), "rb") as f:
        return pickle.load(f)


def save_to_disk(fname, obj, compressed=True, verbose=False):
    """Returns the path written.  compressed=True: `<fname>.snappy` (python-snappy present) or `<fname>.pbz2`."""
    if not compressed:
        return save_pickle(fname, obj, verbose)
    obj = _to_cpu(obj)
    if _snappy is not None:
        out = _with_ext(fname, ".snappy")
        with open(out, "wb") as f:
            f.write(_snappy.compress(pickle.dumps(obj)))
    else:
        out = _with_ext(fname, ".pbz2")
        with bz2.BZ2File(out, "w") as f:
            pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)
    if verbose:
        print(f"Saved to {out}.")
    return out


def load_from_disk(fname, compressed=True, verbose=False):
    if not compressed:
        return load_pickle(fname, verbose)
    base = os.path.splitext(fname)[0] if os.path.splitext(fname)[-1] in (".snappy", ".pbz2") else fname
    if os.path.exists(base + ".snappy"):
        if _snappy is None:
            raise ImportError(f"{base}.snappy needs python-snappy to be read")
        with open(base + ".snappy", "rb") as f:
            return pickle.loads(_snappy.uncompress(f.read()))
    with bz2.BZ2File(base + ".pbz2", "rb") as f:
        return pickle.load(f)
