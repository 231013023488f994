"""Read the reference's pickled TLeague model files (data/models/*.model) without TLeague: every non-numpy class is replaced
by a permissive stub.  Returns the object tree; `arrays(obj)` lists the numpy arrays in traversal order."""
import pickle

import numpy as np


class _Meta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Meta(name, (Stub,), {})
        setattr(cls, name, sub)
        return sub


class Stub(metaclass=_Meta):
    def __init__(self, *a, **k):
        self.args, self.kw = a, k

    def __setstate__(self, st):
        self.__dict__.update(st if isinstance(st, dict) else {"state": st})


_cache = {}


class _U(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("numpy") or module in ("builtins", "collections", "_codecs", "copyreg"):
            return super().find_class(module, name)
        key = (module, name.split(".")[0])
        if key not in _cache:
            _cache[key] = _Meta(key[1], (Stub,), {})
        cls = _cache[key]
        for p in name.split(".")[1:]:
            cls = getattr(cls, p)
        return cls


def load(path):
    try:
        with open(path, "rb") as f:
            return _U(f).load()
    except pickle.UnpicklingError:
        # joblib.dump with numpy arrays stored raw after the pickle opcodes (environmental_level_hole.model)
        import inspect
        from joblib import numpy_pickle

        class _J(numpy_pickle.NumpyUnpickler):
            def find_class(self, module, name):
                if module.startswith("numpy") or module.startswith("joblib") or module in ("builtins", "collections", "_codecs", "copyreg"):
                    return super().find_class(module, name)
                return _U.find_class(self, module, name)
        with open(path, "rb") as f:
            kw = {"ensure_native_byte_order": False} if "ensure_native_byte_order" in inspect.signature(numpy_pickle.NumpyUnpickler.__init__).parameters else {}
            return _J(path, f, **kw).load()


def walk(o, depth=0, name="root", out=None):
    out = [] if out is None else out
    if isinstance(o, np.ndarray):
        out.append(("  " * depth + name, o.shape, o.dtype))
    elif isinstance(o, (list, tuple)):
        out.append(("  " * depth + name, type(o).__name__, len(o)))
        for i, x in enumerate(o):
            walk(x, depth + 1, "[%d]" % i, out)
    elif isinstance(o, dict):
        for k, v in o.items():
            walk(v, depth + 1, str(k), out)
    elif hasattr(o, "__dict__"):
        out.append(("  " * depth + name, type(o).__name__, ""))
        for k, v in o.__dict__.items():
            walk(v, depth + 1, k, out)
    else:
        out.append(("  " * depth + name, repr(o)[:80], ""))
    return out


if __name__ == "__main__":
    import sys
    for line in walk(load(sys.argv[1])):
        print(*line)
