"""A minimal restatement of jsonpickle 1.4.1's Pickler / Unpickler — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference writes its checkpoints with `json.dump(jsonpickle.encode(drone))` (evolve/evolution.py:85-87,
116-118) and reads them with `jsonpickle.decode(json.load(f), classes=Drone)` (fly_best_drone.py:11-14).
jsonpickle (pinned 1.4.1, requirements_pip_list.txt:25) is a third-party package that is NOT installed in this
image and cannot be fetched, so its published algorithm is restated here for the object kinds a Drone contains,
to test learning_drones_b200/checkpoint.py in both directions:

  * `encode(obj)`  — what jsonpickle.encode produces: {"py/object": "module.Class", **__dict__} for plain
    instances; {"py/reduce": [f, args, state]} for objects whose class overrides __reduce__ (numpy ndarray,
    dtype and scalars); {"py/tuple": [...]}, {"py/b64": ...} for bytes, {"py/type": ...}, {"py/function": ...};
    dict keys stringified; a repeated object (same id()) becomes {"py/id": n}, n counting lists, instances and
    reduced objects in order of first appearance (pre-order, 0-based).
  * `decode(text, classes)` — the inverse: instances are created with cls.__new__(cls) and filled through
    __dict__ (no __init__ call), reduced objects are rebuilt with f(*args) + __setstate__(state), {"py/id": n}
    resolves against the same numbering.

"parity unpinned": there is no jsonpickle here to check this restatement against; it follows the library's
pickler.py / unpickler.py control flow as published (tags, reference numbering, __reduce__ handling).
"""
import base64
import importlib
import json
import types

PRIMITIVES = (str, bool, int, float, type(None))


def _importable_name(cls):
    module = cls.__module__
    if module == "numpy._core.multiarray":  # numpy >= 2 moved the C module; 1.17 wrote numpy.core.multiarray
        module = "numpy.core.multiarray"
    return f"{module}.{cls.__qualname__}"


def _overrides_reduce(obj):
    cls = type(obj)
    if cls in (list, dict, tuple, set, frozenset, object) or issubclass(cls, type):
        return False
    return getattr(cls, "__reduce__", None) is not getattr(object, "__reduce__") or \
        getattr(cls, "__reduce_ex__", None) is not getattr(object, "__reduce_ex__")


class Pickler:
    def __init__(self):
        self._objs = {}
        self._keep = []  # keep flattened objects alive so that id() values are not reused

    def _mkref(self, obj):
        if id(obj) in self._objs:
            return False
        self._objs[id(obj)] = len(self._objs)
        self._keep.append(obj)
        return True

    def flatten(self, obj):
        if type(obj) in PRIMITIVES:
            return obj
        if isinstance(obj, bytes):
            return {"py/b64": base64.b64encode(obj).decode("ascii")}
        if type(obj) is list:
            if not self._mkref(obj):
                return {"py/id": self._objs[id(obj)]}
            return [self.flatten(v) for v in obj]
        if type(obj) is tuple:
            return {"py/tuple": [self.flatten(v) for v in obj]}
        if type(obj) is dict:
            return {(k if isinstance(k, str) else repr(k)): self.flatten(v) for k, v in obj.items()}
        if isinstance(obj, type):
            return {"py/type": _importable_name(obj)}
        if isinstance(obj, (types.FunctionType, types.BuiltinFunctionType)):
            module = "numpy.core.multiarray" if obj.__module__ == "numpy._core.multiarray" else obj.__module__
            return {"py/function": f"{module}.{obj.__name__}"}
        # object instance
        if not self._mkref(obj):
            return {"py/id": self._objs[id(obj)]}
        if _overrides_reduce(obj):
            rv = list(obj.__reduce__())
            rv += [None] * (5 - len(rv))
            flat = [self.flatten(v) for v in rv]
            last = len(flat) - 1
            while last >= 2 and flat[last] is None:
                last -= 1
            return {"py/reduce": flat[: last + 1]}
        data = {"py/object": _importable_name(type(obj))}
        for k, v in obj.__dict__.items():
            data[k if isinstance(k, str) else repr(k)] = self.flatten(v)
        return data


def encode(obj) -> str:
    return json.dumps(Pickler().flatten(obj))


class Unpickler:
    def __init__(self, classes=None):
        classes = classes or []
        if not isinstance(classes, (list, tuple)):
            classes = [classes]
        self._classes = {_importable_name(c): c for c in classes}
        self._objs = []

    def _load(self, name):
        if name in self._classes:
            return self._classes[name]
        module, _, attr = name.rpartition(".")
        try:
            return getattr(importlib.import_module(module), attr)
        except (ImportError, AttributeError):
            return None

    def restore(self, obj):
        if isinstance(obj, list):
            parent = []
            self._objs.append(parent)
            parent.extend(self.restore(v) for v in obj)
            return parent
        if not isinstance(obj, dict):
            return obj
        if "py/b64" in obj:
            return base64.b64decode(obj["py/b64"])
        if "py/id" in obj:
            return self._objs[obj["py/id"]]
        if "py/type" in obj:
            return self._load(obj["py/type"])
        if "py/function" in obj:
            return self._load(obj["py/function"])
        if "py/tuple" in obj:
            return tuple(self.restore(v) for v in obj["py/tuple"])
        if "py/reduce" in obj:
            slot = len(self._objs)
            self._objs.append(None)  # placeholder: children are numbered after their parent
            rv = [self.restore(v) for v in obj["py/reduce"]]
            rv += [None] * (5 - len(rv))
            f, args, state = rv[0], rv[1], rv[2]
            inst = f(*args)
            if state:
                inst.__setstate__(state)
            self._objs[slot] = inst
            return inst
        if "py/object" in obj:
            cls = self._load(obj["py/object"])
            if cls is None:  # class not importable: jsonpickle leaves the dict as it is
                self._objs.append(obj)
                return obj
            inst = cls.__new__(cls)
            self._objs.append(inst)
            for k, v in obj.items():
                if k != "py/object":
                    setattr(inst, k, self.restore(v))
            return inst
        return {k: self.restore(v) for k, v in obj.items()}


def decode(text: str, classes=None):
    return Unpickler(classes).restore(json.loads(text))
