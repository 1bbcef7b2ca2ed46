"""Best-drone checkpoints in the REFERENCE's format, written and read by hand.

The reference stores `json.dump(jsonpickle.encode(drone))` in evolution/best_drone_<t>.json and
winning_drone_<t>.json (evolve/evolution.py:85-87,116-118) — a JSON *string* whose content is jsonpickle's text
for a `drone.drone.Drone` — and replays it with `jsonpickle.decode(json.load(f), classes=Drone)`, `reset()`,
`enable_visualization()`, `fly()` (fly_best_drone.py:11-20).  jsonpickle is neither required nor imported here:

  * `save_drone(drone, path)` emits that document directly: {"py/object": "drone.drone.Drone", ...} with the
    attribute set and order of the reference's Drone.__init__ (drone/drone.py:32-76: quadcopter, dh, sensors,
    controller, distance_to_coin, reward, coins, flight_time, crashed, position, velocity, angle, angle_vel, thrust),
    sub-objects tagged with the reference's module paths, numpy arrays in the form jsonpickle gives objects that
    define __reduce__: {"py/reduce": [numpy.core.multiarray._reconstruct, (ndarray, (0,), b"b"), (1, shape, dtype,
    False, raw bytes)]}.  Deliberate simplifications, all decoding to equal objects: no {"py/id"} back-references
    (every array carries its dtype in full), Sensor values as plain floats, `quadcopter` and `dh` in their
    just-constructed state (fly_best_drone.py's reset() / enable_visualization() rebuild both before flying).
  * `load_drone(path)` reads the reference's files (and ours): it walks the JSON for the brain — controller.weights,
    layer_spec — and the counters, decoding every array form jsonpickle can have produced (py/reduce with inline or
    back-referenced dtype, the jsonpickle.ext.numpy handler's {"py/object": "numpy.ndarray", "values": ...}, plain
    lists), and builds a learning_drones_b200 Drone WITHOUT drawing from numpy's global RNG.

A plain JSON format (fmt="plain": standard library only, bit-exact floats) remains for tools that do not want
jsonpickle tags.  Tested against the reference's own classes through a restatement of jsonpickle's pickler /
unpickler (oracle/mini_jsonpickle.py, tests/test_host_logic.py) and by re-flying decoded drones.
"""
import base64
import json
import time

import numpy as np

from .drone.quadcopter import parameters as P

PLAIN_FORMAT = "learning-drones-b200/drone/1"
_DTYPE_STATE = {"py/tuple": [3, "<", None, None, None, -1, -1, 0]}


# ---- writing: jsonpickle's text for a reference Drone -------------------------------------------------------------
def _b64(raw: bytes):
    return {"py/b64": base64.b64encode(raw).decode("ascii")}


def _dtype_code(dt: np.dtype) -> str:
    return f"{dt.kind}{dt.itemsize}"  # 'f8', 'f4', 'i8' — first argument of numpy.dtype.__reduce__


def encode_ndarray(a) -> dict:
    """numpy.ndarray.__reduce__ -> (numpy.core.multiarray._reconstruct, (ndarray, (0,), b'b'),
    (1, shape, dtype, is_fortran, raw bytes)) as jsonpickle flattens it (tuples -> py/tuple, bytes -> py/b64)."""
    a = np.ascontiguousarray(a)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    dtype = {"py/reduce": [{"py/type": "numpy.dtype"}, {"py/tuple": [_dtype_code(a.dtype), 0, 1]}, _DTYPE_STATE]}
    return {"py/reduce": [
        {"py/function": "numpy.core.multiarray._reconstruct"},
        {"py/tuple": [{"py/type": "numpy.ndarray"}, {"py/tuple": [0]}, _b64(b"b")]},
        {"py/tuple": [1, {"py/tuple": [int(s) for s in a.shape]}, dtype, False, _b64(a.tobytes())]},
    ]}


def _quadcopter_doc():
    """QuadcopterPhysics.__init__ (drone/quadcopter/physical_model.py:53-73) with drone/quadcopter/parameters.py."""
    G = P.gravity * (4 * P.mass_motor + P.mass_center)
    moments = np.zeros([3, 1])
    if P.mass_payload != 0:
        force_payload = P.gravity * P.mass_payload
        G += force_payload
        moments[:, 0] = [P.x_payload * force_payload, P.y_payload * force_payload, 0]
    return {
        "py/object": "drone.quadcopter.physical_model.QuadcopterPhysics",
        "m_c": P.mass_center, "m_m": P.mass_motor, "m_p": P.mass_payload, "r_m": P.radius_motor_center,
        "c_f": P.coef_force, "c_m": P.coef_moment, "c_w": P.coef_wind,
        "Rot": encode_ndarray(np.identity(3, dtype=np.float32)),
        "G": G,
        "moments_payload": encode_ndarray(moments),
        "inverse_inertia_vector": encode_ndarray(np.array([[1 / P.I_x, 1 / P.I_y, 1 / P.I_z]]).T),
        "angle_control_to_thrust": encode_ndarray(np.array([[1, -1, -1], [-1, -1, 1], [-1, 1, -1], [1, 1, 1]], dtype=np.int64)),
    }


def _datahandler_doc():
    """DataHandler.__init__ with visualize=False, n_servers=0 (datahandling/data_scraper.py:47-92); no directory is created."""
    z = lambda c: encode_ndarray(np.zeros([1, c], dtype=np.float32))
    return {"py/object": "datahandling.data_scraper.DataHandler", "n_servers": 0, "printouts": True,
            "dir_name": "./evolution/" + time.strftime("%d_%m_%Y_%H_%M_%S") + "/", "visualize": False,
            "time": z(1), "thrusters": z(4), "rotation": z(3), "translation": z(3), "wind": z(2), "pid": z(6), "setpoints": z(6)}


def drone_to_jsonpickle_doc(drone) -> dict:
    c = drone.controller
    sensors = [{"py/object": "drone.sensor.sensor.Sensor", "current_acceleration": float(s.current_acceleration),
                "last_acceleration": float(s.last_acceleration), "previous_pos": float(s.previous_pos),
                "previous_vel": float(s.previous_vel), "delta_t": float(s.delta_t), "verbose": bool(s.verbose)} for s in drone.sensors]
    return {
        "py/object": "drone.drone.Drone",
        "quadcopter": _quadcopter_doc(),
        "dh": _datahandler_doc(),
        "sensors": sensors,
        "controller": {"py/object": "evolve.brain.NeuralNetwork",
                       "weights": [encode_ndarray(np.asarray(w, dtype=np.float64)) for w in c.weights],
                       "n_layers": int(c.n_layers),
                       "layer_spec": {repr(int(k)): int(v) for k, v in c.layer_spec.items()}},  # jsonpickle stringifies int keys
        "distance_to_coin": float(drone.distance_to_coin),
        "reward": int(drone.reward),
        "coins": int(drone.coins),
        "flight_time": int(drone.flight_time),
        "crashed": bool(drone.crashed),
        "position": encode_ndarray(drone.position), "velocity": encode_ndarray(drone.velocity),
        "angle": encode_ndarray(drone.angle), "angle_vel": encode_ndarray(drone.angle_vel),
        "thrust": encode_ndarray(np.asarray(drone.thrust, dtype=np.float64)),
    }


def encode_drone(drone) -> str:
    """The string `jsonpickle.encode(drone)` stands for in evolve/evolution.py:85."""
    return json.dumps(drone_to_jsonpickle_doc(drone))


# ---- reading: any array form jsonpickle can have written ---------------------------------------------------------------
def _walk_b64(node):
    if isinstance(node, dict):
        if "py/b64" in node:
            return base64.b64decode(node["py/b64"])
        if "py/b85" in node:
            return base64.b85decode(node["py/b85"])
    return None


def _dtype_of(node, n_items, n_bytes):
    """dtype position of an ndarray state: an inline numpy.dtype reduce, or a {"py/id"} back-reference to an earlier
    dtype object (jsonpickle numbers objects per document; resolved here from the item size)."""
    if isinstance(node, dict) and "py/reduce" in node:
        args = node["py/reduce"][1]
        code = (args.get("py/tuple") if isinstance(args, dict) else args)[0]
        order = "<"
        if len(node["py/reduce"]) > 2 and node["py/reduce"][2]:
            st = node["py/reduce"][2]
            st = st.get("py/tuple") if isinstance(st, dict) else st
            order = st[1] if st[1] in ("<", ">") else "<"
        return np.dtype(order + code if np.dtype(code).itemsize > 1 else code)
    size = n_bytes // max(n_items, 1)
    return np.dtype({8: "<f8", 4: "<f4"}.get(size, "<f8"))


def decode_ndarray(node) -> np.ndarray:
    if isinstance(node, list):
        return np.array([decode_ndarray(v) if isinstance(v, (list, dict)) else v for v in node], dtype=np.float64)
    if not isinstance(node, dict):
        raise ValueError(f"not an encoded array: {node!r}")
    if "py/reduce" in node:  # ndarray.__reduce__ form
        state = node["py/reduce"][2]
        state = state.get("py/tuple") if isinstance(state, dict) else state
        shape = state[1].get("py/tuple") if isinstance(state[1], dict) else state[1]
        raw = _walk_b64(state[4])
        if raw is None:
            raise ValueError("ndarray state without binary data")
        n_items = int(np.prod(shape)) if len(shape) else 1
        dt = _dtype_of(state[2], n_items, len(raw))
        return np.frombuffer(raw, dtype=dt).reshape(shape, order="F" if state[3] else "C").astype(dt.newbyteorder("="), copy=True)
    if str(node.get("py/object", "")).endswith("ndarray"):  # jsonpickle.ext.numpy handler
        dt = np.dtype(node.get("dtype", "float64"))
        vals = node.get("values")
        raw = _walk_b64(vals)
        if raw is not None or isinstance(vals, str):
            raw = raw if raw is not None else base64.b64decode(vals)
            arr = np.frombuffer(raw, dtype=dt)
            if "shape" in node:
                shape = node["shape"].get("py/tuple") if isinstance(node["shape"], dict) else node["shape"]
                arr = arr.reshape(shape)
            return arr.copy()
        return np.array(vals, dtype=dt)
    raise ValueError(f"unknown array encoding with keys {sorted(node)}")


def decode_scalar(node, default=None):
    """A number as jsonpickle writes it: a JSON number, or numpy.core.multiarray.scalar(dtype, bytes) in py/reduce
    form for numpy scalars (e.g. distance_to_coin = np.linalg.norm(...), the Sensor values)."""
    if isinstance(node, (int, float)):
        return node
    if isinstance(node, dict) and "py/reduce" in node:
        args = node["py/reduce"][1]
        args = args.get("py/tuple") if isinstance(args, dict) else args
        raw = _walk_b64(args[1])
        if raw is not None:
            return np.frombuffer(raw, dtype=_dtype_of(args[0], 1, len(raw)))[0].item()
    return default


def _new_drone(weights, layer_spec=None):
    """A learning_drones_b200 Drone carrying `weights`, built without touching numpy's global RNG
    (Drone() / NeuralNetwork() draw their initial weights from it like the reference does)."""
    from .drone.drone import Drone
    from .evolve.brain import NeuralNetwork

    nn = NeuralNetwork.__new__(NeuralNetwork)
    nn.weights = [np.array(w, dtype=np.float64) for w in weights]
    nn.n_layers = len(nn.weights)
    if layer_spec is None:
        layer_spec = {i + 1: w.shape[0] for i, w in enumerate(nn.weights)}
        layer_spec[0] = nn.weights[0].shape[1] - 1
    nn.layer_spec = {int(k): int(v) for k, v in layer_spec.items()}
    d = Drone.__new__(Drone)
    d.port = 65432
    d.controller = nn
    d.reset()
    return d


def drone_from_jsonpickle_doc(doc: dict):
    if not (isinstance(doc, dict) and "controller" in doc):
        raise ValueError("not a jsonpickle-encoded Drone (no `controller`)")
    tag = str(doc.get("py/object", "drone.drone.Drone"))
    if not tag.endswith("Drone"):
        raise ValueError(f"expected a Drone document, found {tag}")
    ctl = doc["controller"]
    d = _new_drone([decode_ndarray(w) for w in ctl["weights"]], ctl.get("layer_spec"))
    d.reward, d.coins, d.flight_time = (int(decode_scalar(doc.get(k, 0), 0)) for k in ("reward", "coins", "flight_time"))
    d.crashed = bool(decode_scalar(doc.get("crashed", False), False))
    d.distance_to_coin = float(decode_scalar(doc.get("distance_to_coin", np.inf), np.inf))
    return d


def decode_drone(text: str):
    """Counterpart of `jsonpickle.decode(jsonstring, classes=Drone)` (fly_best_drone.py:13).  Like that script's next
    line, the returned Drone is in the reset state with the saved brain; the saved counters are kept for inspection."""
    return drone_from_jsonpickle_doc(json.loads(text))


# ---- plain format ----------------------------------------------------------------------------------------------------
def drone_to_dict(drone) -> dict:
    c = drone.controller
    return {
        "format": PLAIN_FORMAT,
        "layer_spec": {str(k): int(v) for k, v in c.layer_spec.items()},
        "weights": [np.asarray(w, dtype=np.float64).tolist() for w in c.weights],
        "reward": int(drone.reward),
        "coins": int(drone.coins),
        "flight_time": int(drone.flight_time),
        "crashed": bool(drone.crashed),
        "distance_to_coin": float(drone.distance_to_coin),
    }


def drone_from_dict(doc: dict):
    if doc.get("format") != PLAIN_FORMAT:
        raise ValueError(f"not a {PLAIN_FORMAT} document")
    d = _new_drone(doc["weights"], doc["layer_spec"])
    d.reward, d.coins, d.flight_time = int(doc["reward"]), int(doc["coins"]), int(doc["flight_time"])
    d.crashed, d.distance_to_coin = bool(doc["crashed"]), float(doc["distance_to_coin"])
    return d


def save_drone(drone, path, fmt: str = "jsonpickle"):
    """fmt='jsonpickle' (default): the reference's file, `json.dump` of jsonpickle's text (evolution.py:85-87);
    fmt='plain': a plain JSON object."""
    if fmt == "jsonpickle":
        payload = encode_drone(drone)
    elif fmt == "plain":
        payload = drone_to_dict(drone)
    else:
        raise ValueError("fmt must be 'jsonpickle' or 'plain'")
    with open(path, "w") as f:
        json.dump(payload, f)
    return path


def load_drone(path):
    """Reads either format: the reference's files hold a JSON *string* (jsonpickle text), plain ones a JSON object."""
    with open(path) as f:
        payload = json.load(f)
    if isinstance(payload, str):
        return decode_drone(payload)
    if isinstance(payload, dict) and "py/object" in payload:  # jsonpickle text that was dumped without the second encoding
        return drone_from_jsonpickle_doc(payload)
    return drone_from_dict(payload)
