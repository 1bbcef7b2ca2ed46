"""Best-drone checkpoints (reference evolve/evolution.py:85-87,116-118 and fly_best_drone.py:11-14).

The reference stores `json.dump(jsonpickle.encode(drone))`.  jsonpickle is a third-party package:
when it is importable, `save_drone(..., fmt="jsonpickle")` / `load_drone` write and read exactly
that (a JSON string holding jsonpickle's encoding of the Drone object); when it is not, asking for
that format raises ImportError — nothing is emulated.  The default format is a plain JSON document
that needs nothing but the standard library and round-trips every float64 bit for bit
(Python's float repr is the shortest string that parses back to the same double).
"""
import json

import numpy as np

PLAIN_FORMAT = "learning-drones-b200/drone/1"


def _jsonpickle():
    try:
        import jsonpickle
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError("the reference's checkpoint format needs the `jsonpickle` package; "
                          "use fmt='plain' (standard library only)") from e
    return jsonpickle


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
    from .drone.drone import Drone

    if doc.get("format") != PLAIN_FORMAT:
        raise ValueError(f"not a {PLAIN_FORMAT} document")
    d = Drone()
    d.controller.layer_spec = {int(k): int(v) for k, v in doc["layer_spec"].items()}
    d.controller.n_layers = len(d.controller.layer_spec) - 1
    d.controller.weights = [np.array(w, dtype=np.float64) for w in doc["weights"]]
    d.reward, d.coins, d.flight_time = int(doc["reward"]), int(doc["coins"]), int(doc["flight_time"])
    d.crashed, d.distance_to_coin = bool(doc["crashed"]), float(doc["distance_to_coin"])
    return d


def save_drone(drone, path, fmt: str = "plain"):
    """fmt='plain' (default) or 'jsonpickle' (the reference's file, readable by its fly_best_drone.py)."""
    if fmt == "jsonpickle":
        payload = _jsonpickle().encode(drone)  # evolution.py:85
    elif fmt == "plain":
        payload = drone_to_dict(drone)
    else:
        raise ValueError("fmt must be 'plain' or 'jsonpickle'")
    with open(path, "w") as f:
        json.dump(payload, f)  # evolution.py:86-87
    return path


def load_drone(path):
    """Reads either format: the reference's files hold a JSON *string* (jsonpickle text), ours a JSON object."""
    with open(path) as f:
        payload = json.load(f)
    if isinstance(payload, str):
        from .drone.drone import Drone

        return _jsonpickle().decode(payload, classes=Drone)  # fly_best_drone.py:13
    return drone_from_dict(payload)
