"""Trajectory tap in the reference DataHandler's on-disk layout.

The reference logs per-step data through `DataHandler.new_data` and writes it with `_save_npy` /
`_save_csv` (datahandling/data_scraper.py:94-101,270-307).  `save_trace` writes the same files
(time / thrust_values / rotation / translation / wind / pid / setpoints .npy as float32 with the
leading all-zero row, and Results.csv with the same header and degree conversion) from a trace
returned by `population.trace_drone`, so the reference's plotting scripts can read a GPU flight.
No sockets, no visualisation.
"""
import os

import numpy as np

from .objects.coins import Coin

CSV_HEADER = ("Time, Roll, Pitch, Yaw, X, Y, Z, T1, T2, T3, T4, Wind X, Wind Y,"
              " PID Roll, PID Pitch, PID Yaw, PID X, PID Y, PID Z, Set Roll,"
              " Set Pitch, Set Yaw, Set X, Set Y, Set Z")


def coin_schedule(trace, initial_pos=(0.0, 0.0, 10.0)):
    """Coin position seen by the brain at every step of a fly() trace (flight/__init__.py:25,47-58):
    the test runs on the PRE-step position, the first coin has radius 2, later ones radius 1."""
    coin = Coin(np.array([[5.0], [5.0], [10.0]]), 2, 2000)
    pos = np.asarray(initial_pos, dtype=np.float64).reshape(3, 1)
    out = np.zeros((len(trace), 3))
    for t, row in enumerate(trace):
        if coin.is_in_vicinity(pos):
            coin = Coin(coin.position + np.array([[5.0], [2.5], [0.0]]), 1, value=1000)
        out[t] = coin.position.ravel()
        pos = np.asarray(row[0:3], dtype=np.float64).reshape(3, 1)
    return out


def save_trace(dir_name, trace, delta_t=0.01, wind=(0.0, 0.0), initial_pos=(0.0, 0.0, 10.0)):
    """trace: (steps, 28) rows of population.trace_drone (pos, vel, angle, angle_vel, acc, acc_last, thrust)."""
    trace = np.asarray(trace, dtype=np.float64)
    n = trace.shape[0]
    os.makedirs(dir_name, exist_ok=True)
    z = lambda c: np.zeros([1, c], dtype=np.float32)  # DataHandler.__init__ starts every array with a zero row
    time = np.concatenate([z(1), (np.arange(n) * delta_t).astype(np.float32)[:, None]])
    thrusters = np.concatenate([z(4), trace[:, 24:28].astype(np.float32)])
    rotation = np.concatenate([z(3), trace[:, 6:9].astype(np.float32)])
    translation = np.concatenate([z(3), trace[:, 0:3].astype(np.float32)])
    wind_a = np.concatenate([z(2), np.tile(np.asarray(wind, dtype=np.float32), (n, 1))])
    pid = np.concatenate([z(6), np.zeros((n, 6), dtype=np.float32)])  # fly() logs zeros (drone/drone.py:119)
    setpoints = np.concatenate([z(6), np.concatenate([np.zeros((n, 3)), coin_schedule(trace, initial_pos)], axis=1).astype(np.float32)])
    for name, a in (("time", time), ("thrust_values", thrusters), ("rotation", rotation), ("translation", translation),
                    ("wind", wind_a), ("pid", pid), ("setpoints", setpoints)):
        np.save(os.path.join(dir_name, name), a)
    results = np.concatenate([time, rotation * 180 / np.pi, translation, thrusters, wind_a, pid, setpoints], axis=1)
    np.savetxt(os.path.join(dir_name, "Results.csv"), results, delimiter=",", header=CSV_HEADER)
    return results
