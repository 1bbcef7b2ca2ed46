"""fly() — drop-in for the reference's flight/__init__.py:12-99, plus fly_population().

`fly(drone, ...)` keeps the reference signature, return value `(reward, real_time, idx)` and
in-place mutation of the Drone.  Internally it is a population of one; evolution loops should
call `fly_population(drones, ...)` once per generation instead of N times fly() — same results,
one kernel launch.
"""
import numpy as np

from .._abi import STATUS_NONFINITE, make_params
from ..drone.drone import Drone
from ..population import fly_population_host


def fly_population(drones, timesteps: int = 2500, delta_t: float = 0.01, visualize=False, n_servers=0,
                   punish_time=False, reward_time=True):
    """Fly every Drone of `drones` (one generation, evolve/evolution.py:64-73).

    Returns [(reward, real_time, idx), ...] in population order and mutates the drones exactly
    like the reference's fly() does.
    """
    if visualize or n_servers:
        raise NotImplementedError("the B200 fitness path is headless (visualize=False, n_servers=0)")
    drones = list(drones)
    if not drones:
        return []
    starts = [d.start_state() for d in drones]
    params = make_params(timesteps=timesteps, reward_time=reward_time, punish_time=punish_time)
    genomes = np.stack([d.controller.genome() for d in drones])
    rows = np.stack([s[0] for s in starts])
    # every Drone freshly reset (the evolution loop, evolve/evolution.py:91-92): the kernel's own initial state;
    # otherwise each drone continues from the state it holds (constructor arguments, a previous flight)
    default_row = np.zeros(rows.shape[1])
    default_row[0:3], default_row[3:6] = np.ravel(params.init_pos), np.ravel(params.init_vel)
    default_row[6:9], default_row[9:12] = np.ravel(params.init_angle), np.ravel(params.init_angle_vel)
    default_row[24:28] = np.ravel(params.init_thrust)
    custom = not all(np.array_equal(r[:18], default_row[:18]) and np.array_equal(r[24:], default_row[24:]) for r in rows)
    records, state = fly_population_host(genomes, params, want_state=True, init_state=rows if custom else None)
    out = []
    for i, d in enumerate(drones):
        if records[i]["status"] & STATUS_NONFINITE:
            # reference: REWARD_COIN_DISTANCE(nan) -> int(nan) raises (drone/drone.py:14)
            raise ValueError("cannot convert float NaN to integer")
        d.load_result(records[i], state[i], before=starts[i][1:])
        if timesteps <= 0:
            # reference: `real_time` is unbound when the loop body never runs (flight/__init__.py:99)
            raise UnboundLocalError("cannot access local variable 'real_time' where it is not associated with a value")
        out.append((d.reward, (int(records[i]["flight_time"]) - 1) * delta_t, i))  # real_time = last loop index * delta_t
    return out


def fly(drone: Drone, timesteps: int = 2500, delta_t: float = 0.01, visualize=False, n_servers=0,
        punish_time=False, reward_time=True, idx=None):
    reward, real_time, _ = fly_population(
        [drone], timesteps=timesteps, delta_t=delta_t, visualize=visualize, n_servers=n_servers,
        punish_time=punish_time, reward_time=reward_time,
    )[0]
    return reward, real_time, idx
