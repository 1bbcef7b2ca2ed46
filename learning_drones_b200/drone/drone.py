"""Drone — state / brain / reward container with the reference's attribute names
(drone/drone.py:20-76,167-218).  The physics, integrator, controller and reward rules that the
reference implements in Drone.measure / translate_input_to_thrust run inside the CUDA episode
kernel; fly() (learning_drones_b200.flight) writes the outcome back onto these attributes.

Differences from the reference, on purpose: no DataHandler is constructed (the reference's
constructor creates ./evolution/<timestamp>/ on disk, data_scraper.py:58-66) and there is no
visualisation socket.
"""
import numpy as np

from .._abi import (  # noqa: F401  (re-exported constants of drone/drone.py:10-17)
    ANGLE_LIMIT,
    ANGLE_VEL_LIMIT,
    REWARD_CRASHED,
    REWARD_TIME_PASSED,
    REWARD_UNSTABLE,
    STATE_LEN,
)
from ..evolve.brain import NeuralNetwork
from .quadcopter import parameters as P

delta_t = 0.01
REWARD_COIN_DISTANCE = lambda d: -100 * int(d)  # noqa: E731


class Sensor:
    """Per-DOF integrator state (reference drone/sensor/sensor.py:25-31); filled in by fly()."""

    def __init__(self, delta_t, previous_pos, previous_vel, verbose=False):
        self.current_acceleration = 0.0
        self.last_acceleration = 0.0
        self.previous_pos = previous_pos
        self.previous_vel = previous_vel
        self.delta_t = delta_t
        self.verbose = verbose


def _col(v, n):
    return np.array(v, dtype=np.float64).reshape(n, 1)


class Drone:
    def __init__(
        self,
        port: int = 65432,
        visualize=False,
        n_servers=0,
        initial_pos=None,
        initial_vel=None,
        initial_angle=None,
        initial_angle_vel=None,
    ):
        if visualize or n_servers:
            raise NotImplementedError("the B200 fitness path is headless (visualize=False, n_servers=0)")
        self.port = port
        self.controller = NeuralNetwork(layer_spec={1: 10, 2: 10, 3: 4})
        # reference drone/drone.py:55-76: the CONSTRUCTOR arguments set the state the first flight starts from
        # (attributes and the six Sensors); reset() below always goes back to the module-level defaults
        self.distance_to_coin = np.inf
        self.reward = 0
        self.coins = 0
        self.flight_time = 0
        self.crashed = False
        self.position = _col(P.initial_pos if initial_pos is None else initial_pos, 3)
        self.velocity = _col(P.initial_vel if initial_vel is None else initial_vel, 3)
        self.angle = _col(P.initial_angle if initial_angle is None else initial_angle, 3)
        self.angle_vel = _col(P.initial_angle_vel if initial_angle_vel is None else initial_angle_vel, 3)
        self.thrust = _col(P.initial_thrust, 4)
        self.sensors = [Sensor(delta_t, self.position[i, 0], self.velocity[i, 0]) for i in range(3)]
        self.sensors += [Sensor(delta_t, self.angle[i, 0], self.angle_vel[i, 0]) for i in range(3)]

    # reference drone/drone.py:177-218 (always resets to the module-level initial state)
    def reset(self):
        self.distance_to_coin = np.inf
        self.reward = 0
        self.crashed = False
        self.coins = 0
        self.flight_time = 0
        self.position = _col(P.initial_pos, 3)
        self.velocity = _col(P.initial_vel, 3)
        self.angle = _col(P.initial_angle, 3)
        self.angle_vel = _col(P.initial_angle_vel, 3)
        self.thrust = _col(P.initial_thrust, 4)
        self.sensors = [Sensor(delta_t, self.position[i, 0], self.velocity[i, 0]) for i in range(3)]
        self.sensors += [Sensor(delta_t, self.angle[i, 0], self.angle_vel[i, 0]) for i in range(3)]

    def reset_reward(self):
        self.reward = 0

    # reference drone/drone.py:167-172 — same draws from the global numpy stream, in the same order
    def mutate(self, mutation_rate):
        for i, layer in enumerate(self.controller.weights):
            for j, row in enumerate(layer):
                for k, _ in enumerate(row):
                    if np.random.random_sample() < mutation_rate:
                        self.controller.weights[i][j][k] += np.random.normal(0.0, 0.3)

    # -- single steps, for code that drives the episode loop itself (reference drone.py:83-86,129-165).
    #    Both run on the GPU with the episode kernel's device functions (ld_drone_step_host).
    def _state_row(self):
        s = np.zeros(STATE_LEN)
        s[0:3], s[3:6] = self.position.ravel(), self.velocity.ravel()
        s[6:9], s[9:12] = self.angle.ravel(), self.angle_vel.ravel()
        s[12:18] = [x.current_acceleration for x in self.sensors]
        s[18:24] = [x.last_acceleration for x in self.sensors]
        s[24:28] = np.asarray(self.thrust, dtype=np.float64).ravel()
        return s

    def _load_state_row(self, s):
        self.position, self.velocity = s[0:3].reshape(3, 1).copy(), s[3:6].reshape(3, 1).copy()
        self.angle, self.angle_vel = s[6:9].reshape(3, 1).copy(), s[9:12].reshape(3, 1).copy()
        self.thrust = s[24:28].reshape(4, 1).copy()
        pos, vel = np.concatenate([s[0:3], s[6:9]]), np.concatenate([s[3:6], s[9:12]])
        for i, sensor in enumerate(self.sensors):
            sensor.previous_pos, sensor.previous_vel = float(pos[i]), float(vel[i])
            sensor.current_acceleration, sensor.last_acceleration = float(s[12 + i]), float(s[18 + i])

    def measure(self, wind_speed):
        """One physics + integrator step and the crash / unstable rules; returns True when crashed."""
        import ctypes as C

        from .. import _lib
        from .._abi import make_params

        state = self._state_row()[None, :].copy()
        flags = np.zeros((1, 2), dtype=np.int32)
        p = make_params(timesteps=1, wind_speed=wind_speed)
        _lib.check(_lib.load().ld_drone_step_host(C.byref(p), 1, None, state.ctypes.data, None, flags.ctypes.data, 1))
        self._load_state_row(state[0])
        if flags[0, 0]:
            self.reward += REWARD_CRASHED
            self.crashed = True
            return True
        self.reward += REWARD_UNSTABLE * int(flags[0, 1])
        return False

    def translate_input_to_thrust(self, coin_position):
        import ctypes as C

        from .. import _lib
        from .._abi import make_params

        state = self._state_row()[None, :].copy()
        genome = self.controller.genome()[None, :].copy()
        coin = np.asarray(coin_position, dtype=np.float64).reshape(1, 3).copy()
        p = make_params(timesteps=1)
        _lib.check(_lib.load().ld_drone_step_host(C.byref(p), 2, genome.ctypes.data, state.ctypes.data, coin.ctypes.data, None, 1))
        self.thrust = state[0, 24:28].reshape(4, 1).copy()

    # -- used by flight.fly -------------------------------------------------------------------
    def start_state(self):
        """(state row, reward, coins, flight_time, crashed) an episode starts from: whatever the Drone holds, as
        in the reference where fly() continues from the Drone's state and adds to its counters
        (flight/__init__.py:35-45).  The integrator reads the Sensors (previous_pos / previous_vel / accelerations),
        the physics and the brain read the attributes (drone/drone.py:83-86,131-148); both are in the row."""
        s = self._state_row()
        pos = np.array([x.previous_pos for x in self.sensors], dtype=np.float64)
        vel = np.array([x.previous_vel for x in self.sensors], dtype=np.float64)
        attr_pos, attr_vel = np.concatenate([s[0:3], s[6:9]]), np.concatenate([s[3:6], s[9:12]])
        if not (np.array_equal(pos, attr_pos) and np.array_equal(vel, attr_vel)):
            raise ValueError("Drone attributes and Sensor state disagree (position/velocity/angle were edited without "
                             "the sensors): the episode kernel keeps ONE copy of the kinematic state")
        return s, int(self.reward), int(self.coins), int(self.flight_time), bool(self.crashed)

    def load_result(self, record, state_row, before=(0, 0, 0, False)):
        """`before`: (reward, coins, flight_time, crashed) the Drone held when the episode started — the reference
        adds to them (`drone.reward +=`, `drone.coins += 1`, `drone.flight_time += 1`; `crashed` is only ever set)."""
        self.reward = before[0] + int(record["reward"])
        self.coins = before[1] + int(record["coins"])
        self.flight_time = before[2] + int(record["flight_time"])
        self.crashed = before[3] or bool(record["status"] & 1)
        self.distance_to_coin = float(record["distance_to_coin"])
        s = np.asarray(state_row, dtype=np.float64).reshape(STATE_LEN)
        self._load_state_row(s)
