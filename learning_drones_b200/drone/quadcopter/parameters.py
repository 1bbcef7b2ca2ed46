"""Physical constants and initial conditions of the simulated quadcopter.

Exports the same module-level names, with the same values, as the reference's
drone/quadcopter/parameters.py:13-52 so that `from ...parameters import *` call sites keep
working.  The reference's optional payload table (`iom_collection.npy`, parameters.py:54-63) is
disabled there (`payload_set = None`) and is not reproduced.
"""
import numpy as np

_SCALARS = {
    "payload_set": None,
    # geometry and mass [kg, m, kg m^2]
    "mass_center": 0.200, "mass_motor": 0.050, "radius_motor_center": 0.1,
    "I_x": 2160e-6, "I_y": 2160e-6, "I_z": 4210e-6,
    # environment
    "gravity": 9.81,        # m/s^2
    "delta_time": 0.05,     # s — unused by fly(); the integrator step is drone.drone.delta_t
    "coef_force": 5,        # N per unit thrust command
    "coef_moment": 0.5,     # N
    "coef_wind": 0.1,       # kg/m
    "mass_payload": 0, "x_payload": 0, "y_payload": 0,
    # generic PID constants of the reference module (kp, ti, td and what follows from them)
    "kp": 1, "ti": 1, "td": 1,
}
_SCALARS["ki"] = _SCALARS["kp"] / _SCALARS["ti"]
_SCALARS["kd"] = _SCALARS["kp"] * _SCALARS["td"]
_SCALARS["limitRange"] = [0, 1]

# column vectors with the dtypes the reference creates (integer thrust / angular velocity!)
_VECTORS = {
    "initial_thrust": np.zeros((4, 1), dtype=np.int64),
    "initial_wind_speed": np.zeros((3, 1), dtype=np.float64),
    "initial_angle": np.zeros((3, 1), dtype=np.int64) * np.pi / 180,
    "initial_angle_vel": np.zeros((3, 1), dtype=np.int64),
    "initial_pos": np.array([[0.0], [0.0], [10.0]]),
    "initial_vel": np.zeros((3, 1), dtype=np.float64),
}

globals().update(_SCALARS)
globals().update(_VECTORS)
__all__ = ["np", *_SCALARS, *_VECTORS]
