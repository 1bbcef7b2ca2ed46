"""Physical constants and initial conditions of the simulated quadcopter.

Same names and values as the reference's drone/quadcopter/parameters.py:13-52
(payload table disabled there, `payload_set = None`, parameters.py:13), so
`from ...parameters import *` call sites keep working.
"""
import numpy as np

payload_set = None

# geometry and mass
mass_center = 0.200  # kg
mass_motor = 0.050  # kg
radius_motor_center = 0.1  # m
I_x = 2160e-6  # kg m^2
I_y = 2160e-6
I_z = 4210e-6

gravity = 9.81  # m/s^2
delta_time = 0.05  # s (unused by fly(); the integrator step is drone.drone.delta_t)

coef_force = 5  # N
coef_moment = 0.5  # N
coef_wind = 0.1  # kg/m
mass_payload = 0
x_payload = 0
y_payload = 0

kp = 1
ti = 1
td = 1
ki = kp / ti
kd = kp * td
limitRange = [0, 1]

initial_thrust = np.array([[0, 0, 0, 0]]).T
initial_wind_speed = np.array([[0.0, 0.0, 0.0]]).T
initial_angle = np.array([[0, 0, 0]]).T * np.pi / 180
initial_angle_vel = np.array([[0, 0, 0]]).T
initial_pos = np.array([[0.0, 0.0, 10.0]]).T
initial_vel = np.array([[0.0, 0.0, 0.0]]).T
