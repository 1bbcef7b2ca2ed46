"""ctypes mirror of include/ldrones.h and the host-side derivation of `ld_params`.

Derived constants are computed here with python floats in exactly the order the
reference computes them (QuadcopterPhysics.__init__,
reference drone/quadcopter/physical_model.py:53-73), so the device receives the
reference's own roundings instead of re-deriving them.
"""
import ctypes as C
import math

import numpy as np

from .drone.quadcopter import parameters as P

LD_ABI_VERSION = 3
SLICE_STAGGER = -1000000  # ld_params.slice_steps = LD_SLICE_STAGGER
N_IN, N_H1, N_H2, N_OUT = 15, 10, 10, 4
LAYER_SHAPES = ((N_H1, N_IN + 1), (N_H2, N_H1 + 1), (N_OUT, N_H2 + 1))
GENOME_LEN = sum(r * c for r, c in LAYER_SHAPES)  # 314
STATE_LEN = 28
PID_GAINS = 18

STATUS_CRASHED = 1
STATUS_NONFINITE = 2
STATUS_FINISHED = 4
PID_MAX_COINS = 8

# drone/drone.py:10-17 of the reference
DELTA_T = 0.01
REWARD_CRASHED = -3000
REWARD_TIME_PASSED = 2
REWARD_COIN_DISTANCE_FACTOR = -100
REWARD_UNSTABLE = -1
ANGLE_LIMIT = np.pi / 4
ANGLE_VEL_LIMIT = 20


class LdParams(C.Structure):
    _fields_ = [
        ("c_f", C.c_double),
        ("c_m", C.c_double),
        ("r_m", C.c_double),
        ("sqrt2_r_m", C.c_double),
        ("inv_I", C.c_double * 3),
        ("moments_payload", C.c_double * 3),
        ("neg_G32", C.c_float),
        ("m_total32", C.c_float),
        ("wind_cw32", C.c_float * 3),
        ("_pad0", C.c_float * 3),
        ("dt", C.c_double),
        ("init_pos", C.c_double * 3),
        ("init_vel", C.c_double * 3),
        ("init_angle", C.c_double * 3),
        ("init_angle_vel", C.c_double * 3),
        ("init_thrust", C.c_double * 4),
        ("angle_limit", C.c_double),
        ("angle_vel_limit", C.c_double),
        ("z_min", C.c_double),
        ("z_max", C.c_double),
        ("reward_time_passed", C.c_int32),
        ("reward_crashed", C.c_int32),
        ("reward_unstable", C.c_int32),
        ("reward_coin_distance", C.c_int32),
        ("coin0_pos", C.c_double * 3),
        ("coin_step", C.c_double * 3),
        ("coin0_radius", C.c_double),
        ("coinN_radius", C.c_double),
        ("coin0_thr2", C.c_double),
        ("coinN_thr2", C.c_double),
        ("coin0_value", C.c_int32),
        ("coinN_value", C.c_int32),
        ("sigmoid_beta", C.c_double),
        ("timesteps", C.c_int32),
        ("reward_time", C.c_int32),
        ("punish_time", C.c_int32),
        ("refill_batch", C.c_int32),
        ("slice_steps", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class LdRecord(C.Structure):
    _fields_ = [
        ("reward", C.c_int64),
        ("coins", C.c_int32),
        ("flight_time", C.c_int32),
        ("status", C.c_int32),
        ("unstable", C.c_int32),
        ("distance_to_coin", C.c_double),
    ]


RECORD_DTYPE = np.dtype(
    [
        ("reward", "<i8"),
        ("coins", "<i4"),
        ("flight_time", "<i4"),
        ("status", "<i4"),
        ("unstable", "<i4"),
        ("distance_to_coin", "<f8"),
    ]
)
assert RECORD_DTYPE.itemsize == C.sizeof(LdRecord) == 32


def sqrt_threshold(radius: float) -> float:
    """Smallest double s with sqrt(s) >= radius (math.sqrt is correctly rounded).

    `sqrt(s) < radius` (reference objects/coins.py:19-21) is then equivalent to
    `s < sqrt_threshold(radius)` for every double s >= 0, which lets the device
    skip the square root in the per-step coin test without changing its truth value.
    """
    r = float(radius)
    if not r > 0.0:
        return 0.0
    s = r * r
    while math.sqrt(s) >= r:
        s = math.nextafter(s, 0.0)
    while math.sqrt(s) < r:
        s = math.nextafter(s, math.inf)
    return s


def _vec(v, n):
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.shape[0] != n:
        raise ValueError(f"expected {n} values, got shape {np.shape(v)}")
    return [float(x) for x in a]


def make_params(
    timesteps: int = 2500,
    reward_time: bool = True,
    punish_time: bool = False,
    wind_speed=None,
    initial_pos=None,
    initial_vel=None,
    initial_angle=None,
    initial_angle_vel=None,
    initial_thrust=None,
    mass_center=P.mass_center,
    mass_motor=P.mass_motor,
    radius_motor_center=P.radius_motor_center,
    coef_force=P.coef_force,
    coef_moment=P.coef_moment,
    coef_wind=P.coef_wind,
    gravity=P.gravity,
    mass_payload=P.mass_payload,
    x_payload=P.x_payload,
    y_payload=P.y_payload,
    I_x=P.I_x,
    I_y=P.I_y,
    I_z=P.I_z,
) -> LdParams:
    """Build `ld_params` for one population evaluation (the arguments of the reference's
    fly() + the module constants it reads)."""
    timesteps = int(timesteps)
    if timesteps < 0:
        raise ValueError("timesteps must be >= 0")
    p = LdParams()
    # QuadcopterPhysics.__init__ (physical_model.py:53-73), same expression order
    G = gravity * (4 * mass_motor + mass_center)
    moments_payload = [0.0, 0.0, 0.0]
    if mass_payload != 0:
        force_payload = gravity * mass_payload
        G += force_payload
        moments_payload = [x_payload * force_payload, y_payload * force_payload, 0.0]
    p.c_f, p.c_m, p.r_m = float(coef_force), float(coef_moment), float(radius_motor_center)
    p.sqrt2_r_m = float(np.sqrt(2) * radius_motor_center)  # physical_model.py:109
    p.inv_I[:] = [1 / I_x, 1 / I_y, 1 / I_z]
    p.moments_payload[:] = moments_payload
    p.neg_G32 = float(np.float32(-1 * G))  # physical_model.py:101
    p.m_total32 = float(np.float32(mass_center + 4 * mass_motor + mass_payload))  # :114
    wind = np.asarray(P.initial_wind_speed if wind_speed is None else wind_speed, dtype=np.float64).reshape(-1)
    p.wind_cw32[:] = [float(x) for x in (wind.astype(np.float32) * np.float32(coef_wind))]  # :103
    p.dt = DELTA_T
    p.init_pos[:] = _vec(P.initial_pos if initial_pos is None else initial_pos, 3)
    p.init_vel[:] = _vec(P.initial_vel if initial_vel is None else initial_vel, 3)
    p.init_angle[:] = _vec(P.initial_angle if initial_angle is None else initial_angle, 3)
    p.init_angle_vel[:] = _vec(P.initial_angle_vel if initial_angle_vel is None else initial_angle_vel, 3)
    p.init_thrust[:] = _vec(P.initial_thrust if initial_thrust is None else initial_thrust, 4)
    p.angle_limit = float(ANGLE_LIMIT)
    p.angle_vel_limit = float(ANGLE_VEL_LIMIT)
    p.z_min, p.z_max = 0.0, 80.0
    p.reward_time_passed = REWARD_TIME_PASSED
    p.reward_crashed = REWARD_CRASHED
    p.reward_unstable = REWARD_UNSTABLE
    p.reward_coin_distance = REWARD_COIN_DISTANCE_FACTOR
    # flight/__init__.py:25,52
    p.coin0_pos[:] = [5.0, 5.0, 10.0]
    p.coin_step[:] = [5.0, 2.5, 0.0]
    p.coin0_radius, p.coinN_radius = 2.0, 1.0
    p.coin0_thr2, p.coinN_thr2 = sqrt_threshold(2.0), sqrt_threshold(1.0)
    p.coin0_value, p.coinN_value = 2000, 1000
    p.sigmoid_beta = 0.4  # evolve/brain.py:37
    p.timesteps = timesteps
    p.reward_time = int(bool(reward_time))
    p.punish_time = int(bool(punish_time))
    return p


class LdPidParams(C.Structure):
    _fields_ = [
        ("n_coins", C.c_int32),
        ("timesteps", C.c_int32),
        ("step_reward", C.c_int32),
        ("_pad", C.c_int32),
        ("coin_pos", (C.c_double * 3) * PID_MAX_COINS),
        ("coin_thr2", C.c_double * PID_MAX_COINS),
        ("coin_value", C.c_int32 * PID_MAX_COINS),
        ("integral_range", C.c_double),
        ("out_lo", C.c_double * 6),
        ("out_hi", C.c_double * 6),
        ("base_thrust", C.c_double),
        ("yaw_setpoint", C.c_double),
    ]


# flight/collect_coins.py:253-258
PID_TARGET_COINS = ((0.0, 0.0, 5.0), (25.0, 25.0, 5.0), (-10.0, 35.0, 5.0), (0.0, 0.0, 5.0))


def make_pid_params(timesteps: int = 3500, coins=PID_TARGET_COINS, coin_radius=1, coin_value=1000,
                    gravity=P.gravity, mass_motor=P.mass_motor, mass_center=P.mass_center, coef_force=P.coef_force) -> LdPidParams:
    """Episode constants of the reference's flight/collect_coins.py main loop (:125-188, :253-277)."""
    coins = [tuple(float(v) for v in c) for c in coins]
    if not 1 <= len(coins) <= PID_MAX_COINS:
        raise ValueError(f"1..{PID_MAX_COINS} target coins")
    q = LdPidParams()
    q.n_coins, q.timesteps, q.step_reward = len(coins), int(timesteps), -1
    for i, c in enumerate(coins):
        q.coin_pos[i][:] = c
        q.coin_thr2[i] = sqrt_threshold(coin_radius)
        q.coin_value[i] = int(coin_value)
    q.integral_range = 2.0
    lim4, lim6 = float(np.pi / 4), float(np.pi / 6)
    #            x      y      z    roll   pitch  yaw     (collect_coins.py:125-188; z keeps PID's default [-1, 1])
    q.out_lo[:] = [-lim4, -lim4, -1.0, -lim4, -lim4, -lim6]
    q.out_hi[:] = [lim4, lim4, 1.0, lim4, lim4, lim6]
    G = gravity * (4 * mass_motor + mass_center)
    q.base_thrust = float(G / (4 * coef_force))  # onboard_computer.py:97
    q.yaw_setpoint = 0.0
    return q


def paramset_to_pid_gains(paramset18) -> np.ndarray:
    """(.., 18) gains in choose_paramset.convert_paramset_2_float order -> (.., 6, 3) per-controller
    (kp, ki, kd) for x, y, z, roll, pitch, yaw, wired like collect_coins.py:125-188 — including
    its use of ki_y for the z controller (:181)."""
    g = np.asarray(paramset18, dtype=np.float64)
    if g.shape[-1] != PID_GAINS:
        raise ValueError(f"expected {PID_GAINS} gains, got {g.shape}")
    kp_x, kp_y, kp_z, ki_x, ki_y, ki_z, kd_x, kd_y, kd_z, kp_r, kp_p, kp_w, ki_r, ki_p, ki_w, kd_r, kd_p, kd_w = np.moveaxis(g, -1, 0)
    rows = [(kp_x, ki_x, kd_x), (kp_y, ki_y, kd_y), (kp_z, ki_y, kd_z), (kp_r, ki_r, kd_r), (kp_p, ki_p, kd_p), (kp_w, ki_w, kd_w)]
    return np.ascontiguousarray(np.stack([np.stack(r, axis=-1) for r in rows], axis=-2))


class LdGaParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("generation", C.c_uint32),
        ("_pad", C.c_uint32),
        ("p_tournament", C.c_double),
        ("p_crossover", C.c_double),
        ("mutation_rate", C.c_double),
        ("mutation_sigma", C.c_double),
    ]


def make_ga_params(seed: int, generation: int, p_tournament: float = 0.6, p_crossover: float = 0.05,
                   mutation_rate: float = 0.05, mutation_sigma: float = 0.3) -> LdGaParams:
    """Hyper-parameters of evolve/evolution.py:16-20 and drone/drone.py:172."""
    g = LdGaParams()
    g.seed, g.generation = int(seed) & (2**64 - 1), int(generation)
    g.p_tournament, g.p_crossover = float(p_tournament), float(p_crossover)
    g.mutation_rate, g.mutation_sigma = float(mutation_rate), float(mutation_sigma)
    return g


MAX_PEERS = 16


class LdPeerTable(C.Structure):
    _fields_ = [
        ("shard", C.c_void_p * MAX_PEERS),
        ("shard_rows", C.c_int64),
        ("world", C.c_int32),
        ("_pad", C.c_int32),
    ]


def make_peer_table(shard_ptrs, shard_rows: int) -> LdPeerTable:
    """shard_ptrs[r]: address, valid on the calling rank's device, of rank r's genome shard."""
    if not 1 <= len(shard_ptrs) <= MAX_PEERS:
        raise ValueError(f"1..{MAX_PEERS} peers")
    t = LdPeerTable()
    for r, ptr in enumerate(shard_ptrs):
        t.shard[r] = int(ptr)
    t.shard_rows, t.world = int(shard_rows), len(shard_ptrs)
    return t
