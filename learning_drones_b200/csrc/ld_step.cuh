// ld_step.cuh — per-drone episode arithmetic for sm_100a (device only).
//
// One drone lives in one lane's registers for its whole episode.  Every function cites the
// reference code it reproduces.  Compile with -fmad=false: python-float expressions of the
// reference are evaluated unfused in source order; the products that the reference sends to
// OpenBLAS (np.dot) are explicit fma() chains in the order measured for OpenBLAS 0.3.30
// (see oracle/ld_oracle.c header and DESIGN.md "parity contract").
//
// Loop-carried split of Drone.measure(): the trigonometry of calculate_accelerations depends
// only on the angles, which are final once the integrator has run, so `trig_update` for step
// t+1 is issued right after the integrator of step t, in the same basic block as the brain of
// step t — its long dependent chains (range reduction, polynomials, reciprocal) fill the issue
// slots the MLP's fma chains leave open.  The arithmetic performed is unchanged.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ld_math.cuh"
#include "ldrones.h"

namespace ld {

constexpr int kIn = LD_N_IN, kH1 = LD_N_H1, kH2 = LD_N_H2, kOut = LD_N_OUT;
constexpr int kL1 = kH1 * (kIn + 1);  // 160
constexpr int kL2 = kH2 * (kH1 + 1);  // 110
constexpr int kL3 = kOut * (kH2 + 1); // 44
static_assert(kL1 + kL2 + kL3 == LD_GENOME_LEN, "genome layout");

// sin/cos of (roll, pitch, yaw) and the float32-rounded entries of translational_matrix
// (geometry.py:62-69) for the CURRENT angles
struct Trig {
    double sr, cr, sp, cp, sy, cy;
    float t01, t02, t11, t12, t21, t22;
};

struct Drone {
    double x[6];   // Drone.position (0..2) and Drone.angle roll,pitch,yaw (3..5): the six Sensor.previous_pos
    double v[6];   // Drone.velocity (0..2) and Drone.angle_vel (3..5): the six Sensor.previous_vel
    double acc[6]; // Sensor.current_acceleration (x y z roll pitch yaw)
    double thrust[4];
    double coin[3], thr2; // current coin centre and squared-radius threshold (ld_params.coin*_thr2)
    Trig trig;
    int reward; // running reward without the terminal distance term (fits int32 for timesteps < 2^27)
    int coins, flight_time, unstable, coin_value, status;
};

// geometry.py:26-40 (sin/cos of the three angles) and :62-69 (Tr entries, float32)
__device__ __forceinline__ void trig_finish(Trig& g, double tp, double rc) {
    g.t01 = __double2float_rn(g.sr * tp);
    g.t02 = __double2float_rn(g.cr * tp);
    g.t11 = __double2float_rn(g.cr);
    g.t12 = __double2float_rn(-g.sr);
    g.t21 = __double2float_rn(div_by(g.sr, g.cp, rc));
    g.t22 = __double2float_rn(div_by(g.cr, g.cp, rc));
}

// straight-line part: valid for |angle| <= kTrigFastMax (always, in practice)
__device__ __forceinline__ void trig_fast(Drone& d) {
    Trig& g = d.trig;
    sincos_fast(d.x[3], g.sr, g.cr);
    sincos_fast(d.x[4], g.sp, g.cp);
    sincos_fast(d.x[5], g.sy, g.cy);
    const double rc = recip(g.cp);
    trig_finish(g, div_by(g.sp, g.cp, rc), rc); // np.tan(pitch) = sin/cos to ~2 ulp
}

// |angle| beyond the fast reduction's range (never seen in practice: angles are unwrapped but
// episodes end long before): redo everything with CUDA's libm (Payne-Hanek reduction).  Kept apart
// from trig_fast so that the straight-line part can share a basic block with the MLP: the episode
// kernel folds `trig_is_big` into the vote that ends its hot loop and calls `trig_slow` per lane
// outside of it.  Must run before the next physics_step.
__device__ __forceinline__ bool trig_is_big(const Drone& d) {
    // !(|angle| <= kTrigFastMax) for any of the three angles, NaN included, decided on the high words (integer pipe):
    // kTrigFastMax = 105615.0 has a zero low word, and the fast reduction is valid well beyond it, so an angle whose
    // high word equals the bound's is still handled by the straight-line path.
    const unsigned a = (unsigned)__double2hiint(d.x[3]) & 0x7fffffffu, b = (unsigned)__double2hiint(d.x[4]) & 0x7fffffffu,
                   c = (unsigned)__double2hiint(d.x[5]) & 0x7fffffffu;
#ifdef LD_DIAG_NO_SLOW_TRIG // timing diagnostic only (wrong results beyond the bound)
    return false;
#else
    return max(a, max(b, c)) > kTrigFastMaxHi;
#endif
}

__device__ __forceinline__ void trig_slow(Drone& d) {
    Trig& g = d.trig;
    sincos(d.x[3], &g.sr, &g.cr);
    sincos(d.x[4], &g.sp, &g.cp);
    sincos(d.x[5], &g.sy, &g.cy);
    trig_finish(g, tan(d.x[4]), recip(g.cp));
}

// warp-converged form: one warp-uniform, normally-not-taken branch
__device__ __forceinline__ void trig_fallback(Drone& d) {
    const bool big = trig_is_big(d);
    if (__any_sync(0xffffffffu, big)) {
        if (big) trig_slow(d);
    }
}

__device__ __forceinline__ void trig_update(Drone& d) {
    trig_fast(d);
    trig_fallback(d);
}

__device__ __forceinline__ void drone_reset(const ld_params& p, Drone& d) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        d.x[i] = p.init_pos[i];
        d.v[i] = p.init_vel[i];
        d.x[3 + i] = p.init_angle[i];
        d.v[3 + i] = p.init_angle_vel[i];
        d.coin[i] = p.coin0_pos[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) d.acc[i] = 0.0; // sensor.py:26-27
#pragma unroll
    for (int i = 0; i < 4; ++i) d.thrust[i] = p.init_thrust[i];
    d.thr2 = p.coin0_thr2;
    d.coin_value = p.coin0_value;
    d.reward = 0;
    d.coins = d.flight_time = d.unstable = d.status = 0;
    // NOTE: d.trig is stale until trig_update(d) runs (warp-converged) — the kernel does that
}

// ||coin - pos||^2 the way np.linalg.norm forms it: x.dot(x) = fma chain in k order (coins.py:19)
__device__ __forceinline__ double coin_dist2(const Drone& d) {
    double x = d.coin[0] - d.x[0], y = d.coin[1] - d.x[1], z = d.coin[2] - d.x[2];
    return fma(z, z, fma(y, y, x * x));
}

// flight/__init__.py:41-58 — time reward, flight_time, coin pick-up on the PRE-step position.
// `live` masks the bookkeeping for lanes that are not flying a drone.
__device__ __forceinline__ void step_rewards_and_coin(const ld_params& p, Drone& d, bool live, int time_reward) {
    const bool hit = live && (coin_dist2(d) < d.thr2); // == (sqrt(dist2) < radius), see _abi.sqrt_threshold
    d.reward += live ? time_reward : 0;
    d.flight_time += live ? 1 : 0;
    d.reward += hit ? d.coin_value : 0;
    d.coins += hit ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) d.coin[i] = hit ? d.coin[i] + p.coin_step[i] : d.coin[i];
    d.thr2 = hit ? p.coinN_thr2 : d.thr2;
    d.coin_value = hit ? p.coinN_value : d.coin_value;
}

// Drone.measure (drone/drone.py:129-165) minus the trigonometry: calculate_accelerations
// (physical_model.py:93-120) + 6x Sensor.measure_acceleration/velocity_verlet
// (sensor.py:42-49,62-73,85-93) + crash rule.  `rcp_m` = 1 / double(m_total32).
// Returns true when the crash box is left; `unstable` gets the number of -1 penalties.
// `dt` = ld_params.dt for a lane that flies a drone.  The episode kernel passes 0 for lanes without one: all lanes
// execute every step, and a lane that waits for its next drone would otherwise fly a "ghost" of the reset state with
// the previous drone's brain — tumbling for thousands of steps until its angles leave the fast trigonometric range and
// drag the whole warp through the libm branch on every step (measured: 8.5 % of the warp-steps and 17 % of the kernel
// time of a GA generation).  With dt = 0 the ghost stays at the finite reset state.
template <bool EXTRAS>
__device__ __forceinline__ bool physics_step(const ld_params& p, Drone& d, double rcp_m, double* a_last_out, int& unstable, double dt) {
    const Trig& g = d.trig;
    const double sr = g.sr, cr = g.cr, sp = g.sp, cp = g.cp, sy = g.sy, cy = g.cy;
    // B = Rz.(Ry.Rx); rotation_matrix() returns B^T.  Zero/one entries folded (exact).
    const double A01 = sp * sr, A02 = sp * cr, A21 = cp * sr, A22 = cp * cr;
    const double nsr = -sr, nsy = -sy;
    const double B00 = cy * cp, B01 = fma(nsy, cr, cy * A01), B02 = fma(nsy, nsr, cy * A02);
    const double B10 = sy * cp, B11 = fma(cy, cr, sy * A01), B12 = fma(cy, nsr, sy * A02);
    const double B20 = -sp, B21 = A21, B22 = A22;

    double T[4], R[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T[i] = d.thrust[i] * p.c_f;
        R[i] = d.thrust[i] * p.c_m;
    }
    const double sumT = ((T[0] + T[1]) + T[2]) + T[3];
    // forces: float32 vector [0,0,sum T] += Rot.[0,0,-G]f32 + Rot.(wind*c_w)f32   (:98-104)
    const double gg = (double)p.neg_G32;
    double G0 = B20 * gg, G1 = B21 * gg, G2 = B22 * gg; // Rot[i][2] = B[2][i]
    if (EXTRAS) {
        const double w0 = (double)p.wind_cw32[0], w1 = (double)p.wind_cw32[1], w2 = (double)p.wind_cw32[2];
        G0 = G0 + fma(B20, w2, fma(B10, w1, B00 * w0));
        G1 = G1 + fma(B21, w2, fma(B11, w1, B01 * w0));
        G2 = G2 + fma(B22, w2, fma(B12, w1, B02 * w0));
    }
    const float F0 = __double2float_rn(G0);
    const float F1 = __double2float_rn(G1);
    const float F2 = __double2float_rn((double)__double2float_rn(sumT) + G2);
    // lin_acc = forces / m_total in float32 (:114).  float(double(F) * RN(1/m)) IS the correctly
    // rounded float32 quotient: the product is within 2^-52 of F/m while F/m (24-bit operands)
    // stays >= 2^-50 away from every float32 rounding midpoint.  Then Rot.T . lin_acc in the
    // C-contiguous dgemv order (:118).
    const double l0 = (double)__double2float_rn((double)F0 * rcp_m);
    const double l1 = (double)__double2float_rn((double)F1 * rcp_m);
    const double l2 = (double)__double2float_rn((double)F2 * rcp_m);
    double a[6];
    a[0] = fma(B02, l2, fma(B00, l0, B01 * l1));
    a[1] = fma(B12, l2, fma(B10, l0, B11 * l1));
    a[2] = fma(B22, l2, fma(B20, l0, B21 * l1));
    // moments (:107-111) and body angular acceleration (:115)
    double L = p.r_m * (T[0] + T[3] - T[1] - T[2]);
    double M = p.r_m * (T[2] + T[3] - T[0] - T[1]);
    double N = p.sqrt2_r_m * (R[1] + R[3] - R[0] - R[2]);
    if (EXTRAS) {
        L = L + p.moments_payload[0];
        M = M + p.moments_payload[1];
        N = N + p.moments_payload[2];
    }
    const double rb0 = L * p.inv_I[0], rb1 = M * p.inv_I[1], rb2 = N * p.inv_I[2];
    // Tr . rot_acc (:120), Tr entries float32 (geometry.py:62-69)
    a[3] = fma((double)g.t02, rb2, rb0 + (double)g.t01 * rb1);
    a[4] = fma((double)g.t12, rb2, (double)g.t11 * rb1);
    a[5] = fma((double)g.t22, rb2, (double)g.t21 * rb1);

    // integrator, expressions as sensor.py:63-72 (old velocity in delta_x).  0.5*a*dt*dt is
    // evaluated as (a*dt)*(0.5*dt) and 0.5*(a0+a1)*dt as (a0+a1)*(0.5*dt): scaling by 0.5 is
    // exact, so both are bit-identical to the reference's left-to-right products.
    const double hdt = 0.5 * dt;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double a_last = d.acc[i];
        const double delta_x = d.v[i] * dt + (a[i] * dt) * hdt;
        const double delta_v = (a_last + a[i]) * hdt;
        d.x[i] = d.x[i] + delta_x;
        d.v[i] = d.v[i] + delta_v;
        if (a_last_out) a_last_out[i] = a_last;
        d.acc[i] = a[i];
    }

    const double z = d.x[2];
    const bool crashed = (z <= p.z_min) || (z > p.z_max); // drone.py:150-154
    int u = 0;                                           // drone.py:159-164
    u += fabs(d.x[3]) > p.angle_limit;
    u += fabs(d.x[4]) > p.angle_limit;
    u += fabs(d.v[5]) > p.angle_vel_limit;
    unstable = u;
    return crashed;
}

// flight/__init__.py:83-86 / :97-98 — terminal distance penalty; returns the final python-int reward
__device__ __forceinline__ long long finish(const ld_params& p, Drone& d, double& dist) {
    dist = sqrt(coin_dist2(d));
    long long reward = d.reward;
    if (isfinite(dist))
        reward += (long long)p.reward_coin_distance * (long long)dist; // -100 * int(d)
    else
        d.status |= LD_STATUS_NONFINITE; // the reference raises ValueError in int(nan)
    return reward;
}

// evolve/brain.py:41-54 — one dense layer: y_j = sum_k W[j][k] x[k] + W[j][n_in] * 1, k-ordered fma chain.
// `W(i)` returns flattened weight i of this layer (row-major, bias last).
template <int N_OUT_, int N_IN_, typename WFn>
__device__ __forceinline__ void dense(WFn W, const double* x, double* y) {
#pragma unroll
    for (int j = 0; j < N_OUT_; ++j) y[j] = W(j * (N_IN_ + 1)) * x[0];
#pragma unroll
    for (int k = 1; k < N_IN_; ++k) {
#pragma unroll
        for (int j = 0; j < N_OUT_; ++j) y[j] = fma(W(j * (N_IN_ + 1) + k), x[k], y[j]);
    }
#pragma unroll
    for (int j = 0; j < N_OUT_; ++j) y[j] = y[j] + W(j * (N_IN_ + 1) + N_IN_); // fma(b, 1.0, acc)
}

// brain.py:26-30 : `x < 0 -> 0.0 else 1.0 * x`, decided on the sign bit (integer pipe).  Differs
// from the fp compare only for -0.0 (-> +0.0, indistinguishable downstream) and negative NaN.
__device__ __forceinline__ double relu(double s) {
    const int hi = __double2hiint(s);
    return hi < 0 ? 0.0 : s;
}

} // namespace ld
