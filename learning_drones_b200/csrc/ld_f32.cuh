// ld_f32.cuh — fp32 "fast mode" of the MLP-controlled episode (SURVEY §7 step 5, §8b ld_fly_mlp_f32).
//
// NOT a parity mode: the dynamics are chaotic and float32 changes the integer reward of ~40 % of
// random drones; this kernel is judged statistically against the fp64 kernel
// (tests/test_gpu_f32.py) and is never used for the headline numbers.  Same episode semantics
// (flight/__init__.py:12-99, drone/drone.py:129-165, physical_model.py:75-122, brain.py:41-76),
// all state and weights in float32, CUDA float libm, Kahan-compensated integrators so that the
// 2,000 small position increments do not drown in float32 rounding.
//
// Layout: a brain is 314 floats = 1256 B, so half a Tensor-Memory lane (256 columns) + 58 floats of
// shared memory hold it, and the SM can keep 8 warps (2 per sub-partition, 256 drones) resident:
// while one warp sits in a dependent chain the other issues — the latency hiding the fp64 kernel
// cannot have because its brains fill the whole TMEM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ldrones.h"

namespace ldf {

constexpr int kIn = LD_N_IN, kH1 = LD_N_H1, kH2 = LD_N_H2, kOut = LD_N_OUT;
constexpr int kL1 = kH1 * (kIn + 1), kL2 = kH2 * (kH1 + 1), kL3 = kOut * (kH2 + 1);
constexpr int kWarps = 8;                 // warps w and w+4 share TMEM lane quadrant w%4, 256 columns each
constexpr int kThreads = kWarps * 32;     // 256 drones resident per SM
constexpr int kTmemFloats = 256;          // stored 0..255 (k-major order, see ldrones.cu stored_to_genome)
constexpr int kChunk = 32;                // floats per tcgen05.ld x32
constexpr int kChunks = kTmemFloats / kChunk; // 8
constexpr int kSmemFloats = LD_GENOME_LEN - kTmemFloats; // 58
constexpr int kRowBytes = kSmemFloats * 4; // 232 B: 29 x 8 B, odd => conflict-free LDS.64
constexpr int kSmemUsed = kThreads * kRowBytes;
static_assert(kSmemFloats % 2 == 0 && (kRowBytes / 8) % 2 == 1, "shared-memory row layout");

struct DroneF {
    float x[6], v[6];   // position+angles, velocities+rates
    float cx[6], cv[6]; // Kahan compensation terms of the twelve integrators
    float acc[6];
    float thrust[4];
    float coin[3], r2;  // current coin and its squared radius
    float sr, cr, sp, cp, sy, cy, tp; // trigonometry of the current angles
    int reward, coins, flight_time, unstable, coin_value, status;
};

struct ParamsF { // ld_params narrowed once per kernel
    float c_f, c_m, r_m, sqrt2_r_m, inv_I[3], mp[3], neg_G, inv_m, wind[3], dt, hdt;
    float angle_limit, angle_vel_limit, z_min, z_max, coin_step[3], r2_0, r2_n, neg_beta;
};

__device__ __forceinline__ ParamsF narrow(const ld_params& p) {
    ParamsF f;
    f.c_f = (float)p.c_f; f.c_m = (float)p.c_m; f.r_m = (float)p.r_m; f.sqrt2_r_m = (float)p.sqrt2_r_m;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        f.inv_I[i] = (float)p.inv_I[i];
        f.mp[i] = (float)p.moments_payload[i];
        f.wind[i] = p.wind_cw32[i];
        f.coin_step[i] = (float)p.coin_step[i];
    }
    f.neg_G = p.neg_G32; f.inv_m = 1.0f / p.m_total32; f.dt = (float)p.dt; f.hdt = 0.5f * (float)p.dt;
    f.angle_limit = (float)p.angle_limit; f.angle_vel_limit = (float)p.angle_vel_limit;
    f.z_min = (float)p.z_min; f.z_max = (float)p.z_max;
    f.r2_0 = (float)(p.coin0_radius * p.coin0_radius); f.r2_n = (float)(p.coinN_radius * p.coinN_radius);
    f.neg_beta = -(float)p.sigmoid_beta;
    return f;
}

__device__ __forceinline__ void trig(DroneF& d) {
    sincosf(d.x[3], &d.sr, &d.cr);
    sincosf(d.x[4], &d.sp, &d.cp);
    sincosf(d.x[5], &d.sy, &d.cy);
    d.tp = __fdividef(d.sp, d.cp);
}

__device__ __forceinline__ void reset(const ld_params& p, const ParamsF& f, DroneF& d) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        d.x[i] = (float)p.init_pos[i];
        d.v[i] = (float)p.init_vel[i];
        d.x[3 + i] = (float)p.init_angle[i];
        d.v[3 + i] = (float)p.init_angle_vel[i];
        d.coin[i] = (float)p.coin0_pos[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) d.acc[i] = d.cx[i] = d.cv[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) d.thrust[i] = (float)p.init_thrust[i];
    d.r2 = f.r2_0;
    d.coin_value = p.coin0_value;
    d.reward = d.coins = d.flight_time = d.unstable = d.status = 0;
}

__device__ __forceinline__ float dist2(const DroneF& d) {
    const float x = d.coin[0] - d.x[0], y = d.coin[1] - d.x[1], z = d.coin[2] - d.x[2];
    return fmaf(z, z, fmaf(y, y, x * x));
}

// flight/__init__.py:41-58
__device__ __forceinline__ void rewards_and_coin(const ld_params& p, const ParamsF& f, DroneF& d, bool live, int time_reward) {
    const bool hit = live && dist2(d) < d.r2;
    d.reward += (live ? time_reward : 0) + (hit ? d.coin_value : 0);
    d.flight_time += live ? 1 : 0;
    d.coins += hit ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) d.coin[i] = hit ? d.coin[i] + f.coin_step[i] : d.coin[i];
    d.r2 = hit ? f.r2_n : d.r2;
    d.coin_value = hit ? p.coinN_value : d.coin_value;
}

// compensated x += dx
__device__ __forceinline__ void kahan_add(float& x, float& c, float dx) {
    const float y = dx - c, t = x + y;
    c = (t - x) - y;
    x = t;
}

// physical_model.py:75-122 + sensor.py:62-93 + drone.py:150-164; returns crashed
template <bool EXTRAS>
__device__ __forceinline__ bool physics(const ParamsF& f, DroneF& d, int& unstable) {
    const float sr = d.sr, cr = d.cr, sp = d.sp, cp = d.cp, sy = d.sy, cy = d.cy;
    const float A01 = sp * sr, A02 = sp * cr, A21 = cp * sr, A22 = cp * cr;
    const float B00 = cy * cp, B01 = fmaf(-sy, cr, cy * A01), B02 = fmaf(sy, sr, cy * A02);
    const float B10 = sy * cp, B11 = fmaf(cy, cr, sy * A01), B12 = fmaf(-cy, sr, sy * A02);
    const float B20 = -sp, B21 = A21, B22 = A22;
    float T[4], R[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T[i] = d.thrust[i] * f.c_f;
        R[i] = d.thrust[i] * f.c_m;
    }
    float F0 = B20 * f.neg_G, F1 = B21 * f.neg_G, F2 = ((T[0] + T[1]) + T[2]) + T[3] + B22 * f.neg_G;
    if (EXTRAS) {
        F0 += fmaf(B20, f.wind[2], fmaf(B10, f.wind[1], B00 * f.wind[0]));
        F1 += fmaf(B21, f.wind[2], fmaf(B11, f.wind[1], B01 * f.wind[0]));
        F2 += fmaf(B22, f.wind[2], fmaf(B12, f.wind[1], B02 * f.wind[0]));
    }
    const float l0 = F0 * f.inv_m, l1 = F1 * f.inv_m, l2 = F2 * f.inv_m;
    float a[6];
    a[0] = fmaf(B02, l2, fmaf(B00, l0, B01 * l1));
    a[1] = fmaf(B12, l2, fmaf(B10, l0, B11 * l1));
    a[2] = fmaf(B22, l2, fmaf(B20, l0, B21 * l1));
    float L = f.r_m * (T[0] + T[3] - T[1] - T[2]), M = f.r_m * (T[2] + T[3] - T[0] - T[1]);
    float N = f.sqrt2_r_m * (R[1] + R[3] - R[0] - R[2]);
    if (EXTRAS) { L += f.mp[0]; M += f.mp[1]; N += f.mp[2]; }
    const float rb0 = L * f.inv_I[0], rb1 = M * f.inv_I[1], rb2 = N * f.inv_I[2];
    const float icp = __fdividef(1.0f, cp);
    a[3] = fmaf(cr * d.tp, rb2, fmaf(sr * d.tp, rb1, rb0));
    a[4] = fmaf(-sr, rb2, cr * rb1);
    a[5] = fmaf(cr * icp, rb2, (sr * icp) * rb1);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        kahan_add(d.x[i], d.cx[i], fmaf(d.v[i], f.dt, (a[i] * f.dt) * f.hdt));
        kahan_add(d.v[i], d.cv[i], (d.acc[i] + a[i]) * f.hdt);
        d.acc[i] = a[i];
    }
    const float z = d.x[2];
    int u = 0;
    u += fabsf(d.x[3]) > f.angle_limit;
    u += fabsf(d.x[4]) > f.angle_limit;
    u += fabsf(d.v[5]) > f.angle_vel_limit;
    unstable = u;
    return (z <= f.z_min) || (z > f.z_max);
}

struct MlpF {
    float in[kIn], h1[kH1], h2[kH2], out[kOut];
};

// stored index S (k-major per layer, same permutation as the fp64 kernel)
template <int S>
__device__ __forceinline__ void mac(MlpF& m, float w) {
    if (S < kL1) {
        constexpr int k = S / kH1, j = S % kH1;
        if (k == 0) m.h1[j] = w * m.in[0];
        else if (k < kIn) m.h1[j] = fmaf(w, m.in[k < kIn ? k : 0], m.h1[j]);
        else m.h1[j] = fmaxf(m.h1[j] + w, 0.0f);
    } else if (S < kL1 + kL2) {
        constexpr int t = S - kL1 >= 0 ? S - kL1 : 0, k = t / kH2, j = t % kH2;
        if (k == 0) m.h2[j] = w * m.h1[0];
        else if (k < kH1) m.h2[j] = fmaf(w, m.h1[k < kH1 ? k : 0], m.h2[j]);
        else m.h2[j] = fmaxf(m.h2[j] + w, 0.0f);
    } else {
        constexpr int t = S - kL1 - kL2 >= 0 ? S - kL1 - kL2 : 0, k = t / kOut, j = t % kOut;
        if (k == 0) m.out[j] = w * m.h2[0];
        else if (k < kH2) m.out[j] = fmaf(w, m.h2[k < kH2 ? k : 0], m.out[j]);
        else m.out[j] = m.out[j] + w;
    }
}

template <int C, int E = 0>
__device__ __forceinline__ void chunk(MlpF& m, const uint32_t (&v)[32]) {
    if constexpr (E < kChunk) {
        mac<C * kChunk + E>(m, __uint_as_float(v[E]));
        chunk<C, E + 1>(m, v);
    }
}

template <int P = 0>
__device__ __forceinline__ void smem_tail(MlpF& m, const float2* __restrict__ wrow) {
    if constexpr (P < kSmemFloats / 2) {
        const float2 w = wrow[P];
        mac<kTmemFloats + 2 * P>(m, w.x);
        mac<kTmemFloats + 2 * P + 1>(m, w.y);
        smem_tail<P + 1>(m, wrow);
    }
}

} // namespace ldf
