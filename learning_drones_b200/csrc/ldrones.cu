// ldrones.cu — libldrones.so: sm_100a kernels + the C-ABI of include/ldrones.h.
//
// Episode kernel layout:
//   * persistent grid, one CTA of 3 warps per SM; one drone per lane, its whole episode in
//     registers; lanes pull the next drone from a global queue head when theirs ends.
//   * each lane's brain: layers 1+2 (270 doubles) in its own shared-memory row of 2160 B
//     (135 x 16 B: an odd number of 16-B units, so a warp's LDS.128 at a common row offset hits
//     every bank group exactly once per quarter-warp), layer 3 (44 doubles) in registers.
//     96 lanes x 2160 B = 202.5 KiB of the SM's 227 KiB.
//   * HBM traffic per drone-episode: its 2512-B genome row in, a 32-B record (+224-B state) out.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ld_step.cuh"

namespace {

using namespace ld;

constexpr int kWarpsPerCta = 3;
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr int kL3Smem = 32;                           // leading part of layer 3 that also fits in shared memory
constexpr int kL3Regs = kL3 - kL3Smem;                // 12 trailing layer-3 weights stay in registers
constexpr int kSmemDoubles = kL1 + kL2 + kL3Smem;     // 302 per lane
constexpr int kRowBytes = kSmemDoubles * 8;           // 2416
constexpr int kSmemBytes = kCtaThreads * kRowBytes;   // 231936 of the SM's 232448
static_assert((kRowBytes / 16) % 2 == 1, "row stride must be an odd number of 16-B units (bank-conflict-free LDS.128)");
static_assert(kL1 % 2 == 0 && kSmemDoubles % 2 == 0, "double2 staging");

thread_local char g_err[512] = "";

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess)
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    else
        snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

#define LD_CUDA(call)                                                   \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) return fail(LD_ERR_CUDA, #call, e_);     \
    } while (0)

__device__ __forceinline__ void write_state(const Drone& d, const double* a_last, double* row) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        row[i] = d.x[i];
        row[3 + i] = d.v[i];
        row[6 + i] = d.x[3 + i];
        row[9 + i] = d.v[3 + i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        row[12 + i] = d.acc[i];
        row[18 + i] = a_last[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) row[24 + i] = d.thrust[i];
}

// evolve/brain.py:56-76 on this lane's staged weights
__device__ __forceinline__ void brain(Drone& d, const double2* __restrict__ wrow, const double (&w3)[kL3Regs], double neg_beta,
                                      bool keep_old) {
    double in[kIn];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        in[i] = d.x[i];
        in[3 + i] = d.v[i];
        in[6 + i] = d.x[3 + i];
        in[9 + i] = d.v[3 + i];
        in[12 + i] = d.coin[i];
    }
    double h1[kH1], h2[kH2], out[kOut];
    // weights are fetched as 16-B pairs (LDS.128); after unrolling every index is a constant
    auto smem_w = [&](int i) {
        const double2 w = wrow[i >> 1];
        return (i & 1) ? w.y : w.x;
    };
    dense<kH1, kIn>([&](int i) { return smem_w(i); }, in, h1);
#pragma unroll
    for (int j = 0; j < kH1; ++j) h1[j] = relu(h1[j]);
    dense<kH2, kH1>([&](int i) { return smem_w(kL1 + i); }, h1, h2);
#pragma unroll
    for (int j = 0; j < kH2; ++j) h2[j] = relu(h2[j]);
    dense<kOut, kH2>([&](int i) { return i < kL3Smem ? smem_w(kL1 + kL2 + (i < kL3Smem ? i : 0)) : w3[i >= kL3Smem ? i - kL3Smem : 0]; }, h2, out);
#pragma unroll
    for (int j = 0; j < kOut; ++j) {
        const double th = sigmoid_fast(out[j], neg_beta);
        // on the crash step the reference never reaches the brain (flight/__init__.py:80-87)
        d.thrust[j] = keep_old ? d.thrust[j] : th;
    }
}

enum Mode { kPlain = 0, kState = 1, kTrace = 2 };

// flight/__init__.py:83-87,97-99: terminal distance penalty, then publish the drone's outcome
template <int MODE>
__device__ __forceinline__ void publish(const ld_params& p, Drone& d, const double* a_last, long long idx,
                                        ld_record* __restrict__ records, double* __restrict__ final_state) {
    double dist;
    const long long reward = finish(p, d, dist);
    // 32-B record = two 16-B stores
    int4 lo, hi;
    lo.x = (int)(unsigned)(reward & 0xffffffffLL);
    lo.y = (int)(reward >> 32);
    lo.z = d.coins;
    lo.w = d.flight_time;
    hi.x = d.status;
    hi.y = d.unstable;
    hi.z = __double2loint(dist);
    hi.w = __double2hiint(dist);
    int4* out = reinterpret_cast<int4*>(records + idx);
    out[0] = lo;
    out[1] = hi;
    if (MODE != kPlain) {
        if (final_state) write_state(d, a_last, final_state + idx * LD_STATE_LEN);
    }
}


// One population evaluation = the loop of evolve/evolution.py:64-73 around flight/__init__.py:fly().
// MODE kState: also write the final kinematic state; kTrace: n == 1, dump the state after every
// measure() (debug tap).  All 32 lanes execute every step (idle lanes run masked) so that the
// step body is one straight-line block.
template <bool EXTRAS, int MODE>
__global__ void __launch_bounds__(kCtaThreads, 1)
fly_mlp_f64_kernel(const __grid_constant__ ld_params p, const double* __restrict__ genomes, long long n,
                   ld_record* __restrict__ records, double* __restrict__ final_state,
                   unsigned long long* __restrict__ queue_head, double* __restrict__ trace) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* wrow = reinterpret_cast<double2*>(smem_raw + (size_t)threadIdx.x * kRowBytes);
    const unsigned lane = threadIdx.x & 31u;
    const double neg_beta = -p.sigmoid_beta;
    const double rcp_m = 1.0 / (double)p.m_total32;
    const int time_reward = (p.reward_time ? p.reward_time_passed : 0) - (p.punish_time ? p.reward_time_passed : 0);
    const int H = p.timesteps;

    Drone d;
    drone_reset(p, d);
    double w3[kL3Regs];
#pragma unroll
    for (int i = 0; i < kL3Regs; ++i) w3[i] = 0.0;
    long long idx = -1; // drone this lane is flying; -1 = idle
    int t = 0;
    bool more = true; // queue not yet exhausted (warp-uniform)

    for (;;) {
        // ---- refill idle lanes from the queue (warp-aggregated fetch) -------------------------
        const unsigned idle = __ballot_sync(0xffffffffu, idx < 0);
        if (idle && more) {
            unsigned long long base = 0;
            const int leader = __ffs(idle) - 1;
            if ((int)lane == leader) base = atomicAdd(queue_head, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (idx < 0) {
                const long long mine = (long long)base + __popc(idle & ((1u << lane) - 1u));
                if (mine < n) {
                    idx = mine;
                    const double2* __restrict__ src = reinterpret_cast<const double2*>(genomes + idx * LD_GENOME_LEN);
#pragma unroll 9
                    for (int i = 0; i < kSmemDoubles / 2; ++i) wrow[i] = __ldg(src + i);
#pragma unroll
                    for (int i = 0; i < kL3Regs / 2; ++i) {
                        const double2 w = __ldg(src + kSmemDoubles / 2 + i);
                        w3[2 * i] = w.x;
                        w3[2 * i + 1] = w.y;
                    }
                    drone_reset(p, d);
                    t = 0;
                }
            }
            more = (long long)base + __popc(idle) < n;
            trig_update(d); // warp-converged; recomputing it for lanes in flight is idempotent
        }
        if (__ballot_sync(0xffffffffu, idx >= 0) == 0u) break;

        if (H <= 0) { // fly(timesteps=0): the loop body never runs (warp-uniform, cold)
            if (idx >= 0) {
                const double zeros[6] = {0, 0, 0, 0, 0, 0};
                publish<MODE>(p, d, zeros, idx, records, final_state);
                idx = -1;
            }
            continue;
        }
        // ---- one timestep of flight/__init__.py:39-93 ------------------------------------------
        const bool flying = idx >= 0;
        const bool live = flying;
        double a_last[6];
        int u;
        step_rewards_and_coin(p, d, live, time_reward);
        const bool crashed = physics_step<EXTRAS>(p, d, rcp_m, MODE != kPlain ? a_last : nullptr, u) && live;
        d.reward += crashed ? p.reward_crashed : ((live ? u : 0) * p.reward_unstable);
        d.unstable += (live && !crashed) ? u : 0;
        d.status |= crashed ? LD_STATUS_CRASHED : 0;
        if (MODE == kTrace) {
            if (live) write_state(d, a_last, trace + (size_t)t * LD_STATE_LEN);
        }
        // next step's trigonometry and this step's brain: independent, interleaved by the scheduler
        trig_update(d);
        brain(d, wrow, w3, neg_beta, MODE != kPlain && crashed);
        if (MODE == kTrace) {
            if (live && !crashed) {
#pragma unroll
                for (int i = 0; i < 4; ++i) trace[(size_t)t * LD_STATE_LEN + 24 + i] = d.thrust[i];
            }
        }
        t += live ? 1 : 0;
        if (flying && (crashed || t >= H)) {
            publish<MODE>(p, d, a_last, idx, records, final_state);
            idx = -1;
        }
    }
}

// ---- population statistics: warp-shuffle reductions -----------------------------------------
__device__ __forceinline__ unsigned long long pack_key(long long reward, long long i) {
    long long r = reward < -2147483648LL ? -2147483648LL : (reward > 2147483647LL ? 2147483647LL : reward);
    return ((unsigned long long)(r + 2147483648LL) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
}

__global__ void __launch_bounds__(256) reduce_records_kernel(const ld_record* __restrict__ rec, long long n,
                                                            unsigned long long* __restrict__ stats) {
    long long steps = 0, coins = 0, crashed = 0;
    unsigned long long best = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(rec + i));     // reward (8 B), coins, flight_time
        const int2 b = __ldg(reinterpret_cast<const int2*>(rec + i) + 2); // status, unstable
        const long long reward = ((long long)(unsigned)a.y << 32) | (unsigned)a.x;
        steps += a.w;
        coins += a.z;
        crashed += b.x & LD_STATUS_CRASHED;
        const unsigned long long k = pack_key(reward, i);
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        steps += __shfl_xor_sync(0xffffffffu, steps, o);
        coins += __shfl_xor_sync(0xffffffffu, coins, o);
        crashed += __shfl_xor_sync(0xffffffffu, crashed, o);
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + 0, (unsigned long long)steps);
        atomicMax(stats + 1, best);
        atomicAdd(stats + 3, (unsigned long long)coins);
        atomicAdd(stats + 4, (unsigned long long)crashed);
    }
}

__global__ void finalize_stats_kernel(unsigned long long* stats, long long n) {
    const unsigned long long k = stats[1];
    if (n > 0) {
        stats[1] = (unsigned long long)((long long)(k >> 32) - 2147483648LL);
        stats[2] = 0xffffffffull - (k & 0xffffffffull);
    } else {
        stats[1] = 0;
        stats[2] = (unsigned long long)-1LL;
    }
}

// ---- fp64 pipe peak: independent DFMA chains, registers only (roofline denominator) ----------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456) out[0] = s; // keep the chains live
}

// ---- host side --------------------------------------------------------------------------------
struct DeviceInfo {
    int device = -1;
    int sms = 0;
    bool attrs_set = false;
};
thread_local DeviceInfo g_dev;

int ensure_device() {
    int dev = 0;
    LD_CUDA(cudaGetDevice(&dev));
    if (g_dev.device == dev && g_dev.attrs_set) return LD_OK;
    cudaDeviceProp prop;
    LD_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(LD_ERR_NO_DEVICE, "libldrones is built for sm_100a only (compute capability 10.x required)");
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<false, kPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<true, kPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<false, kState>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<true, kState>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<false, kTrace>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<true, kTrace>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    g_dev.device = dev;
    g_dev.sms = prop.multiProcessorCount;
    g_dev.attrs_set = true;
    return LD_OK;
}

bool has_extras(const ld_params* p) {
    for (int i = 0; i < 3; ++i)
        if (p->wind_cw32[i] != 0.0f || p->moments_payload[i] != 0.0) return true;
    return false;
}

int check_params(const ld_params* p) {
    if (!p) return fail(LD_ERR_INVALID_ARG, "params is NULL");
    if (p->timesteps < 0) return fail(LD_ERR_INVALID_ARG, "params.timesteps < 0");
    if (!(p->dt > 0.0)) return fail(LD_ERR_INVALID_ARG, "params.dt must be > 0");
    if (!(p->m_total32 > 0.0f)) return fail(LD_ERR_INVALID_ARG, "params.m_total32 must be > 0");
    return LD_OK;
}

int launch_fly(const ld_params* p, const double* genomes, int64_t n, ld_record* records, double* final_state,
               int64_t* work_counter, double* trace, cudaStream_t s) {
    int rc = ensure_device();
    if (rc) return rc;
    if (n == 0) return LD_OK;
    LD_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(int64_t), s));
    long long ctas = (n + kCtaThreads - 1) / kCtaThreads;
    if (ctas > g_dev.sms) ctas = g_dev.sms;
    dim3 grid((unsigned)ctas), block(kCtaThreads);
    auto* head = reinterpret_cast<unsigned long long*>(work_counter);
    const bool ex = has_extras(p);
#define LD_LAUNCH(EX, MD) \
    fly_mlp_f64_kernel<EX, MD><<<grid, block, kSmemBytes, s>>>(*p, genomes, n, records, final_state, head, trace)
    const int mode = trace ? kTrace : (final_state ? kState : kPlain);
    if (ex) {
        if (mode == kTrace) LD_LAUNCH(true, kTrace);
        else if (mode == kState) LD_LAUNCH(true, kState);
        else LD_LAUNCH(true, kPlain);
    } else {
        if (mode == kTrace) LD_LAUNCH(false, kTrace);
        else if (mode == kState) LD_LAUNCH(false, kState);
        else LD_LAUNCH(false, kPlain);
    }
#undef LD_LAUNCH
    LD_CUDA(cudaGetLastError());
    return LD_OK;
}

// grow-only device workspace for the *_host entry point (per host thread)
struct Workspace {
    int device = -1;
    void* genomes[2] = {nullptr, nullptr};
    void* records[2] = {nullptr, nullptr};
    void* state[2] = {nullptr, nullptr};
    int64_t* counters = nullptr;
    size_t cap_drones = 0;
    bool cap_state = false;
    cudaStream_t streams[2] = {nullptr, nullptr};
};
thread_local Workspace g_ws;

void ws_release() {
    for (int i = 0; i < 2; ++i) {
        if (g_ws.genomes[i]) cudaFree(g_ws.genomes[i]);
        if (g_ws.records[i]) cudaFree(g_ws.records[i]);
        if (g_ws.state[i]) cudaFree(g_ws.state[i]);
        if (g_ws.streams[i]) cudaStreamDestroy(g_ws.streams[i]);
        g_ws.genomes[i] = g_ws.records[i] = g_ws.state[i] = nullptr;
        g_ws.streams[i] = nullptr;
    }
    if (g_ws.counters) cudaFree(g_ws.counters);
    g_ws.counters = nullptr;
    g_ws.cap_drones = 0;
    g_ws.cap_state = false;
    g_ws.device = -1;
}

int ws_reserve(size_t chunk, bool want_state) {
    int dev = 0;
    LD_CUDA(cudaGetDevice(&dev));
    if (g_ws.device == dev && g_ws.cap_drones >= chunk && (g_ws.cap_state || !want_state)) return LD_OK;
    ws_release();
    for (int i = 0; i < 2; ++i) {
        LD_CUDA(cudaMalloc(&g_ws.genomes[i], chunk * LD_GENOME_LEN * sizeof(double)));
        LD_CUDA(cudaMalloc(&g_ws.records[i], chunk * sizeof(ld_record)));
        if (want_state) LD_CUDA(cudaMalloc(&g_ws.state[i], chunk * LD_STATE_LEN * sizeof(double)));
        LD_CUDA(cudaStreamCreateWithFlags(&g_ws.streams[i], cudaStreamNonBlocking));
    }
    LD_CUDA(cudaMalloc((void**)&g_ws.counters, 2 * sizeof(int64_t)));
    g_ws.device = dev;
    g_ws.cap_drones = chunk;
    g_ws.cap_state = want_state;
    return LD_OK;
}

} // namespace

extern "C" {

int ld_abi_version(void) { return LD_ABI_VERSION; }

const char* ld_last_error(void) { return g_err; }

int ld_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LD_ERR_NO_DEVICE, "cudaGetDeviceCount", e);
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int64_t ld_wave_size(void) {
    if (ensure_device() != LD_OK) return -1;
    return (int64_t)g_dev.sms * kCtaThreads;
}

int ld_fly_mlp_f64_device(const ld_params* params, const double* genomes, int64_t n, ld_record* records,
                          double* final_state, int64_t* work_counter, void* stream) {
    int rc = check_params(params);
    if (rc) return rc;
    if (n < 0) return fail(LD_ERR_INVALID_ARG, "n < 0");
    if (n > 0 && (!genomes || !records || !work_counter)) return fail(LD_ERR_INVALID_ARG, "NULL device buffer");
    if ((reinterpret_cast<uintptr_t>(genomes) & 15u) || (reinterpret_cast<uintptr_t>(records) & 15u))
        return fail(LD_ERR_INVALID_ARG, "genomes and records must be 16-byte aligned");
    return launch_fly(params, genomes, n, records, final_state, work_counter, nullptr, (cudaStream_t)stream);
}

int ld_fly_mlp_f64_trace_device(const ld_params* params, const double* genome, ld_record* record, double* trace,
                                void* stream) {
    int rc = check_params(params);
    if (rc) return rc;
    if (!genome || !record || !trace) return fail(LD_ERR_INVALID_ARG, "NULL device buffer");
    rc = ensure_device();
    if (rc) return rc;
    int64_t* counter = nullptr;
    LD_CUDA(cudaMalloc((void**)&counter, sizeof(int64_t)));
    rc = launch_fly(params, genome, 1, record, nullptr, counter, trace, (cudaStream_t)stream);
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(counter);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(LD_ERR_CUDA, "trace kernel", e);
    return LD_OK;
}

int ld_fly_mlp_f64_host(const ld_params* params, const double* genomes, int64_t n, ld_record* records,
                        double* final_state) {
    int rc = check_params(params);
    if (rc) return rc;
    if (n < 0) return fail(LD_ERR_INVALID_ARG, "n < 0");
    if (n == 0) return LD_OK;
    if (!genomes || !records) return fail(LD_ERR_INVALID_ARG, "NULL host buffer");
    rc = ensure_device();
    if (rc) return rc;
    // chunk = whole waves so every chunk fills the machine; two buffers so the H2D copy of chunk
    // i+1 and the D2H copy of chunk i-1 overlap the kernel of chunk i
    const int64_t wave = (int64_t)g_dev.sms * kCtaThreads;
    int64_t chunk = n;
    if (n > 2 * wave) {
        int64_t waves = (n + wave - 1) / wave;
        int64_t per = waves >= 8 ? (waves + 3) / 4 : (waves + 1) / 2;
        if (per > 8) per = 8;
        chunk = per * wave;
    }
    rc = ws_reserve((size_t)chunk, final_state != nullptr);
    if (rc) return rc;
    int b = 0;
    for (int64_t off = 0; off < n; off += chunk, b ^= 1) {
        const int64_t m = (n - off < chunk) ? (n - off) : chunk;
        cudaStream_t s = g_ws.streams[b];
        LD_CUDA(cudaMemcpyAsync(g_ws.genomes[b], genomes + off * LD_GENOME_LEN, (size_t)m * LD_GENOME_LEN * sizeof(double),
                                cudaMemcpyHostToDevice, s));
        rc = launch_fly(params, (const double*)g_ws.genomes[b], m, (ld_record*)g_ws.records[b],
                        final_state ? (double*)g_ws.state[b] : nullptr, g_ws.counters + b, nullptr, s);
        if (rc) return rc;
        LD_CUDA(cudaMemcpyAsync(records + off, g_ws.records[b], (size_t)m * sizeof(ld_record), cudaMemcpyDeviceToHost, s));
        if (final_state)
            LD_CUDA(cudaMemcpyAsync(final_state + off * LD_STATE_LEN, g_ws.state[b], (size_t)m * LD_STATE_LEN * sizeof(double),
                                    cudaMemcpyDeviceToHost, s));
    }
    LD_CUDA(cudaStreamSynchronize(g_ws.streams[0]));
    LD_CUDA(cudaStreamSynchronize(g_ws.streams[1]));
    return LD_OK;
}

int ld_reduce_records_device(const ld_record* records, int64_t n, int64_t* stats5, void* stream) {
    if (n < 0 || !stats5 || (n > 0 && !records)) return fail(LD_ERR_INVALID_ARG, "bad arguments");
    if (n > 0xffffffffLL) return fail(LD_ERR_INVALID_ARG, "n must fit in 32 bits");
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    LD_CUDA(cudaMemsetAsync(stats5, 0, 5 * sizeof(int64_t), s));
    auto* st = reinterpret_cast<unsigned long long*>(stats5);
    if (n > 0) {
        long long blocks = (n + 255) / 256;
        if (blocks > 8LL * g_dev.sms) blocks = 8LL * g_dev.sms;
        reduce_records_kernel<<<(unsigned)blocks, 256, 0, s>>>(records, n, st);
    }
    finalize_stats_kernel<<<1, 1, 0, s>>>(st, n);
    LD_CUDA(cudaGetLastError());
    return LD_OK;
}

void ld_release_workspace(void) { ws_release(); }

int ld_measure_fp64_peak(double* tflops_out, double* dfma_per_clk_per_sm_out) {
    if (!tflops_out) return fail(LD_ERR_INVALID_ARG, "NULL output");
    int rc = ensure_device();
    if (rc) return rc;
    double* scratch = nullptr;
    LD_CUDA(cudaMalloc((void**)&scratch, sizeof(double)));
    cudaEvent_t e0, e1;
    LD_CUDA(cudaEventCreate(&e0));
    LD_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = g_dev.sms * 8, threads = 256;
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        LD_CUDA(cudaEventRecord(e0, 0));
        fp64_peak_kernel<<<blocks, threads>>>(scratch, iters, 1.0 + rep);
        LD_CUDA(cudaEventRecord(e1, 0));
        LD_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        LD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(scratch);
    const double fmas = (double)blocks * threads * iters * 16.0 * 8.0;
    *tflops_out = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
    if (dfma_per_clk_per_sm_out) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_dev.device);
        *dfma_per_clk_per_sm_out = fmas / (best_ms * 1e-3) / ((double)khz * 1e3) / g_dev.sms;
    }
    return LD_OK;
}

} // extern "C"
