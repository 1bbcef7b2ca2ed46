// ldrones.cu — libldrones.so: sm_100a kernels + the C-ABI of include/ldrones.h.
//
// Episode kernel layout:
//   * persistent grid, one CTA of 4 warps per SM (one warp per SM sub-partition); one drone per
//     lane, its whole episode in registers; lanes pull the next drone from a global queue head.
//   * each lane's brain (314 fp64 weights) must be re-read every timestep and nothing else is
//     (measured: ~40 B/clk/warp from shared memory, ~54 B/clk/warp from Tensor Memory).  256 weights
//     per lane therefore sit in Tensor Memory (all 512 columns of the lane) and are streamed with
//     tcgen05.ld (x16 loads on alternating scoreboards, no wait::ld between them: mlp_tmem_stream), 58 in
//     the lane's shared-memory row; what binds is the fp64 pipe + dependency latency of the one warp a
//     sub-partition has room for.
//   * long episodes are flown in segments: at a segment boundary a drone whose lane is wanted by
//     a waiting drone is parked in HBM and re-queued (round robin), so a population of survivors
//     ends together instead of in whole-episode waves (struct Sched).
//   * HBM traffic per drone-episode: its 2512-B genome row in, a 32-B record (+224-B state) out.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ld_f32.cuh"
#include "ld_step.cuh"

namespace {

using namespace ld;

// ---- where one lane's 314 brain weights live ----------------------------------------------------
// Stored order is k-major per layer (all neurons' weight for input k next to each other), so any
// run of consecutive stored weights feeds up to 10 independent fma chains:
//   s in [0,160)   : layer 1, s = k*10 + j      <-> genome[j*16 + k]
//   s in [160,270) : layer 2, s-160 = k*10 + j  <-> genome[160 + j*11 + k]
//   s in [270,314) : layer 3, s-270 = k*4 + j   <-> genome[270 + j*11 + k]
// s < 256 sits in Tensor Memory: lane = the thread's TMEM lane, columns 2s (low word), 2s+1 (high
// word); read per step as 32 tcgen05.ld.32x32b.x16 loads of 8 doubles (staging writes and re-reads it in
// x32 chunks of 16).  s >= 256 (58 doubles) sits in this lane's 464-B shared-memory row (29 x 16 B: odd,
// bank-conflict-free LDS.128).
//
// Two CTA layouts (template parameter WARPS of the episode kernel):
//   4 warps: one warp per SM sub-partition, every lane owns its whole TMEM lane (256 doubles) + a 464-B row;
//   5 warps (-DLD_LAYOUT5, measured dead end): warps 0 and 4 SHARE sub-partition 0 — each owns half of its TMEM lanes
//            (columns 0..255 / 256..511, 128 doubles) and keeps the other 186 doubles in a 1488-B shared-memory row
//            (93 x 16 B, odd) — while warps 1..3 keep the 4-warp layout.  160 brains are all that Tensor Memory + shared
//            memory of an SM hold (256 KiB + 227 KiB = 192 brains minus the staging buffers).  The two warps do hide each
//            other's stalls, but a sub-partition takes only ~40 B/clk from shared memory: 3,642 cycles per step for each of
//            the two against 2,000 for a warp alone — the SM's rate is unchanged and 40 % of the lanes fly at half speed
//            (65,536 survivors: 8.02 ms against 6.85 ms; DESIGN.md 4, scripts/layout_probe.py).
constexpr int kWarpsPerCta = 4;                       // the default layout: warp w owns TMEM lanes 32w..32w+31
constexpr int kCtaThreads = kWarpsPerCta * 32;        // 128 drones resident per SM
constexpr int kTmemCols = 512;                        // the whole Tensor Memory of the SM
constexpr int kChunk = 16;                            // doubles per tcgen05.ld / tcgen05.st x32 (staging)
// where one lane's brain lives: TM doubles in Tensor Memory (stored index s < TM), the rest in its shared-memory row
template <int TM>
struct BrainLay {
    static constexpr int kTm = TM;                       // 256 (whole TMEM lane) or 128 (half)
    static constexpr int kChunks = TM / kChunk;          // 16 or 8 x32 chunks
    static constexpr int kSm = LD_GENOME_LEN - TM;       // 58 or 186
    static constexpr int kRowBytes = kSm * 8;            // 464 or 1488
    static_assert(TM % kChunk == 0 && kSm % 2 == 0, "whole chunks, double2 rows");
    static_assert((kRowBytes / 16) % 2 == 1, "row stride must be an odd number of 16-B units (bank-conflict-free LDS.128)");
};
// Brain staging: the rows of the lanes that get a new drone are copied into a per-warp staging buffer with
// coalesced 16-B cp.async (a 2512-B row = 157 pieces), kStageRows rows at a time; the owner lanes then read their
// row in stored order (the genome -> stored permutation costs nothing at LDS time; rows 2512 B apart are
// bank-conflict-free for LDS.64/.128 because 157 is odd).
#ifndef LD_STAGE_ROWS
#define LD_STAGE_ROWS 8
#endif
constexpr int kGenomeBytes = LD_GENOME_LEN * 8;            // 2512
constexpr int kRowPieces = kGenomeBytes / 16;              // 157
#ifndef LD_STAGE_BUFFERS
#define LD_STAGE_BUFFERS 1 // 2 (220 KB of shared memory): the next group's rows arrive while this group's are written — a whole-warp
#endif                     // hand-over drops from 27.5k to 22-24k cycles, but the headline kernel gets 1.1 % SLOWER (7.08 vs 7.00 ms)
constexpr int kStageBuffers = LD_STAGE_BUFFERS;
static_assert(kGenomeBytes % 16 == 0 && kRowPieces % 2 == 1, "16-B pieces, odd row stride");
template <int WARPS>
struct CtaLay {
    static_assert(WARPS == 4 || WARPS == 5, "layouts");
    static constexpr int kWarps = WARPS;
    static constexpr int kThreads = WARPS * 32;
    static constexpr int kHalfWarps = WARPS == 5 ? 2 : 0; // warps 0 and 4 (both on SM sub-partition 0)
    static constexpr int kHalfRowsBytes = kHalfWarps * 32 * BrainLay<128>::kRowBytes;
    static constexpr int kRowsBytes = kHalfRowsBytes + (WARPS - kHalfWarps) * 32 * BrainLay<256>::kRowBytes; // 59392 / 139776
    static constexpr int kStageRows = WARPS == 5 ? 6 : LD_STAGE_ROWS;
    static constexpr int kStageGroupBytes = kStageRows * kGenomeBytes;   // 20096 / 15072
    static constexpr int kStageBytesPerWarp = kStageBuffers * kStageGroupBytes;
    // Tensor Memory is one allocation per SM: the request is more than half of the SM's shared memory, so a
    // second CTA (e.g. of a concurrent launch on another stream) can never become resident next to us
    static constexpr int kSmemBytes = kRowsBytes + WARPS * kStageBytesPerWarp; // 139776 / 215136
    static_assert(kSmemBytes > 116 * 1024 && kSmemBytes <= 227 * 1024, "one CTA per SM");
    __host__ __device__ static constexpr bool half(unsigned warp) { return WARPS == 5 && (warp == 0 || warp == 4); }
};
constexpr int kSmemBytes = CtaLay<4>::kSmemBytes;
constexpr int kRefillBatch = 6;                       // idle lanes that trigger a refill

__host__ __device__ constexpr int stored_to_genome(int s) {
    return s < kL1 ? (s % kH1) * (kIn + 1) + s / kH1
                   : s < kL1 + kL2 ? kL1 + ((s - kL1) % kH2) * (kH1 + 1) + (s - kL1) / kH2
                                   : kL1 + kL2 + ((s - kL1 - kL2) % kOut) * (kH2 + 1) + (s - kL1 - kL2) / kOut;
}

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

// ---- tcgen05 (Tensor Memory) wrappers -----------------------------------------------------------
#define LD_R32(v, o) \
    v[o + 0], v[o + 1], v[o + 2], v[o + 3], v[o + 4], v[o + 5], v[o + 6], v[o + 7], v[o + 8], v[o + 9], v[o + 10], v[o + 11], \
    v[o + 12], v[o + 13], v[o + 14], v[o + 15], v[o + 16], v[o + 17], v[o + 18], v[o + 19], v[o + 20], v[o + 21], v[o + 22], \
    v[o + 23], v[o + 24], v[o + 25], v[o + 26], v[o + 27], v[o + 28], v[o + 29], v[o + 30], v[o + 31]

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld32(uint32_t (&v)[32], uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// x16 form (8 doubles) for the hot loop's weight stream, see mlp_tmem_stream
__device__ __forceinline__ void tmem_ld16(uint32_t (&v)[16], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld_all() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tcgen05.ld is asynchronous: the registers are valid only after wait::ld.  Passing them through
// the wait as read-write operands gives the compiler a true dependency, so no consumer can be
// scheduled above it.
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

__device__ __forceinline__ void tmem_st32(const uint32_t (&v)[32], uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}

__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- the work queue of one launch -----------------------------------------------------------------
// Slots 0..n-1 are the drones themselves (fresh); slot n+i is the i-th drone that was parked at a
// segment boundary (requeue[i] = drone + 1, written once, 0 = not yet).  Lanes claim slots in order
// with one atomicAdd on `head`; a lane whose slot is not filled yet polls it every kPollSteps steps
// and retires when `done` says that every drone has landed.
constexpr int kParkDoubles = 32; // 29 used: x6 v6 acc6 thrust4 coin3 thr2 + 3 packed int pairs
constexpr int kPollSteps = 32;
struct Sched {
    unsigned long long* head; // claimed slots
    unsigned long long* tail; // parked drones appended so far              (NULL when cap == 0)
    unsigned long long* done; // drones whose record has been published      (NULL when cap == 0)
    int* requeue;             // [cap]
    double* park;             // [n][kParkDoubles]
    long long cap;            // 0 = segments off
    int slice;                // steps per segment, 1 <= slice <= max(timesteps, 1)
    int growth;               // 1: equal segments; 2: each segment twice as long as the one before (boundaries s, 3s, 7s, ...)
    int stagger;              // 1: LD_SLICE_STAGGER — only a warp's FIRST batch has a boundary, at H * (warp + 1) / warps
    // host-buffer entry point: genome rows [0, *arrived) are in HBM, the copy engine raises the count
    // after every chunk while the kernel is already flying the first ones (NULL = everything is there)
    const unsigned long long* arrived;
    int* error;               // set to 1 if a warp gave up waiting (kWaitLimitNs); never NULL (work_counter[1] of the device API)
    // optional per-drone start: [n][LD_STATE_LEN] rows (pos vel angle angle_vel acc_current acc_last thrust) that
    // replace ld_params.init_* — fly() on Drones that are not in the reset state (flight/__init__.py:35-39)
    const double* init_state;
    // debug (ld_debug_fly_profile_device): per-phase clock64 totals over all warps, see kProf*; NULL in production
    unsigned long long* prof;
    unsigned long long* prof_times; // [n][2] globaltimer ns of every drone's first start and of its landing, or NULL
};
enum { kProfClaim = 0, kProfStage, kProfWait, kProfHot, kProfPost, kProfStageEvents, kProfFreshLanes, kProfHotEntries, kProfWarpSteps,
       kProfLiveLaneSteps, kProfWarps, kProfStageState, kProfStageCopy, kProfStageTmem, kProfStageGroups, kProfSlowTrigSteps, kProfCount };
static_assert(LD_PROFILE_SLOTS >= kProfCount, "profile slots");
// a warp that has had nothing to fly and nothing become ready for this long gives up (never in a healthy run: it
// turns a lost queue entry into an error instead of a hung GPU).  Measured from the moment the warp began to wait
// without progress, so long launches, debuggers and time-sliced GPUs do not trip it.
constexpr unsigned long long kWaitLimitNs = 20ull * 1000 * 1000 * 1000;

// Checked build (-DLD_CHECKED -> libldrones_checked.so, run by tests/test_gpu_soak.py): compute-sanitizer is closed on
// this GPU pool, so the index arithmetic of the queue / park / re-queue / staging protocol carries its own bounds
// checks; a violation ORs a code into the launch's error word (work_counter[1]), which every entry point reports.
#ifdef LD_CHECKED
#define LD_CHECK(cond, code)                 \
    do {                                     \
        if (!(cond)) atomicOr(q.error, code); \
    } while (0)
#else
#define LD_CHECK(cond, code) \
    do {                     \
    } while (0)
#endif
enum { kErrGaveUp = 1, kErrSlotRange = 2, kErrDroneRange = 4, kErrRequeueOverflow = 8, kErrPublishRange = 16, kErrStageSlot = 32 };

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void park_drone(const Drone& d, double* __restrict__ row) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        row[i] = d.x[i];
        row[6 + i] = d.v[i];
        row[12 + i] = d.acc[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) row[18 + i] = d.thrust[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) row[22 + i] = d.coin[i];
    row[25] = d.thr2;
    row[26] = __hiloint2double(d.coins, d.reward);
    row[27] = __hiloint2double(d.unstable, d.flight_time);
    row[28] = __hiloint2double(d.status, d.coin_value);
}

// L2 loads: another SM wrote the row (and this SM may hold a stale L1 line of an earlier segment)
__device__ __forceinline__ void unpark_drone(Drone& d, const double* __restrict__ row) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        d.x[i] = __ldcg(row + i);
        d.v[i] = __ldcg(row + 6 + i);
        d.acc[i] = __ldcg(row + 12 + i);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) d.thrust[i] = __ldcg(row + 18 + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) d.coin[i] = __ldcg(row + 22 + i);
    d.thr2 = __ldcg(row + 25);
    const double a = __ldcg(row + 26), b = __ldcg(row + 27), c = __ldcg(row + 28);
    d.reward = __double2loint(a);
    d.coins = __double2hiint(a);
    d.flight_time = __double2loint(b);
    d.unstable = __double2hiint(b);
    d.coin_value = __double2loint(c);
    d.status = __double2hiint(c);
}

// Drone attributes + Sensor accelerations as fly() finds them on a Drone that is not freshly reset
__device__ __forceinline__ void load_start_state(Drone& d, const double* __restrict__ row) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        d.x[i] = __ldg(row + i);
        d.v[i] = __ldg(row + 3 + i);
        d.x[3 + i] = __ldg(row + 6 + i);
        d.v[3 + i] = __ldg(row + 9 + i);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) d.acc[i] = __ldg(row + 12 + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) d.thrust[i] = __ldg(row + 24 + i);
}

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

// evolve/brain.py:41-54, one stored weight at a time.  S is a compile-time stored index; every
// neuron's sum is a k-ordered fma chain that starts with the k=0 product and ends with `+ bias`
// (the trailing 1 of np.vstack((x, [[1]]))), followed by its activation.
struct Mlp {
    double in[kIn];
    double h1[kH1], h2[kH2], out[kOut];
};

template <int S>
__device__ __forceinline__ void mlp_mac(Mlp& m, double w) {
    if (S < kL1) {
        constexpr int k = S / kH1, j = S % kH1;
        if (k == 0) m.h1[j] = w * m.in[0];
        else if (k < kIn) m.h1[j] = fma(w, m.in[k < kIn ? k : 0], m.h1[j]);
        else m.h1[j] = relu(m.h1[j] + w);
    } else if (S < kL1 + kL2) {
        constexpr int t = S - kL1 >= 0 ? S - kL1 : 0, k = t / kH2, j = t % kH2;
        if (k == 0) m.h2[j] = w * m.h1[0];
        else if (k < kH1) m.h2[j] = fma(w, m.h1[k < kH1 ? k : 0], m.h2[j]);
        else m.h2[j] = relu(m.h2[j] + w);
    } else {
        constexpr int t = S - kL1 - kL2 >= 0 ? S - kL1 - kL2 : 0, k = t / kOut, j = t % kOut;
        if (k == 0) m.out[j] = w * m.h2[0];
        else if (k < kH2) m.out[j] = fma(w, m.h2[k < kH2 ? k : 0], m.out[j]);
        else m.out[j] = m.out[j] + w;
    }
}

template <int TM, int P = 0>
__device__ __forceinline__ void mlp_smem_tail(Mlp& m, const double2* __restrict__ wrow) {
    if constexpr (P < BrainLay<TM>::kSm / 2) {
        const double2 w = wrow[P];
        mlp_mac<TM + 2 * P>(m, w.x);
        mlp_mac<TM + 2 * P + 1>(m, w.y);
        mlp_smem_tail<TM, P + 1>(m, wrow);
    }
}

#ifndef LD_TMEM_X16
#define LD_TMEM_X16 1
#endif
#if LD_TMEM_X16
// The hot loop's weight stream in x16 loads (8 doubles), one load ahead in the source and NO tcgen05.wait::ld between the
// loads: tcgen05.wait::ld waits for EVERY outstanding load, so with it only one load can usefully be in flight, and ptxas
// issues a tcgen05.ld only ~25 statically scheduled cycles before its first use while an x32 load takes ~35 cycles to land
// — with one warp per scheduler that was ~155 stalled cycles per warp-step.  Without the wait ptxas tracks every load on a
// hardware scoreboard of its own (it alternates two), puts the scoreboard wait on the load's first consumer itself, and two
// x16 loads (~27 cycles each) overlap.  PTX asks for a wait::ld between a tcgen05.ld and the use of its registers; what is
// executed is SASS, so the build VERIFIES the property on the SASS of every episode-kernel instantiation
// (build.verify_tmem_scoreboards: between a tcgen05.ld and the first instruction that waits on its scoreboard nothing
// touches its destination registers) and refuses the library otherwise; the bit-exact parity tests run on top of that.
// One real wait::ld closes every step's stream.
typedef uint32_t TmemChunk[16];
constexpr int kLoadDoubles = 8;
template <int C, int E = 0>
__device__ __forceinline__ void mlp_load_chunk(Mlp& m, const uint32_t (&v)[16]) {
    if constexpr (E < kLoadDoubles) {
        mlp_mac<C * kLoadDoubles + E>(m, __hiloint2double((int)v[2 * E + 1], (int)v[2 * E]));
        mlp_load_chunk<C, E + 1>(m, v);
    }
}
template <int TM, int C>
__device__ __forceinline__ void mlp_tmem_stream(Mlp& m, uint32_t taddr, uint32_t (&cur)[16], uint32_t (&nxt)[16]) {
    constexpr int kLoads = TM / kLoadDoubles;
    if constexpr (C + 1 < kLoads) tmem_ld16(nxt, taddr + 16 * (C + 1));
    mlp_load_chunk<C>(m, cur);
    if constexpr (C + 1 < kLoads) mlp_tmem_stream<TM, C + 1>(m, taddr, nxt, cur);
}
__device__ __forceinline__ void tmem_ld_first(TmemChunk& v, uint32_t taddr) { tmem_ld16(v, taddr); }
#else
// -DLD_TMEM_X16=0: the stream of the first half of round 2 (profiles/r02_b_*), kept for A/B runs: x32 loads, three rotating
// buffers, a wait::ld before every chunk (chunks C..15 with the next two chunks' loads requested while chunk C is consumed)
typedef uint32_t TmemChunk[32];
__device__ __forceinline__ void tmem_ld_first(TmemChunk& v, uint32_t taddr) { tmem_ld32(v, taddr); }
template <int C, int E = 0>
__device__ __forceinline__ void mlp_chunk(Mlp& m, const uint32_t (&v)[32]) {
    if constexpr (E < kChunk) {
        mlp_mac<C * kChunk + E>(m, __hiloint2double((int)v[2 * E + 1], (int)v[2 * E]));
        mlp_chunk<C, E + 1>(m, v);
    }
}
template <int TM, int C>
__device__ __forceinline__ void mlp_tmem_chunks(Mlp& m, uint32_t taddr, uint32_t (&cur)[32], uint32_t (&nxt)[32], uint32_t (&nx2)[32]) {
    if constexpr (C + 2 < BrainLay<TM>::kChunks) tmem_ld32(nx2, taddr + 32 * (C + 2));
    mlp_chunk<C>(m, cur);
    if constexpr (C + 1 < BrainLay<TM>::kChunks) {
        tmem_wait_ld(nxt);
        mlp_tmem_chunks<TM, C + 1>(m, taddr, nxt, nx2, cur);
    }
}
#endif

// evolve/brain.py:56-76 on this lane's staged weights (warp-converged: tcgen05.ld is .sync.aligned)
template <int TM>
__device__ __forceinline__ void brain(Drone& d, uint32_t taddr, const double2* __restrict__ wrow, double neg_beta, bool keep_old,
                                      TmemChunk& a) {
    Mlp m;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        m.in[i] = d.x[i];
        m.in[3 + i] = d.v[i];
        m.in[6 + i] = d.x[3 + i];
        m.in[9 + i] = d.v[3 + i];
        m.in[12 + i] = d.coin[i];
    }
    // chunk 0 was requested before the physics of this step (its latency hides behind it)
#if LD_TMEM_X16
    TmemChunk b;
    mlp_tmem_stream<TM, 0>(m, taddr, a, b);
    mlp_smem_tail<TM>(m, wrow);
    tmem_wait_ld_all(); // nothing is in flight any more: every chunk's consumers have waited on its scoreboard
#else
    uint32_t b[32], c[32];
    tmem_ld32(b, taddr + 32);
    tmem_wait_ld(a);
    mlp_tmem_chunks<TM, 0>(m, taddr, a, b, c);
    mlp_smem_tail<TM>(m, wrow);
#endif
#pragma unroll
    for (int j = 0; j < kOut; ++j) {
        const double th = sigmoid_fast(m.out[j], neg_beta);
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

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Tensor-Memory part of the brains of one staging group: read-modify-write of whole 16-double chunks (tcgen05.st
// writes all 32 lanes, so lanes that keep their drone write back what they read), four chunks per round so that
// the waits and the LDS latency are paid four times instead of sixteen.  `mine` = this lane's staged row (genome
// order) if it is in the group.  Warp-converged.
constexpr int kStageBatch = 2; // 4 spills the hot loop's registers (255 + 1.5 KB of spill stores): measured 30 % slower
template <int TM, int C0 = 0>
__device__ __forceinline__ void stage_tmem(uint32_t taddr, bool in_group, const double* __restrict__ mine) {
    if constexpr (C0 < BrainLay<TM>::kChunks) {
        uint32_t v[kStageBatch][32];
#pragma unroll
        for (int q = 0; q < kStageBatch; ++q) tmem_ld32(v[q], taddr + 32 * (C0 + q));
#pragma unroll
        for (int q = 0; q < kStageBatch; ++q) tmem_wait_ld(v[q]);
        if (in_group) {
#pragma unroll
            for (int q = 0; q < kStageBatch; ++q) {
#pragma unroll
                for (int e = 0; e < kChunk; ++e) {
                    const double w = mine[stored_to_genome((C0 + q) * kChunk + e)];
                    v[q][2 * e] = (uint32_t)__double2loint(w);
                    v[q][2 * e + 1] = (uint32_t)__double2hiint(w);
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kStageBatch; ++q) tmem_st32(v[q], taddr + 32 * (C0 + q));
        stage_tmem<TM, C0 + kStageBatch>(taddr, in_group, mine);
    }
}
static_assert(BrainLay<256>::kChunks % kStageBatch == 0 && BrainLay<128>::kChunks % kStageBatch == 0, "staging rounds");

// One population evaluation = the loop of evolve/evolution.py:64-73 around flight/__init__.py:fly().
// MODE kState: also write the final kinematic state; kTrace: n == 1, dump the state after every
// measure() (debug tap).  All 32 lanes execute every step (idle lanes run masked) so that the
// step body is one straight-line, warp-converged block.
// The whole life of one warp of the episode kernel.  TM = this warp's brain layout (BrainLay): `taddr` is the first of its
// TM/... Tensor-Memory columns in its own lane quadrant, `wrow` this lane's shared-memory row, `stage_buf` the warp's staging
// buffer.
template <bool EXTRAS, int MODE, bool PROF, int WARPS, int TM>
__device__ __forceinline__ void fly_warp(const ld_params& p, const double* __restrict__ genomes, long long n,
                                         ld_record* __restrict__ records, double* __restrict__ final_state, const Sched& q,
                                         double* __restrict__ trace, const uint32_t taddr, double2* __restrict__ wrow,
                                         unsigned char* __restrict__ stage_buf, const unsigned lane, const unsigned warp) {
    constexpr int kStageRows = CtaLay<WARPS>::kStageRows, kStageGroupBytes = CtaLay<WARPS>::kStageGroupBytes;
    const double neg_beta = -p.sigmoid_beta;
    const double rcp_m = 1.0 / (double)p.m_total32;
    const int time_reward = (p.reward_time ? p.reward_time_passed : 0) - (p.punish_time ? p.reward_time_passed : 0);
    const int H = p.timesteps;
    const long long n_slots = n + q.cap;
    const int batch = p.refill_batch > 0 ? p.refill_batch : kRefillBatch;

    unsigned long long wait_since = 0; // 0 = the warp is not waiting (warp-uniform)
    unsigned long long pf[kProfCount] = {}; // PROF only
    long long pt = PROF ? clock64() : 0;
#define LD_PROF_LAP(slot)                          \
    if (PROF) {                                    \
        const long long now_ = clock64();          \
        pf[slot] += (unsigned long long)(now_ - pt); \
        pt = now_;                                 \
    }

    Drone d;
    drone_reset(p, d);
    long long idx = -1;  // drone this lane is flying; -1 = none
    long long slot = -1; // claimed queue slot whose drone is not staged yet; -1 = none
    int t = 0, t_end = 0;
    bool more = true; // slots left to claim (warp-uniform)
    // LD_SLICE_STAGGER: the boundary of this warp's first batch, spread evenly over the horizon across the grid's warps
    const int total_warps = (int)gridDim.x * WARPS;
    int first_boundary = q.stagger ? max(1, (int)(((long long)H * ((long long)blockIdx.x * WARPS + warp + 1)) / total_warps)) : 0;

    for (;;) {
        // ---- claim queue slots for idle lanes (warp-aggregated, batched) ---------------------------
        const unsigned occupied = __ballot_sync(0xffffffffu, idx >= 0 || slot >= 0);
        const unsigned idle = ~occupied;
        if (more && idle != 0u && (__popc(idle) >= batch || occupied == 0u)) {
            unsigned long long base = 0;
            const int leader = __ffs(idle) - 1;
            if ((int)lane == leader) base = atomicAdd(q.head, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            const long long mine = (long long)base + __popc(idle & ((1u << lane) - 1u));
            if ((idle >> lane) & 1u) slot = mine < n_slots ? mine : -1;
            LD_CHECK(slot < n_slots, kErrSlotRange);
            more = (long long)base + __popc(idle) < n_slots;
        }
        // ---- whose slot is filled? -------------------------------------------------------------------
        bool ready = false, resume = false, drained = false;
        long long who = 0;
        if (slot >= 0) {
            if (slot < n) {
                // plain (coherent) loads of the row after this acquire see what the copy engine wrote before
                // it raised the count
                who = slot;
                ready = q.arrived == nullptr || slot < (long long)ld_acquire_sys(q.arrived);
            } else {
                const int v = ld_acquire(q.requeue + (slot - n));
                if (v != 0) {
                    who = v - 1;
                    ready = resume = true;
                } else if ((long long)ld_acquire(q.done) >= n) {
                    slot = -1; // every drone has landed: nothing will ever be parked again
                    drained = true;
                }
            }
        }
        LD_CHECK(!ready || (who >= 0 && who < n), kErrDroneRange);
        if (__any_sync(0xffffffffu, drained)) more = false;
        LD_PROF_LAP(kProfClaim)
        // ---- stage the brains of the lanes that got a drone ------------------------------------------
        if (__any_sync(0xffffffffu, ready)) {
            wait_since = 0; // progress
            const double* __restrict__ row = genomes + who * LD_GENOME_LEN;
            if (ready) {
                idx = who;
                slot = -1;
                if (resume) unpark_drone(d, q.park + who * kParkDoubles);
                else {
                    drone_reset(p, d);
                    if (q.init_state) load_start_state(d, q.init_state + who * LD_STATE_LEN);
                    if (PROF && q.prof_times) q.prof_times[2 * who] = globaltimer_ns();
                }
                t = d.flight_time;
                t_end = min(H, q.growth * t + q.slice);
                if (first_boundary > 0 && !resume) t_end = first_boundary;
            }
            first_boundary = 0; // warp-uniform: only the first batch of the warp
            long long ps = PROF ? clock64() : 0; // sub-phases of the staging (PROF only)
            if (PROF) pf[kProfStageState] += (unsigned long long)(ps - pt);
            // brains: kStageRows lanes at a time; with two staging buffers the next group's rows are requested before
            // this group's are written to Tensor Memory
            auto take_group = [](unsigned& todo) {
                unsigned group = 0;
#pragma unroll
                for (int j = 0; j < kStageRows; ++j) {
                    const unsigned low = todo & (0u - todo);
                    group |= low;
                    todo ^= low;
                }
                return group;
            };
            // coalesced asynchronous copy of a group's rows (L2 -> shared memory, no registers, one round trip)
            auto request_rows = [&](unsigned group, int buf) {
                unsigned g = group;
#pragma unroll
                for (int j = 0; j < kStageRows; ++j) {
                    const bool have = g != 0u; // warp-uniform
                    const int src = have ? __ffs(g) - 1 : 0;
                    g &= g - 1u;
                    const unsigned long long rp = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)row, src);
                    if (have) {
                        const uint32_t dst = smem_u32(stage_buf) + buf * kStageGroupBytes + j * kGenomeBytes;
#pragma unroll
                        for (int i = 0; i < (kRowPieces + 31) / 32; ++i) {
                            const int piece = (int)lane + 32 * i;
                            if (piece < kRowPieces) cp_async16(dst + piece * 16, reinterpret_cast<const char*>(rp) + piece * 16);
                        }
                    }
                }
                cp_async_commit();
            };
            unsigned todo = __ballot_sync(0xffffffffu, ready);
            unsigned group = take_group(todo);
            int buf = 0;
            request_rows(group, 0);
            while (group != 0u) { // warp-uniform
                unsigned next = 0;
                if (kStageBuffers > 1) {
                    next = take_group(todo);
                    if (next != 0u) {
                        request_rows(next, buf ^ 1);
                        cp_async_wait_group<1>(); // everything but the group just requested
                    } else {
                        cp_async_wait_group<0>();
                    }
                } else {
                    cp_async_wait_group<0>();
                }
                __syncwarp();
                if (PROF) {
                    const long long now_ = clock64();
                    pf[kProfStageCopy] += (unsigned long long)(now_ - ps);
                    pf[kProfStageGroups] += 1;
                    ps = now_;
                }
                const bool in_group = (group >> lane) & 1u;
                LD_CHECK(__popc(group) <= kStageRows && __popc(group & ((1u << lane) - 1u)) < kStageRows + (in_group ? 0 : 32), kErrStageSlot);
                const double* __restrict__ mine = reinterpret_cast<const double*>(
                    stage_buf + (size_t)buf * kStageGroupBytes + (size_t)__popc(group & ((1u << lane) - 1u)) * kGenomeBytes);
                if (in_group) {
#pragma unroll
                    for (int i = 0; i < BrainLay<TM>::kSm / 2; ++i)
                        wrow[i] = make_double2(mine[stored_to_genome(TM + 2 * i)], mine[stored_to_genome(TM + 2 * i + 1)]);
                }
                stage_tmem<TM>(taddr, in_group, mine);
                __syncwarp(); // everyone is done with this buffer before it is requested again
                if (PROF) {
                    const long long now_ = clock64();
                    pf[kProfStageTmem] += (unsigned long long)(now_ - ps);
                    ps = now_;
                }
                if (kStageBuffers > 1) {
                    group = next;
                    buf ^= 1;
                } else {
                    group = take_group(todo);
                    if (group != 0u) request_rows(group, 0);
                }
            }
            tmem_wait_st();
            trig_update(d); // warp-converged; recomputing it for lanes in flight is idempotent
            if (PROF) {
                pf[kProfStageEvents] += 1;
                pf[kProfFreshLanes] += __popc(__ballot_sync(0xffffffffu, ready));
            }
            LD_PROF_LAP(kProfStage)
        }
        const unsigned flying = __ballot_sync(0xffffffffu, idx >= 0);
        if (flying == 0u) {
            const unsigned pending = __ballot_sync(0xffffffffu, slot >= 0);
            if (pending == 0u && !more) break;
            if (pending != 0u) { // wait for rows to arrive, for a drone to be parked, or for the last one to land
                __nanosleep(500);
                const unsigned long long now = globaltimer_ns();
                if (wait_since == 0) wait_since = now;
                else if (now - wait_since > kWaitLimitNs) {
                    atomicOr(q.error, (int)kErrGaveUp); // reported by the host entry points and through work_counter[1] of the device API
                    slot = -1;
                    more = false;
                }
            }
            LD_PROF_LAP(kProfWait)
            continue;
        }

        if (H <= 0) { // fly(timesteps=0): the loop body never runs (warp-uniform, cold)
            if (idx >= 0) {
                const double zeros[6] = {0, 0, 0, 0, 0, 0};
                publish<MODE>(p, d, zeros, idx, records, final_state);
                if (q.cap) atomicAdd(q.done, 1ull);
                idx = -1;
            }
            continue;
        }
        // ---- hot loop: timesteps of flight/__init__.py:39-93 until some lane's segment ends ---------
        // `live` cannot change inside: lanes only leave (publish / park) or join (stage) outside of it.
        // Lanes that wait for their slot count kPollSteps steps and then have it looked at again.
        wait_since = 0; // some lane of the warp is flying
        const bool live = idx >= 0;
        const bool counting = live || slot >= 0;
        if (!live) {
            t = 0;
            t_end = kPollSteps;
        }
        const double dt_lane = live ? p.dt : 0.0; // lanes without a drone do not move (see physics_step)
        double a_last[6];
        bool crashed, done, big;
        int prof_iters = 0;
        do {
            int u;
            if (PROF) ++prof_iters;
            TmemChunk w0;
            tmem_ld_first(w0, taddr); // first brain weights for this step's MLP, in flight during the physics
            step_rewards_and_coin(p, d, live, time_reward);
            crashed = physics_step<EXTRAS>(p, d, rcp_m, MODE != kPlain ? a_last : nullptr, u, dt_lane) && live;
            d.reward += crashed ? p.reward_crashed : ((live ? u : 0) * p.reward_unstable);
            d.unstable += (live && !crashed) ? u : 0;
            d.status |= crashed ? LD_STATUS_CRASHED : 0;
            if (MODE == kTrace) {
                if (live) write_state(d, a_last, trace + (size_t)t * LD_STATE_LEN);
            }
            // next step's trigonometry and this step's brain: independent, one basic block, interleaved
            // by the scheduler
            trig_fast(d);
            brain<TM>(d, taddr, wrow, neg_beta, MODE != kPlain && crashed, w0);
            if (MODE == kTrace) {
                if (live && !crashed) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) trace[(size_t)t * LD_STATE_LEN + 24 + i] = d.thrust[i];
                }
            }
            t += counting ? 1 : 0;
            done = counting && (crashed || t >= t_end);
            // one vote per step: the large-angle case (a few tumbling drones in ten thousand) shares it
            big = trig_is_big(d);
            if (!__any_sync(0xffffffffu, done || big)) continue;
            if (big) trig_slow(d);
            if (PROF) pf[kProfSlowTrigSteps] += __any_sync(0xffffffffu, big) ? 1 : 0;
            if (__any_sync(0xffffffffu, done)) break;
        } while (true);
        if (PROF) {
            pf[kProfHotEntries] += 1;
            pf[kProfWarpSteps] += prof_iters;
            pf[kProfLiveLaneSteps] += (unsigned long long)prof_iters * __popc(__ballot_sync(0xffffffffu, live));
        }
        LD_PROF_LAP(kProfHot)
        if (done && live) {
            bool leave = true;
            LD_CHECK(idx >= 0 && idx < n, kErrPublishRange);
            if (crashed || t >= H) {
                publish<MODE>(p, d, a_last, idx, records, final_state);
                if (q.cap) atomicAdd(q.done, 1ull);
                if (PROF && q.prof_times) q.prof_times[2 * idx + 1] = globaltimer_ns();
            } else {
                // segment boundary: give the lane up only if some drone is waiting for one, i.e. if the next
                // slot to be claimed is filled (a fresh row that has arrived, or a parked drone)
                bool wanted = false;
                if (q.cap) {
                    const long long next = (long long)*(volatile unsigned long long*)q.head;
                    if (next < n) wanted = q.arrived == nullptr || next < (long long)*(volatile const unsigned long long*)q.arrived;
                    else wanted = next < n + (long long)*(volatile unsigned long long*)q.tail;
                }
                if (wanted) {
                    park_drone(d, q.park + idx * kParkDoubles);
                    __threadfence();
                    const unsigned long long at = atomicAdd(q.tail, 1ull);
                    LD_CHECK((long long)at < q.cap, kErrRequeueOverflow);
                    st_release(q.requeue + at, (int)idx + 1);
                } else {
                    t_end = min(H, q.growth * t + q.slice);
                    leave = false;
                }
            }
            if (leave) {
                idx = -1;
                // the lane keeps stepping (masked) until it gets its next drone: park it in the initial state so
                // that a wildly tumbling leftover cannot drag the warp into the large-angle fallback branch
                drone_reset(p, d);
            }
        }
        LD_PROF_LAP(kProfPost)
    }
#undef LD_PROF_LAP
    if (PROF) {
        pf[kProfWarps] = 1;
        if (lane == 0 && q.prof) {
#pragma unroll
            for (int i = 0; i < kProfCount; ++i) atomicAdd(q.prof + i, pf[i]);
        }
    }
}

template <bool EXTRAS, int MODE, bool PROF = false, int WARPS = 4>
__global__ void __launch_bounds__(CtaLay<WARPS>::kThreads, 1)
fly_mlp_f64_kernel(const __grid_constant__ ld_params p, const double* __restrict__ genomes, long long n,
                   ld_record* __restrict__ records, double* __restrict__ final_state,
                   const __grid_constant__ Sched q, double* __restrict__ trace) {
    using L = CtaLay<WARPS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned char* stage_buf = smem_raw + L::kRowsBytes + (size_t)warp * L::kStageBytesPerWarp; // this warp's kStageRows staged rows

    // the CTA owns the SM's whole Tensor Memory (512 columns x 128 lanes x 32 bit = 256 KiB)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_slot;
    // a warp can only reach the TMEM lanes of its own SM sub-partition (warp % 4)
    const uint32_t quadrant = tmem_base + (((warp & 3u) * 32u) << 16);
    if (L::half(warp)) {
        // warps 0 and 4 share sub-partition 0: columns 0..255 / 256..511 of lanes 0..31; rows of 1488 B at the start
        unsigned char* row = smem_raw + (size_t)((warp == 4 ? 32u : 0u) + lane) * BrainLay<128>::kRowBytes;
        fly_warp<EXTRAS, MODE, PROF, WARPS, 128>(p, genomes, n, records, final_state, q, trace, quadrant + (warp == 4 ? 256u : 0u),
                                                 reinterpret_cast<double2*>(row), stage_buf, lane, warp);
    } else {
        const unsigned full_index = WARPS == 5 ? (warp - 1u) * 32u + lane : threadIdx.x;
        unsigned char* row = smem_raw + L::kHalfRowsBytes + (size_t)full_index * BrainLay<256>::kRowBytes;
        fly_warp<EXTRAS, MODE, PROF, WARPS, 256>(p, genomes, n, records, final_state, q, trace, quadrant, reinterpret_cast<double2*>(row),
                                                 stage_buf, lane, warp);
    }

    // every warp is done with its Tensor-Memory lanes before the allocation is returned
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// ---- fp32 fast mode: same episode, float32 state and brain, 8 warps per SM (see ld_f32.cuh) ---------
constexpr int kSmemBytesF32 = 120 * 1024; // > half of the SM's shared memory: one CTA (one TMEM owner) per SM
static_assert(ldf::kSmemUsed <= kSmemBytesF32, "fp32 shared-memory rows");

template <int C = 0>
__device__ __forceinline__ void stage_tmem_f32(uint32_t taddr, bool mine_fresh, const double* __restrict__ row) {
    if constexpr (C < ldf::kChunks) {
        uint32_t v[32];
        tmem_ld32(v, taddr + 32 * C);
        tmem_wait_ld(v);
        if (mine_fresh) {
#pragma unroll
            for (int e = 0; e < ldf::kChunk; ++e) v[e] = __float_as_uint((float)__ldg(row + stored_to_genome(C * ldf::kChunk + e)));
        }
        __syncwarp();
        tmem_st32(v, taddr + 32 * C);
        stage_tmem_f32<C + 1>(taddr, mine_fresh, row);
    }
}

template <int C>
__device__ __forceinline__ void mlp_chunks_f32(ldf::MlpF& m, uint32_t taddr, uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
    if constexpr (C + 1 < ldf::kChunks) tmem_ld32(nxt, taddr + 32 * (C + 1));
    ldf::chunk<C>(m, cur);
    if constexpr (C + 1 < ldf::kChunks) {
        tmem_wait_ld(nxt);
        mlp_chunks_f32<C + 1>(m, taddr, nxt, cur);
    }
}

template <bool EXTRAS, bool STATE>
__global__ void __launch_bounds__(ldf::kThreads, 1)
fly_mlp_f32_kernel(const __grid_constant__ ld_params p, const double* __restrict__ genomes, long long n,
                   ld_record* __restrict__ records, double* __restrict__ final_state,
                   unsigned long long* __restrict__ queue_head) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    float2* wrow = reinterpret_cast<float2*>(smem_raw + (size_t)threadIdx.x * ldf::kRowBytes);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_slot;
    // warps w and w+4 share the lane quadrant w%4 and take 256 columns each
    const uint32_t taddr = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 256u;

    const ldf::ParamsF f = ldf::narrow(p);
    const int time_reward = (p.reward_time ? p.reward_time_passed : 0) - (p.punish_time ? p.reward_time_passed : 0);
    const int H = p.timesteps;
    ldf::DroneF d;
    ldf::reset(p, f, d);
    long long idx = -1;
    int t = 0;
    bool more = true;

    auto publish_f32 = [&]() {
        const float dist = sqrtf(ldf::dist2(d));
        long long reward = d.reward;
        if (isfinite(dist)) reward += (long long)p.reward_coin_distance * (long long)dist;
        else d.status |= LD_STATUS_NONFINITE;
        int4 lo, hi;
        lo.x = (int)(unsigned)(reward & 0xffffffffLL);
        lo.y = (int)(reward >> 32);
        lo.z = d.coins;
        lo.w = d.flight_time;
        hi.x = d.status;
        hi.y = d.unstable;
        hi.z = __double2loint((double)dist);
        hi.w = __double2hiint((double)dist);
        int4* out = reinterpret_cast<int4*>(records + idx);
        out[0] = lo;
        out[1] = hi;
        if (STATE && final_state) {
            double* row = final_state + idx * LD_STATE_LEN;
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
                row[18 + i] = 0.0; // previous acceleration is not kept in the fast mode
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) row[24 + i] = d.thrust[i];
        }
    };

    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, idx < 0);
        if (more && (__popc(idle) >= (p.refill_batch > 0 ? p.refill_batch : kRefillBatch) || idle == 0xffffffffu)) {
            unsigned long long base = 0;
            const int leader = __ffs(idle) - 1;
            if ((int)lane == leader) base = atomicAdd(queue_head, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            const long long mine = (long long)base + __popc(idle & ((1u << lane) - 1u));
            const bool fresh = idx < 0 && mine < n;
            const double* __restrict__ row = genomes + (fresh ? mine : 0) * LD_GENOME_LEN;
            if (fresh) {
                unsigned touch = 0;
#pragma unroll
                for (int i = 0; i < LD_GENOME_LEN; i += 16) touch ^= (unsigned)__double2loint(__ldg(row + i));
                if (touch == 0x9e3779b9u) wrow[0].x = 0.0f;
#pragma unroll
                for (int i = 0; i < ldf::kSmemFloats / 2; ++i)
                    wrow[i] = make_float2((float)__ldg(row + stored_to_genome(ldf::kTmemFloats + 2 * i)),
                                          (float)__ldg(row + stored_to_genome(ldf::kTmemFloats + 2 * i + 1)));
                idx = mine;
                ldf::reset(p, f, d);
                t = 0;
            }
            __syncwarp();
            stage_tmem_f32(taddr, fresh, row);
            tmem_wait_st();
            more = (long long)base + __popc(idle) < n;
            ldf::trig(d);
        }
        if (__ballot_sync(0xffffffffu, idx >= 0) == 0u) break;
        if (H <= 0) {
            if (idx >= 0) {
                publish_f32();
                idx = -1;
            }
            continue;
        }
        const bool live = idx >= 0;
        bool crashed, done;
        do {
            int u;
            ldf::rewards_and_coin(p, f, d, live, time_reward);
            crashed = ldf::physics<EXTRAS>(f, d, u) && live;
            d.reward += crashed ? p.reward_crashed : ((live ? u : 0) * p.reward_unstable);
            d.unstable += (live && !crashed) ? u : 0;
            d.status |= crashed ? LD_STATUS_CRASHED : 0;
            ldf::trig(d);
            ldf::MlpF m;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                m.in[i] = d.x[i];
                m.in[3 + i] = d.v[i];
                m.in[6 + i] = d.x[3 + i];
                m.in[9 + i] = d.v[3 + i];
                m.in[12 + i] = d.coin[i];
            }
            uint32_t a[32], b[32];
            tmem_ld32(a, taddr);
            tmem_wait_ld(a);
            mlp_chunks_f32<0>(m, taddr, a, b);
            ldf::smem_tail(m, wrow);
#pragma unroll
            for (int j = 0; j < ldf::kOut; ++j) {
                const float th = __fdividef(1.0f, __expf(f.neg_beta * m.out[j]) + 1.0f); // brain.py:37-39
                d.thrust[j] = (STATE && crashed) ? d.thrust[j] : th;
            }
            t += live ? 1 : 0;
            done = live && (crashed || t >= H);
        } while (!__any_sync(0xffffffffu, done));
        if (done) {
            publish_f32();
            idx = -1;
        }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// ---- PID controller mode: flight/collect_coins.py:278-364 for a batch of gain sets ---------------
struct PidCtl {
    double kp, ki, kd, I, prev_e, prev_pv;
};

// drone/controller/pid.py:77-106 ("signChange"), :132-134 (accumulate), :136-142 (clamp)
__device__ __forceinline__ double pid_calculate(PidCtl& c, double set, double pv, double dt, double rc_dt, double range,
                                                double lo, double hi) {
    const double error = set - pv;
    if (c.prev_e * error < 0.0) c.I = 0.0;
    if (fabs(error) < range) c.I = c.I + error;
    const double diff = (c.prev_pv != 0.0) ? pv - c.prev_pv : 0.0;
    double out = c.kp * error + c.ki * c.I * dt - div_by(c.kd * diff, dt, rc_dt);
    c.prev_pv = pv;
    if (error != 0.0) c.prev_e = error;
    out = out < lo ? lo : (out > hi ? hi : out);
    return out;
}

template <bool EXTRAS>
__global__ void __launch_bounds__(128)
fly_pid_f64_kernel(const __grid_constant__ ld_params p, const __grid_constant__ ld_pid_params q,
                   const double* __restrict__ gains, long long n, ld_record* __restrict__ records,
                   double* __restrict__ final_state) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool exists = i < n;
    const double* __restrict__ g = gains + (exists ? i : 0) * 18;
    PidCtl pid[6]; // x y z roll pitch yaw
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        pid[k].kp = __ldg(g + 3 * k);
        pid[k].ki = __ldg(g + 3 * k + 1);
        pid[k].kd = __ldg(g + 3 * k + 2);
        pid[k].I = pid[k].prev_e = pid[k].prev_pv = 0.0;
    }
    const double rcp_m = 1.0 / (double)p.m_total32;
    const double dt = p.dt, rc_dt = 1.0 / p.dt;
    Drone d;
    drone_reset(p, d);
    trig_update(d);
    long long score = q.timesteps; // collect_coins.py:277
    int target = 0, steps = 0, coins = 0;
    bool active = exists;
    double a_last[6] = {0, 0, 0, 0, 0, 0};

    auto publish_pid = [&](int status) {
#pragma unroll
        for (int k = 0; k < 3; ++k) d.coin[k] = q.coin_pos[target][k];
        const double dist = sqrt(coin_dist2(d));
        int4 lo, hi;
        lo.x = (int)(unsigned)(score & 0xffffffffLL);
        lo.y = (int)(score >> 32);
        lo.z = coins;
        lo.w = steps;
        hi.x = status;
        hi.y = 0;
        hi.z = __double2loint(dist);
        hi.w = __double2hiint(dist);
        int4* out = reinterpret_cast<int4*>(records + i);
        out[0] = lo;
        out[1] = hi;
        if (final_state) write_state(d, a_last, final_state + i * LD_STATE_LEN);
    };

    for (int t = 0; t < q.timesteps; ++t) {
        if (!__any_sync(0xffffffffu, active)) break;
        // :279-292 — step cost, coin test on the current target, exit when all targets are collected
        score += active ? q.step_reward : 0;
        steps += active ? 1 : 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) d.coin[k] = q.coin_pos[target][k];
        const bool hit = active && (coin_dist2(d) < q.coin_thr2[target]);
        score += hit ? q.coin_value[target] : 0;
        coins += hit ? 1 : 0;
        const int next = target + (hit ? 1 : 0);
        if (active && next >= q.n_coins) {
            publish_pid(LD_STATUS_FINISHED);
            active = false;
        }
        target = next < q.n_coins ? next : target; // :294-311, setpoints follow the new target
        double set[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) set[k] = q.coin_pos[target][k];
        // :328-341 — physics + integrator (no crash rule in this script)
        int u;
        physics_step<EXTRAS>(p, d, rcp_m, a_last, u, p.dt);
        trig_update(d);
        // drone/onboard_computer.py:50-85 — cascaded PIDs
        double lin_out[3], rot_out[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) lin_out[k] = pid_calculate(pid[k], set[k], d.x[k], dt, rc_dt, q.integral_range, q.out_lo[k], q.out_hi[k]);
        const double rset[3] = {-lin_out[1], lin_out[0], q.yaw_setpoint};
#pragma unroll
        for (int k = 0; k < 3; ++k)
            rot_out[k] = pid_calculate(pid[3 + k], rset[k], d.x[3 + k], dt, rc_dt, q.integral_range, q.out_lo[3 + k], q.out_hi[3 + k]);
        // :87-118 — mixer: tilt-compensated base thrust, double clip to [0, 1]
        const double base = q.base_thrust + lin_out[2];
        const double cc = d.trig.cr * d.trig.cp; // math.cos(roll) * math.cos(pitch) of the NEW angles
        double bt = base * div_by(1.0, cc, recip(cc));
        bt = bt > 1.0 ? 1.0 : bt;
        bt = bt < 0.0 ? 0.0 : bt;
        const double r0 = rot_out[0], r1 = rot_out[1], r2 = rot_out[2];
        const double des[4] = {(r0 - r1) - r2, (-r0 - r1) + r2, (-r0 + r1) - r2, (r0 + r1) + r2}; // physical_model.py:64-66
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double th = bt + des[j];
            th = th > 1.0 ? 1.0 : th;
            th = th < 0.0 ? 0.0 : th;
            d.thrust[j] = th;
        }
    }
    if (active) publish_pid(0);
}

// ---- GA step: evolve/util.py:5-78 + drone/drone.py:167-172, one warp per pair of children --------
struct U4 {
    uint32_t v[4];
};

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    U4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * 0x1.0p-53;
}

constexpr uint32_t kStreamSelect = 1, kStreamMutate = 2, kStreamNoise = 3;

// P2P: the parent population is sharded over the GPUs of the box; `peers.shard[r]` is rank r's shard
// mapped into this device's address space (NVLink peer memory), rows are fetched with plain loads.
template <bool P2P>
__global__ void __launch_bounds__(256)
ga_step_kernel(const __grid_constant__ ld_ga_params ga, const double* __restrict__ parents, const __grid_constant__ ld_peer_table peers,
               const ld_record* __restrict__ records, long long n, double* __restrict__ children, long long lo, long long hi) {
    __shared__ unsigned short mut_list[8][2 * LD_GENOME_LEN]; // per warp: (gene << 1 | child) of the genes to mutate
    const uint32_t k0 = (uint32_t)ga.seed, k1 = (uint32_t)(ga.seed >> 32), gen = ga.generation;
    const unsigned lane = threadIdx.x & 31u;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long c = lo + 2 * ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); c < hi; c += 2 * warps) {
        const uint32_t pair = (uint32_t)(c >> 1);
        // selection: every lane derives the same draws (counter-based), no shuffles needed
        const U4 a = philox4x32_10(pair, 0, kStreamSelect, gen, k0, k1), b = philox4x32_10(pair, 1, kStreamSelect, gen, k0, k1);
        const U4 t = philox4x32_10(pair, 2, kStreamSelect, gen, k0, k1);
        const long long idx[4] = {(long long)(u53(a.v[0], a.v[1]) * (double)n), (long long)(u53(a.v[2], a.v[3]) * (double)n),
                                  (long long)(u53(b.v[0], b.v[1]) * (double)n), (long long)(u53(b.v[2], b.v[3]) * (double)n)};
        const double r[2] = {u53(t.v[0], t.v[1]), u53(t.v[2], t.v[3])};
        long long win[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const long long i1 = idx[2 * w], i2 = idx[2 * w + 1];
            const long long s1 = __ldg(reinterpret_cast<const long long*>(records + i1));
            const long long s2 = __ldg(reinterpret_cast<const long long*>(records + i2));
            const long long better = s1 > s2 ? i1 : i2, worse = s1 > s2 ? i2 : i1;
            win[w] = r[w] < ga.p_tournament ? better : worse;
        }
        int cut[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const U4 x = philox4x32_10(pair, 3 + l, kStreamSelect, gen, k0, k1);
            const int lo_g = l == 0 ? 0 : (l == 1 ? kL1 : kL1 + kL2), genes = l == 0 ? kL1 : (l == 1 ? kL2 : kL3);
            // cut position in genome coordinates; "no crossover" = cut at the layer's end
            cut[l] = u53(x.v[0], x.v[1]) < ga.p_crossover ? lo_g + (int)(u53(x.v[2], x.v[3]) * (double)genes) : lo_g + genes;
        }
        const double* __restrict__ pa;
        const double* __restrict__ pb;
        if (P2P) {
            const long long ra = win[0] / peers.shard_rows, rb = win[1] / peers.shard_rows;
            pa = peers.shard[ra] + (win[0] - ra * peers.shard_rows) * LD_GENOME_LEN;
            pb = peers.shard[rb] + (win[1] - rb * peers.shard_rows) * LD_GENOME_LEN;
        } else {
            pa = parents + win[0] * LD_GENOME_LEN;
            pb = parents + win[1] * LD_GENOME_LEN;
        }
        double* __restrict__ ca = children + (c - lo) * LD_GENOME_LEN;
        double* __restrict__ cb = children + (c + 1 - lo) * LD_GENOME_LEN;
        // Pass 1 — copy with crossover.  Each lane handles groups of 4 consecutive genes (32 B per row, 16-B
        // vector accesses; rows are 16-B aligned).  One Philox block per child and group gives the four
        // `random_sample() < mutation_rate` decisions (uniform = word * 2^-32); mutated genes are only LISTED here.
        constexpr int kGroups = (LD_GENOME_LEN + 3) / 4; // 79, the last group has 2 genes
        unsigned short* list = mut_list[threadIdx.x >> 5];
        int n_mut = 0; // warp-uniform
        for (int q0 = 0; q0 < kGroups; q0 += 32) {
            const int q = q0 + (int)lane;
            const bool have = q < kGroups;
            const int g0 = 4 * q, cnt = !have ? 0 : ((LD_GENOME_LEN - g0) < 4 ? (LD_GENOME_LEN - g0) : 4); // 4, or 2 for the last group
            U4 m[2];
            m[0] = philox4x32_10((uint32_t)c, (uint32_t)q, kStreamMutate, gen, k0, k1);
            m[1] = philox4x32_10((uint32_t)(c + 1), (uint32_t)q, kStreamMutate, gen, k0, k1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (2 * h < cnt) {
                    // peer memory is not covered by the read-only (non-coherent) path's assumptions: plain loads there
                    const double2* qa = reinterpret_cast<const double2*>(pa + g0) + h;
                    const double2* qb = reinterpret_cast<const double2*>(pb + g0) + h;
                    const double2 ta = P2P ? *qa : __ldg(qa), tb = P2P ? *qb : __ldg(qb);
                    const int ga0 = g0 + 2 * h, ga1 = ga0 + 1;
                    const bool s0 = ga0 >= cut[ga0 < kL1 ? 0 : (ga0 < kL1 + kL2 ? 1 : 2)]; // tail of a crossed layer comes
                    const bool s1 = ga1 >= cut[ga1 < kL1 ? 0 : (ga1 < kL1 + kL2 ? 1 : 2)]; // from the other child
                    reinterpret_cast<double2*>(ca + g0)[h] = make_double2(s0 ? tb.x : ta.x, s1 ? tb.y : ta.y);
                    reinterpret_cast<double2*>(cb + g0)[h] = make_double2(s0 ? ta.x : tb.x, s1 ? ta.y : tb.y);
                }
            }
#pragma unroll
            for (int w = 0; w < 2; ++w) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool hit = j < cnt && ((double)m[w].v[j] * 0x1.0p-32 < ga.mutation_rate); // random_sample() < rate
                    const unsigned vote = __ballot_sync(0xffffffffu, hit);
                    if (hit) list[n_mut + __popc(vote & ((1u << lane) - 1u))] = (unsigned short)(((g0 + j) << 1) | w);
                    n_mut += __popc(vote);
                }
            }
        }
        __syncwarp(); // the children rows and the list written above are visible to the whole warp
        // Pass 2 — mutation (drone/drone.py:167-172), 32 listed genes at a time: one Philox block per mutated
        // gene for its Box-Muller N(0, sigma) draw, added to the child gene.
        for (int i = (int)lane; i < n_mut; i += 32) {
            const int item = list[i], g = item >> 1, w = item & 1;
            const U4 x = philox4x32_10((uint32_t)(c + w), (uint32_t)g, kStreamNoise, gen, k0, k1);
            const double u1 = ((double)x.v[0] + 0.5) * 0x1.0p-32, u2 = ((double)x.v[1] + 0.5) * 0x1.0p-32;
            const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
            double* dst = (w ? cb : ca) + g;
            *dst = *dst + ga.mutation_sigma * z;
        }
        __syncwarp(); // the list is reused by this warp's next pair
    }
}

// ---- single calls of Drone.measure / Drone.translate_input_to_thrust (drone/drone.py:83-86,129-165) ----
// For code that steps a Drone by hand instead of calling fly().  One thread per drone, brain read
// straight from global memory; same device functions as the episode kernel, so stepping by hand
// reproduces fly() bit for bit.  state rows: pos vel angle angle_vel acc_current acc_last thrust.
template <int S, int E = 0>
__device__ __forceinline__ void mlp_global_block(Mlp& m, const double* __restrict__ genome) {
    if constexpr (E < 16 && S + E < LD_GENOME_LEN) {
        mlp_mac<S + E>(m, __ldg(genome + stored_to_genome(S + E)));
        mlp_global_block<S, E + 1>(m, genome);
    }
}

template <int S = 0>
__device__ __forceinline__ void mlp_global(Mlp& m, const double* __restrict__ genome) {
    if constexpr (S < LD_GENOME_LEN) {
        mlp_global_block<S>(m, genome);
        mlp_global<S + 16>(m, genome);
    }
}

template <bool EXTRAS>
__global__ void __launch_bounds__(128)
drone_step_kernel(const __grid_constant__ ld_params p, int what, const double* __restrict__ genomes, double* __restrict__ state,
                  const double* __restrict__ coin, int* __restrict__ flags, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool exists = i < n;
    double* row = state + (exists ? i : 0) * LD_STATE_LEN;
    Drone d;
    drone_reset(p, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        d.x[k] = row[k];
        d.v[k] = row[3 + k];
        d.x[3 + k] = row[6 + k];
        d.v[3 + k] = row[9 + k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) d.acc[k] = row[12 + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) d.thrust[k] = row[24 + k];
    if (what == 1) { // Drone.measure(wind_speed)
        trig_update(d); // warp-converged: every thread of the block gets here
        double a_last[6];
        int u;
        const bool crashed = physics_step<EXTRAS>(p, d, 1.0 / (double)p.m_total32, a_last, u, p.dt);
        if (exists) {
            write_state(d, a_last, row);
            flags[2 * i] = crashed ? 1 : 0;
            flags[2 * i + 1] = crashed ? 0 : u; // the unstable rules are skipped on a crash (drone.py:150-154)
        }
    } else { // Drone.translate_input_to_thrust(coin_position)
        Mlp m;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            m.in[k] = d.x[k];
            m.in[3 + k] = d.v[k];
            m.in[6 + k] = d.x[3 + k];
            m.in[9 + k] = d.v[3 + k];
            m.in[12 + k] = coin[(exists ? i : 0) * 3 + k];
        }
        mlp_global(m, genomes + (exists ? i : 0) * LD_GENOME_LEN);
        if (exists) {
#pragma unroll
            for (int k = 0; k < kOut; ++k) row[24 + k] = sigmoid_fast(m.out[k], -p.sigmoid_beta);
        }
    }
}

// ---- accuracy tap for ld_math.cuh (tests/test_gpu_math.py) --------------------------------------
__global__ void debug_math_kernel(int fn, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ o0,
                                  double* __restrict__ o1, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (fn == 0) {
            double s, c;
            sincos_fast(x[i], s, c);
            o0[i] = s;
            o1[i] = c;
        } else if (fn == 1) {
            o0[i] = exp_fast(x[i]);
        } else if (fn == 2) {
            o0[i] = sigmoid_fast(x[i], -0.4);
        } else if (fn == 3) {
            const double rc = recip(y[i]);
            o0[i] = div_by(x[i], y[i], rc);
            o1[i] = rc;
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

// evolve/evolution.py:77-83 on the device (one block): keep the best drone seen so far
__global__ void __launch_bounds__(320) track_best_kernel(const long long* __restrict__ stats, const double* __restrict__ population,
                                                         const __grid_constant__ ld_peer_table peers, long long generation,
                                                         long long* __restrict__ best2, double* __restrict__ best_genome,
                                                         long long* __restrict__ history7) {
    const long long score = stats[1], index = stats[2];
    const bool improved = best2[1] == 0 || score > best2[0];
    if (improved && threadIdx.x < LD_GENOME_LEN) {
        const double* row = population ? population + index * LD_GENOME_LEN
                                       : peers.shard[index / peers.shard_rows] + (index % peers.shard_rows) * LD_GENOME_LEN;
        best_genome[threadIdx.x] = row[threadIdx.x];
    }
    __syncthreads(); // every thread has read best2 before it changes
    if (threadIdx.x == 0) {
        if (improved) {
            best2[0] = score;
            best2[1] = 1;
        }
        if (history7) {
            history7[0] = generation;
            history7[1] = score;
            history7[2] = index;
            history7[3] = improved ? score : best2[0];
            history7[4] = stats[0];
            history7[5] = stats[3];
            history7[6] = stats[4];
        }
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
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<false, kPlain, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
#ifdef LD_LAYOUT5
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<false, kPlain, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, CtaLay<5>::kSmemBytes));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<true, kPlain, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, CtaLay<5>::kSmemBytes));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<false, kState, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, CtaLay<5>::kSmemBytes));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<true, kState, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, CtaLay<5>::kSmemBytes));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f64_kernel<false, kPlain, true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, CtaLay<5>::kSmemBytes));
#endif
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f32_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesF32));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f32_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesF32));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f32_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesF32));
    LD_CUDA(cudaFuncSetAttribute(fly_mlp_f32_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesF32));
    // the segment queue / parking area of every launch comes from the device's default stream-ordered pool:
    // let it keep freed blocks instead of returning them to the driver at every synchronisation
    cudaMemPool_t pool;
    LD_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ull;
    LD_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    g_dev.device = dev;
    g_dev.sms = prop.multiProcessorCount;
    g_dev.attrs_set = true;
    return LD_OK;
}

// CTA layout of the fp64 episode kernel: 4 warps per SM.  The 5-warp layout (top of this file) is a measured dead end
// that stays reproducible: build with -DLD_LAYOUT5 and run with LD_WARPS_PER_SM=5 (scripts/layout_probe.py).
int layout_warps() {
#ifdef LD_LAYOUT5
    static const int v = [] {
        const char* e = getenv("LD_WARPS_PER_SM");
        return (e && atoi(e) == 5) ? 5 : 4;
    }();
    return v;
#else
    return kWarpsPerCta;
#endif
}

bool has_extras(const ld_params* p) {
    for (int i = 0; i < 3; ++i)
        if (p->wind_cw32[i] != 0.0f || p->moments_payload[i] != 0.0) return true;
    return false;
}

int check_params(const ld_params* p) {
    if (!p) return fail(LD_ERR_INVALID_ARG, "params is NULL");
    if (p->timesteps < 0) return fail(LD_ERR_INVALID_ARG, "params.timesteps < 0");
    if (p->timesteps > (1 << 27)) return fail(LD_ERR_INVALID_ARG, "params.timesteps > 2^27 (the running reward is kept in 32 bits)");
    if (!(p->dt > 0.0)) return fail(LD_ERR_INVALID_ARG, "params.dt must be > 0");
    if (!(p->m_total32 > 0.0f)) return fail(LD_ERR_INVALID_ARG, "params.m_total32 must be > 0");
    return LD_OK;
}

// ld_params.slice_steps == 0: the automatic schedule.  Segments pay when drones outnumber the lanes by a small
// factor W: with few waves the end of the launch is ragged (whole-episode quantisation for survivors, a tail of
// late-started long fliers for mixed populations); beyond ~16 waves that tail is a few percent and the
// hand-overs (~15 us each) cost more than they bring.  The schedule doubles: boundaries at H/4 and 3H/4, so a
// drone is handed over at most twice, short-lived drones free their lane after a short first segment, every
// drone gets started early, and the last segment (H/4) bounds the raggedness of the end.  Measured on B200
// (65,536 drones, kernel ms unsegmented -> this schedule): all survivors 7.79 -> 6.97, 90 % 8.58 -> 6.49,
// 50 % 5.59 -> 4.30, 25 % 4.82 -> 3.77.  Returns the first segment's length, 0 = no segments.
int auto_first_segment(long long n, long long lanes, int H) {
    constexpr int kMinSlice = 125;
    if (n <= lanes || n > 16 * lanes || H < 4 * kMinSlice) return 0;
    // a population of survivors that fills its last wave to within 1.5 % needs no help (131,072 drones = 6.92
    // waves: 13.70 ms unsegmented, 13.92 ms with the schedule); callers that know their population is mixed ask
    // for segments explicitly (GpuEvolution does)
    const long long waves = (n + lanes - 1) / lanes;
    if ((double)(waves * lanes) <= 1.015 * (double)n) return 0;
    return (H + 3) / 4;
}

// work_counter: int64[2] of device scratch — [0] queue head, [1] error flag (raised if a warp gave up waiting)
int launch_fly(const ld_params* p, const double* genomes, int64_t n, ld_record* records, double* final_state,
               int64_t* work_counter, double* trace, cudaStream_t s, bool f32 = false,
               const unsigned long long* arrived = nullptr, const double* init_state = nullptr, unsigned long long* prof = nullptr,
               unsigned long long* prof_times = nullptr) {
    int rc = ensure_device();
    if (rc) return rc;
    if (n == 0) return LD_OK;
    LD_CUDA(cudaMemsetAsync(work_counter, 0, 2 * sizeof(int64_t), s));
    if (f32) {
        long long ctas = (n + ldf::kThreads - 1) / ldf::kThreads;
        if (ctas > g_dev.sms) ctas = g_dev.sms;
        auto* head = reinterpret_cast<unsigned long long*>(work_counter);
        const bool ex = has_extras(p), st = final_state != nullptr;
#define LD_LAUNCH32(EX, ST) \
    fly_mlp_f32_kernel<EX, ST><<<(unsigned)ctas, ldf::kThreads, kSmemBytesF32, s>>>(*p, genomes, n, records, final_state, head)
        if (ex) { if (st) LD_LAUNCH32(true, true); else LD_LAUNCH32(true, false); }
        else { if (st) LD_LAUNCH32(false, true); else LD_LAUNCH32(false, false); }
#undef LD_LAUNCH32
        LD_CUDA(cudaGetLastError());
        return LD_OK;
    }
    const int warps = trace ? 4 : layout_warps(); // the trajectory tap flies one drone: default layout
    const int cta_threads = warps * 32;
    long long ctas = (n + cta_threads - 1) / cta_threads;
    if (ctas > g_dev.sms) ctas = g_dev.sms;
    dim3 grid((unsigned)ctas), block(cta_threads);
    const bool ex = has_extras(p);
    const int H = p->timesteps;
    // Episode segments (ld_params.slice_steps): only when drones outnumber the lanes, otherwise nobody ever
    // waits for a lane.  Queue + parking area come from the stream-ordered pool and go back after the kernel.
    Sched q;
    memset(&q, 0, sizeof(q));
    q.head = reinterpret_cast<unsigned long long*>(work_counter);
    q.slice = H > 1 ? H : 1;
    q.arrived = arrived;
    q.error = reinterpret_cast<int*>(work_counter + 1);
    q.init_state = init_state;
    q.prof = prof;
    q.prof_times = prof_times;
    // slice_steps > 0: equal segments; <= -2: segments that double in length, the first one -slice_steps long;
    // 0: automatic (doubling from H/4 when it pays); -1: off; LD_SLICE_STAGGER: staggered first wave
    // The staggered schedule needs every lane to start at the same moment: a host-buffer launch starts on an empty genome
    // matrix and its lanes start as their rows arrive (measured: 8.0 ms end to end against 7.7 ms with the automatic
    // schedule), so it falls back to the automatic one there.
    const int requested = (p->slice_steps == LD_SLICE_STAGGER && arrived != nullptr) ? 0 : p->slice_steps;
    const bool stagger = requested == LD_SLICE_STAGGER;
    const bool doubling = !stagger && (requested <= -2 || requested == 0);
    const int slice = stagger ? 0 : (requested == 0 ? auto_first_segment(n, ctas * cta_threads, H) : (requested <= -2 ? -requested : requested));
    q.growth = doubling ? 2 : 1;
    void* scratch = nullptr;
    const bool queue_ok = !trace && n > ctas * cta_threads && n < 0x7fffffffLL;
    if (queue_ok && stagger && H > 1) {
        q.cap = ctas * cta_threads; // only drones of the first wave are ever parked, once each
        q.stagger = 1;              // q.slice stays H: no other boundary
    } else if (queue_ok && slice > 0 && slice < H) {
        long long segments = ((long long)H + slice - 1) / slice;
        if (doubling) {
            segments = 1;
            for (long long b = slice; b < H; b = 2 * b + slice) ++segments;
        }
        q.cap = (segments - 1) * n;
        q.slice = slice;
    }
    if (q.cap > 0) {
        const size_t queue_bytes = (16 + (size_t)q.cap * sizeof(int) + 255) & ~(size_t)255;
        LD_CUDA(cudaMallocAsync(&scratch, queue_bytes + (size_t)n * kParkDoubles * sizeof(double), s));
        LD_CUDA(cudaMemsetAsync(scratch, 0, queue_bytes, s));
        q.tail = reinterpret_cast<unsigned long long*>(scratch);
        q.done = q.tail + 1;
        q.requeue = reinterpret_cast<int*>(q.tail + 2);
        q.park = reinterpret_cast<double*>(static_cast<char*>(scratch) + queue_bytes);
    }
#define LD_LAUNCH_W(EX, MD, PR, W) \
    fly_mlp_f64_kernel<EX, MD, PR, W><<<grid, block, CtaLay<W>::kSmemBytes, s>>>(*p, genomes, n, records, final_state, q, trace)
#define LD_LAUNCH(EX, MD) \
    LD_LAUNCH_W(EX, MD, false, 4)
    const int mode = trace ? kTrace : (final_state ? kState : kPlain);
    if (prof) {
        if (ex || mode != kPlain) return fail(LD_ERR_INVALID_ARG, "the profile build exists for the plain kernel without wind / payload / final state");
#ifdef LD_LAYOUT5
        if (warps == 5) LD_LAUNCH_W(false, kPlain, true, 5);
        else
#endif
            LD_LAUNCH_W(false, kPlain, true, 4);
#ifdef LD_LAYOUT5
    } else if (warps == 5) {
        if (ex) { if (mode == kState) LD_LAUNCH_W(true, kState, false, 5); else LD_LAUNCH_W(true, kPlain, false, 5); }
        else { if (mode == kState) LD_LAUNCH_W(false, kState, false, 5); else LD_LAUNCH_W(false, kPlain, false, 5); }
#endif
    } else if (ex) {
        if (mode == kTrace) LD_LAUNCH(true, kTrace);
        else if (mode == kState) LD_LAUNCH(true, kState);
        else LD_LAUNCH(true, kPlain);
    } else {
        if (mode == kTrace) LD_LAUNCH(false, kTrace);
        else if (mode == kState) LD_LAUNCH(false, kState);
        else LD_LAUNCH(false, kPlain);
    }
#undef LD_LAUNCH
#undef LD_LAUNCH_W
    cudaError_t launched = cudaGetLastError();
    if (scratch) LD_CUDA(cudaFreeAsync(scratch, s));
    if (launched != cudaSuccess) return fail(LD_ERR_CUDA, "fly_mlp_f64_kernel launch", launched);
    return LD_OK;
}

// grow-only device workspace for the *_host entry point (per host thread)
constexpr int kHostStreams = 8;
struct Workspace {
    int device = -1;
    void* genomes[kHostStreams] = {};
    void* records[kHostStreams] = {};
    void* state[kHostStreams] = {};
    int64_t* counters = nullptr;
    size_t cap_drones = 0;
    bool cap_state = false;
    cudaStream_t streams[kHostStreams] = {};
};
thread_local Workspace g_ws;

void ws_release() {
    for (int i = 0; i < kHostStreams; ++i) {
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
    for (int i = 0; i < kHostStreams; ++i) {
        LD_CUDA(cudaMalloc(&g_ws.genomes[i], chunk * LD_GENOME_LEN * sizeof(double)));
        LD_CUDA(cudaMalloc(&g_ws.records[i], chunk * sizeof(ld_record)));
        if (want_state) LD_CUDA(cudaMalloc(&g_ws.state[i], chunk * LD_STATE_LEN * sizeof(double)));
        LD_CUDA(cudaStreamCreateWithFlags(&g_ws.streams[i], cudaStreamNonBlocking));
    }
    LD_CUDA(cudaMalloc((void**)&g_ws.counters, 2 * kHostStreams * sizeof(int64_t)));
    g_ws.device = dev;
    g_ws.cap_drones = chunk;
    g_ws.cap_state = want_state;
    return LD_OK;
}

// ---- fp64 host-buffer path: ONE persistent launch, rows streamed in behind it ---------------------
// The kernel starts on an empty genome matrix; the copy engine delivers it in chunks and raises
// `arrived` after each one (an 8-byte copy from a pinned table, stream-ordered behind the chunk).  Lanes
// whose drone is not there yet wait in the queue logic of the kernel.  Compared with chunk-sized
// launches this keeps every SM in one work queue, so segments and refills balance the whole population.
constexpr int64_t kHostBatchMax = 4ll << 20; // drones per launch (10.5 GB of genomes)
constexpr int kHostChunksMax = 4096;
struct StreamWs {
    int device = -1;
    double* genomes = nullptr;
    ld_record* records = nullptr;
    double* state = nullptr;
    double* init = nullptr;
    unsigned long long* ctl = nullptr;      // [0] arrived, [1] queue head, [2] error flag
    unsigned long long* h_counts = nullptr; // pinned [kHostChunksMax]
    int* h_error = nullptr;                 // pinned
    size_t cap_drones = 0, cap_state = 0, cap_init = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t armed = nullptr;
    bool in_flight = false;  // a submission has been enqueued and not yet waited for
    long long sequence = 0;  // ticket check of the pipelined entry points
};
// Two workspaces per host thread: while one submission's kernel runs, the next one's genome rows stream into the
// other (ld_fly_mlp_f64_host_submit / _wait, and the batches of a population larger than one launch).
constexpr int kHostPipelineDepth = 2;
thread_local StreamWs g_sws_slots[kHostPipelineDepth];
static_assert(kHostPipelineDepth == LD_HOST_PIPELINE_DEPTH, "header constant");

void sws_release_slot(StreamWs& w) {
    if (w.genomes) cudaFree(w.genomes);
    if (w.records) cudaFree(w.records);
    if (w.state) cudaFree(w.state);
    if (w.init) cudaFree(w.init);
    if (w.ctl) cudaFree(w.ctl);
    if (w.h_counts) cudaFreeHost(w.h_counts);
    if (w.h_error) cudaFreeHost(w.h_error);
    if (w.compute) cudaStreamDestroy(w.compute);
    if (w.copy) cudaStreamDestroy(w.copy);
    if (w.armed) cudaEventDestroy(w.armed);
    w = StreamWs();
}

void sws_release() {
    for (int i = 0; i < kHostPipelineDepth; ++i) sws_release_slot(g_sws_slots[i]);
}

int sws_reserve(StreamWs& w, size_t drones, bool want_state, bool want_init = false) {
    int dev = 0;
    LD_CUDA(cudaGetDevice(&dev));
    if (w.device != dev) {
        sws_release_slot(w);
        LD_CUDA(cudaMalloc((void**)&w.ctl, 4 * sizeof(unsigned long long)));
        LD_CUDA(cudaHostAlloc((void**)&w.h_counts, kHostChunksMax * sizeof(unsigned long long), cudaHostAllocDefault));
        LD_CUDA(cudaHostAlloc((void**)&w.h_error, sizeof(int), cudaHostAllocDefault));
        LD_CUDA(cudaStreamCreateWithFlags(&w.compute, cudaStreamNonBlocking));
        LD_CUDA(cudaStreamCreateWithFlags(&w.copy, cudaStreamNonBlocking));
        LD_CUDA(cudaEventCreateWithFlags(&w.armed, cudaEventDisableTiming));
        w.device = dev;
    }
    if (w.cap_drones < drones) {
        if (w.genomes) cudaFree(w.genomes);
        if (w.records) cudaFree(w.records);
        if (w.state) cudaFree(w.state);
        if (w.init) cudaFree(w.init);
        w.genomes = nullptr; w.records = nullptr; w.state = nullptr; w.init = nullptr;
        w.cap_drones = w.cap_state = w.cap_init = 0;
        LD_CUDA(cudaMalloc((void**)&w.genomes, drones * LD_GENOME_LEN * sizeof(double)));
        LD_CUDA(cudaMalloc((void**)&w.records, drones * sizeof(ld_record)));
        w.cap_drones = drones;
    }
    if (want_state && w.cap_state < drones) {
        if (w.state) cudaFree(w.state);
        w.state = nullptr; w.cap_state = 0;
        LD_CUDA(cudaMalloc((void**)&w.state, w.cap_drones * LD_STATE_LEN * sizeof(double)));
        w.cap_state = w.cap_drones;
    }
    if (want_init && w.cap_init < drones) {
        if (w.init) cudaFree(w.init);
        w.init = nullptr; w.cap_init = 0;
        LD_CUDA(cudaMalloc((void**)&w.init, w.cap_drones * LD_STATE_LEN * sizeof(double)));
        w.cap_init = w.cap_drones;
    }
    return LD_OK;
}

// the blocking and the secondary *_host entry points work on slot 0
int sws_reserve(size_t drones, bool want_state, bool want_init = false) { return sws_reserve(g_sws_slots[0], drones, want_state, want_init); }

// One launch over rows [0, m) of the host matrix `genomes`, in two halves: `enqueue` queues every copy and the
// kernel on the workspace's streams and returns; `finish` waits for the records to be in the caller's buffers.
int fly_streamed_enqueue(StreamWs& w, const ld_params* params, const double* genomes, const double* init_state, int64_t m,
                         ld_record* records, double* final_state) {
    if (w.in_flight) return fail(LD_ERR_INVALID_ARG, "host workspace busy: wait for the submission in flight first");
    const int64_t wave = (int64_t)g_dev.sms * layout_warps() * 32;
    // chunks: eighths of a wave (2368 drones = 5.9 MB, the first lanes start after ~0.1 ms) until the first
    // wave is in, then half waves (every chunk costs one extra 8-byte copy on the copy engine)
    const int64_t small = wave / 8 > 1024 ? wave / 8 : 1024;
    int64_t big = 4 * small;
    if ((m + big - 1) / big + 8 > kHostChunksMax) big = (m + kHostChunksMax - 9) / (kHostChunksMax - 8);
    // Every copy is queued BEFORE the kernel is launched: the copy engine works through them on its own, so
    // the kernel can never be left waiting for a copy the host has not issued yet (tools that make launches
    // synchronous — profilers, CUDA_LAUNCH_BLOCKING — would otherwise stall it until kWaitLimitNs).
    LD_CUDA(cudaMemsetAsync(w.ctl, 0, 4 * sizeof(unsigned long long), w.copy));
    if (init_state) // start rows are small (224 B per drone): all of them go first
        LD_CUDA(cudaMemcpyAsync(w.init, init_state, (size_t)m * LD_STATE_LEN * sizeof(double), cudaMemcpyHostToDevice, w.copy));
    LD_CUDA(cudaEventRecord(w.armed, w.copy));
    LD_CUDA(cudaStreamWaitEvent(w.compute, w.armed, 0)); // the kernel only waits for the counters to be reset
    int k = 0;
    for (int64_t off = 0; off < m; ++k) {
        const int64_t chunk = off < wave ? small : big;
        const int64_t c = (m - off < chunk) ? (m - off) : chunk;
        LD_CUDA(cudaMemcpyAsync(w.genomes + off * LD_GENOME_LEN, genomes + off * LD_GENOME_LEN, (size_t)c * LD_GENOME_LEN * sizeof(double),
                                cudaMemcpyHostToDevice, w.copy));
        off += c;
        w.h_counts[k] = (unsigned long long)off;
        LD_CUDA(cudaMemcpyAsync(w.ctl, w.h_counts + k, sizeof(unsigned long long), cudaMemcpyHostToDevice, w.copy));
    }
    w.in_flight = true; // from here on `finish` must run before the caller's buffers may go away
    int rc = launch_fly(params, w.genomes, m, w.records, final_state ? w.state : nullptr, reinterpret_cast<int64_t*>(w.ctl + 1),
                        nullptr, w.compute, false, w.ctl, init_state ? w.init : nullptr);
    if (rc) return rc;
    LD_CUDA(cudaMemcpyAsync(records, w.records, (size_t)m * sizeof(ld_record), cudaMemcpyDeviceToHost, w.compute));
    if (final_state)
        LD_CUDA(cudaMemcpyAsync(final_state, w.state, (size_t)m * LD_STATE_LEN * sizeof(double), cudaMemcpyDeviceToHost, w.compute));
    LD_CUDA(cudaMemcpyAsync(w.h_error, w.ctl + 2, sizeof(int), cudaMemcpyDeviceToHost, w.compute));
    return LD_OK;
}

int fly_streamed_finish(StreamWs& w) {
    if (!w.in_flight) return LD_OK;
    const cudaError_t e1 = cudaStreamSynchronize(w.copy), e2 = cudaStreamSynchronize(w.compute);
    w.in_flight = false;
    if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(LD_ERR_CUDA, "host-buffer episode launch", e1 != cudaSuccess ? e1 : e2);
    if (*w.h_error) {
        snprintf(g_err, sizeof(g_err), "episode kernel error word %d (1 = gave up waiting for its work queue; >= 2 = bounds check of a checked build)",
                 *w.h_error);
        return LD_ERR_CUDA;
    }
    return LD_OK;
}

// blocking form; on an enqueue error nothing may still read the caller's buffers when we return
int fly_streamed_batch(StreamWs& w, const ld_params* params, const double* genomes, const double* init_state, int64_t m,
                       ld_record* records, double* final_state) {
    const int rc = fly_streamed_enqueue(w, params, genomes, init_state, m, records, final_state);
    if (rc) {
        if (w.in_flight) {
            cudaStreamSynchronize(w.copy);
            cudaStreamSynchronize(w.compute);
            w.in_flight = false;
        }
        return rc;
    }
    return fly_streamed_finish(w);
}

// grow-only device scratch of the secondary *_host entry points (per host thread): no allocation per call
constexpr int kScratchSlots = 6;
struct Scratch {
    int device = -1;
    void* buf[kScratchSlots] = {};
    size_t cap[kScratchSlots] = {};
};
thread_local Scratch g_scratch;

void scratch_release() {
    for (int i = 0; i < kScratchSlots; ++i) {
        if (g_scratch.buf[i]) cudaFree(g_scratch.buf[i]);
        g_scratch.buf[i] = nullptr;
        g_scratch.cap[i] = 0;
    }
    g_scratch.device = -1;
}

int scratch_get(int slot, size_t bytes, void** out) {
    int dev = 0;
    LD_CUDA(cudaGetDevice(&dev));
    if (g_scratch.device != dev) {
        scratch_release();
        g_scratch.device = dev;
    }
    if (g_scratch.cap[slot] < bytes) {
        if (g_scratch.buf[slot]) cudaFree(g_scratch.buf[slot]);
        g_scratch.buf[slot] = nullptr;
        g_scratch.cap[slot] = 0;
        size_t want = bytes < 4096 ? 4096 : bytes + bytes / 4;
        LD_CUDA(cudaMalloc(&g_scratch.buf[slot], want));
        g_scratch.cap[slot] = want;
    }
    *out = g_scratch.buf[slot];
    return LD_OK;
}

// ---- generic brain forward (evolve/brain.py:41-54 for ANY layer_spec) and the two activations ----------------
// One thread per input row; same arithmetic as the episode kernel's mlp_mac (k-ordered fma chain that starts
// with the k = 0 product and ends with `+ bias`, ReLU on the sign bit, sigmoid_fast), so a 15-10-10-4 brain gives
// the episode kernel's thrust bit for bit.
constexpr int kBrainMaxLayers = 8, kBrainMaxWidth = 64;
struct BrainSpec {
    int n_layers;
    int size[kBrainMaxLayers + 1]; // size[0] = inputs
};

__global__ void __launch_bounds__(128)
brain_forward_kernel(const __grid_constant__ BrainSpec spec, const double* __restrict__ weights, long long weights_stride,
                     const double* __restrict__ inputs, long long n, double* __restrict__ outputs, double neg_beta) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x[kBrainMaxWidth], y[kBrainMaxWidth];
    const double* w = weights + i * weights_stride;
    for (int k = 0; k < spec.size[0]; ++k) x[k] = inputs[i * spec.size[0] + k];
    for (int l = 0; l < spec.n_layers; ++l) {
        const int n_in = spec.size[l], n_out = spec.size[l + 1];
        const bool last = l == spec.n_layers - 1;
        for (int j = 0; j < n_out; ++j) {
            const double* row = w + (long long)j * (n_in + 1);
            double acc = row[0] * x[0];
            for (int k = 1; k < n_in; ++k) acc = fma(row[k], x[k], acc);
            acc = acc + row[n_in];
            y[j] = last ? sigmoid_fast(acc, neg_beta) : relu(acc);
        }
        for (int j = 0; j < n_out; ++j) x[j] = y[j];
        w += (long long)n_out * (n_in + 1);
    }
    const int n_last = spec.size[spec.n_layers];
    for (int j = 0; j < n_last; ++j) outputs[i * n_last + j] = x[j];
}

// kind 0: NeuralNetwork._relu (brain.py:26-30) `0.0 if x < a else b * x`; kind 1: sigmoid (brain.py:37-39), beta = a
__global__ void activation_kernel(int kind, const double* __restrict__ x, long long n, double* __restrict__ out, double a, double b) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = kind == 0 ? (x[i] < a ? 0.0 : b * x[i]) : sigmoid_fast(x[i], -a);
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
    return (int64_t)g_dev.sms * layout_warps() * 32;
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

int ld_debug_fly_profile_device(const ld_params* params, const double* genomes, int64_t n, ld_record* records, int64_t* work_counter,
                                uint64_t* prof, uint64_t* times, void* stream) {
    int rc = check_params(params);
    if (rc) return rc;
    if (n <= 0 || !genomes || !records || !work_counter || !prof) return fail(LD_ERR_INVALID_ARG, "bad arguments");
    LD_CUDA(cudaMemsetAsync(prof, 0, LD_PROFILE_SLOTS * sizeof(uint64_t), (cudaStream_t)stream));
    return launch_fly(params, genomes, n, records, nullptr, work_counter, nullptr, (cudaStream_t)stream, false, nullptr, nullptr,
                      reinterpret_cast<unsigned long long*>(prof), reinterpret_cast<unsigned long long*>(times));
}

int ld_fly_mlp_f64_trace_device(const ld_params* params, const double* genome, ld_record* record, double* trace,
                                void* stream) {
    int rc = check_params(params);
    if (rc) return rc;
    if (!genome || !record || !trace) return fail(LD_ERR_INVALID_ARG, "NULL device buffer");
    rc = ensure_device();
    if (rc) return rc;
    int64_t* counter = nullptr;
    rc = scratch_get(0, 2 * sizeof(int64_t), (void**)&counter);
    if (rc) return rc;
    rc = launch_fly(params, genome, 1, record, nullptr, counter, trace, (cudaStream_t)stream);
    if (rc) return rc;
    int err = 0;
    cudaError_t e = cudaMemcpyAsync(&err, counter + 1, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return fail(LD_ERR_CUDA, "trace kernel", e);
    if (err) return fail(LD_ERR_CUDA, "episode kernel gave up waiting for its work queue");
    return LD_OK;
}

static int fly_mlp_host_impl(const ld_params* params, const double* genomes, const double* init_state, int64_t n, ld_record* records,
                             double* final_state, bool f32) {
    int rc = check_params(params);
    if (rc) return rc;
    if (n < 0) return fail(LD_ERR_INVALID_ARG, "n < 0");
    if (n == 0) return LD_OK;
    if (!genomes || !records) return fail(LD_ERR_INVALID_ARG, "NULL host buffer");
    rc = ensure_device();
    if (rc) return rc;
    if (!f32) {
        const int64_t batch = n < kHostBatchMax ? n : kHostBatchMax;
        for (int i = 0; i < kHostPipelineDepth; ++i)
            if (g_sws_slots[i].in_flight) return fail(LD_ERR_INVALID_ARG, "a pipelined submission is in flight on this thread: wait for it first");
        if (n <= batch) {
            rc = sws_reserve((size_t)batch, final_state != nullptr, init_state != nullptr);
            if (rc) return rc;
            return fly_streamed_batch(g_sws_slots[0], params, genomes, init_state, n, records, final_state);
        }
        // a population larger than one launch: batches alternate between the two workspaces, so batch k+1's rows
        // stream in while batch k flies
        int slot = 0;
        for (int64_t off = 0; off < n && rc == LD_OK; off += batch, slot ^= 1) {
            StreamWs& w = g_sws_slots[slot];
            rc = fly_streamed_finish(w); // the batch that used this workspace two rounds ago
            if (rc == LD_OK) rc = sws_reserve(w, (size_t)batch, final_state != nullptr, init_state != nullptr);
            if (rc == LD_OK)
                rc = fly_streamed_enqueue(w, params, genomes + off * LD_GENOME_LEN, init_state ? init_state + off * LD_STATE_LEN : nullptr,
                                          (n - off < batch) ? (n - off) : batch, records + off,
                                          final_state ? final_state + off * LD_STATE_LEN : nullptr);
        }
        for (int i = 0; i < kHostPipelineDepth; ++i) {
            StreamWs& w = g_sws_slots[i];
            if (w.in_flight && rc != LD_OK) { // error path: drain without overwriting the first error
                cudaStreamSynchronize(w.copy);
                cudaStreamSynchronize(w.compute);
                w.in_flight = false;
            } else if (rc == LD_OK) {
                rc = fly_streamed_finish(w);
            }
        }
        return rc;
    }
    // fp32 mode: pipeline over kHostStreams streams / buffers.  A chunk is a quarter of a wave (37 of 148 SMs):
    // kernels of different chunks are resident on disjoint SMs at the same time (a CTA owns its SM's
    // Tensor Memory), so the machine fills up after four small H2D copies instead of one big one
    // and every later copy / record read-back hides behind the kernels.  Big populations use
    // whole-wave chunks (fewer launches).
    const int64_t cta = f32 ? ldf::kThreads : kCtaThreads;
    const int64_t wave = (int64_t)g_dev.sms * cta;
    const int div = 4; // measured on B200 with 8 streams: quarter-wave chunks give the best end-to-end rate
    int64_t chunk = (n > 16 * wave) ? wave : ((int64_t)((g_dev.sms + div - 1) / div) * cta);
    if (chunk > n) chunk = n;
    rc = ws_reserve((size_t)chunk, final_state != nullptr);
    if (rc) return rc;
    int b = 0;
    for (int64_t off = 0; off < n; b = (b + 1) % kHostStreams) {
        const int64_t m = (n - off < chunk) ? (n - off) : chunk;
        cudaStream_t s = g_ws.streams[b];
        LD_CUDA(cudaMemcpyAsync(g_ws.genomes[b], genomes + off * LD_GENOME_LEN, (size_t)m * LD_GENOME_LEN * sizeof(double),
                                cudaMemcpyHostToDevice, s));
        rc = launch_fly(params, (const double*)g_ws.genomes[b], m, (ld_record*)g_ws.records[b],
                        final_state ? (double*)g_ws.state[b] : nullptr, g_ws.counters + 2 * b, nullptr, s, f32);
        if (rc) return rc;
        LD_CUDA(cudaMemcpyAsync(records + off, g_ws.records[b], (size_t)m * sizeof(ld_record), cudaMemcpyDeviceToHost, s));
        if (final_state)
            LD_CUDA(cudaMemcpyAsync(final_state + off * LD_STATE_LEN, g_ws.state[b], (size_t)m * LD_STATE_LEN * sizeof(double),
                                    cudaMemcpyDeviceToHost, s));
        off += m;
    }
    for (int i = 0; i < kHostStreams; ++i) LD_CUDA(cudaStreamSynchronize(g_ws.streams[i]));
    return LD_OK;
}

int ld_fly_mlp_f64_host(const ld_params* params, const double* genomes, int64_t n, ld_record* records, double* final_state) {
    return fly_mlp_host_impl(params, genomes, nullptr, n, records, final_state, false);
}

int ld_fly_mlp_f64_from_host(const ld_params* params, const double* genomes, const double* init_state, int64_t n, ld_record* records,
                             double* final_state) {
    return fly_mlp_host_impl(params, genomes, init_state, n, records, final_state, false);
}

int ld_fly_mlp_f64_host_submit(const ld_params* params, const double* genomes, const double* init_state, int64_t n, ld_record* records,
                               double* final_state, int64_t* ticket) {
    int rc = check_params(params);
    if (rc) return rc;
    if (!ticket) return fail(LD_ERR_INVALID_ARG, "NULL ticket");
    if (n <= 0 || n > kHostBatchMax) return fail(LD_ERR_INVALID_ARG, "submit takes 1 .. 4 Mi drones per submission");
    if (!genomes || !records) return fail(LD_ERR_INVALID_ARG, "NULL host buffer");
    rc = ensure_device();
    if (rc) return rc;
    int slot = -1;
    for (int i = 0; i < kHostPipelineDepth && slot < 0; ++i)
        if (!g_sws_slots[i].in_flight) slot = i;
    if (slot < 0) return fail(LD_ERR_INVALID_ARG, "pipeline full: LD_HOST_PIPELINE_DEPTH submissions are in flight, wait for one");
    StreamWs& w = g_sws_slots[slot];
    rc = sws_reserve(w, (size_t)n, final_state != nullptr, init_state != nullptr);
    if (rc) return rc;
    rc = fly_streamed_enqueue(w, params, genomes, init_state, n, records, final_state);
    if (rc) {
        if (w.in_flight) {
            cudaStreamSynchronize(w.copy);
            cudaStreamSynchronize(w.compute);
            w.in_flight = false;
        }
        return rc;
    }
    *ticket = (++w.sequence << 8) | slot;
    return LD_OK;
}

int ld_fly_mlp_f64_host_wait(int64_t ticket) {
    const int slot = (int)(ticket & 0xff);
    if (slot < 0 || slot >= kHostPipelineDepth || g_sws_slots[slot].sequence != (ticket >> 8) || !g_sws_slots[slot].in_flight)
        return fail(LD_ERR_INVALID_ARG, "unknown or already completed ticket (tickets belong to the submitting host thread)");
    return fly_streamed_finish(g_sws_slots[slot]);
}

int ld_fly_mlp_f64_from_device(const ld_params* params, const double* genomes, const double* init_state, int64_t n, ld_record* records,
                               double* final_state, int64_t* work_counter, void* stream) {
    int rc = check_params(params);
    if (rc) return rc;
    if (n < 0) return fail(LD_ERR_INVALID_ARG, "n < 0");
    if (n > 0 && (!genomes || !records || !work_counter)) return fail(LD_ERR_INVALID_ARG, "NULL device buffer");
    if ((reinterpret_cast<uintptr_t>(genomes) & 15u) || (reinterpret_cast<uintptr_t>(records) & 15u))
        return fail(LD_ERR_INVALID_ARG, "genomes and records must be 16-byte aligned");
    return launch_fly(params, genomes, n, records, final_state, work_counter, nullptr, (cudaStream_t)stream, false, nullptr, init_state);
}

int ld_fly_mlp_f32_host(const ld_params* params, const double* genomes, int64_t n, ld_record* records, double* final_state) {
    return fly_mlp_host_impl(params, genomes, nullptr, n, records, final_state, true);
}

int ld_fly_mlp_f32_device(const ld_params* params, const double* genomes, int64_t n, ld_record* records,
                          double* final_state, int64_t* work_counter, void* stream) {
    int rc = check_params(params);
    if (rc) return rc;
    if (n < 0) return fail(LD_ERR_INVALID_ARG, "n < 0");
    if (n > 0 && (!genomes || !records || !work_counter)) return fail(LD_ERR_INVALID_ARG, "NULL device buffer");
    if (reinterpret_cast<uintptr_t>(records) & 15u) return fail(LD_ERR_INVALID_ARG, "records must be 16-byte aligned");
    return launch_fly(params, genomes, n, records, final_state, work_counter, nullptr, (cudaStream_t)stream, true);
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

int ld_track_best_device(const int64_t* stats5, const double* population, const ld_peer_table* peers, int64_t generation,
                         int64_t* best2, double* best_genome, int64_t* history7, void* stream) {
    if (!stats5 || !best2 || !best_genome || (!population && !peers)) return fail(LD_ERR_INVALID_ARG, "NULL argument");
    ld_peer_table table;
    memset(&table, 0, sizeof(table));
    if (!population) {
        if (peers->world < 1 || peers->world > LD_MAX_PEERS || peers->shard_rows < 1) return fail(LD_ERR_INVALID_ARG, "bad peer table");
        for (int r = 0; r < peers->world; ++r)
            if (!peers->shard[r]) return fail(LD_ERR_INVALID_ARG, "peer table: NULL shard pointer");
        table = *peers;
    }
    int rc = ensure_device();
    if (rc) return rc;
    track_best_kernel<<<1, 320, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(stats5), population, table, generation,
                                                          reinterpret_cast<long long*>(best2), best_genome,
                                                          reinterpret_cast<long long*>(history7));
    LD_CUDA(cudaGetLastError());
    return LD_OK;
}

int ld_fly_pid_f64_device(const ld_params* params, const ld_pid_params* pid, const double* pid_gains, int64_t n,
                          ld_record* records, double* final_state, void* stream) {
    int rc = check_params(params);
    if (rc) return rc;
    if (!pid || pid->n_coins < 1 || pid->n_coins > LD_PID_MAX_COINS || pid->timesteps < 0)
        return fail(LD_ERR_INVALID_ARG, "bad ld_pid_params");
    if (n < 0) return fail(LD_ERR_INVALID_ARG, "n < 0");
    if (n == 0) return LD_OK;
    if (!pid_gains || !records) return fail(LD_ERR_INVALID_ARG, "NULL device buffer");
    if (reinterpret_cast<uintptr_t>(records) & 15u) return fail(LD_ERR_INVALID_ARG, "records must be 16-byte aligned");
    rc = ensure_device();
    if (rc) return rc;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (has_extras(params))
        fly_pid_f64_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(*params, *pid, pid_gains, n, records, final_state);
    else
        fly_pid_f64_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(*params, *pid, pid_gains, n, records, final_state);
    LD_CUDA(cudaGetLastError());
    return LD_OK;
}

int ld_fly_pid_f64_host(const ld_params* params, const ld_pid_params* pid, const double* pid_gains, int64_t n,
                        ld_record* records, double* final_state) {
    int rc = check_params(params);
    if (rc) return rc;
    if (!pid || pid->n_coins < 1 || pid->n_coins > LD_PID_MAX_COINS || pid->timesteps < 0)
        return fail(LD_ERR_INVALID_ARG, "bad ld_pid_params");
    if (n < 0) return fail(LD_ERR_INVALID_ARG, "n < 0");
    if (n == 0) return LD_OK;
    if (!pid_gains || !records) return fail(LD_ERR_INVALID_ARG, "NULL host buffer");
    rc = ensure_device();
    if (rc) return rc;
    // validated: now the (grow-only, reused) device workspace and the work stream
    rc = sws_reserve(1, false);
    if (rc) return rc;
    double *d_gains = nullptr, *d_state = nullptr;
    ld_record* d_rec = nullptr;
    if ((rc = scratch_get(1, (size_t)n * LD_PID_GAINS * sizeof(double), (void**)&d_gains))) return rc;
    if ((rc = scratch_get(2, (size_t)n * sizeof(ld_record), (void**)&d_rec))) return rc;
    if (final_state && (rc = scratch_get(3, (size_t)n * LD_STATE_LEN * sizeof(double), (void**)&d_state))) return rc;
    cudaStream_t s = g_sws_slots[0].compute;
    cudaError_t e = cudaMemcpyAsync(d_gains, pid_gains, (size_t)n * LD_PID_GAINS * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) rc = ld_fly_pid_f64_device(params, pid, d_gains, n, d_rec, d_state, s);
    if (e == cudaSuccess && rc == LD_OK) e = cudaMemcpyAsync(records, d_rec, (size_t)n * sizeof(ld_record), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && rc == LD_OK && final_state)
        e = cudaMemcpyAsync(final_state, d_state, (size_t)n * LD_STATE_LEN * sizeof(double), cudaMemcpyDeviceToHost, s);
    const cudaError_t e_sync = cudaStreamSynchronize(s); // also on failure: nothing may still read the caller's buffers
    if (rc) return rc;
    if (e != cudaSuccess || e_sync != cudaSuccess) return fail(LD_ERR_CUDA, "ld_fly_pid_f64_host", e != cudaSuccess ? e : e_sync);
    return LD_OK;
}

int ld_ga_step_device(const ld_ga_params* ga, const double* parents, const ld_record* records, int64_t n, double* children,
                      int64_t lo, int64_t hi, void* stream) {
    if (!ga || !parents || !records || !children) return fail(LD_ERR_INVALID_ARG, "NULL argument");
    if (n < 2 || (n & 1)) return fail(LD_ERR_INVALID_ARG, "population size must be even and >= 2 (evolve/util.py:14-24 pairs slots)");
    if (lo < 0 || hi > n || lo > hi || (lo & 1) || (hi & 1)) return fail(LD_ERR_INVALID_ARG, "children range [lo, hi) must be even-aligned inside [0, n]");
    if (lo == hi) return LD_OK;
    int rc = ensure_device();
    if (rc) return rc;
    long long pairs = (hi - lo) / 2;
    long long blocks = (pairs + 7) / 8;
    if (blocks > 16LL * g_dev.sms) blocks = 16LL * g_dev.sms;
    ld_peer_table none;
    memset(&none, 0, sizeof(none));
    ga_step_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*ga, parents, none, records, n, children, lo, hi);
    LD_CUDA(cudaGetLastError());
    return LD_OK;
}

int ld_ga_step_p2p_device(const ld_ga_params* ga, const ld_peer_table* peers, const ld_record* records, int64_t n, double* children,
                          int64_t lo, int64_t hi, void* stream) {
    if (!ga || !peers || !records || !children) return fail(LD_ERR_INVALID_ARG, "NULL argument");
    if (peers->world < 1 || peers->world > LD_MAX_PEERS || peers->shard_rows < 2 || (peers->shard_rows & 1) ||
        peers->shard_rows * peers->world != n)
        return fail(LD_ERR_INVALID_ARG, "peer table: n must equal world * shard_rows with an even shard_rows");
    for (int r = 0; r < peers->world; ++r)
        if (!peers->shard[r]) return fail(LD_ERR_INVALID_ARG, "peer table: NULL shard pointer");
    if (lo < 0 || hi > n || lo > hi || (lo & 1) || (hi & 1)) return fail(LD_ERR_INVALID_ARG, "children range [lo, hi) must be even-aligned inside [0, n]");
    if (lo == hi) return LD_OK;
    int rc = ensure_device();
    if (rc) return rc;
    long long pairs = (hi - lo) / 2;
    long long blocks = (pairs + 7) / 8;
    if (blocks > 16LL * g_dev.sms) blocks = 16LL * g_dev.sms;
    ga_step_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*ga, nullptr, *peers, records, n, children, lo, hi);
    LD_CUDA(cudaGetLastError());
    return LD_OK;
}

int ld_drone_step_host(const ld_params* params, int what, const double* genomes, double* state, const double* coin, int32_t* flags,
                       int64_t n) {
    int rc = check_params(params);
    if (rc) return rc;
    if (what != 1 && what != 2) return fail(LD_ERR_INVALID_ARG, "what must be 1 (measure) or 2 (translate_input_to_thrust)");
    if (n < 0) return fail(LD_ERR_INVALID_ARG, "n < 0");
    if (n == 0) return LD_OK;
    if (!state || (what == 1 && !flags) || (what == 2 && (!genomes || !coin))) return fail(LD_ERR_INVALID_ARG, "NULL host buffer");
    rc = ensure_device();
    if (rc) return rc;
    rc = sws_reserve(1, false);
    if (rc) return rc;
    double *d_state = nullptr, *d_gen = nullptr, *d_coin = nullptr;
    int* d_flags = nullptr;
    if ((rc = scratch_get(1, (size_t)n * LD_STATE_LEN * sizeof(double), (void**)&d_state))) return rc;
    if (what == 1) {
        if ((rc = scratch_get(2, (size_t)n * 2 * sizeof(int), (void**)&d_flags))) return rc;
    } else {
        if ((rc = scratch_get(3, (size_t)n * LD_GENOME_LEN * sizeof(double), (void**)&d_gen))) return rc;
        if ((rc = scratch_get(4, (size_t)n * 3 * sizeof(double), (void**)&d_coin))) return rc;
    }
    cudaStream_t s = g_sws_slots[0].compute;
    cudaError_t e = cudaMemcpyAsync(d_state, state, (size_t)n * LD_STATE_LEN * sizeof(double), cudaMemcpyHostToDevice, s);
    if (what == 2) {
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_gen, genomes, (size_t)n * LD_GENOME_LEN * sizeof(double), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_coin, coin, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s);
    }
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((n + 127) / 128);
        if (has_extras(params)) drone_step_kernel<true><<<blocks, 128, 0, s>>>(*params, what, d_gen, d_state, d_coin, d_flags, n);
        else drone_step_kernel<false><<<blocks, 128, 0, s>>>(*params, what, d_gen, d_state, d_coin, d_flags, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(state, d_state, (size_t)n * LD_STATE_LEN * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && what == 1) e = cudaMemcpyAsync(flags, d_flags, (size_t)n * 2 * sizeof(int), cudaMemcpyDeviceToHost, s);
    const cudaError_t e_sync = cudaStreamSynchronize(s);
    if (e != cudaSuccess || e_sync != cudaSuccess) return fail(LD_ERR_CUDA, "ld_drone_step_host", e != cudaSuccess ? e : e_sync);
    return LD_OK;
}

int ld_brain_forward_host(const int32_t* layer_sizes, int32_t n_layers, const double* weights, int64_t weights_stride, const double* inputs,
                          int64_t n, double* outputs, double sigmoid_beta) {
    if (!layer_sizes || n_layers < 1 || n_layers > kBrainMaxLayers) return fail(LD_ERR_INVALID_ARG, "1..8 layers");
    BrainSpec spec;
    memset(&spec, 0, sizeof(spec));
    spec.n_layers = n_layers;
    int64_t total = 0;
    for (int l = 0; l <= n_layers; ++l) {
        if (layer_sizes[l] < 1 || layer_sizes[l] > kBrainMaxWidth) return fail(LD_ERR_INVALID_ARG, "layer widths must be 1..64");
        spec.size[l] = layer_sizes[l];
        if (l > 0) total += (int64_t)layer_sizes[l] * (layer_sizes[l - 1] + 1);
    }
    if (n < 0) return fail(LD_ERR_INVALID_ARG, "n < 0");
    if (n == 0) return LD_OK;
    if (!weights || !inputs || !outputs) return fail(LD_ERR_INVALID_ARG, "NULL host buffer");
    if (weights_stride != 0 && weights_stride < total) return fail(LD_ERR_INVALID_ARG, "weights_stride shorter than one brain");
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = sws_reserve(1, false))) return rc;
    const int64_t n_brains = weights_stride == 0 ? 1 : n; // stride 0: one brain shared by every input row
    const int n_in = spec.size[0], n_out = spec.size[n_layers];
    double *d_w = nullptr, *d_x = nullptr, *d_y = nullptr;
    if ((rc = scratch_get(1, (size_t)n_brains * (weights_stride ? weights_stride : total) * sizeof(double), (void**)&d_w))) return rc;
    if ((rc = scratch_get(2, (size_t)n * n_in * sizeof(double), (void**)&d_x))) return rc;
    if ((rc = scratch_get(3, (size_t)n * n_out * sizeof(double), (void**)&d_y))) return rc;
    cudaStream_t s = g_sws_slots[0].compute;
    cudaError_t e = cudaMemcpyAsync(d_w, weights, (size_t)n_brains * (weights_stride ? weights_stride : total) * sizeof(double),
                                    cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, inputs, (size_t)n * n_in * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        brain_forward_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(spec, d_w, weights_stride, d_x, n, d_y, -sigmoid_beta);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(outputs, d_y, (size_t)n * n_out * sizeof(double), cudaMemcpyDeviceToHost, s);
    const cudaError_t e_sync = cudaStreamSynchronize(s);
    if (e != cudaSuccess || e_sync != cudaSuccess) return fail(LD_ERR_CUDA, "ld_brain_forward_host", e != cudaSuccess ? e : e_sync);
    return LD_OK;
}

int ld_activation_host(int kind, const double* x, int64_t n, double* out, double a, double b) {
    if ((kind != 0 && kind != 1) || n < 0 || (n > 0 && (!x || !out))) return fail(LD_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return LD_OK;
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = sws_reserve(1, false))) return rc;
    double *d_x = nullptr, *d_y = nullptr;
    if ((rc = scratch_get(1, (size_t)n * sizeof(double), (void**)&d_x))) return rc;
    if ((rc = scratch_get(2, (size_t)n * sizeof(double), (void**)&d_y))) return rc;
    cudaStream_t s = g_sws_slots[0].compute;
    cudaError_t e = cudaMemcpyAsync(d_x, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        long long blocks = (n + 255) / 256;
        activation_kernel<<<(unsigned)(blocks > 1184 ? 1184 : blocks), 256, 0, s>>>(kind, d_x, n, d_y, a, b);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_y, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s);
    const cudaError_t e_sync = cudaStreamSynchronize(s);
    if (e != cudaSuccess || e_sync != cudaSuccess) return fail(LD_ERR_CUDA, "ld_activation_host", e != cudaSuccess ? e : e_sync);
    return LD_OK;
}

void ld_release_workspace(void) {
    ws_release();
    sws_release();
    scratch_release();
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
}

int ld_debug_math_device(int fn, const double* x, const double* y, double* out0, double* out1, int64_t n, void* stream) {
    if (fn < 0 || fn > 3 || n < 0 || (n > 0 && (!x || !out0))) return fail(LD_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return LD_OK;
    debug_math_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(fn, x, y ? y : x, out0, out1 ? out1 : out0, n);
    LD_CUDA(cudaGetLastError());
    return LD_OK;
}

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
