/*
 * ldrones.h — C-ABI of libldrones.so, the B200 (sm_100a) population fitness
 * evaluator for mbecker12/learning-drones.
 *
 * The reference has no FFI: its seam is the Python call
 *     fly(drone, timesteps, delta_t, visualize, n_servers, punish_time,
 *         reward_time, idx) -> (reward, real_time, idx)
 * (reference flight/__init__.py:12-21,87,99), invoked once per drone per
 * generation from evolve/evolution.py:64.  The entry points below are what a
 * ctypes binding for that call site binds instead: one
 * call evaluates the whole population.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in any signature.
 *   - "device pointer" = address in the current CUDA device's global memory,
 *     borrowed from the caller (e.g. torch.Tensor.data_ptr()); the library
 *     never allocates, frees or retains caller buffers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream). *_device entry points are stream-ordered and asynchronous.
 *   - every entry point returns 0 on success or a negative ld_status;
 *     ld_last_error() returns a thread-local message for the last failure.
 *     Nothing throws across the ABI.
 *   - a "genome" is one drone's brain flattened exactly like the reference's
 *     crossover does (evolve/util.py:58-59: layer.flatten(), layer after
 *     layer): W1 (10x16) | W2 (10x11) | W3 (4x11), row = neuron, bias is the
 *     LAST column (evolve/brain.py:45-46).  314 doubles = 2512 B per drone.
 *     A population is the row-major matrix genomes[n][LD_GENOME_LEN].
 */
#ifndef LDRONES_H
#define LDRONES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LD_ABI_VERSION 3
#define LD_SLICE_STAGGER (-1000000) /* ld_params.slice_steps */

/* brain shape fixed by drone/drone.py:65 : NeuralNetwork({1:10, 2:10, 3:4}), n_inputs=15 */
#define LD_N_IN 15
#define LD_N_H1 10
#define LD_N_H2 10
#define LD_N_OUT 4
#define LD_GENOME_LEN (LD_N_H1 * (LD_N_IN + 1) + LD_N_H2 * (LD_N_H1 + 1) + LD_N_OUT * (LD_N_H2 + 1)) /* 314 */

/* PID mode: 18 gains per drone, order of choose_paramset.py:284-332
 * (kp_x,kp_y,kp_z, ki_x,ki_y,ki_z, kd_x,kd_y,kd_z, kp_rol,kp_pit,kp_yaw,
 *  ki_rol,ki_pit,ki_yaw, kd_rol,kd_pit,kd_yaw) */
#define LD_PID_GAINS 18
#define LD_PID_MAX_COINS 8

/* final_state row: pos(3) vel(3) angle(3) angle_vel(3) acc_current(6) acc_last(6) thrust(4) */
#define LD_STATE_LEN 28

typedef enum ld_status {
    LD_OK = 0,
    LD_ERR_INVALID_ARG = -1, /* NULL pointer, negative size, bad params */
    LD_ERR_CUDA = -2,        /* a CUDA runtime call failed; see ld_last_error() */
    LD_ERR_NO_DEVICE = -3    /* no usable sm_100 device */
} ld_status;
/* There is no collective entry point in this ABI: the one exchange step of the path, the all-gather of the 32-byte
 * fitness records, is torch.distributed's NCCL all-gather on the caller's side
 * (learning_drones_b200/distributed.py); the SURVEY §8b sketch's `ld_allgather_records` is struck.  The irregular
 * exchange — the GA step's parent fetch — IS in the library, fused into the kernel over peer memory
 * (ld_ga_step_p2p_device). */

/* status bits of ld_record.status */
#define LD_STATUS_CRASHED 1   /* drone/drone.py:150-154 */
#define LD_STATUS_NONFINITE 2 /* z became NaN: reference would raise in int(nan), drone.py:14 */
#define LD_STATUS_FINISHED 4  /* PID mode: every target coin collected (collect_coins.py:289-292) */

/*
 * Everything fly() reads that is not the brain.  Derived constants are
 * computed by the HOST exactly the way the reference computes them (python
 * floats), so the device never re-derives them with different rounding.
 */
typedef struct ld_params {
    /* QuadcopterPhysics (drone/quadcopter/physical_model.py:53-73) */
    double c_f;                /* coef_force   parameters.py:31 */
    double c_m;                /* coef_moment  parameters.py:32 */
    double r_m;                /* radius_motor_center */
    double sqrt2_r_m;          /* np.sqrt(2) * r_m       physical_model.py:109 */
    double inv_I[3];           /* 1/I_x, 1/I_y, 1/I_z    physical_model.py:63 */
    double moments_payload[3]; /* physical_model.py:62,68-73 (zeros without payload) */
    float neg_G32;             /* float32(-1 * G), G = gravity*(4 m_m + m_c) (+payload)  :61,101 */
    float m_total32;           /* float32(m_c + 4 m_m + m_p)                             :114 */
    float wind_cw32[3];        /* float32(wind) * float32(c_w)                           :103 */
    float _pad0[3];
    double dt;                 /* integrator step, drone/drone.py:10 (NOT fly()'s delta_t) */

    /* initial state (parameters.py:47-52, Drone.__init__ / reset) */
    double init_pos[3], init_vel[3], init_angle[3], init_angle_vel[3], init_thrust[4];

    /* reward rules (drone/drone.py:12-17) and crash box (:150) */
    double angle_limit;     /* pi/4 */
    double angle_vel_limit; /* 20 */
    double z_min, z_max;    /* crash iff z <= z_min or z > z_max : 0, 80 */
    int32_t reward_time_passed;   /* 2 */
    int32_t reward_crashed;       /* -3000 */
    int32_t reward_unstable;      /* -1 */
    int32_t reward_coin_distance; /* -100  (times int(distance)) */

    /* coin schedule (flight/__init__.py:25,52) */
    double coin0_pos[3];   /* (5,5,10) */
    double coin_step[3];   /* (5,2.5,0) */
    double coin0_radius;   /* 2 */
    double coinN_radius;   /* 1 */
    double coin0_thr2;     /* smallest s with sqrt(s) >= coin0_radius (host-computed; lets the
                              device test ||d||^2 < thr2 instead of sqrt(||d||^2) < r, same truth value) */
    double coinN_thr2;
    int32_t coin0_value;   /* 2000 */
    int32_t coinN_value;   /* 1000 */

    double sigmoid_beta;   /* 0.4, evolve/brain.py:37 */

    /* episode control (fly() arguments) */
    int32_t timesteps;     /* 2500 default */
    int32_t reward_time;   /* bool */
    int32_t punish_time;   /* bool */
    int32_t refill_batch;  /* tuning: idle lanes per warp that trigger a queue refill; 0 = library default */
    int32_t slice_steps;   /* scheduling only, never changes results: when more drones wait than the GPU has
                              lanes, an fp64 episode is flown in segments and the drone goes back to the queue in
                              between (round robin), so that long-lived drones end together instead of in
                              whole-episode waves and short-lived ones free their lane early.
                              > 0: equal segments of this many steps; <= -2: segments that double in length,
                              the first one -slice_steps steps long; 0 = library default (doubling from
                              timesteps/4 for populations of 1..16 waves); -1 = off (every drone keeps its lane
                              until its episode ends); LD_SLICE_STAGGER = staggered first wave, for populations
                              known to consist of full-horizon drones: only the first batch of every warp has a
                              boundary, spread evenly over the horizon across the warps, so the lanes run out of phase
                              and the drones parked there fill the end of the launch longest-remaining-first —
                              one hand-over per lane instead of two per drone (device-resident launches; a
                              host-buffer launch, whose lanes start as their rows arrive, uses the default) */
    int32_t reserved0;
} ld_params;

/* what fly() leaves on the Drone besides its kinematic state */
typedef struct ld_record {
    int64_t reward;          /* drone.reward (python int) */
    int32_t coins;           /* drone.coins */
    int32_t flight_time;     /* drone.flight_time = number of loop iterations entered */
    int32_t status;          /* LD_STATUS_* bits; bit0 == drone.crashed */
    int32_t unstable;        /* number of -1 penalties applied (drone.py:159-164) */
    double distance_to_coin; /* drone.distance_to_coin */
} ld_record;

/* ---- library ------------------------------------------------------------- */
int ld_abi_version(void);
const char* ld_last_error(void);
/* number of CUDA devices usable by this build (compute capability 10.x); <0 on error */
int ld_device_count(void);
/* resident drones per launch wave on the current device (SMs x lanes per SM) */
int64_t ld_wave_size(void);

/* ---- MLP-controlled episode: replaces the population loop around fly() ----
 * reference: evolve/evolution.py:64-73 calling flight/__init__.py:12-99 per drone.
 *
 * genomes      [n][LD_GENOME_LEN] doubles
 * records      [n] ld_record                        (written)
 * final_state  [n][LD_STATE_LEN] doubles or NULL    (written if not NULL)
 * work_counter TWO int64 of scratch device memory: [0] is the work-queue head, [1] an error flag.  The call
 *              zeroes both itself (stream-ordered); after the launch has completed, work_counter[1] != 0 means that
 *              a warp gave up waiting for its work queue (20 s without progress — never in a healthy run) and the
 *              records are incomplete.  The *_host entry points check it and return LD_ERR_CUDA.
 */
int ld_fly_mlp_f64_device(const ld_params* params, const double* genomes, int64_t n,
                          ld_record* records, double* final_state,
                          int64_t* work_counter, void* stream);

/* Same computation with HOST buffers: copies in, runs, copies out, synchronises.
 * This is the call the reference-facing plugin makes (bench e2e).  One persistent
 * launch per <= 4 Mi drones that starts on an empty genome matrix: the chunked H2D
 * copies are queued first and raise an arrival counter the kernel's work queue
 * waits on, so lanes start flying as their rows land (pinned host memory overlaps
 * fully; pageable memory is staged by the driver and overlaps less). */
int ld_fly_mlp_f64_host(const ld_params* params, const double* genomes, int64_t n,
                        ld_record* records, double* final_state);

/* The same with a per-drone START state: init_state [n][LD_STATE_LEN] rows (pos vel angle angle_vel acc_current
 * acc_last thrust; acc_last is ignored) replace ld_params.init_*.  This is fly() on Drones that are not in the reset
 * state — the reference continues from whatever position / velocity / angle / thrust / sensor accelerations the Drone
 * holds (flight/__init__.py:35-39, drone/drone.py:20-76 constructor arguments).  Counters start at 0: the reference
 * accumulates reward, coins and flight_time on top of the Drone's previous values, which is an integer addition the
 * caller does (learning_drones_b200/flight/__init__.py).  init_state == NULL is ld_fly_mlp_f64_{host,device}. */
int ld_fly_mlp_f64_from_host(const ld_params* params, const double* genomes, const double* init_state, int64_t n,
                             ld_record* records, double* final_state);
int ld_fly_mlp_f64_from_device(const ld_params* params, const double* genomes, const double* init_state, int64_t n,
                               ld_record* records, double* final_state, int64_t* work_counter, void* stream);

/* Pipelined form of the host-buffer call, for jobs that evaluate one population after another (islands, parameter
 * sweeps, batches of a population that does not fit one launch): `submit` queues the copies and the launch and
 * returns at once with a ticket; `wait` blocks until that submission's records (and final_state) are in the caller's
 * buffers and reports its status.  Up to LD_HOST_PIPELINE_DEPTH submissions may be in flight per host thread: while
 * one flies, the next one's genome rows are already streaming in, so a step costs max(kernel, copy) instead of their
 * sum.  Caller buffers (pinned for real overlap) must stay valid and untouched until `wait` returns; tickets belong to
 * the submitting thread; init_state may be NULL.  1 <= n <= 4 Mi drones per submission. */
#define LD_HOST_PIPELINE_DEPTH 2
int ld_fly_mlp_f64_host_submit(const ld_params* params, const double* genomes, const double* init_state, int64_t n,
                               ld_record* records, double* final_state, int64_t* ticket);
int ld_fly_mlp_f64_host_wait(int64_t ticket);

/* fp32 FAST MODE: same call contract and the same fp64 genome matrix / record
 * layout, but state, brain and arithmetic are float32 on the device.  The dynamics are chaotic:
 * float32 changes the integer outcome of a large fraction of random drones, so this
 * mode carries NO parity claim and is compared with the fp64 mode statistically only
 * (tests/test_gpu_f32.py).  final_state[.][18..23] (previous accelerations) is written as 0. */
int ld_fly_mlp_f32_device(const ld_params* params, const double* genomes, int64_t n,
                          ld_record* records, double* final_state,
                          int64_t* work_counter, void* stream);
int ld_fly_mlp_f32_host(const ld_params* params, const double* genomes, int64_t n,
                        ld_record* records, double* final_state);

/* Debug tap: fly ONE drone and dump its state after every
 * measure() call: trace[t][LD_STATE_LEN], t < timesteps; rows after the final
 * step are left untouched. Device pointers. */
int ld_fly_mlp_f64_trace_device(const ld_params* params, const double* genome,
                                ld_record* record, double* trace, void* stream);

/* ---- single steps of a Drone, for code that drives the episode loop itself ----------------------
 * what = 1: Drone.measure(wind_speed)              (drone/drone.py:129-165; wind via ld_params)
 *           state rows [n][LD_STATE_LEN] are updated in place; flags[i] = {crashed, number of -1 penalties}
 * what = 2: Drone.translate_input_to_thrust(coin)  (drone/drone.py:83-86 -> evolve/brain.py:56-76)
 *           state[.][24..27] (thrust) is replaced by the brain's output for coin[n][3]
 * Host buffers; same device arithmetic as the episode kernel. */
int ld_drone_step_host(const ld_params* params, int what, const double* genomes, double* state, const double* coin,
                       int32_t* flags, int64_t n);

/* NeuralNetwork.feed_forward for ANY layer_spec (evolve/brain.py:10,41-54; the reference's own default is
 * {1: 8, 2: 10, 3: 4}): layer_sizes[0..n_layers] = inputs, then neurons per layer (<= 8 layers, widths <= 64);
 * weights = per layer (n_out, n_in + 1) row-major, bias LAST, layer after layer; row i of inputs [n][sizes[0]] goes
 * through the brain at weights + i * weights_stride (weights_stride == 0: one brain for every row); hidden ReLU,
 * output sigmoid(beta).  Same device arithmetic as the episode kernel (k-ordered fma chains), so the Drone's
 * 15-10-10-4 shape reproduces the episode kernel's thrust bit for bit.  Host buffers. */
int ld_brain_forward_host(const int32_t* layer_sizes, int32_t n_layers, const double* weights, int64_t weights_stride,
                          const double* inputs, int64_t n, double* outputs, double sigmoid_beta);
/* NeuralNetwork._relu / relu (kind 0: out = x < a ? 0.0 : b * x, brain.py:26-35) and sigmoid (kind 1:
 * out = 1 / (exp(-a x) + 1), brain.py:37-39) on the device.  Host buffers. */
int ld_activation_host(int kind, const double* x, int64_t n, double* out, double a, double b);

/* ---- PID-controlled episode (second controller mode) ------------
 * The main loop of flight/collect_coins.py:278-364 for a batch of PID gain sets: per step
 * score-1, coin test on the current target (+value, next target), exit when all targets are
 * collected, physics + integrator (same as the MLP mode, no crash rule), cascaded PIDs
 * (drone/controller/pid.py:77-106,132-142 "signChange"; drone/onboard_computer.py:50-118). */
typedef struct ld_pid_params {
    int32_t n_coins;   /* 4 */
    int32_t timesteps; /* 3500; the score starts at this value (collect_coins.py:276-277) */
    int32_t step_reward; /* -1 */
    int32_t _pad;
    double coin_pos[LD_PID_MAX_COINS][3]; /* collect_coins.py:253-258 */
    double coin_thr2[LD_PID_MAX_COINS];   /* sqrt-threshold of each radius (see ld_params.coin0_thr2) */
    int32_t coin_value[LD_PID_MAX_COINS];
    double integral_range;        /* 2 */
    double out_lo[6], out_hi[6];  /* output limits of the x y z roll pitch yaw PIDs (collect_coins.py:125-188) */
    double base_thrust;           /* G / (4 c_f)  onboard_computer.py:97 */
    double yaw_setpoint;          /* 0 */
} ld_pid_params;

/* pid_gains [n][6][3]: (kp, ki, kd) of the x, y, z, roll, pitch, yaw controllers of every drone.
 * records: reward = final score, coins, flight_time = loop iterations entered, status
 * (LD_STATUS_FINISHED), distance_to_coin = distance to the current target.  Device pointers. */
int ld_fly_pid_f64_device(const ld_params* params, const ld_pid_params* pid, const double* pid_gains, int64_t n,
                          ld_record* records, double* final_state, void* stream);
int ld_fly_pid_f64_host(const ld_params* params, const ld_pid_params* pid, const double* pid_gains, int64_t n,
                        ld_record* records, double* final_state);

/* ---- GA step: evolve/util.py:5-78 + drone/drone.py:167-172 on the genome matrix --
 * For every pair slot (c, c+1), c even, of the NEW population: four uniform parent indices, two
 * 2-way tournaments on `reward` (with probability p_tournament the better one wins, else the
 * worse; "better" = strictly greater reward, ties -> second index, util.py:30-45), children =
 * copies of the winners, then per LAYER with probability p_crossover a single-point crossover
 * of the C-order-flattened layer (util.py:48-78), then per weight with probability
 * mutation_rate `+= N(0, mutation_sigma)` (drone.py:167-172).
 * Randomness is Philox4x32-10 keyed by (seed, generation, slot): every rank derives the same
 * selections without communication.  It is NOT numpy's MT19937 stream (out of scope).
 *
 * parents      [n][LD_GENOME_LEN]   whole current population (device)
 * records      [n]                  fitness of the current population (reward field is used)
 * children     [hi-lo][LD_GENOME_LEN] rows lo..hi-1 of the new population (lo, hi even)
 * Elitism (evolve/evolution.py:105-107: slot 0 <- unmutated all-time best) is a row copy the
 * caller does afterwards. */
typedef struct ld_ga_params {
    uint64_t seed;
    uint32_t generation;
    uint32_t _pad;
    double p_tournament;   /* 0.6  evolve/evolution.py:19 */
    double p_crossover;    /* 0.05 :20 */
    double mutation_rate;  /* 0.05 :18 */
    double mutation_sigma; /* 0.3  drone/drone.py:172 */
} ld_ga_params;

int ld_ga_step_device(const ld_ga_params* ga, const double* parents, const ld_record* records, int64_t n,
                      double* children, int64_t lo, int64_t hi, void* stream);

/* Same GA step with the parent population SHARDED over the GPUs of one box: rank r owns rows
 * [r*shard_rows, (r+1)*shard_rows) at peers.shard[r], a pointer valid on THIS device (its own
 * shard or a peer's buffer mapped over NVLink, e.g. torch symmetric memory buffer_ptrs).  The
 * kernel loads winner rows straight from the owning GPU's HBM while it applies crossover and
 * mutation — no all-gather of genomes.  The caller guarantees that no peer overwrites its parent
 * shard before every rank has finished this step (the next generation's record all-gather is
 * such a point). */
#define LD_MAX_PEERS 16
typedef struct ld_peer_table {
    const double* shard[LD_MAX_PEERS];
    int64_t shard_rows; /* rows per rank; n == world * shard_rows, shard_rows even */
    int32_t world;
    int32_t _pad;
} ld_peer_table;

int ld_ga_step_p2p_device(const ld_ga_params* ga, const ld_peer_table* peers, const ld_record* records, int64_t n,
                          double* children, int64_t lo, int64_t hi, void* stream);

/* ---- population statistics (warp-shuffle reductions) ----------------------
 * stats[0]=sum flight_time (active drone-steps), [1]=max reward, [2]=argmax reward
 * (lowest index wins ties, like np.argmax at evolve/evolution.py:77),
 * [3]=sum coins, [4]=crashed count.  Device pointers. */
int ld_reduce_records_device(const ld_record* records, int64_t n, int64_t* stats5, void* stream);

/* ---- best-of-run tracking on the device: evolve/evolution.py:77-83 ----------
 * Reads stats5 of the generation just flown (ld_reduce_records_device).  If there is no global best yet
 * (best2[1] == 0) or stats5[1] > best2[0]: best2 = {stats5[1], 1} and best_genome[LD_GENOME_LEN] = row
 * stats5[2] of the population — `population` [n][LD_GENOME_LEN], or, when it is NULL, the sharded
 * population behind `peers` (row r lives on rank r / shard_rows; read over NVLink peer memory).
 * history7 (may be NULL) receives {generation, best reward, best index, global best reward, active
 * drone-steps, coins, crashed} so that a host loop can run many generations without synchronising.
 * Device pointers, stream-ordered. */
int ld_track_best_device(const int64_t* stats5, const double* population, const ld_peer_table* peers, int64_t generation,
                         int64_t* best2, double* best_genome, int64_t* history7, void* stream);

/* Roofline denominator: measured fp64 (non-tensor) FMA throughput of the current device in
 * TFLOP/s (FMA = 2 flop), from a register-only DFMA-chain microbenchmark; optionally the same
 * as DFMA per max-clock cycle per SM. */
int ld_measure_fp64_peak(double* tflops_out, double* dfma_per_clk_per_sm_out);

/* Accuracy tap for the kernel's own elementary functions (csrc/ld_math.cuh); device pointers.
 * fn 0: out0=sin(x) out1=cos(x) | 1: out0=exp(x) | 2: out0=1/(exp(-0.4x)+1) | 3: out0=x/y out1=1/y */
int ld_debug_math_device(int fn, const double* x, const double* y, double* out0, double* out1, int64_t n, void* stream);

/* Where the episode kernel spends its time (development tap): same launch as ld_fly_mlp_f64_device (no final state,
 * no wind / payload) with an instrumented instantiation that sums, over all warps, clock64 cycles per phase and a
 * few event counts into prof[LD_PROFILE_SLOTS] (device, zeroed by the call): 0 claim (queue atomics, slot polls),
 * 1 stage (brain rows -> Tensor Memory / shared memory), 2 wait (nothing to fly), 3 hot loop, 4 publish / park,
 * 5 staging events, 6 lanes staged, 7 hot-loop entries, 8 warp-steps, 9 live lane-steps, 10 warps; of the staging:
 * 11 drone state (reset / unpark), 12 row copies into shared memory, 13 Tensor-Memory / shared-memory writes, 14 groups, 15 warp-steps that ran the
 * large-angle (libm) trigonometry.
 * times (device, [n][2], may be NULL): %globaltimer ns at which every drone first got a lane and at which it landed. */
#define LD_PROFILE_SLOTS 16
int ld_debug_fly_profile_device(const ld_params* params, const double* genomes, int64_t n, ld_record* records,
                                int64_t* work_counter, uint64_t* prof, uint64_t* times, void* stream);

/* frees the device workspaces that the *_host entry points keep between calls (per host thread) */
void ld_release_workspace(void);

#ifdef __cplusplus
}
#endif
#endif /* LDRONES_H */
