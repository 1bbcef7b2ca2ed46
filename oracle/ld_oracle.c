/*
 * ld_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement (plain C, fp64 with the reference's two fp32 rounding
 * points) of learning-drones' population fitness evaluation path.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this; the product (learning_drones_b200) never does.
 *
 * Parity pin: the reference's own test-suite holds no runnable test or golden
 * vector for this path, so the oracle is pinned against outputs
 * of the reference itself, run in the build container by
 * oracle/gen_golden.py and committed under tests/golden/.
 *
 * Arithmetic conventions restated from the reference (numpy 2.3.5/OpenBLAS
 * 0.3.30, measured in the build container):
 *   - python-float / np.float64 scalar expressions: IEEE double, evaluated in
 *     source order, NO contraction (compile with -ffp-contract=off).
 *   - np.sum over 4 elements: sequential left-to-right.
 *   - np.dot on 3x3 / 3-vectors / the MLP matvec goes to OpenBLAS, which uses
 *     FMA.  For the 3-wide products it is bitwise an FMA chain in k order
 *     (measured); for dgemv the lane order is host-CPU specific, so the oracle
 *     uses the k-ordered FMA chain there too.  `use_fma = 0` switches every
 *     dot to unfused mul+add so tests can bound the sensitivity.
 *   - np.sin/np.cos == libm here; np.exp/np.tan differ from libm by <=1 ulp on
 *     a few % of arguments, so bitwise equality with the reference is not
 *     attainable; see DESIGN.md "parity contract".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ldrones.h"

#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
    double pos[3], vel[3], ang[3], rate[3]; /* Drone.position/.velocity/.angle/.angle_vel */
    double a_cur[6], a_last[6];             /* Sensor.current_/last_acceleration, x y z roll pitch yaw */
    double thrust[4];
    double coin[3], coin_radius;
    int32_t coin_value;
    int64_t reward;
    int32_t coins, flight_time, status, unstable;
    double distance;
} od_drone;

static inline double dot_step(double a, double b, double acc, int use_fma) {
    return use_fma ? fma(a, b, acc) : acc + a * b;
}

/* np.dot(A, v) with A C-contiguous 3x3 and v (3,1): OpenBLAS 0.3.30's dgemv_t kernel
 * evaluates each row as fma(a2,v2, fma(a0,v0, a1*v1)) — measured bitwise on 9000/9000 random
 * rows in the build container (x86-64 AVX-512 host).  The F-ordered case (A a .T view) and the
 * 3x3 dgemm are plain k-ordered chains (dot_n). */
static inline double dot3_rowmajor(const double* a, const double* v, int use_fma) {
    if (!use_fma) return (a[0] * v[0] + a[1] * v[1]) + a[2] * v[2];
    return fma(a[2], v[2], fma(a[0], v[0], a[1] * v[1]));
}

/* sum_k a[k]*b[k], k = 0..n-1, OpenBLAS-like: first product, then chain */
static inline double dot_n(const double* a, const double* b, int n, int use_fma) {
    double acc = a[0] * b[0];
    for (int k = 1; k < n; ++k) acc = dot_step(a[k], b[k], acc, use_fma);
    return acc;
}

/* drone/quadcopter/geometry.py:15-43 — rotation_matrix() returns (Rz.(Ry.Rx)).T ; we return
 * B = Rz.(Ry.Rx) (so Rot = B^T), built with the same two 3x3 np.dot calls. */
static void rotation_B(double roll, double pitch, double yaw, double B[3][3], int use_fma) {
    double cr = cos(roll), sr = sin(roll), cp = cos(pitch), sp = sin(pitch), cy = cos(yaw), sy = sin(yaw);
    double Rx[3][3] = {{1, 0, 0}, {0, cr, -1 * sr}, {0, sr, cr}};
    double Ry[3][3] = {{cp, 0, sp}, {0, 1, 0}, {-1 * sp, 0, cp}};
    double Rz[3][3] = {{cy, -1 * sy, 0}, {sy, cy, 0}, {0, 0, 1}};
    double A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double col[3] = {Rx[0][j], Rx[1][j], Rx[2][j]};
            A[i][j] = dot_n(Ry[i], col, 3, use_fma);
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double col[3] = {A[0][j], A[1][j], A[2][j]};
            B[i][j] = dot_n(Rz[i], col, 3, use_fma);
        }
}

/* drone/quadcopter/physical_model.py:75-122 — calculate_accelerations(..., lin_acc_drone_2_lab=True) */
static void accelerations(const ld_params* p, const double ang[3], const double thrust[4], double lin[3],
                          double rot[3], int use_fma) {
    double B[3][3];
    rotation_B(ang[0], ang[1], ang[2], B, use_fma); /* :90, Rot = B^T */
    double T[4], R[4];
    for (int i = 0; i < 4; ++i) {
        T[i] = thrust[i] * p->c_f; /* :93 */
        R[i] = thrust[i] * p->c_m; /* :94 */
    }
    double sumT = ((T[0] + T[1]) + T[2]) + T[3]; /* np.sum, :98 */
    float forces[3] = {0.0f, 0.0f, (float)sumT};  /* float32 array, :98 */
    double g32[3] = {0.0, 0.0, (double)p->neg_G32};
    double w32[3] = {(double)p->wind_cw32[0], (double)p->wind_cw32[1], (double)p->wind_cw32[2]};
    for (int i = 0; i < 3; ++i) {
        double rot_row[3] = {B[0][i], B[1][i], B[2][i]}; /* Rot[i][:] */
        double G_rot = dot_n(rot_row, g32, 3, use_fma);   /* :100-102 */
        double W_rot = dot_n(rot_row, w32, 3, use_fma);   /* :103 */
        forces[i] = (float)((double)forces[i] + (G_rot + W_rot)); /* :104, f64 add, f32 store */
    }
    double L = p->r_m * (T[0] + T[3] - T[1] - T[2]);       /* :107 */
    double M = p->r_m * (T[2] + T[3] - T[0] - T[1]);       /* :108 */
    double N = p->sqrt2_r_m * (R[1] + R[3] - R[0] - R[2]); /* :109 */
    double moments[3] = {L + p->moments_payload[0], M + p->moments_payload[1], N + p->moments_payload[2]}; /* :110-111 */
    double lin32[3];
    for (int i = 0; i < 3; ++i) lin32[i] = (double)(float)(forces[i] / p->m_total32); /* :114, float32 divide */
    double rb[3];
    for (int i = 0; i < 3; ++i) rb[i] = moments[i] * p->inv_I[i]; /* :115 */
    for (int i = 0; i < 3; ++i) lin[i] = dot3_rowmajor(B[i], lin32, use_fma); /* :118, Rot.T = B, C-contiguous */
    /* geometry.py:52-71 — translational_matrix, every entry rounded to float32 */
    double roll = ang[0], pitch = ang[1];
    double Tr[3][3] = {
        {1.0, (double)(float)(sin(roll) * tan(pitch)), (double)(float)(cos(roll) * tan(pitch))},
        {0.0, (double)(float)cos(roll), (double)(float)(-sin(roll))},
        {0.0, (double)(float)(sin(roll) / cos(pitch)), (double)(float)(cos(roll) / cos(pitch))}};
    for (int i = 0; i < 3; ++i) rot[i] = dot3_rowmajor(Tr[i], rb, use_fma); /* :120, Tr C-contiguous */
}

/* objects/coins.py:17-25 — np.linalg.norm(coin - pos) = sqrt(x.dot(x)) */
static double coin_distance(const od_drone* d, int use_fma) {
    double x[3] = {d->coin[0] - d->pos[0], d->coin[1] - d->pos[1], d->coin[2] - d->pos[2]};
    return sqrt(dot_n(x, x, 3, use_fma));
}

/* drone/sensor/sensor.py:42-49, 62-73, 85-93 with delta_t from drone/drone.py:10 */
static void integrate(const ld_params* p, od_drone* d, const double lin[3], const double rot[3]) {
    double dt = p->dt;
    double* x[6] = {&d->pos[0], &d->pos[1], &d->pos[2], &d->ang[0], &d->ang[1], &d->ang[2]};
    double* v[6] = {&d->vel[0], &d->vel[1], &d->vel[2], &d->rate[0], &d->rate[1], &d->rate[2]};
    for (int i = 0; i < 6; ++i) {
        d->a_last[i] = d->a_cur[i];
        d->a_cur[i] = i < 3 ? lin[i] : rot[i - 3];
        double delta_x = *v[i] * dt + 0.5 * d->a_cur[i] * dt * dt; /* :63-66, uses the OLD velocity */
        double delta_v = 0.5 * (d->a_last[i] + d->a_cur[i]) * dt;  /* :70-72 */
        *x[i] = *x[i] + delta_x;
        *v[i] = *v[i] + delta_v;
    }
}

/* drone/drone.py:129-165 — returns 1 when crashed */
static int measure(const ld_params* p, od_drone* d, int use_fma) {
    double lin[3], rot[3];
    accelerations(p, d->ang, d->thrust, lin, rot, use_fma);
    integrate(p, d, lin, rot);
    double z = d->pos[2];
    if (z <= p->z_min || z > p->z_max) { /* :150 */
        d->reward += p->reward_crashed;
        d->status |= LD_STATUS_CRASHED;
        return 1;
    }
    if (fabs(d->ang[0]) > p->angle_limit) { d->reward += p->reward_unstable; d->unstable++; }
    if (fabs(d->ang[1]) > p->angle_limit) { d->reward += p->reward_unstable; d->unstable++; }
    if (fabs(d->rate[2]) > p->angle_vel_limit) { d->reward += p->reward_unstable; d->unstable++; }
    return 0;
}

/* evolve/brain.py:41-76 — input stack, 3x (append 1, W.x, relu|sigmoid) */
static void brain(const ld_params* p, const double* g, od_drone* d, int use_fma) {
    double x0[LD_N_IN + 1], h1[LD_N_H1 + 1], h2[LD_N_H2 + 1];
    for (int i = 0; i < 3; ++i) {
        x0[i] = d->pos[i];
        x0[3 + i] = d->vel[i];
        x0[6 + i] = d->ang[i];
        x0[9 + i] = d->rate[i];
        x0[12 + i] = d->coin[i];
    }
    x0[LD_N_IN] = 1.0;
    const double* W1 = g;
    const double* W2 = W1 + LD_N_H1 * (LD_N_IN + 1);
    const double* W3 = W2 + LD_N_H2 * (LD_N_H1 + 1);
    for (int j = 0; j < LD_N_H1; ++j) {
        double s = dot_n(W1 + j * (LD_N_IN + 1), x0, LD_N_IN + 1, use_fma);
        h1[j] = s < 0 ? 0.0 : 1.0 * s; /* brain.py:26-30 */
    }
    h1[LD_N_H1] = 1.0;
    for (int j = 0; j < LD_N_H2; ++j) {
        double s = dot_n(W2 + j * (LD_N_H1 + 1), h1, LD_N_H1 + 1, use_fma);
        h2[j] = s < 0 ? 0.0 : 1.0 * s;
    }
    h2[LD_N_H2] = 1.0;
    for (int j = 0; j < LD_N_OUT; ++j) {
        double s = dot_n(W3 + j * (LD_N_H2 + 1), h2, LD_N_H2 + 1, use_fma);
        d->thrust[j] = 1 / (exp(-p->sigmoid_beta * s) + 1); /* brain.py:37-39 */
    }
}

static void drone_reset(const ld_params* p, od_drone* d) {
    memset(d, 0, sizeof(*d));
    for (int i = 0; i < 3; ++i) {
        d->pos[i] = p->init_pos[i];
        d->vel[i] = p->init_vel[i];
        d->ang[i] = p->init_angle[i];
        d->rate[i] = p->init_angle_vel[i];
        d->coin[i] = p->coin0_pos[i];
    }
    for (int i = 0; i < 4; ++i) d->thrust[i] = p->init_thrust[i];
    d->coin_radius = p->coin0_radius;
    d->coin_value = p->coin0_value;
    d->distance = INFINITY; /* drone.py:67 */
}

static void state_row(const od_drone* d, double* row) {
    memcpy(row + 0, d->pos, 24);
    memcpy(row + 3, d->vel, 24);
    memcpy(row + 6, d->ang, 24);
    memcpy(row + 9, d->rate, 24);
    memcpy(row + 12, d->a_cur, 48);
    memcpy(row + 18, d->a_last, 48);
    memcpy(row + 24, d->thrust, 32);
}

static void finish(const ld_params* p, od_drone* d, int use_fma) {
    double dist = coin_distance(d, use_fma);
    if (isfinite(dist))
        d->reward += (int64_t)p->reward_coin_distance * (int64_t)dist; /* -100*int(d), drone.py:14 */
    else
        d->status |= LD_STATUS_NONFINITE; /* python: int(nan) raises ValueError */
    d->distance = dist;
}

/* flight/__init__.py:12-99.  `start` (may be NULL): a state row the Drone holds when fly() is entered — constructor
 * arguments (drone/drone.py:55-76) or what a previous flight left (the reference never resets inside fly(), :35-39);
 * its counters are added by the caller (integer addition). */
static void fly(const ld_params* p, const double* g, od_drone* d, double* trace, int use_fma, const double* start) {
    drone_reset(p, d);
    if (start) {
        memcpy(d->pos, start + 0, 24);
        memcpy(d->vel, start + 3, 24);
        memcpy(d->ang, start + 6, 24);
        memcpy(d->rate, start + 9, 24);
        memcpy(d->a_cur, start + 12, 48);
        memcpy(d->a_last, start + 18, 48);
        memcpy(d->thrust, start + 24, 32);
    }
    for (int t = 0; t < p->timesteps; ++t) {
        if (p->punish_time) d->reward -= p->reward_time_passed; /* :41-42 */
        if (p->reward_time) d->reward += p->reward_time_passed; /* :43-44 */
        d->flight_time += 1;
        if (coin_distance(d, use_fma) < d->coin_radius) { /* :47 */
            d->reward += d->coin_value;
            d->coins += 1;
            for (int i = 0; i < 3; ++i) d->coin[i] = d->coin[i] + p->coin_step[i]; /* :52 */
            d->coin_radius = p->coinN_radius;
            d->coin_value = p->coinN_value;
        }
        int crashed = measure(p, d, use_fma); /* :79 */
        if (trace) state_row(d, trace + (size_t)t * LD_STATE_LEN);
        if (crashed) { /* :80-87 */
            finish(p, d, use_fma);
            return;
        }
        brain(p, g, d, use_fma); /* :89 */
        if (trace) memcpy(trace + (size_t)t * LD_STATE_LEN + 24, d->thrust, 32);
    }
    finish(p, d, use_fma); /* :97-98 */
}

static void emit(const od_drone* d, ld_record* rec, double* final_state) {
    rec->reward = d->reward;
    rec->coins = d->coins;
    rec->flight_time = d->flight_time;
    rec->status = d->status;
    rec->unstable = d->unstable;
    rec->distance_to_coin = d->distance;
    if (final_state) state_row(d, final_state);
}

/* ---- exported (ctypes) ---------------------------------------------------- */

typedef struct {
    const ld_params* p;
    const double* genomes;
    int64_t n;
    ld_record* records;
    double* final_state;
    int use_fma;
    atomic_llong* next;
    const double* init_state;
} od_job;

#define OD_CHUNK 16

static void* od_worker(void* arg) {
    od_job* j = (od_job*)arg;
    for (;;) {
        int64_t lo = atomic_fetch_add(j->next, OD_CHUNK);
        if (lo >= j->n) break;
        int64_t hi = lo + OD_CHUNK < j->n ? lo + OD_CHUNK : j->n;
        for (int64_t i = lo; i < hi; ++i) {
            od_drone d;
            fly(j->p, j->genomes + i * LD_GENOME_LEN, &d, NULL, j->use_fma, j->init_state ? j->init_state + i * LD_STATE_LEN : NULL);
            emit(&d, j->records + i, j->final_state ? j->final_state + i * LD_STATE_LEN : NULL);
        }
    }
    return NULL;
}

int ld_oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* population loop of evolve/evolution.py:64-73, fanned out over `nthreads` host threads
 * (<=0: all online cores); drones are independent so the result does not depend on it */
int ld_oracle_fly_mlp_from(const ld_params* p, const double* genomes, const double* init_state, int64_t n, ld_record* records,
                           double* final_state, int use_fma, int nthreads);

int ld_oracle_fly_mlp(const ld_params* p, const double* genomes, int64_t n, ld_record* records,
                      double* final_state, int use_fma, int nthreads) {
    return ld_oracle_fly_mlp_from(p, genomes, NULL, n, records, final_state, use_fma, nthreads);
}

int ld_oracle_fly_mlp_from(const ld_params* p, const double* genomes, const double* init_state, int64_t n, ld_record* records,
                           double* final_state, int use_fma, int nthreads) {
    if (!p || !genomes || !records || n < 0) return LD_ERR_INVALID_ARG;
    if (nthreads <= 0) nthreads = ld_oracle_max_threads();
    if (nthreads > 256) nthreads = 256;
    atomic_llong next = 0;
    od_job job = {p, genomes, n, records, final_state, use_fma, &next, init_state};
    pthread_t tid[256];
    int started = 0;
    for (int t = 1; t < nthreads; ++t)
        if (pthread_create(&tid[started], NULL, od_worker, &job) == 0) ++started;
    od_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
    return LD_OK;
}

int ld_oracle_trace_mlp(const ld_params* p, const double* genome, ld_record* record, double* trace, int use_fma) {
    if (!p || !genome || !record || !trace) return LD_ERR_INVALID_ARG;
    od_drone d;
    fly(p, genome, &d, trace, use_fma, NULL);
    emit(&d, record, NULL);
    return LD_OK;
}

/* ---- PID mode: flight/collect_coins.py:278-364 ---------------------------------------------- */
typedef struct {
    double kp, ki, kd, set, I, prev_e, prev_pv, lo, hi;
} od_pid;

/* drone/controller/pid.py:77-106 (calculate_error_sign), :132-134, :136-142 */
static double pid_calculate(od_pid* c, double pv, double dt, double range) {
    double error = c->set - pv;
    if (c->prev_e * error < 0) c->I = 0;
    if (fabs(error) < range) c->I += error;
    double diff = (c->prev_pv != 0) ? pv - c->prev_pv : 0;
    double out = c->kp * error + c->ki * c->I * dt - c->kd * diff / dt;
    c->prev_pv = pv;
    if (error != 0) c->prev_e = error;
    if (out < c->lo) out = c->lo;
    else if (out > c->hi) out = c->hi;
    return out;
}

static void fly_pid(const ld_params* p, const ld_pid_params* q, const double* gains, od_drone* d, double* trace, int use_fma) {
    drone_reset(p, d);
    od_pid pid[6]; /* x y z roll pitch yaw */
    for (int i = 0; i < 6; ++i) {
        pid[i].kp = gains[3 * i];
        pid[i].ki = gains[3 * i + 1];
        pid[i].kd = gains[3 * i + 2];
        pid[i].set = 0;
        pid[i].I = pid[i].prev_e = pid[i].prev_pv = 0;
        pid[i].lo = q->out_lo[i];
        pid[i].hi = q->out_hi[i];
    }
    static const double A[4][3] = {{1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {1, 1, 1}}; /* physical_model.py:64-66 */
    int target = 0;
    d->reward = q->timesteps; /* collect_coins.py:277 */
    for (int i = 0; i < 3; ++i) d->coin[i] = q->coin_pos[0][i];
    for (int t = 0; t < q->timesteps; ++t) {
        d->reward += q->step_reward; /* :279 */
        d->flight_time += 1;
        int next = target;
        double x[3] = {q->coin_pos[target][0] - d->pos[0], q->coin_pos[target][1] - d->pos[1], q->coin_pos[target][2] - d->pos[2]};
        if (dot_n(x, x, 3, use_fma) < q->coin_thr2[target]) { /* :284-287, sqrt(s) < r <=> s < thr2 */
            d->reward += q->coin_value[target];
            d->coins += 1;
            next = target + 1;
        }
        if (next >= q->n_coins) { /* :289-292 */
            d->status |= LD_STATUS_FINISHED;
            break;
        }
        target = next; /* :294-311 */
        for (int i = 0; i < 3; ++i) d->coin[i] = q->coin_pos[target][i];
        double lin[3], rot[3];
        accelerations(p, d->ang, d->thrust, lin, rot, use_fma); /* :328-330 */
        integrate(p, d, lin, rot);                              /* :335-341 */
        /* onboard_computer.py:50-85 */
        double lin_out[3], rot_out[3];
        for (int i = 0; i < 3; ++i) {
            pid[i].set = d->coin[i];
            lin_out[i] = pid_calculate(&pid[i], d->pos[i], p->dt, q->integral_range);
        }
        pid[3].set = -lin_out[1];
        pid[4].set = lin_out[0];
        pid[5].set = q->yaw_setpoint;
        for (int i = 0; i < 3; ++i) rot_out[i] = pid_calculate(&pid[3 + i], d->ang[i], p->dt, q->integral_range);
        /* onboard_computer.py:87-118 */
        double base = q->base_thrust + lin_out[2];
        double ratio = 1 / (cos(d->ang[0]) * cos(d->ang[1]));
        double bt = base * ratio;
        if (bt > 1) bt = 1;
        if (bt < 0) bt = 0;
        for (int j = 0; j < 4; ++j) {
            double desired = (A[j][0] * rot_out[0] + A[j][1] * rot_out[1]) + A[j][2] * rot_out[2];
            double th = bt + desired;
            if (th > 1) th = 1;
            if (th < 0) th = 0;
            d->thrust[j] = th;
        }
        if (trace) state_row(d, trace + (size_t)t * LD_STATE_LEN);
    }
    d->distance = coin_distance(d, use_fma);
}

typedef struct {
    const ld_params* p;
    const ld_pid_params* q;
    const double* gains;
    int64_t n;
    ld_record* records;
    double* final_state;
    int use_fma;
    atomic_llong* next;
} od_pid_job;

static void* od_pid_worker(void* arg) {
    od_pid_job* j = (od_pid_job*)arg;
    for (;;) {
        int64_t i = atomic_fetch_add(j->next, 1);
        if (i >= j->n) break;
        od_drone d;
        fly_pid(j->p, j->q, j->gains + i * 18, &d, NULL, j->use_fma);
        emit(&d, j->records + i, j->final_state ? j->final_state + i * LD_STATE_LEN : NULL);
    }
    return NULL;
}

int ld_oracle_fly_pid(const ld_params* p, const ld_pid_params* q, const double* gains, int64_t n, ld_record* records,
                      double* final_state, double* trace, int use_fma, int nthreads) {
    if (!p || !q || !gains || !records || n < 0 || q->n_coins < 1 || q->n_coins > LD_PID_MAX_COINS) return LD_ERR_INVALID_ARG;
    if (trace) {
        od_drone d;
        fly_pid(p, q, gains, &d, trace, use_fma);
        emit(&d, records, final_state);
        return LD_OK;
    }
    if (nthreads <= 0) nthreads = ld_oracle_max_threads();
    if (nthreads > 256) nthreads = 256;
    atomic_llong next = 0;
    od_pid_job job = {p, q, gains, n, records, final_state, use_fma, &next};
    pthread_t tid[256];
    int started = 0;
    for (int t = 1; t < nthreads; ++t)
        if (pthread_create(&tid[started], NULL, od_pid_worker, &job) == 0) ++started;
    od_pid_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
    return LD_OK;
}

/* ---- GA step: evolve/util.py:5-78, drone/drone.py:167-172 with the library's Philox stream ---- */
typedef struct { uint32_t v[4]; } od_u4;

static od_u4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    od_u4 o = {{c0, c1, c2, c3}};
    return o;
}

static double u53(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53; }

enum { OD_STREAM_SELECT = 1, OD_STREAM_MUTATE = 2, OD_STREAM_NOISE = 3 };

static const int od_layer_off[4] = {0, 160, 270, 314};

int ld_oracle_ga_step(const ld_ga_params* ga, const double* parents, const int64_t* rewards, int64_t n, double* children,
                      int64_t lo, int64_t hi, int32_t* picks /* [hi-lo][5] or NULL: parent, cut0..2 (-1 none), n_mutated */) {
    if (!ga || !parents || !rewards || !children || n < 2 || (n & 1) || lo < 0 || hi > n || (lo & 1) || (hi & 1)) return LD_ERR_INVALID_ARG;
    const uint32_t k0 = (uint32_t)ga->seed, k1 = (uint32_t)(ga->seed >> 32), gen = ga->generation;
    for (int64_t c = lo; c < hi; c += 2) {
        const uint32_t pair = (uint32_t)(c >> 1);
        od_u4 a = philox4x32_10(pair, 0, OD_STREAM_SELECT, gen, k0, k1), b = philox4x32_10(pair, 1, OD_STREAM_SELECT, gen, k0, k1);
        od_u4 t = philox4x32_10(pair, 2, OD_STREAM_SELECT, gen, k0, k1);
        int64_t idx[4] = {(int64_t)(u53(a.v[0], a.v[1]) * (double)n), (int64_t)(u53(a.v[2], a.v[3]) * (double)n),
                          (int64_t)(u53(b.v[0], b.v[1]) * (double)n), (int64_t)(u53(b.v[2], b.v[3]) * (double)n)};
        double r[2] = {u53(t.v[0], t.v[1]), u53(t.v[2], t.v[3])};
        int64_t win[2];
        for (int w = 0; w < 2; ++w) { /* util.py:30-45 */
            int64_t i1 = idx[2 * w], i2 = idx[2 * w + 1];
            int64_t better = rewards[i1] > rewards[i2] ? i1 : i2, worse = rewards[i1] > rewards[i2] ? i2 : i1;
            win[w] = r[w] < ga->p_tournament ? better : worse;
        }
        double* ch[2] = {children + (c - lo) * LD_GENOME_LEN, children + (c + 1 - lo) * LD_GENOME_LEN};
        memcpy(ch[0], parents + win[0] * LD_GENOME_LEN, LD_GENOME_LEN * sizeof(double));
        memcpy(ch[1], parents + win[1] * LD_GENOME_LEN, LD_GENOME_LEN * sizeof(double));
        int cuts[3] = {-1, -1, -1};
        for (int l = 0; l < 3; ++l) { /* util.py:48-78 */
            od_u4 x = philox4x32_10(pair, 3 + l, OD_STREAM_SELECT, gen, k0, k1);
            if (u53(x.v[0], x.v[1]) < ga->p_crossover) {
                int genes = od_layer_off[l + 1] - od_layer_off[l];
                int cut = (int)(u53(x.v[2], x.v[3]) * (double)genes);
                cuts[l] = cut;
                for (int g = od_layer_off[l] + cut; g < od_layer_off[l + 1]; ++g) {
                    double tmp = ch[0][g];
                    ch[0][g] = ch[1][g];
                    ch[1][g] = tmp;
                }
            }
        }
        /* drone.py:167-172 — per gene `random_sample() < rate` with the uniform v * 2^-32 taken from one Philox
         * block per group of 4 genes, a second block per mutated gene for its Box-Muller N(0, sigma) draw.
         * Pinned to the reference by draw replay: oracle/gen_ga_replay.py, tests/golden/ga_reference_replay.npz */
        for (int w = 0; w < 2; ++w) {
            int nm = 0;
            od_u4 m = {{0, 0, 0, 0}};
            for (int g = 0; g < LD_GENOME_LEN; ++g) {
                if (g % 4 == 0) m = philox4x32_10((uint32_t)(c + w), (uint32_t)(g / 4), OD_STREAM_MUTATE, gen, k0, k1);
                if ((double)m.v[g % 4] * 0x1.0p-32 < ga->mutation_rate) { /* random_sample() < mutation_rate, 32-bit uniform */
                    od_u4 x = philox4x32_10((uint32_t)(c + w), (uint32_t)g, OD_STREAM_NOISE, gen, k0, k1);
                    double u1 = ((double)x.v[0] + 0.5) * 0x1.0p-32, u2 = ((double)x.v[1] + 0.5) * 0x1.0p-32;
                    double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2); /* Box-Muller */
                    ch[w][g] = ch[w][g] + ga->mutation_sigma * z;
                    ++nm;
                }
            }
            if (picks) {
                int32_t* pk = picks + (c + w - lo) * 5;
                pk[0] = (int32_t)win[w]; pk[1] = cuts[0]; pk[2] = cuts[1]; pk[3] = cuts[2]; pk[4] = nm;
            }
        }
    }
    return LD_OK;
}

/* single calculate_accelerations() call, for unit tests against the reference */
int ld_oracle_accelerations(const ld_params* p, const double* angle3, const double* thrust4, double* lin3,
                            double* rot3, int use_fma) {
    accelerations(p, angle3, thrust4, lin3, rot3, use_fma);
    return LD_OK;
}
