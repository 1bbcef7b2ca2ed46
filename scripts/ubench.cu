// ubench.cu — B200 micro-benchmarks that size the episode kernel (dev tool, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o scripts/ubench scripts/ubench.cu
// Measures, with 1 warp per SMSP as in the episode kernel:
//   dependent DFMA latency, DFMA issue interval with N independent chains,
//   LDS.128 latency / throughput at the kernel's row stride, tcgen05.ld (TMEM) throughput.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

template <int CHAINS>
__global__ void dfma_chains(double* out, long long* cyc, int iters) {
    double a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double m = 1.0 - 1e-9, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 123.0) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// LDS.128 at the kernel's layout: each lane its own row of ROWB bytes
template <int ROWB, int INFLIGHT>
__global__ void lds_bench(double* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* row = reinterpret_cast<double2*>(smem + (size_t)threadIdx.x * ROWB);
    for (int i = 0; i < ROWB / 16; ++i) row[i] = make_double2(threadIdx.x, i);
    __syncthreads();
    unsigned acc = 0;
    const int4* rowi = reinterpret_cast<const int4*>(row);
    int base = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        int4 v[INFLIGHT];
#pragma unroll
        for (int i = 0; i < INFLIGHT; ++i) v[i] = rowi[base + i];
        unsigned x[INFLIGHT];
#pragma unroll
        for (int i = 0; i < INFLIGHT; ++i) x[i] = v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
#pragma unroll
        for (int w = 1; w < INFLIGHT; w *= 2)
#pragma unroll
            for (int i = 0; i + w < INFLIGHT; i += 2 * w) x[i] ^= x[i + w];
        acc ^= x[0];
        base = (base + INFLIGHT) & 63;
    }
    long long t1 = clock64();
    if (acc == 123u) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// dependent LDS chain: address depends on loaded value -> pure latency
__global__ void lds_latency(int* out, long long* cyc, int iters) {
    __shared__ int ring[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) ring[i] = (i + 32) % 1024;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) p = ring[p];
    long long t1 = clock64();
    if (p == -1) out[0] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// ---- TMEM (tcgen05) -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NWARPS, int NLOADS>
__global__ void __launch_bounds__(NWARPS * 32, 1) tmem_bench(double* out, long long* cyc, int iters, int check) {
    __shared__ uint32_t tmem_base_holder;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_holder)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_base_holder;
    // this warp's lane quadrant: lanes 32*(warp%4) .. +31 ; address = (lane << 16) | column
    const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16);
    // fill 512 columns: value = lane*1000 + column
    for (int col = 0; col < 512; col += 16) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (threadIdx.x & 127) * 1000 + col + i;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr + col),
                     "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                     "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    __syncthreads();
    unsigned acc = 0;
    int bad = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int col = 0; col < 512; col += 32 * NLOADS) {
            uint32_t v[NLOADS][32];
#pragma unroll
            for (int l = 0; l < NLOADS; ++l) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[l][0]), "=r"(v[l][1]), "=r"(v[l][2]), "=r"(v[l][3]), "=r"(v[l][4]), "=r"(v[l][5]), "=r"(v[l][6]), "=r"(v[l][7]),
                      "=r"(v[l][8]), "=r"(v[l][9]), "=r"(v[l][10]), "=r"(v[l][11]), "=r"(v[l][12]), "=r"(v[l][13]), "=r"(v[l][14]),
                      "=r"(v[l][15]), "=r"(v[l][16]), "=r"(v[l][17]), "=r"(v[l][18]), "=r"(v[l][19]), "=r"(v[l][20]), "=r"(v[l][21]),
                      "=r"(v[l][22]), "=r"(v[l][23]), "=r"(v[l][24]), "=r"(v[l][25]), "=r"(v[l][26]), "=r"(v[l][27]), "=r"(v[l][28]),
                      "=r"(v[l][29]), "=r"(v[l][30]), "=r"(v[l][31])
                    : "r"(taddr + col + 32 * l));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int l = 0; l < NLOADS; ++l) {
                if (check) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (v[l][i] != (threadIdx.x & 127) * 1000u + col + 32 * l + i) ++bad;
                } else {
                    // touch every register with a shallow xor tree
                    unsigned x = 0;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) x ^= (v[l][i] ^ v[l][i + 1]) ^ (v[l][i + 2] ^ v[l][i + 3]);
                    acc ^= x;
                }
            }
        }
    }
    long long t1 = clock64();
    if (acc == 1) out[0] = 1.0;
    if (check && bad) printf("TMEM mismatch thread %d: %d bad\n", threadIdx.x, bad);
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
    double* out; long long* cyc; long long h;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, 64));
    const int iters = 2000;
#define RUN_DFMA(C)                                                                                          \
    dfma_chains<C><<<148, 96>>>(out, cyc, iters); CK(cudaDeviceSynchronize());                               \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));                                                      \
    printf("DFMA %2d chains, 1 warp/SMSP: %.2f cycles per DFMA (per chain step %.2f)\n", C, (double)h / (iters * 8.0 * C), (double)h / (iters * 8.0));
    RUN_DFMA(1) RUN_DFMA(2) RUN_DFMA(3) RUN_DFMA(4) RUN_DFMA(5) RUN_DFMA(6) RUN_DFMA(8) RUN_DFMA(10) RUN_DFMA(12)
    lds_latency<<<1, 32>>>((int*)out, cyc, 10000); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("LDS.32 dependent latency: %.1f cycles\n", (double)h / 10000);
    CK(cudaFuncSetAttribute(lds_bench<2416, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231936));
    CK(cudaFuncSetAttribute(lds_bench<2416, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231936));
    CK(cudaFuncSetAttribute(lds_bench<2416, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231936));
    CK(cudaFuncSetAttribute(lds_bench<2416, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231936));
#define RUN_LDS(F, T)                                                                                        \
    lds_bench<2416, F><<<148, T, 231936>>>(out, cyc, iters); CK(cudaDeviceSynchronize());                    \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));                                                      \
    printf("LDS.128 row-strided, %2d in flight, %d warps/SM: %.2f cycles per LDS.128 per warp\n", F, T / 32, (double)h / (iters * (double)F));
    RUN_LDS(1, 32) RUN_LDS(4, 32) RUN_LDS(8, 32) RUN_LDS(16, 32) RUN_LDS(1, 96) RUN_LDS(4, 96) RUN_LDS(8, 96) RUN_LDS(16, 96)
#define RUN_TMEM(W, NL, CHECK)                                                                               \
    tmem_bench<W, NL><<<148, W * 32>>>(out, cyc, 200, CHECK); CK(cudaDeviceSynchronize());                   \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));                                                      \
    printf("TMEM ld 32x32b.x32 (4 KB/warp each), %d warps/SM, %d loads per wait, check=%d: %.1f cycles per load per warp => %.1f B/clk/warp, %.1f B/clk/SM\n", W, NL, CHECK, \
           (double)h / (200 * 16.0), 4096.0 / ((double)h / (200 * 16.0)), W * 4096.0 / ((double)h / (200 * 16.0)));
    RUN_TMEM(1, 1, 1) RUN_TMEM(1, 1, 0) RUN_TMEM(1, 2, 0) RUN_TMEM(1, 4, 0) RUN_TMEM(2, 2, 0) RUN_TMEM(3, 2, 0) RUN_TMEM(4, 2, 1) RUN_TMEM(4, 2, 0) RUN_TMEM(4, 4, 0)
    printf("done\n");
    return 0;
}
