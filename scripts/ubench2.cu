// ubench2.cu — per-warp shared-memory load throughput by instruction type, and LDS+TMEM overlap (dev tool)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: LDS.32 lane-contiguous, 1: LDS.64 contiguous, 2: LDS.128 contiguous (lane*16), 3: LDS.128 row-strided 2416,
// 4: LDSM.x4 (each lane supplies a 16-B row address, contiguous), 5: LDS.128 row-strided 2512
template <int MODE, int INFLIGHT>
__global__ void lds_kind(unsigned* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    for (int i = threadIdx.x; i < 231936 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(smem)[i] = i * 2654435761u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned acc = 0;
    uint32_t base;
    int step;
    if (MODE == 0) { base = smem_u32(smem) + warp * 8192 + lane * 4; step = 128; }
    else if (MODE == 1) { base = smem_u32(smem) + warp * 16384 + lane * 8; step = 256; }
    else if (MODE == 2 || MODE == 4) { base = smem_u32(smem) + warp * 32768 + lane * 16; step = 512; }
    else if (MODE == 3) { base = smem_u32(smem) + threadIdx.x * 2416; step = 16; }
    else { base = smem_u32(smem) + (threadIdx.x % 92) * 2512; step = 16; }
    int off = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        unsigned x[INFLIGHT];
#pragma unroll
        for (int i = 0; i < INFLIGHT; ++i) {
            uint32_t a = base + (off + i) * step;
            if (MODE == 0) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); x[i] = v; }
            else if (MODE == 1) { unsigned v0, v1; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(a)); x[i] = v0 ^ v1; }
            else if (MODE == 4) { unsigned v0, v1, v2, v3; asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a)); x[i] = (v0 ^ v1) ^ (v2 ^ v3); }
            else { unsigned v0, v1, v2, v3; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a)); x[i] = (v0 ^ v1) ^ (v2 ^ v3); }
        }
#pragma unroll
        for (int w = 1; w < INFLIGHT; w *= 2)
#pragma unroll
            for (int i = 0; i + w < INFLIGHT; i += 2 * w) x[i] ^= x[i + w];
        acc ^= x[0];
        off = (off + INFLIGHT) & 31;
    }
    long long t1 = clock64();
    if (acc == 123u) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// LDS.128 row-strided and tcgen05.ld x16 interleaved in the same warp: are the two paths additive?
template <int NWARPS, int DO_LDS, int DO_TMEM>
__global__ void __launch_bounds__(NWARPS * 32, 1) mix_bench(unsigned* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t holder;
    for (int i = threadIdx.x; i < 200000 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(smem)[i] = i * 2654435761u;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&holder)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = holder + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t base = smem_u32(smem) + threadIdx.x * 1552; // 97 x 16 B rows
    unsigned acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // one "block": 8 LDS.128 (4 KB/warp) and/or two tcgen05.ld x16 (2x2 KB/warp... = 4 KB)
        unsigned v[8][4];
        unsigned t[2][16];
        if (DO_TMEM) {
#pragma unroll
            for (int l = 0; l < 2; ++l)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(t[l][0]), "=r"(t[l][1]), "=r"(t[l][2]), "=r"(t[l][3]), "=r"(t[l][4]), "=r"(t[l][5]), "=r"(t[l][6]), "=r"(t[l][7]),
                               "=r"(t[l][8]), "=r"(t[l][9]), "=r"(t[l][10]), "=r"(t[l][11]), "=r"(t[l][12]), "=r"(t[l][13]), "=r"(t[l][14]), "=r"(t[l][15])
                             : "r"(taddr + ((it * 32 + l * 16) & 511)));
        }
        if (DO_LDS) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3]) : "r"(base + ((it * 8 + i) % 97) * 16));
        }
        if (DO_TMEM) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int l = 0; l < 2; ++l)
#pragma unroll
                for (int i = 0; i < 16; i += 4) acc ^= (t[l][i] ^ t[l][i + 1]) ^ (t[l][i + 2] ^ t[l][i + 3]);
        }
        if (DO_LDS) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc ^= (v[i][0] ^ v[i][1]) ^ (v[i][2] ^ v[i][3]);
        }
    }
    long long t1 = clock64();
    if (acc == 123u) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(holder));
}

int main() {
    unsigned* out; long long* cyc; long long h;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, 64));
    const int iters = 4000;
#define RUN(MODE, F, T, NAME, BYTES)                                                                          \
    CK(cudaFuncSetAttribute(lds_kind<MODE, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231936));       \
    lds_kind<MODE, F><<<148, T, 231936>>>(out, cyc, iters); CK(cudaDeviceSynchronize());                    \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));                                                      \
    printf("%-28s %2d in flight, %d warps/SM: %6.2f cyc/instr/warp  %6.1f B/clk/warp  %6.1f B/clk/SM\n", NAME, F, T / 32, (double)h / (iters * (double)F), BYTES / ((double)h / (iters * (double)F)), (T / 32) * BYTES / ((double)h / (iters * (double)F)));
    RUN(0, 8, 32, "LDS.32 contiguous", 128.0) RUN(1, 8, 32, "LDS.64 contiguous", 256.0) RUN(2, 8, 32, "LDS.128 contiguous", 512.0)
    RUN(3, 8, 32, "LDS.128 row-stride 2416", 512.0) RUN(5, 8, 32, "LDS.128 row-stride 2512", 512.0) RUN(4, 8, 32, "LDSM.x4", 512.0)
    RUN(0, 8, 96, "LDS.32 contiguous", 128.0) RUN(1, 8, 96, "LDS.64 contiguous", 256.0) RUN(2, 8, 96, "LDS.128 contiguous", 512.0)
    RUN(3, 8, 96, "LDS.128 row-stride 2416", 512.0) RUN(4, 8, 96, "LDSM.x4", 512.0)
    RUN(2, 8, 128, "LDS.128 contiguous", 512.0) RUN(4, 8, 128, "LDSM.x4", 512.0) RUN(4, 16, 128, "LDSM.x4", 512.0)
#define RUNMIX(W, L, T)                                                                                       \
    CK(cudaFuncSetAttribute(mix_bench<W, L, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));      \
    mix_bench<W, L, T><<<148, W * 32, 200000>>>(out, cyc, iters); CK(cudaDeviceSynchronize());              \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));                                                      \
    printf("mix %d warps: LDS=%d (4KB/warp/iter) TMEM=%d (4KB/warp/iter): %.1f cycles per iter => %.1f B/clk/warp\n", W, L, T, (double)h / iters, (L + T) * 4096.0 / ((double)h / iters));
    RUNMIX(1, 1, 0) RUNMIX(1, 0, 1) RUNMIX(1, 1, 1) RUNMIX(4, 1, 0) RUNMIX(4, 0, 1) RUNMIX(4, 1, 1)
    printf("done\n");
    return 0;
}
