// ubench3.cu — does a DFMA whose active lanes sit in ONE half of the warp occupy the fp64 pipe for one cycle
// instead of two?  (dev tool; if it did, 2 warps of 16 active lanes per SM sub-partition could hide latency at the
// same on-chip brain capacity.)  Also: fp64 issue rate with 1 / 2 warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

template <int CHAINS>
__global__ void dfma(double* out, long long* cyc, int iters, unsigned active_mask) {
    const unsigned lane = threadIdx.x & 31u;
    double a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double m = 1.0 - 1e-9, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    if ((active_mask >> lane) & 1u) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, c);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double* out; long long* cyc;
    CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, 8));
    const int iters = 2000;
    struct { const char* name; int threads; unsigned mask; } cases[] = {
        {"1 warp/SMSP, 32 lanes", 128, 0xffffffffu}, {"1 warp/SMSP, lanes 0-15", 128, 0x0000ffffu}, {"1 warp/SMSP, even lanes", 128, 0x55555555u},
        {"2 warps/SMSP, 32 lanes", 256, 0xffffffffu}, {"2 warps/SMSP, lanes 0-15", 256, 0x0000ffffu}, {"2 warps/SMSP, lanes 0-7", 256, 0x000000ffu},
        {"4 warps/SMSP, lanes 0-15", 512, 0x0000ffffu}};
    for (auto& cs : cases) {
        long long h = 0;
        dfma<8><<<148, cs.threads>>>(out, cyc, iters, cs.mask);
        CK(cudaDeviceSynchronize());
        dfma<8><<<148, cs.threads>>>(out, cyc, iters, cs.mask);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double warps_per_smsp = cs.threads / 128.0;
        const double instr = (double)iters * 8 * 8 * warps_per_smsp; // warp-instructions per sub-partition
        printf("%-28s: %8lld cycles, %.3f cycles per warp-DFMA per sub-partition\n", cs.name, h, h / instr);
    }
    return 0;
}
