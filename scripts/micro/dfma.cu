// fp64 pipe of one SM: DFMA throughput and dependent-issue latency as a function of warps per SM and independent chains per
// thread.  (What bounds k_step: ~94 fp64 instructions per particle.)   nvcc -arch=sm_100a dfma.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double b = 0.999999, c = 1e-7;
    __syncthreads();
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 8 * 1024 * 148 * 2); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    printf("warps/SM  ILP  cycles per DFMA warp-instruction per SM sub-partition (one block per SM, all 148 SMs)\n");
    for (int warps : {4, 8, 16, 32}) {
        for (int ilp : {1, 2, 4, 8}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (ilp == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
                else if (ilp == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
                else if (ilp == 4) k<4><<<148, warps * 32>>>(out, iters, cyc);
                else k<8><<<148, warps * 32>>>(out, iters, cyc);
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            const double per_smsp = (double)iters * ilp * (warps / 4.0);   // warp instructions issued by one sub-partition
            printf("%5d %5d   %.2f   (chain latency seen by one warp: %.1f cycles per dependent DFMA)\n", warps, ilp, h / per_smsp, (double)h / iters);
        }
    }
    return 0;
}
