// Issue rate of a DFMA / integer mix like k_step's (36 % fp64): what IPC can one SM sub-partition sustain with 3-4 warps?
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NI>
__global__ void k(double* out, int iters, long long* cyc, unsigned seed) {
    double a[NF > 0 ? NF : 1]; unsigned u[NI > 0 ? NI : 1];
#pragma unroll
    for (int i = 0; i < NF; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < NI; ++i) u[i] = seed + threadIdx.x * 7 + i;
    const double b = 0.999999, c = 1e-7;
    __syncthreads();
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int i = 0; i < (NF > NI ? NF : NI); ++i) {
            if (i < NF) a[i] = fma(a[i], b, c);
            if (i < NI) u[i] = u[i] * 0xD2511F53u + (u[i] >> 7);   // IMAD + SHF-ish dependent chain (3 instr)
        }
    }
    long long t1 = clock64();
    double s = 0; unsigned v = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < NI; ++i) v ^= u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + v;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NF, int NI> void run(double* out, long long* cyc, int warps) {
    long long h; const int iters = 2048;
    for (int rep = 0; rep < 2; ++rep) { k<NF, NI><<<148, warps * 32>>>(out, iters, cyc, 12345u); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); }
    const double instr = (double)iters * (NF + 2.0 * NI) * (warps / 4.0);   // per sub-partition (the integer step compiles to 2 instructions: SHF + IMAD)
    printf("warps/SM %2d  DFMA chains %d  int chains %d : %.2f instr/cycle per sub-partition (fp64 share %.0f %%), fp64 pipe %.0f %% busy\n", warps, NF, NI,
           instr / h, 100.0 * NF / (NF + 2.0 * NI), 100.0 * 2.0 * iters * NF * (warps / 4.0) / h);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024 * 148); cudaMalloc(&cyc, 8);
    for (int warps : {12, 16, 24}) {
        run<4, 0>(out, cyc, warps); run<4, 2>(out, cyc, warps); run<4, 4>(out, cyc, warps); run<4, 8>(out, cyc, warps); run<2, 4>(out, cyc, warps); run<0, 8>(out, cyc, warps);
    }
    return 0;
}
