// Latency microbenchmarks behind the resampling look-back design: how long is one round trip to L2 for the kinds of
// access a decoupled look-back uses (strong loads, atomics, a flag handshake between two SMs)?   nvcc -arch=sm_100a lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_cg(const unsigned long long* p) { unsigned long long v; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) { unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

template <int KIND>
__global__ void k_chase(unsigned long long* buf, int iters, long long* out) {
    // dependent chain of loads over a small ring that lives in L2 (every element holds the index of the next)
    unsigned long long i = threadIdx.x;
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
        if (KIND == 0) i = ld_volatile(buf + i);
        else if (KIND == 1) i = ld_relaxed_gpu(buf + i);
        else if (KIND == 2) i = ld_cg(buf + i);
        else if (KIND == 3) i = ld_acquire_gpu(buf + i);
        else i = atomicAdd(buf + i, 0ull);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)i; }
}
// two blocks on (very likely) different SMs bounce a flag: round trip of a publish + poll handshake
template <int KIND>
__global__ void k_pingpong(unsigned long long* flag, int iters, long long* out) {
    const int me = blockIdx.x;
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
        const unsigned long long want = 2ull * k + me;   // block 0 waits for even-ish values written by block 1, and vice versa
        if (me == 0) {
            if (KIND == 0) *(volatile unsigned long long*)flag = 2ull * k + 1; else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"(2ull * k + 1) : "memory");
            while ((KIND == 0 ? ld_volatile(flag) : ld_relaxed_gpu(flag)) != 2ull * k + 2) { }
        } else {
            while ((KIND == 0 ? ld_volatile(flag) : ld_relaxed_gpu(flag)) != 2ull * k + 1) { }
            if (KIND == 0) *(volatile unsigned long long*)flag = 2ull * k + 2; else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"(2ull * k + 2) : "memory");
        }
        (void)want;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && me == 0) out[0] = t1 - t0;
}
int main() {
    const int n = 1 << 14;   // 128 KB ring: L2 resident, larger than nothing; stride of 32 elements (256 B) between hops
    unsigned long long* h = new unsigned long long[n];
    for (int i = 0; i < n; ++i) h[i] = (i + 32 * 37) % n;
    unsigned long long *d, *flag; long long* out; long long ho[2];
    cudaMalloc(&d, n * 8); cudaMalloc(&flag, 8); cudaMalloc(&out, 16);
    cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
    const int iters = 4096;
    const char* names[] = {"ld.volatile", "ld.relaxed.gpu", "ld.global.cg", "ld.acquire.gpu", "atomicAdd(+0)"};
    for (int kind = 0; kind < 5; ++kind) {
        for (int rep = 0; rep < 2; ++rep) {
            if (kind == 0) k_chase<0><<<1, 1>>>(d, iters, out); else if (kind == 1) k_chase<1><<<1, 1>>>(d, iters, out);
            else if (kind == 2) k_chase<2><<<1, 1>>>(d, iters, out); else if (kind == 3) k_chase<3><<<1, 1>>>(d, iters, out);
            else { cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice); k_chase<4><<<1, 1>>>(d, iters, out); }
            cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
        }
        printf("%-16s dependent chain: %.1f cycles per access\n", names[kind], (double)ho[0] / iters);
    }
    for (int kind = 0; kind < 2; ++kind) {
        cudaMemset(flag, 0, 8);
        if (kind == 0) k_pingpong<0><<<2, 32>>>(flag, 2000, out); else k_pingpong<1><<<2, 32>>>(flag, 2000, out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ho, out, 8, cudaMemcpyDeviceToHost);
        printf("flag ping-pong (%s): %.1f cycles per round trip (two publish+observe hops) [%s]\n", kind == 0 ? "volatile" : "relaxed.gpu", (double)ho[0] / 2000, cudaGetErrorString(e));
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("SM clock attr %d kHz\n", clk);
    return 0;
}
