// Counter-based RNG for the device engine: Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random
// numbers: as easy as 1, 2, 3", SC'11) written from the published round function.  It replaces the
// process-global R RNG the reference calls from inside user moves (reference call sites:
// src/pfnonlinbs.cpp:96,107; inst/include/sampler.h:810,819,830).  A draw is addressed by
// (seed; stream, step, global particle id, draw slot), so results do not depend on the launch geometry
// or on how many GPUs the population is split over.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#ifdef __CUDACC__
#include "fastmath.cuh"
#endif

namespace smcb {

struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
        lo = a * b;
        hi = __umulhi(a, b);
#else
        uint64_t p = (uint64_t)a * (uint64_t)b;
        lo = (uint32_t)p;
        hi = (uint32_t)(p >> 32);
#endif
    }

    // ctr = (c0,c1,c2,c3), key = (k0,k1); 10 rounds; out[4]
    __host__ __device__ static inline void rand4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(M0, c0, hi0, lo0);
            mulhilo(M1, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ k0;
            uint32_t n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += W0; k1 += W1;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// Stream identifiers (high half of counter word 3) keep the engine's own draws disjoint from the model's.
enum : uint32_t {
    STREAM_INIT = 1,      // pfInitialise draws
    STREAM_MOVE = 2,      // pfMove draws
    STREAM_MCMC = 3,      // pfMCMC draws
    STREAM_RESAMPLE = 4,  // resampling uniforms
    STREAM_USER = 8
};

// Two 64-bit lanes -> uniforms.  u_open in (0,1) with 52 random bits (safe for log); u_half in [0,1) with 53.
__host__ __device__ inline double u64_to_open01(uint64_t a) { return ((double)(a >> 12) + 0.5) * (1.0 / 4503599627370496.0); }
__host__ __device__ inline double u64_to_unit(uint64_t a) { return (double)(a >> 11) * (1.0 / 9007199254740992.0); }

struct Draw128 { uint64_t a, b; };

__host__ __device__ inline Draw128 philox_draw(uint64_t seed, uint32_t stream, uint32_t step, uint64_t id, uint32_t slot) {
    uint32_t o[4];
    Philox::rand4((uint32_t)id, (uint32_t)(id >> 32), step, (stream << 24) | (slot & 0xFFFFFFu),
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    Draw128 d;
    d.a = ((uint64_t)o[1] << 32) | o[0];
    d.b = ((uint64_t)o[3] << 32) | o[2];
    return d;
}

#ifdef __CUDACC__
// One Philox call -> two independent N(0,1) (Box-Muller in fp64).
__device__ inline void philox_normal2(uint64_t seed, uint32_t stream, uint32_t step, uint64_t pair_id, uint32_t slot,
                                      double& z0, double& z1) {
    Draw128 d = philox_draw(seed, stream, step, pair_id, slot);
    double u1 = u64_to_open01(d.a);
    double u2 = u64_to_unit(d.b);
    double r = fm_sqrt(fm_neg2log(u1));
    double s, c;
    fm_sincos2pi(u2, s, c);
    z0 = r * c;
    z1 = r * s;
}
#endif

}  // namespace smcb
