// Island-partitioned populations: one process per GPU, rank g owns global slots [g N_loc, (g+1) N_loc).
// The reference has no counterpart.  All data-path communication is done by the kernels themselves
// over NVLink peer memory (CUDA IPC-mapped slabs): three tiny all-gathers per step (log-sum-exp tuples, integer weight
// mass, zero-slot / surplus totals) through per-rank mailboxes, and the parents that live on another island are read
// straight from the owner's HBM inside the gather of k_step.  No NCCL call sits on the data path.
// Only the first all-gather is waited for where it is sent (the last block of k_step needs the global sums to close the
// step).  The other two are posted by the last block of the producing kernel and picked up by the blocks of the CONSUMING
// kernel (k_rs_count reads the lower ranks' masses, the next k_step reads the zero-slot / surplus totals), so no kernel
// ends with a serial wait for a peer.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace smcb {

constexpr int ISL_MAX = 8;        // ranks on one NVSwitch box
constexpr int ISL_PAYLOAD = 16;   // doubles per message
enum { ISL_KIND_LSE = 0, ISL_KIND_MASS = 1, ISL_KIND_ZE = 2, ISL_KIND_OBS = 3, ISL_KIND_LSE2 = 4, ISL_KINDS = 5 };

struct IslMsg {                   // one mailbox slot (written by exactly one peer)
    ulonglong2 e[ISL_PAYLOAD];    // .x = the bits of one double, .y = the sequence number of the message it belongs to.  Each
};                                // entry is written and read as ONE aligned 16-byte access, so a reader that sees the expected
                                  // sequence number has the value that came with it: no fence between payload and flag, and the
                                  // sequence number is the full 64-bit one (no stale match after wrap-around)

struct IslandInfo {
    int rank, world;
    long long n_loc, n_glob;
    unsigned char* base[ISL_MAX];  // slab base of every rank in THIS process's address space
    size_t off_xa[4], off_xb[4], off_surplus, off_mailbox;
    unsigned long long epoch;      // run counter, makes mailbox sequence numbers unique across runs
};

__device__ __forceinline__ IslMsg* isl_slot(const IslandInfo& I, int owner, int kind, unsigned long long seq, int src) {
    IslMsg* mb = reinterpret_cast<IslMsg*>(I.base[owner] + I.off_mailbox);
    return mb + ((kind * 2 + (int)(seq & 1ull)) * ISL_MAX + src);
}

__device__ __forceinline__ void isl_st16(ulonglong2* p, unsigned long long x, unsigned long long y) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}
__device__ __forceinline__ ulonglong2 isl_ld16(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}

// Posts `count` doubles (the same values in every lane of the calling warp) into this rank's slot of every rank's mailbox
// (peer stores over NVLink, fire and forget).  Called by all 32 lanes of ONE warp per rank; lane i carries value i.
__device__ __forceinline__ void island_send(const IslandInfo& I, int kind, unsigned long long seq, const double* mine, int count) {
    const int lane = threadIdx.x & 31;
    double x = 0.0;
#pragma unroll
    for (int i = 0; i < ISL_PAYLOAD; ++i) if (i < count && i == lane) x = mine[i];
    if (lane < count)
        for (int h = 0; h < I.world; ++h)
            isl_st16(isl_slot(I, h, kind, seq, I.rank)->e + lane, (unsigned long long)__double_as_longlong(x), seq);
}
// Value i of rank src's message `seq` of this kind, from this rank's own mailbox (local memory): spins until it has arrived.
__device__ __forceinline__ double island_recv_one(const IslandInfo& I, int kind, unsigned long long seq, int src, int i) {
    const ulonglong2* p = isl_slot(I, I.rank, kind, seq, src)->e + i;
    ulonglong2 v = isl_ld16(p);
    while (v.y != seq) v = isl_ld16(p);
    return __longlong_as_double((long long)v.x);
}
// out[h][0..count) = rank h's message, for every rank.  One warp.
__device__ __forceinline__ void island_recv(const IslandInfo& I, int kind, unsigned long long seq, int count, double (*out)[ISL_PAYLOAD]) {
    const int lane = threadIdx.x & 31;
    for (int idx = lane; idx < I.world * count; idx += 32) {
        const int h = idx / count, i = idx - h * count;
        out[h][i] = island_recv_one(I, kind, seq, h, i);
    }
    __syncwarp();
}
// All-gather of `count` doubles per rank (send + receive); out[h][0..count) receives rank h's contribution.
__device__ __forceinline__ void island_allgather(const IslandInfo& I, int kind, unsigned long long seq, const double* mine,
                                                 int count, double (*out)[ISL_PAYLOAD]) {
    island_send(I, kind, seq, mine, count);
    island_recv(I, kind, seq, count, out);
}
// Sum of value 0 of the messages of the ranks BELOW this one (all lanes of the calling warp get it): where this rank's
// cumulative weight starts, from the all-gathered island masses.
__device__ __forceinline__ double island_recv_lower_sum(const IslandInfo& I, int kind, unsigned long long seq) {
    const int lane = threadIdx.x & 31;
    double v = lane < I.rank ? island_recv_one(I, kind, seq, lane, 0) : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // ISL_MAX = 8 ranks: lanes 0..7 (exact: integers < 2^53)
    return __shfl_sync(0xffffffffu, v, 0);
}
// Exclusive prefixes over the ranks of the all-gathered (zero-count slots, surplus-list entries): zoff[h], eoff[h] for
// h = 0..world (entry [world] = totals).  One warp; the tables usually live in shared memory.
__device__ __forceinline__ void island_ze_offsets(const IslandInfo& I, unsigned long long seq, uint32_t* zoff, uint32_t* eoff) {
    const int lane = threadIdx.x & 31;
    uint32_t z = 0u, e = 0u;
    if (lane < I.world) {
        z = (uint32_t)island_recv_one(I, ISL_KIND_ZE, seq, lane, 0);
        e = (uint32_t)island_recv_one(I, ISL_KIND_ZE, seq, lane, 1);
    }
    uint32_t zi = z, ei = e;
#pragma unroll
    for (int o = 1; o < ISL_MAX; o <<= 1) {
        const uint32_t tz = __shfl_up_sync(0xffffffffu, zi, o), te = __shfl_up_sync(0xffffffffu, ei, o);
        if (lane >= o) { zi += tz; ei += te; }
    }
    if (lane < I.world) { zoff[lane] = zi - z; eoff[lane] = ei - e; }
    if (lane == I.world - 1) { zoff[I.world] = zi; eoff[I.world] = ei; }
    __syncwarp();
}

}  // namespace smcb
