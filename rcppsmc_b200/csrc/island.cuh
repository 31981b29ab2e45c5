// Island-partitioned populations: one process per GPU, rank g owns global slots [g N_loc, (g+1) N_loc).
// The reference has no counterpart.  All data-path communication is done by the kernels themselves
// over NVLink peer memory (CUDA IPC-mapped slabs): three tiny all-gathers per step (log-sum-exp tuples, integer weight
// mass, zero-slot / surplus totals) through per-rank mailboxes, and the parents that live on another island are read
// straight from the owner's HBM inside the gather of k_step.  No NCCL call sits on the data path.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace smcb {

constexpr int ISL_MAX = 8;        // ranks on one NVSwitch box
constexpr int ISL_PAYLOAD = 16;   // doubles per message
enum { ISL_KIND_LSE = 0, ISL_KIND_MASS = 1, ISL_KIND_ZE = 2, ISL_KIND_OBS = 3, ISL_KIND_LSE2 = 4, ISL_KINDS = 5 };

struct IslMsg {                   // one mailbox slot (written by exactly one peer)
    unsigned long long seq;       // written last; the reader spins on it
    unsigned long long pad;
    double v[ISL_PAYLOAD];
};

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

// All-gather of `count` doubles per rank.  Called by all 32 lanes of ONE warp per rank; lane h talks to rank h.
// out[h][0..count) receives rank h's contribution (valid in every lane after the call for h < world: read via smem).
__device__ __forceinline__ void island_allgather(const IslandInfo& I, int kind, unsigned long long seq, const double* mine,
                                                 int count, double (*out)[ISL_PAYLOAD]) {
    const int lane = threadIdx.x & 31;
    if (lane < I.world) {
        IslMsg* dst = isl_slot(I, lane, kind, seq, I.rank);   // my slot in rank `lane`'s mailbox (peer store over NVLink)
        for (int i = 0; i < count; ++i) reinterpret_cast<volatile double*>(dst->v)[i] = mine[i];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(&dst->seq) = seq;
        const IslMsg* src = isl_slot(I, I.rank, kind, seq, lane);  // rank `lane`'s slot in my own mailbox
        while (*reinterpret_cast<const volatile unsigned long long*>(&src->seq) != seq) { __nanosleep(64); }
        __threadfence_system();
        for (int i = 0; i < count; ++i) out[lane][i] = reinterpret_cast<const volatile double*>(src->v)[i];
    }
    __syncwarp();
}

}  // namespace smcb
