// Device resampling: weights -> children counts -> the reference's in-place ancestor map, in ONE pass.
//
// Replaces sampler::Resample (reference inst/include/sampler.h:779-868):
//   * cumulative sum (arma::cumsum, :811/:833)      -> single-pass decoupled look-back scan of the
//     normalised weights quantised to the 2^-52 grid (u64 integer scan: exact, order-independent)
//   * two-pointer count loops (:805-843)             -> closed form per parent k:
//         c(k) = #{ j < N : fl(j/N) < fl(W[k] - u_j) },   count[k] = c(k) - c(k-1)
//     evaluated with the reference's own floating-point predicate
//   * count -> index map (:846-857)                  -> second chained look-back scan (zero-count slots)
//     Z[s] = #{s' < s : count[s'] = 0};  E[k] = c(k-1) - k + Z[k] = surplus children of parents < k;
//     parent k writes itself to surplus[E[k] .. E[k]+count[k]-2]; a zero-count slot s takes surplus[Z[s]].
//   * replication (:860-864)                         -> NOT done here: the ancestor of slot s is decoded in
//     the next propagate kernel's gather from (zmask, zbase, surplus): 2 bits + 4 B per zero-count slot.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace smcb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 slots per tile
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

constexpr unsigned long long ST_AGG = 1ull << 62;
constexpr unsigned long long ST_PRE = 2ull << 62;
constexpr unsigned long long ST_VAL = (1ull << 62) - 1;
constexpr double TWO52 = 4503599627370496.0;
constexpr double INV_TWO52 = 1.0 / 4503599627370496.0;

// Ancestor encoding produced by a resampling pass over n slots.
struct AncestorMap {
    uint32_t* zmask;    // [ceil(n/32)] bit b of word g set  <=> slot 32g+b has no child of its own
    uint32_t* zbase;    // [ceil(n/32)] number of zero-count slots before slot 32g
    uint32_t* surplus;  // [n]          surplus[m] = parent of the m-th zero-count slot
};

__device__ __forceinline__ uint32_t decode_ancestor(const AncestorMap& am, long long slot) {
    uint32_t word = __ldg(am.zmask + (slot >> 5));
    uint32_t bit = (uint32_t)slot & 31u;
    if ((word >> bit) & 1u) {
        uint32_t rank = __ldg(am.zbase + (slot >> 5)) + __popc(word & ((1u << bit) - 1u));
        return __ldg(am.surplus + rank);
    }
    return (uint32_t)slot;
}

struct ScanWs {
    unsigned long long* statusA;  // [ntiles] weight-mass chain
    unsigned long long* statusB;  // [ntiles] zero-slot chain
    unsigned long long* statusC;  // [ntiles] surplus chain (count-map kernel only)
};

// Exclusive prefix of `agg` over tiles [0, tile) by decoupled look-back.  Called by all 32 lanes of warp 0.
__device__ __forceinline__ unsigned long long lookback_exclusive(volatile unsigned long long* status, int tile,
                                                                 unsigned long long agg) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) status[0] = ST_PRE | agg;
        return 0ull;
    }
    if (lane == 0) status[tile] = ST_AGG | agg;
    unsigned long long excl = 0ull;
    int base = tile - 1;
    while (true) {
        int idx = base - lane;
        unsigned long long v = ST_PRE;  // virtual predecessors of tile 0: prefix 0
        if (idx >= 0) v = status[idx];
        while (__any_sync(SMCB_FULL, (v >> 62) == 0ull)) {
            if (idx >= 0 && (v >> 62) == 0ull) v = status[idx];
        }
        unsigned pm = __ballot_sync(SMCB_FULL, (v >> 62) == 2ull);
        int first = pm ? (__ffs(pm) - 1) : 31;
        unsigned long long c = (lane <= first) ? (v & ST_VAL) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(SMCB_FULL, c, o);
        excl += c;
        if (pm) break;
        base -= 32;
    }
    if (lane == 0) status[tile] = ST_PRE | (excl + agg);
    return excl;
}

// #{ j in [0,n) : fl((double)j/(double)n) < v } -- the number of children the reference's two-pointer loop has
// handed out once the cumulative weight reaches W with v = fl(W - u) (sampler.h:814,836).
__device__ __forceinline__ long long children_below(double v, long long n, double nd, bool pow2) {
    if (!(v > 0.0)) return 0;
    double g = v * nd;
    double c = ceil(g);
    if (c >= nd) {
        if (pow2 || g - nd >= 9.5367431640625e-07) return n;
    } else if (pow2 || fabs(g - rint(g)) >= 9.5367431640625e-07) {
        return (long long)c;  // exact when n is a power of two; otherwise safely away from a rounding tie
    }
    long long j = (long long)fmin(c, nd);
    while (j < n && ((double)j / nd) < v) ++j;
    while (j > 0 && !(((double)(j - 1) / nd) < v)) --j;
    return j;
}

struct ScanArgs {
    const double* in;         // SRC_LOGW: log-weights; SRC_WEIGHTS: normalised weights
    long long n;
    const double* lse_ptr;    // SRC_LOGW: normalising constant (device scalar)
    const int* enable_ptr;    // if non-null and *enable_ptr == 0 the pass is skipped (ESS above threshold)
    const double* u_ptr;      // SYSTEMATIC: base uniform U in [0,1) (device scalar)
    const double* ustrat;     // STRATIFIED: injected uniforms U[step][0..n] (n+1 per step), or null -> Philox(seed, step)
    const long long* step_ptr;
    unsigned long long seed;
    unsigned int* ticket;     // dynamic tile counter; the caller zeroes it (and the status words) before each pass
    ScanWs ws;
    AncestorMap am;
    int* count_out;           // optional: children counts per parent
    unsigned int* tail_info;  // optional [2]: {sum of counts, number of zero-count slots}
};

enum { SRC_LOGW = 0, SRC_WEIGHTS = 1 };

template <int MODE>
__device__ __forceinline__ long long children_upto(unsigned long long cum, const ScanArgs& a, double nd, double inv_n,
                                                   double dRand, bool pow2, long long step) {
    double W = (double)cum * INV_TWO52;  // exact: cum < 2^53
    if (MODE == SYSTEMATIC) {
        return children_below(W - dRand, a.n, nd, pow2);
    } else {
        // STRATIFIED: child j compares against its own u_j = (1/N) U_j; only the stratum holding W is undecided.
        double g = W * nd;
        double fl = floor(g);
        long long js = (long long)fmin(fl, nd);
        double fr = g - fl;
        long long lo = js - 1 < 0 ? 0 : js - 1;
        long long hi = js + 1 > a.n - 1 ? a.n - 1 : js + 1;
        long long c = lo;
        for (long long j = lo; j <= hi; ++j) {
            bool sure_true = (j < js) && fr >= 9.5367431640625e-07;
            bool sure_false = (j > js) && fr <= 1.0 - 9.5367431640625e-07;
            bool p;
            if (sure_true) p = true;
            else if (sure_false) p = false;
            else {
                double U = a.ustrat ? a.ustrat[step * (a.n + 1) + j]
                                    : u64_to_unit(philox_draw(a.seed, STREAM_RESAMPLE, (uint32_t)step, (uint64_t)j, 0).a);
                double uj = 0.0 + (inv_n - 0.0) * U;
                p = (W - uj) > ((double)j / nd);
            }
            c += p ? 1 : 0;
        }
        return c;
    }
}

// One pass: weights -> (zmask, zbase, surplus).  Persistent blocks pull tiles from a ticket counter.
template <int SRC, int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) k_resample_scan(ScanArgs a) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    __shared__ unsigned long long s_q[SCAN_TILE + SCAN_TILE / 8];
    __shared__ unsigned long long s_warp[SCAN_WARPS];
    __shared__ int s_warpz[SCAN_WARPS];
    __shared__ unsigned long long s_bcast[2];
    __shared__ int s_tile;
    __shared__ unsigned int s_heavy_n;
    __shared__ unsigned int s_heavy[3 * 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n = a.n;
    const double nd = (double)n;
    const double inv_n = 1.0 / nd;
    const bool pow2 = (n & (n - 1)) == 0;
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    const double lse = (SRC == SRC_LOGW) ? *a.lse_ptr : 0.0;
    const double dRand = (MODE == SYSTEMATIC) ? (0.0 + (inv_n - 0.0) * (*a.u_ptr)) : 0.0;
    const long long step = a.step_ptr ? *a.step_ptr : 0;

    while (true) {
        if (tid == 0) { s_tile = (int)atomicAdd(a.ticket, 1u); s_heavy_n = 0; }
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const long long base = (long long)tile * SCAN_TILE;

        // ---- striped (coalesced) load, quantise, transpose to a blocked arrangement through shared memory
#pragma unroll
        for (int r = 0; r < SCAN_ITEMS; ++r) {
            int i = r * SCAN_THREADS + tid;
            long long gi = base + i;
            unsigned long long q = 0ull;
            if (gi < n) {
                double w = (SRC == SRC_LOGW) ? exp(a.in[gi] - lse) : a.in[gi];
                q = (unsigned long long)__double2ll_rn(w * TWO52);
            }
            s_q[i + (i >> 3)] = q;
        }
        __syncthreads();
        unsigned long long cum[SCAN_ITEMS];
        {
            unsigned long long run = 0ull;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) { run += s_q[tid * (SCAN_ITEMS + 1) + k]; cum[k] = run; }
        }
        // ---- block scan of thread totals
        unsigned long long tot = cum[SCAN_ITEMS - 1];
        unsigned long long inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(SMCB_FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned long long warp_off = 0ull, tile_agg = 0ull;
#pragma unroll
        for (int w = 0; w < SCAN_WARPS; ++w) { unsigned long long v = s_warp[w]; if (w < warp) warp_off += v; tile_agg += v; }
        if (warp == 0) {
            unsigned long long ex = lookback_exclusive(a.ws.statusA, tile, tile_agg);
            if (lane == 0) s_bcast[0] = ex;
        }
        __syncthreads();
        const unsigned long long thread_excl = s_bcast[0] + warp_off + (inc - tot);

        // ---- children counts from the closed form, zero-slot flags
        long long cprev = children_upto<MODE>(thread_excl, a, nd, inv_n, dRand, pow2, step);
        const long long cfirst = cprev;
        int cnt[SCAN_ITEMS];
        unsigned zbits = 0u;
        const long long g0 = base + (long long)tid * SCAN_ITEMS;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            long long ck = cprev;
            if (g0 + k < n) ck = children_upto<MODE>(thread_excl + cum[k], a, nd, inv_n, dRand, pow2, step);
            cnt[k] = (int)(ck - cprev);
            if (cnt[k] == 0 && g0 + k < n) zbits |= 1u << k;
            cprev = ck;
        }
        // ---- block scan of zero-slot counts, second look-back
        int ztot = __popc(zbits);
        int zinc = ztot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(SMCB_FULL, zinc, o);
            if (lane >= o) zinc += t;
        }
        if (lane == 31) s_warpz[warp] = zinc;
        __syncthreads();
        int zwarp_off = 0, ztile = 0;
#pragma unroll
        for (int w = 0; w < SCAN_WARPS; ++w) { int v = s_warpz[w]; if (w < warp) zwarp_off += v; ztile += v; }
        if (warp == 0) {
            unsigned long long ex = lookback_exclusive(a.ws.statusB, tile, (unsigned long long)ztile);
            if (lane == 0) s_bcast[1] = ex;
        }
        __syncthreads();
        const long long zexcl = (long long)s_bcast[1] + zwarp_off + (zinc - ztot);

        // ---- zero-slot bitmap: 4 threads x 8 slots = one 32-bit word
        {
            unsigned m = zbits << (8 * (lane & 3));
            m |= __shfl_xor_sync(SMCB_FULL, m, 1);
            m |= __shfl_xor_sync(SMCB_FULL, m, 2);
            if ((lane & 3) == 0 && g0 < n) {
                a.am.zmask[g0 >> 5] = m;
                a.am.zbase[g0 >> 5] = (uint32_t)zexcl;
            }
        }
        // ---- surplus list: parent k with count >= 2 owns surplus[E .. E + count - 2]
        {
            long long ck1 = cfirst;   // c(k-1)
            long long zk = zexcl;     // Z[k]
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                long long gk = g0 + k;
                int e = cnt[k] - 1;
                long long E = ck1 - gk + zk;
                if (e > 0) {
                    int small = e < 16 ? e : 16;
                    for (int r = 0; r < small; ++r) a.am.surplus[E + r] = (uint32_t)gk;
                }
                // heavy parents: the warp shares the rest of the writes
                unsigned hm = __ballot_sync(SMCB_FULL, e > 16);
                while (hm) {
                    int src = __ffs(hm) - 1;
                    hm &= hm - 1;
                    long long hE = __shfl_sync(SMCB_FULL, E, src) + 16;
                    int he = __shfl_sync(SMCB_FULL, e, src) - 16;
                    uint32_t hk = (uint32_t)__shfl_sync(SMCB_FULL, gk, src);
                    if (he > 4096) {  // very heavy: defer to the whole block
                        if (lane == 0) {
                            unsigned slot = atomicAdd(&s_heavy_n, 1u);
                            if (slot < 32) {
                                s_heavy[3 * slot] = hk; s_heavy[3 * slot + 1] = (unsigned)hE; s_heavy[3 * slot + 2] = (unsigned)he;
                                he = 0;
                            }
                        }
                        he = __shfl_sync(SMCB_FULL, he, 0);
                    }
                    for (int r = lane; r < he; r += 32) a.am.surplus[hE + r] = hk;
                }
                ck1 += cnt[k];
                zk += (zbits >> k) & 1u;
            }
        }
        if (a.count_out) {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) if (g0 + k < n) a.count_out[g0 + k] = cnt[k];
        }
        __syncthreads();
        {
            unsigned hn = s_heavy_n < 32 ? s_heavy_n : 32;
            for (unsigned h = 0; h < hn; ++h) {
                uint32_t hk = s_heavy[3 * h]; long long hE = s_heavy[3 * h + 1]; int he = (int)s_heavy[3 * h + 2];
                for (int r = tid; r < he; r += SCAN_THREADS) a.am.surplus[hE + r] = hk;
            }
        }
        // ---- last tile: unfilled zero-count slots copy particle 0 (reference: uRSIndices zero-initialised, :845)
        if (tile == ntiles - 1) {
            __shared__ long long s_tail[2];
            if (tid == SCAN_THREADS - 1) { s_tail[0] = cprev; s_tail[1] = zexcl + ztot; }
            // cprev of the last thread = c(n-1) (tail slots beyond n keep cprev), zero slots = Z total
            __syncthreads();
            long long ctot = s_tail[0], ztotal = s_tail[1];
            long long etot = ctot - n + ztotal;
            for (long long m = etot + tid; m < ztotal; m += SCAN_THREADS) a.am.surplus[m] = 0u;
            if (tid == 0) {
                if (a.tail_info) { a.tail_info[0] = (unsigned)ctot; a.tail_info[1] = (unsigned)ztotal; }
            }
        }
        __syncthreads();
    }
}

// ---- counts -> ancestor map (injected counts, multinomial / residual paths) -----------------------------
// Same map as above but the children counts are given: two independent look-back chains (zero slots, surplus).
struct CountMapArgs {
    const int* count;
    long long n;
    unsigned int* ticket;
    ScanWs ws;
    AncestorMap am;
    unsigned int* tail_info;
};

__global__ void __launch_bounds__(SCAN_THREADS) k_count_map(CountMapArgs a) {
    __shared__ int s_c[SCAN_TILE + SCAN_TILE / 8];
    __shared__ unsigned long long s_ws[SCAN_WARPS];
    __shared__ int s_wz[SCAN_WARPS];
    __shared__ unsigned long long s_bcast[2];
    __shared__ int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n = a.n;
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    while (true) {
        if (tid == 0) s_tile = (int)atomicAdd(a.ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const long long base = (long long)tile * SCAN_TILE;
#pragma unroll
        for (int r = 0; r < SCAN_ITEMS; ++r) {
            int i = r * SCAN_THREADS + tid;
            long long gi = base + i;
            s_c[i + (i >> 3)] = gi < n ? a.count[gi] : 1;  // tail slots: count 1 -> neither zero nor surplus
        }
        __syncthreads();
        int cnt[SCAN_ITEMS];
        unsigned zbits = 0u;
        unsigned long long stot = 0ull;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            cnt[k] = s_c[tid * (SCAN_ITEMS + 1) + k];
            if (cnt[k] <= 0) zbits |= 1u << k;
            if (cnt[k] > 1) stot += (unsigned long long)(cnt[k] - 1);
        }
        int ztot = __popc(zbits), zinc = ztot;
        unsigned long long sinc = stot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(SMCB_FULL, zinc, o);
            unsigned long long u = __shfl_up_sync(SMCB_FULL, sinc, o);
            if (lane >= o) { zinc += t; sinc += u; }
        }
        if (lane == 31) { s_wz[warp] = zinc; s_ws[warp] = sinc; }
        __syncthreads();
        int zwo = 0, ztile = 0;
        unsigned long long swo = 0ull, stile = 0ull;
#pragma unroll
        for (int w = 0; w < SCAN_WARPS; ++w) {
            int v = s_wz[w]; unsigned long long u = s_ws[w];
            if (w < warp) { zwo += v; swo += u; }
            ztile += v; stile += u;
        }
        if (warp == 0) {
            unsigned long long ez = lookback_exclusive(a.ws.statusB, tile, (unsigned long long)ztile);
            unsigned long long es = lookback_exclusive(a.ws.statusC, tile, stile);
            if (lane == 0) { s_bcast[0] = ez; s_bcast[1] = es; }
        }
        __syncthreads();
        long long zk = (long long)s_bcast[0] + zwo + (zinc - ztot);
        long long E = (long long)s_bcast[1] + (long long)swo + (long long)(sinc - stot);
        const long long g0 = base + (long long)tid * SCAN_ITEMS;
        {
            unsigned m = zbits << (8 * (lane & 3));
            m |= __shfl_xor_sync(SMCB_FULL, m, 1);
            m |= __shfl_xor_sync(SMCB_FULL, m, 2);
            if ((lane & 3) == 0 && g0 < n) { a.am.zmask[g0 >> 5] = m; a.am.zbase[g0 >> 5] = (uint32_t)zk; }
        }
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            long long gk = g0 + k;
            int e = cnt[k] - 1;
            if (e > 0) {
                int small = e < 16 ? e : 16;
                for (int r = 0; r < small; ++r) if (E + r < n) a.am.surplus[E + r] = (uint32_t)gk;
            }
            unsigned hm = __ballot_sync(SMCB_FULL, e > 16);
            while (hm) {
                int src = __ffs(hm) - 1;
                hm &= hm - 1;
                long long hE = __shfl_sync(SMCB_FULL, E, src) + 16;
                int he = __shfl_sync(SMCB_FULL, e, src) - 16;
                uint32_t hk = (uint32_t)__shfl_sync(SMCB_FULL, gk, src);
                for (int r = lane; r < he; r += 32) if (hE + r < n) a.am.surplus[hE + r] = hk;
            }
            if (e > 0) E += e;
        }
        if (tile == ntiles - 1) {
            __shared__ long long s_tail;
            if (tid == SCAN_THREADS - 1) s_tail = E;
            __syncthreads();
            // zero-count slots beyond the available surplus children copy particle 0 (sampler.h:845)
            long long etot = s_tail;
            long long ztotal = (long long)s_bcast[0] + ztile;
            for (long long m = etot + tid; m < ztotal; m += SCAN_THREADS) a.am.surplus[m] = 0u;
            if (tid == 0 && a.tail_info) { a.tail_info[0] = (unsigned)etot; a.tail_info[1] = (unsigned)ztotal; }
        }
        __syncthreads();
    }
}

// Materialise uRSIndices (sampler.h:845-857 layout) from the ancestor map.
__global__ void k_decode_indices(AncestorMap am, long long n, uint32_t* idx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = decode_ancestor(am, i);
}

__global__ void k_zero_u64(unsigned long long* p, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0ull;
}

}  // namespace smcb
