// Device resampling: weights -> children counts -> the reference's in-place ancestor map.
//
// Replaces sampler::Resample (reference inst/include/sampler.h:779-868):
//   * cumulative sum (arma::cumsum, :811/:833)      -> exact scan of the normalised weights quantised to the 2^-52 grid
//     (sums of such values below 2 are exact in fp64 in any order, so a parallel scan equals the sequential one bit for bit)
//   * two-pointer count loops (:805-843)             -> closed form per parent k:
//         c(k) = #{ j < N : fl(j/N) < fl(W[k] - u_j) },   count[k] = c(k) - c(k-1)
//     evaluated with the reference's own floating-point predicate
//   * count -> index map (:846-857)                  -> Z[s] = #{s' < s : count[s'] = 0};
//     E[k] = c(k-1) - k + Z[k] = surplus children of parents < k;
//     parent k writes itself to surplus[E[k] .. E[k]+count[k]-2]; a zero-count slot s takes surplus[Z[s]].
//   * replication (:860-864)                         -> NOT done here: the ancestor of slot s is decoded in
//     the next propagate kernel's gather from (zmask, zbase, surplus): 2 bits + 4 B per zero-count slot.
// Two implementations of the SYSTEMATIC / STRATIFIED path: three streaming passes without inter-block waiting (k_rs_mass,
// k_rs_count, k_rs_map: the production path, second half of this file) and ONE pass with decoupled look-back scans
// (k_resample_scan: round 1, kept for A/B measurements and as the injected-counts surface k_count_map).  MULTINOMIAL / RESIDUAL:
// the engine's native definitions (integer scan, iid draws through the exact integer inverse CDF, counts -> map).
#pragma once
#include <functional>

#include "common.cuh"
#include "philox.cuh"
#include "island.cuh"

namespace smcb {

#ifndef SMCB_SCAN_THREADS
#define SMCB_SCAN_THREADS 128
#endif
#ifndef SMCB_SCAN_MINB
#define SMCB_SCAN_MINB 6
#endif
constexpr int SCAN_THREADS = SMCB_SCAN_THREADS;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 1024 slots per tile
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

// Exclusive prefix of `agg` over tiles [0, tile) by decoupled look-back.  Called by all 32 lanes of one warp.
// Each round keeps LB_DEPTH status words per lane in flight (a 32*LB_DEPTH-tile window per L2 round trip) and the
// lanes accumulate privately; the only cross-lane reduction happens once, after the nearest full prefix is found.
#ifndef SMCB_LB_DEPTH
#define SMCB_LB_DEPTH 4
#endif
#ifndef SMCB_LB_SLEEP
#define SMCB_LB_SLEEP 0
#endif
constexpr int LB_DEPTH = SMCB_LB_DEPTH;
#ifdef SMCB_PROFILE_LOOKBACK
__device__ unsigned long long g_lb_stats[8];  // [0] cycles, [1] rounds, [2] calls, [3] stalled polls (experiment build)
#endif
__device__ __forceinline__ unsigned long long lookback_exclusive(volatile unsigned long long* status, int tile,
                                                                 unsigned long long agg) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) status[0] = ST_PRE | agg;
        return 0ull;
    }
    if (lane == 0) status[tile] = ST_AGG | agg;
    unsigned long long acc = 0ull;
    int base = tile - 1;
    bool done = false;
#ifdef SMCB_PROFILE_LOOKBACK
    long long t0 = clock64(); unsigned rounds = 0, stalled = 0;
#endif
    while (!done) {
#ifdef SMCB_PROFILE_LOOKBACK
        ++rounds;
#endif
        unsigned long long v[LB_DEPTH];
#pragma unroll
        for (int r = 0; r < LB_DEPTH; ++r) {
            int idx = base - 32 * r - lane;
            v[r] = 2ull << 62;  // virtual predecessors of tile 0: prefix 0
            if (idx >= 0) v[r] = status[idx];
        }
#pragma unroll
        for (int r = 0; r < LB_DEPTH; ++r) {
            if (done) break;
            if (__any_sync(SMCB_FULL, (v[r] >> 62) == 0ull)) {        // window not published yet: poll again from here
                if (SMCB_LB_SLEEP > 0) __nanosleep(SMCB_LB_SLEEP);
#ifdef SMCB_PROFILE_LOOKBACK
                ++stalled;
#endif
                break;
            }
            unsigned pm = __ballot_sync(SMCB_FULL, (v[r] >> 62) == 2ull);
            int first = pm ? (__ffs(pm) - 1) : 31;
            if (lane <= first) acc += v[r] & ST_VAL;
            base -= 32;
            if (pm) done = true;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SMCB_FULL, acc, o);
    if (lane == 0) status[tile] = ST_PRE | (acc + agg);
#ifdef SMCB_PROFILE_LOOKBACK
    if (lane == 0) {
        atomicAdd(&g_lb_stats[0], (unsigned long long)(clock64() - t0)); atomicAdd(&g_lb_stats[1], (unsigned long long)rounds);
        atomicAdd(&g_lb_stats[2], 1ull); atomicAdd(&g_lb_stats[3], (unsigned long long)stalled);
    }
#endif
    return acc;
}

// Same, for a tile whose aggregate was published earlier (pipelined kernel): walks the predecessors, then publishes the
// inclusive prefix.  Returns the exclusive prefix.
__device__ __forceinline__ unsigned long long lookback_resolve(volatile unsigned long long* status, int tile, unsigned long long agg) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) return 0ull;
    unsigned long long acc = 0ull;
    int base = tile - 1;
    bool done = false;
#ifdef SMCB_PROFILE_LOOKBACK
    long long t0 = clock64(); unsigned rounds = 0, stalled = 0;
#endif
    while (!done) {
#ifdef SMCB_PROFILE_LOOKBACK
        ++rounds;
#endif
        unsigned long long v[LB_DEPTH];
#pragma unroll
        for (int r = 0; r < LB_DEPTH; ++r) {
            int idx = base - 32 * r - lane;
            v[r] = 2ull << 62;  // virtual predecessors of tile 0: prefix 0
            if (idx >= 0) v[r] = status[idx];
        }
#pragma unroll
        for (int r = 0; r < LB_DEPTH; ++r) {
            if (done) break;
            if (__any_sync(SMCB_FULL, (v[r] >> 62) == 0ull)) {        // window not published yet: poll again from here
#ifdef SMCB_PROFILE_LOOKBACK
                ++stalled;
#endif
                break;
            }
            unsigned pm = __ballot_sync(SMCB_FULL, (v[r] >> 62) == 2ull);
            int first = pm ? (__ffs(pm) - 1) : 31;
            if (lane <= first) acc += v[r] & ST_VAL;
            base -= 32;
            if (pm) done = true;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SMCB_FULL, acc, o);
    if (lane == 0) status[tile] = ST_PRE | (acc + agg);
#ifdef SMCB_PROFILE_LOOKBACK
    if (lane == 0) {
        atomicAdd(&g_lb_stats[0], (unsigned long long)(clock64() - t0)); atomicAdd(&g_lb_stats[1], (unsigned long long)rounds);
        atomicAdd(&g_lb_stats[2], 1ull); atomicAdd(&g_lb_stats[3], (unsigned long long)stalled);
    }
#endif
    return acc;
}

// Exact u64 (< 2^53) -> double without the XU-pipe I2F: two magic-number halves and one fused multiply-add.
__device__ __forceinline__ double u64_to_double_exact(unsigned long long v) {
    double hi = __hiloint2double(0x43300000, (int)(v >> 32)) - TWO52;
    double lo = __hiloint2double(0x43300000, (int)(unsigned)v) - TWO52;
    return fma(hi, 4294967296.0, lo);
}

// #{ j in [0,n) : fl((double)j/(double)n) < v } -- the number of children the reference's two-pointer loop has
// handed out once the cumulative weight reaches W with v = fl(W - u) (sampler.h:814,836).
__device__ __forceinline__ long long children_below(double v, long long n, double nd, bool pow2) {
    if (!(v > 0.0)) return 0;
    double g = v * nd;                       // exact when n is a power of two
    if (g >= nd) {
        if (pow2 || g - nd >= 9.5367431640625e-07) return n;
    } else {
        // ceil(g) for 0 < g < 2^31 without F2I/FRND: round to nearest through the 2^52 magic constant
        double t = g + TWO52;
        double r = t - TWO52;
        long long c = (long long)(unsigned)__double2loint(t) + (r < g ? 1 : 0);
        if (pow2 || fabs(g - r) >= 9.5367431640625e-07) return c;  // safely away from a rounding tie
    }
    long long j = (long long)fmin(ceil(g), nd);
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
    const unsigned long long* seed_ptr;   // if non-null the Philox key is read from here (device scalar) instead of `seed`
    unsigned int* ticket;     // dynamic tile counter; the caller zeroes it (and the status words) before each pass
    ScanWs ws;
    AncestorMap am;
    int* count_out;           // optional: children counts per parent
    unsigned int* tail_info;  // optional [2]: {sum of counts, number of zero-count slots}
    int dbg;                  // experiment switches (0 in production): 1 skip look-back A, 2 skip look-back B, 4 skip exp, 8 skip surplus
    // island mode: n = local slots, n_glob = N of the whole population; the scan starts at *mass_offset_ptr
    long long n_glob;
    long long surplus_cap;    // entries the surplus array can hold (0 = n)
    const unsigned long long* mass_offset_ptr;
    long long* zoff_out;      // [world+1] written by the last tile after the all-gather of (zero slots, surplus entries)
    long long* eoff_out;
    unsigned int* done;       // island mode: blocks-finished counter (zero on entry, reset by the kernels)
    unsigned long long* mass_offset_out;   // island mode: written by the mass pass (mass of the lower ranks)
    uint32_t* cslot;          // three-pass path: children handed out up to each slot, resample_cslot_entries(n) entries
    IslandInfo isl;
};

enum { SRC_LOGW = 0, SRC_WEIGHTS = 1 };

// Children handed out once the cumulative weight reaches W (a multiple of 2^-52, exact in fp64).
// POW2: N is a power of two, so (W - u) N is exact and the count is a clamped ceiling -- straight-line fp64 code.
template <int MODE, bool POW2>
__device__ __forceinline__ int children_upto(double W, const ScanArgs& a, long long nthr, double nd, double inv_n,
                                             double dRand, long long step) {
    if (MODE == SYSTEMATIC) {
        if (POW2) {
            // ceil((W - u) N) clamped to [0, N].  The scaling is exact; adding 1.5 * 2^52 rounds to the nearest integer and leaves
            // it, in two's complement, in the low word (valid for |g| < 2^31: g > -1 and N <= 2^30 here), so the clamps are two
            // integer min/max instead of two fp64 ones.
            // N = 2^b: scaling commutes with rounding, so one fused multiply-add gives fl(W - u) N exactly; adding 1.5 * 2^52
            // ROUNDING UP leaves ceil(g) in the low word (two's complement; valid for -1 < g < 2^31), no compare needed
            const double g = fma(W, nd, -(dRand * nd));
            const int c = __double2loint(__dadd_ru(g, 6755399441055744.0));
            return min(max(c, 0), (int)nthr);
        }
        return (int)children_below(W - dRand, nthr, nd, false);
    } else {
        // STRATIFIED: child j compares against its own u_j = (1/N) U_j; only the stratum holding W is undecided.
        double g = W * nd;
        double fl = floor(g);
        long long js = (long long)fmin(fl, nd);
        double fr = g - fl;
        if (fr >= 9.5367431640625e-07 && fr <= 1.0 - 9.5367431640625e-07 && js >= 1 && js <= nthr - 2) {
            // the usual case, straight-line: the loop below would find child js-1 surely below W, child js+1 surely above, and
            // evaluate the reference's predicate for child js alone (same expression, same bits)
            const double U = a.ustrat ? a.ustrat[step * (nthr + 1) + js]
                                      : u64_to_unit(philox_draw(a.seed_ptr ? *a.seed_ptr : a.seed, STREAM_RESAMPLE, (uint32_t)step, (uint64_t)js, 0).a);
            const double uj = 0.0 + (inv_n - 0.0) * U;
            return (int)js + (((W - uj) > ((double)js / nd)) ? 1 : 0);
        }
        long long lo = js - 1 < 0 ? 0 : js - 1;
        long long hi = js + 1 > nthr - 1 ? nthr - 1 : js + 1;
        long long c = lo;
        for (long long j = lo; j <= hi; ++j) {
            bool sure_true = (j < js) && fr >= 9.5367431640625e-07;
            bool sure_false = (j > js) && fr <= 1.0 - 9.5367431640625e-07;
            bool p;
            if (sure_true) p = true;
            else if (sure_false) p = false;
            else {
                double U = a.ustrat ? a.ustrat[step * (nthr + 1) + j]
                                    : u64_to_unit(philox_draw(a.seed_ptr ? *a.seed_ptr : a.seed, STREAM_RESAMPLE, (uint32_t)step, (uint64_t)j, 0).a);
                double uj = 0.0 + (inv_n - 0.0) * U;
                p = (W - uj) > ((double)j / nd);
            }
            c += p ? 1 : 0;
        }
        return (int)c;
    }
}

// Surplus writers shared by the resampling kernels: parent `gk` owns surplus[E .. E+e).  Runs of up to SURPLUS_SMALL
// are written by the owning lane; longer ones are finished by the whole warp (write_surplus_heavy, cold path,
// 16-byte stores for the aligned middle of very long runs).
constexpr int SURPLUS_SMALL = 8;
__device__ __forceinline__ void write_surplus_small(uint32_t* __restrict__ surplus, int E, int e, uint32_t gk, int cap) {
    int m = e < SURPLUS_SMALL ? e : SURPLUS_SMALL;
    const int room = cap - E;              // list positions are < 2^31: 32-bit arithmetic, one address computation
    m = m < room ? m : room;
    uint32_t* p = surplus + E;
    for (int r = 0; r < m; ++r) p[r] = gk;
}
static __device__ __noinline__ void write_surplus_heavy(uint32_t* __restrict__ surplus, long long hE, long long hend, uint32_t hk, long long cap) {
    const int lane = threadIdx.x & 31;
    if (hend > cap) hend = cap;
    if (hend - hE >= 1024) {
        long long a0 = (hE + 3) & ~3ll, a1 = hend & ~3ll;
        for (long long r = hE + lane; r < a0; r += 32) surplus[r] = hk;
        uint4 v = make_uint4(hk, hk, hk, hk);
        for (long long r = a0 + 4 * lane; r < a1; r += 128) *reinterpret_cast<uint4*>(surplus + r) = v;
        for (long long r = a1 + lane; r < hend; r += 32) surplus[r] = hk;
    } else {
        for (long long r = hE + lane; r < hend; r += 32) surplus[r] = hk;
    }
}
// Warp-cooperative completion of the runs flagged in `hbits` (bit k <=> item k of this lane has e > SURPLUS_SMALL).
// E[k], e[k] are recomputed from the lane's running values in this rarely taken path.
template <int ITEMS>
__device__ __forceinline__ void finish_heavy_runs(uint32_t* __restrict__ surplus, unsigned hbits, const int (&cnt)[ITEMS],
                                                  long long cfirst, long long zexcl, unsigned zbits, long long g0, long long cap,
                                                  uint32_t id0 = 0u) {  // cold path: 64-bit arithmetic is fine here
    const int lane = threadIdx.x & 31;
    unsigned any;
    while ((any = __ballot_sync(SMCB_FULL, hbits != 0u)) != 0u) {
        const int src = __ffs(any) - 1;
        long long hE = 0, hend = 0;
        uint32_t hk = 0;
        if (lane == src) {
            const int kk = __ffs(hbits) - 1;
            hbits &= hbits - 1;
            long long ck1 = cfirst, zk = zexcl;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                if (k == kk) { hE = ck1 - (g0 + k) + zk + SURPLUS_SMALL; hend = hE + (cnt[k] - 1 - SURPLUS_SMALL); hk = id0 + (uint32_t)(g0 + k); }
                ck1 += cnt[k];
                zk += (zbits >> k) & 1u;
            }
        }
        hE = __shfl_sync(SMCB_FULL, hE, src);
        hend = __shfl_sync(SMCB_FULL, hend, src);
        hk = __shfl_sync(SMCB_FULL, hk, src);
        write_surplus_heavy(surplus, hE, hend, hk, cap);
    }
}

// Entries per look-back status array: sized for the smallest tile any resampling kernel uses (the warp tile, 256 slots) plus
// two words per group of 32 such tiles (second level of the warp kernel's look-back).
inline long long scan_tiles(long long n) { return (n + 255) / 256 + 2 * ((n + 8191) / 8192) + 2; }

// One pass: weights -> (zmask, zbase, surplus).  Persistent blocks pull 1024-slot tiles from a ticket counter and keep
// FOUR tiles in flight, one per pipeline stage, so that neither look-back chain is ever waited on right after its
// aggregate was published (measured: un-pipelined, each look-back stalled ~10k cycles on predecessors that had not
// reached their publish point yet) and the input never sits on the critical path:
//   S0(tile t+3)  cp.async of the tile's input into shared memory (each thread stages the slots it will read)
//   S1(tile t+2)  staged input -> registers, quantise to the 2^-52 grid, block scan, publish the weight-mass aggregate
//   S2(tile t+1)  look-back A (published a full iteration ago) -> children counts -> publish the zero-slot aggregate
//   S3(tile t)    look-back B (published a full iteration ago) -> zero-slot bitmap + surplus lists
// Stage state lives in shared memory (double-buffered), not registers.
constexpr int SCAN_PAD = SCAN_TILE + SCAN_TILE / SCAN_ITEMS;   // blocked layout with one pad word per thread
struct ScanSmem {
    double stage[SCAN_TILE];               // S0 -> S1: raw input of the next tile, landed by cp.async while the other stages run
    double cum[2][SCAN_PAD];               // S1 -> S2: tile-relative inclusive weight mass per slot (exact: multiples of 2^-52)
    int c[2][SCAN_PAD];                    // S2 -> S3: children handed out up to each slot; pad word = value before the thread's run
    int zrel[2][SCAN_THREADS];             // S2 -> S3: zero-count slots before the thread's run, tile-relative
    double warp_tot[SCAN_WARPS];
    int warp_z[SCAN_WARPS];
    unsigned long long bcast[2];           // S1 -> S2: tile aggregate as fixed point, then the tile's exclusive prefix
    long long zpre;
    int tile[3];
};
constexpr size_t SCAN_SMEM_BYTES = sizeof(ScanSmem);

// The whole scan is done in fp64 ON THE 2^-52 GRID: q = (w + 1) - 1 rounds w to a multiple of 2^-52, and every partial sum
// stays below 2, so all additions are exact and associative -- same bits as the u64 fixed-point scan, but on the (idle)
// fp64 pipe and without integer<->double conversions per slot.  Only the per-tile status words are fixed point.
template <int SRC, int MODE, bool ISLAND, bool POW2>
__device__ __forceinline__ void resample_scan_body(const ScanArgs& a, ScanSmem& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n = a.n;                              // slots scanned by this rank
    const long long nthr = ISLAND ? a.n_glob : n;          // N of the resampling thresholds j/N
    const double nd = (double)nthr;
    const double inv_n = 1.0 / nd;
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    const double lse = (SRC == SRC_LOGW) ? *a.lse_ptr : 0.0;
    const double dRand = (MODE == SYSTEMATIC) ? (0.0 + (inv_n - 0.0) * (*a.u_ptr)) : 0.0;
    const long long step = a.step_ptr ? *a.step_ptr : 0;
    const double mass0 = ISLAND ? u64_to_double_exact(*a.mass_offset_ptr) * INV_TWO52 : 0.0;   // weight mass of lower ranks
    const uint32_t gslot0 = 0u;   // parent ids are local to the island that owns the list (island mode: the reader adds the owner's offset)
    const long long scap = a.surplus_cap > 0 ? a.surplus_cap : n;
    const int n32 = (int)n;                                // n < 2^31
    // children handed out before this rank's first slot (0 on a single GPU)
    const int cstart = ISLAND ? children_upto<MODE, POW2>(mass0, a, nthr, nd, inv_n, dRand, step) : 0;

    // S0: asynchronous copy of a tile's input into shared memory; every thread stages exactly the slots it will read in S1,
    // so S1 needs no barrier, only its own cp.async group to have landed.
    const unsigned sa_stage = (unsigned)__cvta_generic_to_shared(&S.stage[tid]);
    auto prefetch = [&](int t) {
        if (t >= 0) {
            const long long base = (long long)t * SCAN_TILE;
            const bool full = base + SCAN_TILE <= n;
            const int left = full ? SCAN_TILE : (int)(n - base);   // valid slots of this tile
#pragma unroll
            for (int r = 0; r < SCAN_ITEMS; ++r) {
                const int i = r * SCAN_THREADS + tid;
                if (i < left) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa_stage + (unsigned)(r * SCAN_THREADS * 8)), "l"(a.in + base + i) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // tiles held by the stages: t0 is being staged, t1 enters S1 this iteration, t2 is in S2, t3 in S3 (-1 = empty)
    int t2 = -1, t3 = -1, tn;
    if (tid == 0) S.tile[0] = (int)atomicAdd(a.ticket, 1u);
    __syncthreads();
    tn = S.tile[0] < ntiles ? S.tile[0] : -1;
    prefetch(tn);
    __syncthreads();
    for (int it = 0;; ++it) {
        const int b1 = it & 1, b2 = b1 ^ 1;   // cum buffer written by S1 / read by S2; c buffer written by S2 = b2, read by S3 = b1
        if (tid == 0) S.tile[0] = (int)atomicAdd(a.ticket, 1u);
        __syncthreads();
        int t0 = S.tile[0];
        if (t0 >= ntiles) t0 = -1;
        const int t1 = tn;
        if (t0 < 0 && t1 < 0 && t2 < 0 && t3 < 0) break;

        // ================= S1: staged input -> registers, restage the next tile, quantise, block scan, publish =================
        if (t1 >= 0) {
            const long long base = (long long)t1 * SCAN_TILE;
            const bool full = base + SCAN_TILE <= n;
            const int left = full ? SCAN_TILE : (int)(n - base);   // valid slots of this tile
            double* sq = S.cum[b1];
            double raw[SCAN_ITEMS];
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int r = 0; r < SCAN_ITEMS; ++r) raw[r] = S.stage[r * SCAN_THREADS + tid];
            prefetch(t0);
#pragma unroll
            for (int r = 0; r < SCAN_ITEMS; ++r) {
                const int i = r * SCAN_THREADS + tid;
                double q = 0.0;
                if (i < left) {
                    const double w = (SRC == SRC_LOGW) ? fm_exp(raw[r] - lse) : raw[r];
                    q = (w + 1.0) - 1.0;                  // round to the 2^-52 grid (ulp of [1,2))
                }
                sq[i + (i >> 3)] = q;
            }
            __syncthreads();
            double cum[SCAN_ITEMS];
            {
                double run = 0.0;
#pragma unroll
                for (int k = 0; k < SCAN_ITEMS; ++k) { run += sq[tid * (SCAN_ITEMS + 1) + k]; cum[k] = run; }
            }
            const double tot = cum[SCAN_ITEMS - 1];
            double inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double t = __shfl_up_sync(SMCB_FULL, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) S.warp_tot[warp] = inc;
            __syncthreads();
            double off = inc - tot, agg = 0.0;
#pragma unroll
            for (int w = 0; w < SCAN_WARPS; ++w) { double v = S.warp_tot[w]; if (w < warp) off += v; agg += v; }
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) sq[tid * (SCAN_ITEMS + 1) + k] = off + cum[k];
            if (tid == 0) {  // publish: tile 0 knows its prefix, everyone else an aggregate (fixed point, exact)
                const unsigned long long aggq = (unsigned long long)__double2ll_rn(agg * TWO52);
                volatile unsigned long long* st = a.ws.statusA;
                st[t1] = (t1 == 0 ? ST_PRE : ST_AGG) | aggq;
                S.bcast[b1] = aggq;
            }
        }
        __syncthreads();

        // ================= S2: look-back A, children counts, publish zero-slot aggregate =================
        if (t2 >= 0) {
            if (warp == 0) {
                unsigned long long ex = lookback_resolve(a.ws.statusA, t2, S.bcast[b2]);
                if (lane == 0) S.bcast[b2] = ex;
            }
            __syncthreads();
            const double tile_excl = u64_to_double_exact(S.bcast[b2]) * INV_TWO52 + mass0;   // exact
            const int l0 = tid * SCAN_ITEMS;                                                   // slot within the tile
            const bool full = (long long)(t2 + 1) * SCAN_TILE <= n;
            const int nvalid = full ? SCAN_TILE : (int)(n - (long long)t2 * SCAN_TILE);
            const double* sq = S.cum[b2];
            int* sc = S.c[b2];
            const double thread_excl = tile_excl + (tid == 0 ? 0.0 : sq[(tid - 1) * (SCAN_ITEMS + 1) + SCAN_ITEMS - 1]);
            int cprev = children_upto<MODE, POW2>(thread_excl, a, nthr, nd, inv_n, dRand, step) - cstart;
            sc[tid * (SCAN_ITEMS + 1) + SCAN_ITEMS] = cprev;
            unsigned zbits = 0u;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                int ck = cprev;
                if (full || l0 + k < nvalid) {
                    ck = children_upto<MODE, POW2>(tile_excl + sq[tid * (SCAN_ITEMS + 1) + k], a, nthr, nd, inv_n, dRand, step) - cstart;
                    if (ck == cprev) zbits |= 1u << k;
                }
                sc[tid * (SCAN_ITEMS + 1) + k] = ck;
                cprev = ck;
            }
            const int ztot = __popc(zbits);
            int zinc = ztot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(SMCB_FULL, zinc, o);
                if (lane >= o) zinc += t;
            }
            if (lane == 31) S.warp_z[warp] = zinc;
            __syncthreads();
            int zoff = zinc - ztot, zagg = 0;
#pragma unroll
            for (int w = 0; w < SCAN_WARPS; ++w) { int v = S.warp_z[w]; if (w < warp) zoff += v; zagg += v; }
            S.zrel[b2][tid] = zoff;
            if (tid == 0) {
                volatile unsigned long long* st = a.ws.statusB;
                st[t2] = (t2 == 0 ? ST_PRE : ST_AGG) | (unsigned long long)zagg;
                S.tile[1] = zagg;
            }
        }
        __syncthreads();

        // ================= S3: look-back B, bitmap + surplus lists =================
        if (t3 >= 0) {
            if (warp == 0) {
                unsigned long long ex = lookback_resolve(a.ws.statusB, t3, (unsigned long long)S.tile[2]);
                if (lane == 0) S.zpre = (long long)ex;
            }
            __syncthreads();
            const int zprefix = (int)S.zpre;
            const int g0 = t3 * SCAN_TILE + tid * SCAN_ITEMS;          // local slot id (n < 2^31)
            const bool full = (long long)(t3 + 1) * SCAN_TILE <= n;
            const int* sc = S.c[b1];
            const int cfirst = sc[tid * (SCAN_ITEMS + 1) + SCAN_ITEMS];
            int cnt[SCAN_ITEMS];
            unsigned zbits = 0u;
            {
                int cp = cfirst;
#pragma unroll
                for (int k = 0; k < SCAN_ITEMS; ++k) {
                    int ck = sc[tid * (SCAN_ITEMS + 1) + k];
                    cnt[k] = ck - cp;
                    if (cnt[k] == 0 && (full || g0 + k < n32)) zbits |= 1u << k;
                    cp = ck;
                }
            }
            const int zexcl = zprefix + S.zrel[b1][tid];
            {   // zero-slot bitmap: 4 threads x 8 slots = one 32-bit word
                unsigned m = zbits << (8 * (lane & 3));
                m |= __shfl_xor_sync(SMCB_FULL, m, 1);
                m |= __shfl_xor_sync(SMCB_FULL, m, 2);
                if ((lane & 3) == 0 && (full || g0 < n32)) {
                    a.am.zmask[g0 >> 5] = m;
                    a.am.zbase[g0 >> 5] = (uint32_t)zexcl;
                }
            }
            {   // surplus list: parent k with count >= 2 owns surplus[E .. E + count - 2],  E = c(k-1) - k + Z[k]
                int E = cfirst - g0 + zexcl;
                unsigned hbits = 0u;
#pragma unroll
                for (int k = 0; k < SCAN_ITEMS; ++k) {
                    const int e = cnt[k] - 1;
                    write_surplus_small(a.am.surplus, E, e, gslot0 + (uint32_t)(g0 + k), (int)scap);
                    if (e > SURPLUS_SMALL) hbits |= 1u << k;
                    E += cnt[k] - 1 + (int)((zbits >> k) & 1u);      // E[k+1] = E[k] + count[k] - 1 + z[k]
                }
                finish_heavy_runs<SCAN_ITEMS>(a.am.surplus, hbits, cnt, cfirst, zexcl, zbits, g0, scap, gslot0);
            }
            if (a.count_out) {
#pragma unroll
                for (int k = 0; k < SCAN_ITEMS; ++k) if ((long long)g0 + k < n) a.count_out[g0 + k] = cnt[k];
            }
            if (t3 == ntiles - 1) {
                __shared__ long long s_tail[2];
                if (tid == SCAN_THREADS - 1) { s_tail[0] = sc[tid * (SCAN_ITEMS + 1) + SCAN_ITEMS - 1]; s_tail[1] = zexcl + __popc(zbits); }
                __syncthreads();
                const long long ctot = s_tail[0], ztotal = s_tail[1];   // children handed out / zero-count slots on this rank
                const long long etot = ctot - n + ztotal;               // surplus-list entries written by this rank
                if (!ISLAND) {
                    // unfilled zero-count slots copy particle 0 (reference: uRSIndices zero-initialised, sampler.h:845)
                    if (!a.dbg) for (long long m = etot + tid; m < ztotal; m += SCAN_THREADS) a.am.surplus[m] = 0u;
                    if (tid == 0 && a.tail_info) { a.tail_info[0] = (unsigned)ctot; a.tail_info[1] = (unsigned)ztotal; }
                } else if (warp == 0) {
                    // all-gather (zero slots, surplus entries) of every island; exclusive prefixes tell k_step's decode where
                    // the m-th zero-count slot of the GLOBAL order finds its parent
                    __shared__ double s_all[ISL_MAX][ISL_PAYLOAD];
                    double mine[2] = {(double)ztotal, (double)etot};
                    island_allgather(a.isl, ISL_KIND_ZE, (a.isl.epoch << 24) + (unsigned long long)step + 1ull, mine, 2, s_all);
                    if (lane == 0) {
                        long long z = 0, e = 0;
                        for (int h = 0; h < a.isl.world; ++h) {
                            a.zoff_out[h] = z; a.eoff_out[h] = e;
                            z += (long long)s_all[h][0]; e += (long long)s_all[h][1];
                        }
                        a.zoff_out[a.isl.world] = z; a.eoff_out[a.isl.world] = e;
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) S.tile[2] = S.tile[1];   // zero-slot aggregate of the tile moving from S2 to S3
        t3 = t2;
        t2 = t1;
        tn = t0;
    }
}

template <int SRC, int MODE, bool ISLAND = false>
__global__ void __launch_bounds__(SCAN_THREADS, SMCB_SCAN_MINB) k_resample_scan(ScanArgs a) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    if (a.dbg & 16) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanSmem& S = *reinterpret_cast<ScanSmem*>(smem_raw);
    const long long nthr = ISLAND ? a.n_glob : a.n;
    if ((nthr & (nthr - 1)) == 0) resample_scan_body<SRC, MODE, ISLAND, true>(a, S);
    else resample_scan_body<SRC, MODE, ISLAND, false>(a, S);
}

// Host-side launch helpers: the pipelined kernel needs > 48 KB of dynamic shared memory (opt-in attribute).
template <int SRC, int MODE, bool ISLAND = false>
inline int resample_scan_grid(long long n) {
    static bool configured = false;
    if (!configured) {
        SMCB_CUDA(cudaFuncSetAttribute(k_resample_scan<SRC, MODE, ISLAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_SMEM_BYTES));
        configured = true;
    }
    int dev = 0, sms = 0, per_sm = 0;
    SMCB_CUDA(cudaGetDevice(&dev));
    SMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SMCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_resample_scan<SRC, MODE, ISLAND>, SCAN_THREADS, SCAN_SMEM_BYTES));
    if (per_sm < 1) per_sm = 1;
    long long need = (n + SCAN_TILE - 1) / SCAN_TILE, g = (long long)sms * per_sm;
    if (g > need) g = need;
    return (int)(g < 1 ? 1 : g);
}
template <int SRC, int MODE, bool ISLAND = false>
inline void launch_resample_scan(const ScanArgs& a, cudaStream_t stream) {
    int grid = resample_scan_grid<SRC, MODE, ISLAND>(a.n);
    k_resample_scan<SRC, MODE, ISLAND><<<grid, SCAN_THREADS, SCAN_SMEM_BYTES, stream>>>(a);
    SMCB_CUDA(cudaGetLastError());
}

// =====================================================================================================================
// Three-pass resampling without any waiting between thread blocks (the production path for SYSTEMATIC / STRATIFIED).
//
// Same mathematics and the same outputs as the single-pass kernel above; the two prefix dependencies (weight mass, zero-count
// slots) are resolved at kernel boundaries instead of by look-back chains:
//   k_rs_mass   tile masses   (256-slot warp tiles; every tile's exact mass + atomically accumulated masses of 128-tile "supers")
//   k_rs_count  mass prefix of the tile from (supers, tiles) -> exact cumulative weights -> children handed out up to each slot
//               (stored, 4 B per slot) -> zero-slot bitmap, zero-slot totals per tile / super
//   k_rs_map    zero-slot prefix from (supers, tiles) -> rank bases + surplus lists
// A warp owns whole tiles (blocked: lane l holds 8 consecutive slots in registers), all scans are register + shuffle scans,
// there is no shared memory and no __syncthreads.  Every pass streams at HBM speed; nothing ever spins on another block, so
// the passes are indifferent to what else runs on the GPU.  Measured alternatives (profiles/r2_scan_experiments.txt): the
// block-tiled single pass above (177 us per 2^24 slots: barrier + look-back latency bound) and a warp-granular single pass
// with one- and two-level look-back chains (340 - 960 us: every tile couples to the slowest of its ~64 predecessors).
constexpr int RP_ITEMS = 8;
constexpr int RP_TILE = 32 * RP_ITEMS;          // slots per warp tile
constexpr int RP_SUPER = 128;                   // tiles per super
constexpr int RP_THREADS = 256;
static_assert(RP_TILE == 256, "scan_tiles() sizes the per-tile arrays for 256-slot tiles");

// 256-bit global accesses (sm_100): every lane moves one whole 32-byte sector
__device__ __forceinline__ void ld_global_v4(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void ld_global_v8(const uint32_t* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_global_v8(uint32_t* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

enum { SRC_WTILDE = 2 };   // weights exp(lw - c) against a step-wide shift c, normalised by one multiplication (k_step writes them)

struct RsView {            // how the three passes see the workspace of ScanArgs
    double* tile_mass;                 // [ntiles]  exact mass of each tile (multiple of 2^-52)
    unsigned long long* tile_z;        // [ntiles]  zero-count slots of each tile
    unsigned long long* super_mass;    // [nsuper]  fixed-point (2^-52) mass of each super, accumulated atomically (zero on entry)
    unsigned long long* super_z;       // [nsuper]  zero-count slots of each super (zero on entry)
};
__device__ __forceinline__ RsView rs_view(const ScanArgs& a, int ntiles) {
    RsView v;
    v.tile_mass = reinterpret_cast<double*>(a.ws.statusA);
    v.tile_z = a.ws.statusB;
    v.super_mass = a.ws.statusC;
    v.super_z = a.ws.statusC + (ntiles + RP_SUPER - 1) / RP_SUPER;
    return v;
}
template <int SRC>
__device__ __forceinline__ double rs_quantise(double raw, double lse, double inv_s) {
    const double w = (SRC == SRC_LOGW) ? fm_exp(raw - lse) : (SRC == SRC_WTILDE ? raw * inv_s : raw);
    return (w + 1.0) - 1.0;                     // round to the 2^-52 grid (ulp of [1,2)); all sums of such values below 2 are exact
}
// contiguous chunk of tiles owned by a warp: [t0, t1)
__device__ __forceinline__ void rs_chunk(int ntiles, int& t0, int& t1) {
    const int nw = (int)(gridDim.x * (RP_THREADS / 32));
    const int w = (int)(blockIdx.x * (RP_THREADS / 32) + (threadIdx.x >> 5));
    const int per = (ntiles + nw - 1) / nw;
    t0 = min(w * per, ntiles);
    t1 = min(t0 + per, ntiles);
}

#ifndef SMCB_RS_MASS_BPS
#define SMCB_RS_MASS_BPS 12
#endif
template <int SRC, bool ISLAND>
__global__ void __launch_bounds__(RP_THREADS) k_rs_mass(ScanArgs a) {
    pdl_wait_for_predecessor();
    pdl_release_successor();
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    const int lane = threadIdx.x & 31;
    const int n32 = (int)a.n;
    const int ntiles = (n32 + RP_TILE - 1) / RP_TILE;
    const RsView V = rs_view(a, ntiles);
    const double lse = (SRC == SRC_LOGW) ? *a.lse_ptr : 0.0;
    const double inv_s = (SRC == SRC_WTILDE) ? *a.lse_ptr : 1.0;   // SRC_WTILDE: lse_ptr points at 1 / sum of the shifted weights
    const int nw = (int)(gridDim.x * (RP_THREADS / 32));
#ifndef SMCB_RS_MASS_REV
#define SMCB_RS_MASS_REV 1
#endif
    // the next tile's loads are issued before the current tile is reduced (two tiles of a warp in flight)
    auto tile_of = [&](int it) { return SMCB_RS_MASS_REV ? ntiles - 1 - it : it; };
    auto load_full = [&](int tile, double2 (&v)[RP_ITEMS / 2]) {
        const double2* p = reinterpret_cast<const double2*>(a.in + tile * RP_TILE) + lane;
#pragma unroll
        for (int r = 0; r < RP_ITEMS / 2; ++r) v[r] = p[32 * r];
    };
    const int it0 = (int)(blockIdx.x * (RP_THREADS / 32) + (threadIdx.x >> 5));
    double2 vn[RP_ITEMS / 2];
#pragma unroll
    for (int r = 0; r < RP_ITEMS / 2; ++r) vn[r] = make_double2(0.0, 0.0);
    if (it0 < ntiles && tile_of(it0) * RP_TILE + RP_TILE <= n32) load_full(tile_of(it0), vn);
    for (int it = it0; it < ntiles; it += nw) {
        const int tile = tile_of(it);
        const int base = tile * RP_TILE;
        double s = 0.0;
        if (base + RP_TILE <= n32) {   // coalesced 16-byte loads; the order inside a tile does not matter for an exact sum
            double2 v[RP_ITEMS / 2];
#pragma unroll
            for (int r = 0; r < RP_ITEMS / 2; ++r) v[r] = vn[r];
            if (it + nw < ntiles && tile_of(it + nw) * RP_TILE + RP_TILE <= n32) load_full(tile_of(it + nw), vn);
#pragma unroll
            for (int r = 0; r < RP_ITEMS / 2; ++r) s += rs_quantise<SRC>(v[r].x, lse, inv_s) + rs_quantise<SRC>(v[r].y, lse, inv_s);
        } else {
            if (it + nw < ntiles && tile_of(it + nw) * RP_TILE + RP_TILE <= n32) load_full(tile_of(it + nw), vn);
#pragma unroll
            for (int r = 0; r < RP_ITEMS; ++r) {
                const int i = base + r * 32 + lane;
                if (i < n32) s += rs_quantise<SRC>(a.in[i], lse, inv_s);
            }
        }
        s = warp_sum(s);
        if (lane == 0) {
            V.tile_mass[tile] = s;
            atomicAdd(V.super_mass + tile / RP_SUPER, (unsigned long long)__double2ll_rn(s * TWO52));
        }
    }
    // (island mode: the rank's total is posted to the peers by the first block of k_rs_count, from the super masses above)
}

template <int SRC, int MODE, bool ISLAND, bool POW2>
__device__ __forceinline__ void rs_count_body(const ScanArgs& a) {
    const int lane = threadIdx.x & 31;
    const long long n = a.n;
    const int n32 = (int)n;
    const long long nthr = ISLAND ? a.n_glob : n;          // N of the resampling thresholds j/N
    const double nd = (double)nthr;
    const double inv_n = 1.0 / nd;
    const int ntiles = (n32 + RP_TILE - 1) / RP_TILE;
    const RsView V = rs_view(a, ntiles);
    const double lse = (SRC == SRC_LOGW) ? *a.lse_ptr : 0.0;
    const double inv_s = (SRC == SRC_WTILDE) ? *a.lse_ptr : 1.0;
    const double dRand = (MODE == SYSTEMATIC) ? (0.0 + (inv_n - 0.0) * (*a.u_ptr)) : 0.0;
    const long long step = a.step_ptr ? *a.step_ptr : 0;
    // island mode: weight mass of the lower ranks = where this rank's cumulative weight starts.  Every warp reads the masses the
    // other ranks' mass passes posted into this rank's mailbox (local memory; they have usually arrived by now)
    if (ISLAND && blockIdx.x == 0 && threadIdx.x < 32) {
        // this rank's exact weight mass (sum of the super masses the mass pass accumulated), posted to every rank's mailbox
        const int nsuper = (ntiles + RP_SUPER - 1) / RP_SUPER;
        unsigned long long tot = 0ull;
        for (int k = lane; k < nsuper; k += 32) tot += V.super_mass[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(SMCB_FULL, tot, o);
        double mine[1] = {(double)tot};   // < 2^53: exact
        island_send(a.isl, ISL_KIND_MASS, (a.isl.epoch << 24) + (unsigned long long)step + 1ull, mine, 1);
    }
    const double mass0 = ISLAND ? island_recv_lower_sum(a.isl, ISL_KIND_MASS, (a.isl.epoch << 24) + (unsigned long long)step + 1ull) * INV_TWO52 : 0.0;
    const int cstart = ISLAND ? children_upto<MODE, POW2>(mass0, a, nthr, nd, inv_n, dRand, step) : 0;
    int t0, t1;
    rs_chunk(ntiles, t0, t1);
    if (t0 >= t1) return;
    // exact mass in front of the chunk: whole supers, then the tiles of the super the chunk starts in
    double P;
    {
        const int s0 = t0 / RP_SUPER;
        double acc = 0.0;
        for (int k = lane; k < s0; k += 32) acc += u64_to_double_exact(V.super_mass[k]) * INV_TWO52;
        for (int k = s0 * RP_SUPER + lane; k < t0; k += 32) acc += V.tile_mass[k];
        P = warp_sum(acc) + mass0;
    }
    // blocked load of a tile: lane l reads 8 consecutive doubles (16-byte vectors; slabs and tiles are 256-byte aligned)
    auto load_tile = [&](int t, double (&r)[RP_ITEMS]) {
        const int g = t * RP_TILE + lane * RP_ITEMS;
        if (g + RP_ITEMS <= n32) {
            ld_global_v4(a.in + g, r[0], r[1], r[2], r[3]);
            ld_global_v4(a.in + g + 4, r[4], r[5], r[6], r[7]);
        } else {
#pragma unroll
            for (int k = 0; k < RP_ITEMS; ++k) r[k] = g + k < n32 ? a.in[g + k] : 0.0;
        }
    };
    double raw[RP_ITEMS];
    load_tile(t0, raw);
    for (int tile = t0; tile < t1; ++tile) {
        const int g0 = tile * RP_TILE + lane * RP_ITEMS;       // local slot id of the lane's first slot (n < 2^31)
        const int nv = min(max(n32 - g0, 0), RP_ITEMS);         // valid slots of this lane
        double cum[RP_ITEMS];
        {
            double run = 0.0;
#pragma unroll
            for (int k = 0; k < RP_ITEMS; ++k) {
                double q = rs_quantise<SRC>(raw[k], lse, inv_s);
                if (k >= nv) q = 0.0;
                run += q;
                cum[k] = run;
            }
        }
        if (tile + 1 < t1) load_tile(tile + 1, raw);            // the next tile's input is in flight during the arithmetic below
        const double tot = cum[RP_ITEMS - 1];
        double inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(SMCB_FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const double lane_excl = P + (inc - tot);               // exact
        P += __shfl_sync(SMCB_FULL, inc, 31);
        // children handed out up to each slot (closed form with the reference's predicate), relative to this rank's start
        int cprev = children_upto<MODE, POW2>(lane_excl, a, nthr, nd, inv_n, dRand, step) - cstart;
        uint32_t c[RP_ITEMS];
        unsigned zbits = 0u;
#pragma unroll
        for (int k = 0; k < RP_ITEMS; ++k) {
            const int ck = children_upto<MODE, POW2>(lane_excl + cum[k], a, nthr, nd, inv_n, dRand, step) - cstart;
            if (ck == cprev) zbits |= 1u << k;
            c[k] = (uint32_t)ck;
            cprev = ck;
        }
        zbits &= (1u << nv) - 1u;
        // c is padded to whole tiles, so the stores need no guard
        st_global_v8(a.cslot + g0, c);
        {   // zero-slot bitmap: 4 lanes x 8 slots = one 32-bit word
            unsigned m = zbits << (8 * (lane & 3));
            m |= __shfl_xor_sync(SMCB_FULL, m, 1);
            m |= __shfl_xor_sync(SMCB_FULL, m, 2);
            if ((lane & 3) == 0 && g0 < n32) a.am.zmask[g0 >> 5] = m;
        }
        int z = __popc(zbits);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(SMCB_FULL, z, o);
        if (lane == 0) {
            V.tile_z[tile] = (unsigned long long)z;
            atomicAdd(V.super_z + tile / RP_SUPER, (unsigned long long)z);
        }
    }
}

#ifndef SMCB_RS_COUNT_MINB
#define SMCB_RS_COUNT_MINB 3
#endif
template <int SRC, int MODE, bool ISLAND>
__global__ void __launch_bounds__(RP_THREADS, SMCB_RS_COUNT_MINB) k_rs_count(ScanArgs a) {
    pdl_wait_for_predecessor();
    pdl_release_successor();
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    const long long nthr = ISLAND ? a.n_glob : a.n;
    if ((nthr & (nthr - 1)) == 0) rs_count_body<SRC, MODE, ISLAND, true>(a);
    else rs_count_body<SRC, MODE, ISLAND, false>(a);
}

constexpr int RP_WINDOW = 512;                  // surplus entries expanded per shared-memory window (per warp)
#ifndef SMCB_RS_MAP_MINB
#define SMCB_RS_MAP_MINB 4
#endif
template <bool ISLAND>
__global__ void __launch_bounds__(RP_THREADS, SMCB_RS_MAP_MINB) k_rs_map(ScanArgs a) {
    pdl_wait_for_predecessor();
    pdl_release_successor();
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    __shared__ uint32_t s_seg[RP_THREADS / 32][RP_WINDOW];
    const int lane = threadIdx.x & 31;
    const long long n = a.n;
    const int n32 = (int)n;
    const int ntiles = (n32 + RP_TILE - 1) / RP_TILE;
    const RsView V = rs_view(a, ntiles);
    const uint32_t gslot0 = 0u;   // parent ids are local to the island that owns the list (island mode: the reader adds the owner's offset)
    const long long step = a.step_ptr ? *a.step_ptr : 0;
    int t0, t1;
    rs_chunk(ntiles, t0, t1);
    if (t0 < t1) {
        // zero-count slots in front of the chunk, and the children handed out before its first slot
        int Z;
        {
            const int s0 = t0 / RP_SUPER;
            unsigned long long acc = 0ull;
            for (int k = lane; k < s0; k += 32) acc += V.super_z[k];
            for (int k = s0 * RP_SUPER + lane; k < t0; k += 32) acc += V.tile_z[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SMCB_FULL, acc, o);
            Z = (int)acc;
        }
        int clast = t0 > 0 ? (int)a.cslot[t0 * RP_TILE - 1] : 0;    // c of the slot in front of the tile
        uint32_t unext[8];
        ld_global_v8(a.cslot + t0 * RP_TILE + lane * RP_ITEMS, unext);
        for (int tile = t0; tile < t1; ++tile) {
            const int g0 = tile * RP_TILE + lane * RP_ITEMS;
            const int nv = min(max(n32 - g0, 0), RP_ITEMS);
            int c[RP_ITEMS];
#pragma unroll
            for (int k = 0; k < RP_ITEMS; ++k) c[k] = (int)unext[k];
            if (tile + 1 < t1) ld_global_v8(a.cslot + g0 + RP_TILE, unext);   // the next tile's counts are in flight during the work below
            int cfirst = __shfl_up_sync(SMCB_FULL, c[RP_ITEMS - 1], 1);
            if (lane == 0) cfirst = clast;
            clast = __shfl_sync(SMCB_FULL, c[RP_ITEMS - 1], 31);
            int cnt[RP_ITEMS];
            unsigned zbits = 0u;
            {
                int cp = cfirst;
#pragma unroll
                for (int k = 0; k < RP_ITEMS; ++k) {
                    cnt[k] = c[k] - cp;
                    if (cnt[k] == 0) zbits |= 1u << k;
                    cp = c[k];
                }
            }
            zbits &= (1u << nv) - 1u;
            const int ztot = __popc(zbits);
            int zinc = ztot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(SMCB_FULL, zinc, o);
                if (lane >= o) zinc += t;
            }
            const int zexcl = Z + zinc - ztot;
            const int zagg = __shfl_sync(SMCB_FULL, zinc, 31);
            if ((lane & 3) == 0 && g0 < n32) a.am.zbase[g0 >> 5] = (uint32_t)zexcl;
            {   // Surplus lists.  Parent k with count >= 2 owns surplus[E[k] .. E[k] + count - 2], E[k] = c(k-1) - k + Z[k]: the
                // tile's entries are ONE contiguous segment in which every parent's id is repeated count-1 times, in slot
                // order.  The warp expands it cooperatively: each parent drops a marker at the start of its run in a
                // shared-memory window, an inclusive max-scan over the window fills the runs (ids increase along the
                // segment), and the window is written out with coalesced stores.  Cost per entry is the same for a run of
                // 1 and a run of 10^6: skewed weights (the reason a filter resamples) do not slow the pass down.
                int o[RP_ITEMS];                     // start of each slot's run, relative to the tile's segment
                int run = 0;
#pragma unroll
                for (int k = 0; k < RP_ITEMS; ++k) { o[k] = run; run += max(cnt[k] - 1, 0); }
                int einc = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(SMCB_FULL, einc, d);
                    if (lane >= d) einc += t;
                }
                const int lane_off = einc - run;
                const int etile = __shfl_sync(SMCB_FULL, einc, 31);          // entries of this tile
                const long long Etile = __shfl_sync(SMCB_FULL, cfirst - g0 + zexcl, 0);   // E of the tile's first slot (lanes past the end of the array have no valid E of their own)
                uint32_t* seg = s_seg[threadIdx.x >> 5];
                const uint32_t id0 = gslot0 + (uint32_t)(tile * RP_TILE);     // global id of the tile's first slot
                uint32_t carry = 0u;                                         // marker (local slot + 1) in force at the window start
                // A parent with thousands of children (one resampling after a long stretch without: ESS-triggered schemes)
                // owns a run that spans many windows; the windows that lie entirely inside such a run hold ONE value and are
                // written with plain 16-byte stores instead of going through the marker scan (measured on pfLineartBS,
                // 2^20 particles: 1.7 ms -> per resampling pass without this).
                unsigned hbits = 0u;
#pragma unroll
                for (int k = 0; k < RP_ITEMS; ++k) if (cnt[k] - 1 >= 2 * RP_WINDOW) hbits |= 1u << k;
                const bool tile_heavy = __any_sync(SMCB_FULL, hbits != 0u);
                for (int w0 = 0; w0 < etile; w0 += RP_WINDOW) {
                    if (tile_heavy) {
                        int hend = 0; uint32_t hid = 0u;
#pragma unroll
                        for (int k = 0; k < RP_ITEMS; ++k) {
                            const int rs = lane_off + o[k], re = rs + cnt[k] - 1;
                            if (((hbits >> k) & 1u) && rs <= w0 && w0 + RP_WINDOW <= re) { hend = re; hid = (uint32_t)(lane * RP_ITEMS + k + 1); }
                        }
                        const unsigned hb = __ballot_sync(SMCB_FULL, hend != 0);
                        if (hb) {
                            const int src = __ffs(hb) - 1;
                            hend = __shfl_sync(SMCB_FULL, hend, src);
                            hid = __shfl_sync(SMCB_FULL, hid, src);
                            const int wend = w0 + ((hend - w0) / RP_WINDOW) * RP_WINDOW;     // whole windows inside the run
                            write_surplus_heavy(a.am.surplus, Etile + w0, Etile + wend, id0 + hid - 1u, 0x7fffffffffffffffLL);
                            carry = hid;
                            w0 = wend - RP_WINDOW;                                           // the loop adds one window
                            continue;
                        }
                    }
                    const int wlen = min(etile - w0, RP_WINDOW);
                    for (int j = lane; j < wlen; j += 32) seg[j] = 0u;
                    __syncwarp();
                    if (etile <= RP_WINDOW) {   // the usual case, one window: every run starts inside it
#pragma unroll
                        for (int k = 0; k < RP_ITEMS; ++k) if (cnt[k] > 1) seg[lane_off + o[k]] = (uint32_t)(lane * RP_ITEMS + k + 1);
                    } else {
#pragma unroll
                        for (int k = 0; k < RP_ITEMS; ++k) {
                            const int pos = lane_off + o[k] - w0;
                            if (cnt[k] > 1 && pos >= 0 && pos < wlen) seg[pos] = (uint32_t)(lane * RP_ITEMS + k + 1);
                        }
                    }
                    __syncwarp();
                    uint32_t* out = a.am.surplus + (Etile + w0);   // the list cannot overflow: a rank's entries <= children handed out <= N
                    for (int j0 = 0; j0 < wlen; j0 += 32) {
                        const int j = j0 + lane;
                        uint32_t v = j < wlen ? seg[j] : 0u;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) v = max(v, __shfl_up_sync(SMCB_FULL, v, d));   // lanes below d get their own value back
                        v = max(v, carry);
                        carry = __shfl_sync(SMCB_FULL, v, 31);
                        if (j < wlen) out[j] = id0 + v - 1u;
                    }
                    __syncwarp();
                }
            }
            if (a.count_out) {
#pragma unroll
                for (int k = 0; k < RP_ITEMS; ++k) if (k < nv) a.count_out[g0 + k] = cnt[k];
            }
            if (tile == ntiles - 1) {
                // totals of this rank: children handed out, zero-count slots, surplus-list entries
                const long long ctot = clast;
                const long long ztotal = (long long)Z + zagg;
                const long long etot = ctot - n + ztotal;
                if (!ISLAND) {
                    // unfilled zero-count slots copy particle 0 (reference: uRSIndices zero-initialised, sampler.h:845)
                    for (long long m = etot + lane; m < ztotal; m += 32) a.am.surplus[m] = 0u;
                    if (lane == 0 && a.tail_info) { a.tail_info[0] = (unsigned)ctot; a.tail_info[1] = (unsigned)ztotal; }
                } else if (lane == 0) {
                    a.zoff_out[ISL_MAX + 1] = ztotal;           // stash for the block that finishes last (below)
                    a.eoff_out[ISL_MAX + 1] = etot;
                }
            }
            Z += zagg;
        }
    }
    // (island mode: the totals stashed above are posted to the peers by the first block of the next k_step -- by then this
    // whole pass has finished, so their arrival also tells the peers that this rank's surplus list may be read over NVLink)
}

// ---- children counts -> (cslot, zero-slot bitmap, tile / super totals): the front half of the ancestor map for the schemes that
// produce COUNTS (MULTINOMIAL / RESIDUAL draws), in the same streaming structure as the weight passes above; k_rs_map finishes.
struct CountPassArgs {
    const int* count;            // [n] children per parent
    long long n;
    const int* enable_ptr;
    ScanWs ws;                   // statusA: tile totals, statusB: tile zero counts, statusC: super totals | super zero counts (zero on entry)
    AncestorMap am;
    uint32_t* cslot;
};
static __global__ void __launch_bounds__(RP_THREADS) k_cnt_tiles(CountPassArgs a) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    const int lane = threadIdx.x & 31;
    const int n32 = (int)a.n;
    const int ntiles = (n32 + RP_TILE - 1) / RP_TILE;
    unsigned long long* tile_c = a.ws.statusA;
    unsigned long long* super_c = a.ws.statusC;
    const int nw = (int)(gridDim.x * (RP_THREADS / 32));
    for (int tile = (int)(blockIdx.x * (RP_THREADS / 32) + (threadIdx.x >> 5)); tile < ntiles; tile += nw) {
        const int base = tile * RP_TILE;
        int s = 0;
#pragma unroll
        for (int r = 0; r < RP_ITEMS; ++r) {
            const int i = base + r * 32 + lane;
            if (i < n32) s += a.count[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(SMCB_FULL, s, o);
        if (lane == 0) {
            tile_c[tile] = (unsigned long long)s;
            atomicAdd(super_c + tile / RP_SUPER, (unsigned long long)s);
        }
    }
}
static __global__ void __launch_bounds__(RP_THREADS) k_cnt_scan(CountPassArgs a) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    const int lane = threadIdx.x & 31;
    const int n32 = (int)a.n;
    const int ntiles = (n32 + RP_TILE - 1) / RP_TILE;
    const unsigned long long* tile_c = a.ws.statusA;
    const unsigned long long* super_c = a.ws.statusC;
    unsigned long long* tile_z = a.ws.statusB;
    unsigned long long* super_z = a.ws.statusC + (ntiles + RP_SUPER - 1) / RP_SUPER;
    int t0, t1;
    rs_chunk(ntiles, t0, t1);
    if (t0 >= t1) return;
    int C;                                               // children handed out in front of the chunk
    {
        const int s0 = t0 / RP_SUPER;
        unsigned long long acc = 0ull;
        for (int k = lane; k < s0; k += 32) acc += super_c[k];
        for (int k = s0 * RP_SUPER + lane; k < t0; k += 32) acc += tile_c[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SMCB_FULL, acc, o);
        C = (int)acc;
    }
    for (int tile = t0; tile < t1; ++tile) {
        const int g0 = tile * RP_TILE + lane * RP_ITEMS;
        const int nv = min(max(n32 - g0, 0), RP_ITEMS);
        uint32_t c[RP_ITEMS];
        int run = 0;
        unsigned zbits = 0u;
        if (nv == RP_ITEMS) {
            uint32_t v[8];
            ld_global_v8(reinterpret_cast<const uint32_t*>(a.count) + g0, v);
#pragma unroll
            for (int k = 0; k < RP_ITEMS; ++k) { if (v[k] == 0u) zbits |= 1u << k; run += (int)v[k]; c[k] = (uint32_t)run; }
        } else {
#pragma unroll
            for (int k = 0; k < RP_ITEMS; ++k) {
                const int v = k < nv ? a.count[g0 + k] : 0;
                if (k < nv && v == 0) zbits |= 1u << k;
                run += v; c[k] = (uint32_t)run;
            }
        }
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(SMCB_FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const int excl = C + inc - run;
        C += __shfl_sync(SMCB_FULL, inc, 31);
#pragma unroll
        for (int k = 0; k < RP_ITEMS; ++k) c[k] += (uint32_t)excl;
        st_global_v8(a.cslot + g0, c);                   // padded to whole tiles: no guard
        {
            unsigned m = zbits << (8 * (lane & 3));
            m |= __shfl_xor_sync(SMCB_FULL, m, 1);
            m |= __shfl_xor_sync(SMCB_FULL, m, 2);
            if ((lane & 3) == 0 && g0 < n32) a.am.zmask[g0 >> 5] = m;
        }
        int z = __popc(zbits);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(SMCB_FULL, z, o);
        if (lane == 0) {
            tile_z[tile] = (unsigned long long)z;
            atomicAdd(super_z + tile / RP_SUPER, (unsigned long long)z);
        }
    }
}

// Host side: one grid for the three passes (all SMs x resident blocks of the widest kernel, capped by the work available).
inline int resample_pass_grid(long long n, int blocks_per_sm = 3) {
    int dev = 0, sms = 0;
    SMCB_CUDA(cudaGetDevice(&dev));
    SMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long warps_per_block = RP_THREADS / 32;
    long long need = ((n + RP_TILE - 1) / RP_TILE + warps_per_block - 1) / warps_per_block, g = (long long)sms * blocks_per_sm;
    if (g > need) g = need;
    return (int)(g < 1 ? 1 : g);
}
// Entries of ScanArgs::cslot (children handed out up to each slot), padded to whole tiles.
inline size_t resample_cslot_entries(long long n) { return (size_t)((n + RP_TILE - 1) / RP_TILE) * RP_TILE; }

template <int SRC, int MODE, bool ISLAND = false>
inline void launch_resample_passes(const ScanArgs& a, cudaStream_t stream, const std::function<void()>* after = nullptr) {
    const int grid = resample_pass_grid(a.n);
    launch_chained(k_rs_mass<SRC, ISLAND>, resample_pass_grid(a.n, SMCB_RS_MASS_BPS), RP_THREADS, stream, a);   // pure streaming: more warps in flight
    if (after && *after) (*after)();
    launch_chained(k_rs_count<SRC, MODE, ISLAND>, grid, RP_THREADS, stream, a);
    if (after && *after) (*after)();
    launch_chained(k_rs_map<ISLAND>, resample_pass_grid(a.n, SMCB_RS_MAP_MINB), RP_THREADS, stream, a);
}

// ---- counts -> ancestor map (injected counts, multinomial / residual paths) -----------------------------
// Same map as above but the children counts are given: two independent look-back chains (zero slots, surplus).
struct CountMapArgs {
    const int* count;
    long long n;
    const int* enable_ptr;
    unsigned int* ticket;
    ScanWs ws;
    AncestorMap am;
    unsigned int* tail_info;
};

static __global__ void __launch_bounds__(SCAN_THREADS) k_count_map(CountMapArgs a) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
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
        {
            // E[k] in the closed form used by finish_heavy_runs: E = C - k + Z with C = sum of counts before k; here the
            // counts may contain zeros that are NOT reflected in a cumulative-children value, so track E directly.
            // Short runs are written by their own thread; a long run (a parent with many children: skewed weights right before
            // an ESS-triggered resampling) is written by the whole warp, lane-strided, instead of one thread looping over it.
            long long Ek = E;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                const int e = cnt[k] - 1;
                const bool heavy = e > SURPLUS_SMALL;
                if (e > 0 && !heavy)
                    for (int r = 0; r < e; ++r) if (Ek + r < n) a.am.surplus[Ek + r] = (uint32_t)(g0 + k);
                unsigned hv = __ballot_sync(SMCB_FULL, heavy);
                while (hv) {
                    const int src = __ffs(hv) - 1;
                    hv &= hv - 1u;
                    const long long e0 = __shfl_sync(SMCB_FULL, Ek, src);
                    const int ee = __shfl_sync(SMCB_FULL, e, src);
                    const uint32_t par = (uint32_t)__shfl_sync(SMCB_FULL, g0 + k, src);
                    for (int r = lane; r < ee; r += 32) if (e0 + r < n) a.am.surplus[e0 + r] = par;
                }
                if (e > 0) Ek += e;
            }
            E = Ek;
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

// ---- MULTINOMIAL / RESIDUAL: the engine's native definitions (the reference delegates to R's rmultinom,
// sampler.h:790,801, which is absent and inherently sequential) ------------------------------------------
//   q_i = rint(w_i 2^52);  MULTINOMIAL scans q; RESIDUAL takes floor_i = floor(n q_i / 2^52) deterministic children
//   and scans the (right-shifted) 52-bit fractions.  Draw j: m_j = floor(U_j 2^53), t_j = floor(m_j Q / 2^53),
//   ancestor = min{k : C_k > t_j} over the inclusive integer scan C (Q = C_{n-1}).  All integer, all exact.
struct CumArgs {
    const double* in;            // SRC_LOGW: log-weights; SRC_WEIGHTS: normalised weights
    long long n;
    const double* lse_ptr;
    const int* enable_ptr;
    unsigned int* ticket;
    unsigned long long* status;  // [ntiles]
    unsigned long long* cum;     // [n] out: inclusive scan
    int* count;                  // [n] out: zeros (MULTINOMIAL) or floor counts (RESIDUAL)
    unsigned long long* floor_total;  // RESIDUAL: sum of floor counts (device scalar, zeroed by the caller)
    int shift;                   // RESIDUAL: right shift applied to the 52-bit fractions
};

template <int SRC, bool RESID>
__global__ void __launch_bounds__(SCAN_THREADS) k_cum_scan(CumArgs a) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    __shared__ unsigned long long s_q[SCAN_PAD];
    __shared__ unsigned long long s_warp[SCAN_WARPS];
    __shared__ unsigned long long s_bcast;
    __shared__ int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n = a.n;
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    const double lse = (SRC == SRC_LOGW) ? *a.lse_ptr : 0.0;
    unsigned long long fsum = 0ull;
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
            unsigned long long v = 0ull;
            if (gi < n) {
                double w = (SRC == SRC_LOGW) ? fm_exp(a.in[gi] - lse) : a.in[gi];
                unsigned long long q = (unsigned long long)__double2ll_rn(w * TWO52);
                if (RESID) {
                    unsigned long long hi = __umul64hi(q, (unsigned long long)n), lo = q * (unsigned long long)n;
                    unsigned long long fl = (hi << 12) | (lo >> 52);           // floor(n q / 2^52)
                    unsigned long long fr = lo & ((1ull << 52) - 1ull);         // 52-bit fraction
                    a.count[gi] = (int)fl;
                    fsum += fl;
                    v = fr >> a.shift;
                } else {
                    a.count[gi] = 0;
                    v = q;
                }
            }
            s_q[i + (i >> 3)] = v;
        }
        __syncthreads();
        unsigned long long cum[SCAN_ITEMS];
        {
            unsigned long long run = 0ull;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) { run += s_q[tid * (SCAN_ITEMS + 1) + k]; cum[k] = run; }
        }
        const unsigned long long tot = cum[SCAN_ITEMS - 1];
        unsigned long long inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(SMCB_FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned long long off = inc - tot, agg = 0ull;
#pragma unroll
        for (int w = 0; w < SCAN_WARPS; ++w) { unsigned long long v = s_warp[w]; if (w < warp) off += v; agg += v; }
        if (warp == 0) {
            unsigned long long ex = lookback_exclusive(a.status, tile, agg);
            if (lane == 0) s_bcast = ex;
        }
        __syncthreads();
        const unsigned long long excl = s_bcast + off;
        const long long g0 = base + (long long)tid * SCAN_ITEMS;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) if (g0 + k < n) a.cum[g0 + k] = excl + cum[k];
        __syncthreads();
    }
    if (RESID) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fsum += __shfl_xor_sync(SMCB_FULL, fsum, o);
        if (lane == 0 && fsum) atomicAdd(a.floor_total, fsum);
    }
}

// The same integer scan as two streaming passes without inter-block waiting (tile totals, then chunk-local scans from the
// exact prefix), in the structure of k_rs_mass / k_rs_count: 233 -> about 120 us at 2^24 slots.  Tile totals go to `status`
// (statusA), super totals are accumulated atomically in `super` (statusB, zero on entry).
template <int SRC, bool RESID>
__device__ __forceinline__ unsigned long long cum_value(double raw, double lse, long long n, int shift, unsigned long long& fl) {
    const double w = (SRC == SRC_LOGW) ? fm_exp(raw - lse) : raw;
    const unsigned long long q = (unsigned long long)__double2ll_rn(w * TWO52);
    fl = 0ull;
    if (!RESID) return q;
    const unsigned long long hi = __umul64hi(q, (unsigned long long)n), lo = q * (unsigned long long)n;
    fl = (hi << 12) | (lo >> 52);                                   // floor(n q / 2^52)
    return (lo & ((1ull << 52) - 1ull)) >> shift;                   // 52-bit fraction, right-shifted
}
template <int SRC, bool RESID>
__global__ void __launch_bounds__(RP_THREADS) k_cum_tiles(CumArgs a, unsigned long long* super) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    const int lane = threadIdx.x & 31;
    const int n32 = (int)a.n;
    const int ntiles = (n32 + RP_TILE - 1) / RP_TILE;
    const double lse = (SRC == SRC_LOGW) ? *a.lse_ptr : 0.0;
    const int nw = (int)(gridDim.x * (RP_THREADS / 32));
    unsigned long long fsum = 0ull;
    for (int tile = (int)(blockIdx.x * (RP_THREADS / 32) + (threadIdx.x >> 5)); tile < ntiles; tile += nw) {
        const int base = tile * RP_TILE;
        unsigned long long sum = 0ull;
#pragma unroll
        for (int r = 0; r < RP_ITEMS; ++r) {
            const int i = base + r * 32 + lane;
            if (i < n32) {
                unsigned long long fl;
                sum += cum_value<SRC, RESID>(a.in[i], lse, a.n, a.shift, fl);
                a.count[i] = (int)fl;
                fsum += fl;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(SMCB_FULL, sum, o);
        if (lane == 0) {
            a.status[tile] = sum;
            atomicAdd(super + tile / RP_SUPER, sum);
        }
    }
    if (RESID) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fsum += __shfl_xor_sync(SMCB_FULL, fsum, o);
        if (lane == 0 && fsum) atomicAdd(a.floor_total, fsum);
    }
}
template <int SRC, bool RESID>
__global__ void __launch_bounds__(RP_THREADS) k_cum_write(CumArgs a, const unsigned long long* super) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    const int lane = threadIdx.x & 31;
    const int n32 = (int)a.n;
    const int ntiles = (n32 + RP_TILE - 1) / RP_TILE;
    const double lse = (SRC == SRC_LOGW) ? *a.lse_ptr : 0.0;
    int t0, t1;
    rs_chunk(ntiles, t0, t1);
    if (t0 >= t1) return;
    unsigned long long P;
    {
        const int s0 = t0 / RP_SUPER;
        unsigned long long acc = 0ull;
        for (int k = lane; k < s0; k += 32) acc += super[k];
        for (int k = s0 * RP_SUPER + lane; k < t0; k += 32) acc += a.status[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SMCB_FULL, acc, o);
        P = acc;
    }
    for (int tile = t0; tile < t1; ++tile) {
        const int g0 = tile * RP_TILE + lane * RP_ITEMS;
        const int nv = min(max(n32 - g0, 0), RP_ITEMS);
        double raw[RP_ITEMS];
        if (nv == RP_ITEMS) {
            ld_global_v4(a.in + g0, raw[0], raw[1], raw[2], raw[3]);
            ld_global_v4(a.in + g0 + 4, raw[4], raw[5], raw[6], raw[7]);
        } else {
#pragma unroll
            for (int k = 0; k < RP_ITEMS; ++k) raw[k] = k < nv ? a.in[g0 + k] : 0.0;
        }
        unsigned long long c[RP_ITEMS], run = 0ull;
#pragma unroll
        for (int k = 0; k < RP_ITEMS; ++k) {
            unsigned long long fl;
            const unsigned long long v = cum_value<SRC, RESID>(raw[k], lse, a.n, a.shift, fl);
            run += k < nv ? v : 0ull;
            c[k] = run;
        }
        unsigned long long inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(SMCB_FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const unsigned long long excl = P + inc - run;
        P += __shfl_sync(SMCB_FULL, inc, 31);
        if (nv == RP_ITEMS) {
#pragma unroll
            for (int k = 0; k < RP_ITEMS; k += 2) *reinterpret_cast<ulonglong2*>(a.cum + g0 + k) = make_ulonglong2(excl + c[k], excl + c[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < RP_ITEMS; ++k) if (k < nv) a.cum[g0 + k] = excl + c[k];
        }
    }
}

struct DrawArgs {
    const unsigned long long* cum;   // [n] inclusive integer scan
    long long n;
    const int* enable_ptr;
    const unsigned long long* floor_total;  // RESIDUAL: draws = n - *floor_total; null -> n draws
    const double* uniforms;          // injected U[step][0..n), or null -> Philox(seed, step)
    const long long* step_ptr;
    unsigned long long seed;
    const unsigned long long* seed_ptr;   // optional device-resident Philox key (overrides `seed`)
    int* count;                      // [n] in/out: += children per parent
    uint32_t* bucket;                // [2 * nb + 2] search guide, nb = draw_buckets(n)
    int log_nb;
};

// Guide table of the inverse-CDF search.  A draw t falls in bucket t >> s (s chosen so that the total mass spans between nb
// and 2 nb buckets); bucket[b] = first slot whose cumulative mass reaches b << s.  The answer for t (first k with cum[k] > t)
// then lies in [bucket[b], bucket[b + 1]], on average 8-16 slots = one or two cache lines, instead of the ~12 scattered
// lines a full binary search over 2^24 slots misses on.  Neighbouring buckets search neighbouring regions, so building the
// table is cheap.
inline int draw_log_buckets(long long n) {
    int lb = 0;
    while ((16LL << lb) < n) ++lb;   // ~ n / 16 buckets
    return lb;
}
__device__ __forceinline__ int draw_bucket_shift(unsigned long long Q, int log_nb) {
    const int bits = 64 - __clzll((long long)(Q | 1ull));   // Q < 2^bits
    return bits > log_nb + 1 ? bits - (log_nb + 1) : 0;     // nb <= Q >> s < 2 nb
}

static __global__ void __launch_bounds__(256) k_draw_buckets(DrawArgs a) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    const long long n = a.n;
    const unsigned long long Q = a.cum[n - 1];
    const int s = draw_bucket_shift(Q, a.log_nb);
    const long long nb = (long long)(Q >> s) + 2;   // buckets 0 .. (Q >> s) + 1
    for (long long b = (long long)blockIdx.x * 256 + threadIdx.x; b < nb; b += (long long)gridDim.x * 256) {
        const unsigned long long thr = (unsigned long long)b << s;
        long long l = 0, h = n - 1;                   // min k with cum[k] >= thr (n - 1 if none)
        while (l < h) {
            long long mid = (l + h) >> 1;
            if (a.cum[mid] >= thr) h = mid; else l = mid + 1;
        }
        a.bucket[b] = (uint32_t)l;
    }
}

static __global__ void __launch_bounds__(256) k_multinomial_draw(DrawArgs a) {
    if (a.enable_ptr && *a.enable_ptr == 0) return;
    const long long n = a.n;
    const long long ndraws = a.floor_total ? n - (long long)*a.floor_total : n;
    const long long step = a.step_ptr ? *a.step_ptr : 0;
    const unsigned long long Q = a.cum[n - 1];
    const unsigned long long seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    const int s = draw_bucket_shift(Q, a.log_nb);
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < ndraws; j += (long long)gridDim.x * blockDim.x) {
        unsigned long long m;
        if (a.uniforms) m = (unsigned long long)(a.uniforms[step * n + j] * 9007199254740992.0);
        else m = philox_draw(seed, STREAM_RESAMPLE, (uint32_t)step, (uint64_t)j, 2).a >> 11;
        // t = floor(m Q / 2^53)
        unsigned long long hi = __umul64hi(m, Q), lo = m * Q;
        unsigned long long t = (hi << 11) | (lo >> 53);
        const unsigned long long b = t >> s;
        long long l = a.bucket[b], h = a.bucket[b + 1];  // min k with cum[k] > t lies in [l, h]
        while (l < h) {
            long long mid = (l + h) >> 1;
            if (__ldg(a.cum + mid) > t) h = mid; else l = mid + 1;
        }
        atomicAdd(a.count + l, 1);
    }
}

// Right shift of the 52-bit residual fractions: their sum over n slots must stay below 2^61, inside the 62-bit payload of a
// look-back status word (a total of ~2^62 at n = 2^20 used to spill into the flag bits of the last tiles' prefixes).
inline int residual_shift(long long n) {
    int nb = 0;
    while ((1LL << nb) < n) ++nb;
    return nb > 9 ? nb - 9 : 0;
}

struct DrawWs {                      // extra workspace of the MULTINOMIAL / RESIDUAL path
    unsigned long long* cum;         // [n]
    int* count;                      // [n]
    unsigned long long* floor_total; // device scalar (zero on entry)
    unsigned int* ticket2;           // second tile ticket (zero on entry)
    uint32_t* bucket;                // [draw_bucket_entries(n)] search guide of the inverse-CDF draws
    uint32_t* cslot = nullptr;       // [resample_cslot_entries(n)] optional: counts -> map through the streaming passes
};
inline size_t draw_bucket_entries(long long n) { return ((size_t)2 << draw_log_buckets(n)) + 2; }

// weights -> integer scan -> iid draws -> children counts -> ancestor map.  `ticket`, `ticket2`, `floor_total` and the
// three status arrays must be zero on entry.
template <int SRC>
inline void launch_resample_draws(int mode, const double* in, long long n, const double* lse_ptr, const int* enable_ptr,
                                  const double* uniforms, const long long* step_ptr, unsigned long long seed,
                                  unsigned int* ticket, const ScanWs& ws, const DrawWs& dw, const AncestorMap& am,
                                  unsigned int* tail_info, cudaStream_t stream, const unsigned long long* seed_ptr = nullptr) {
    // launch grids of the current device (cached per device: one process may drive several)
    static int grids[64][3] = {};
    int dev = 0;
    SMCB_CUDA(cudaGetDevice(&dev));
    int (&gd)[3] = grids[dev & 63];
    int &grid_cum = gd[0], &grid_map = gd[1], &grid_draw = gd[2];
    if (!grid_cum) {
        int sms = 0, b = 0;
        SMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        SMCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_cum_scan<SRC, false>, SCAN_THREADS, 0));
        grid_cum = sms * (b < 1 ? 1 : b);
        SMCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_count_map, SCAN_THREADS, 0));
        grid_map = sms * (b < 1 ? 1 : b);
        grid_draw = sms * 8;
    }
    const long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    CumArgs c{};
    c.in = in; c.n = n; c.lse_ptr = lse_ptr; c.enable_ptr = enable_ptr; c.ticket = ticket; c.status = ws.statusA;
    c.cum = dw.cum; c.count = dw.count; c.floor_total = dw.floor_total; c.shift = residual_shift(n);
    int g = (int)(tiles < grid_cum ? tiles : grid_cum);
    if (dw.cslot) {   // engine path: streaming passes (tile totals in statusA, super totals in statusB; both free again afterwards)
        const int g1 = resample_pass_grid(n, 8), g2 = resample_pass_grid(n);
        if (mode == RESIDUAL) { k_cum_tiles<SRC, true><<<g1, RP_THREADS, 0, stream>>>(c, ws.statusB); k_cum_write<SRC, true><<<g2, RP_THREADS, 0, stream>>>(c, ws.statusB); }
        else { k_cum_tiles<SRC, false><<<g1, RP_THREADS, 0, stream>>>(c, ws.statusB); k_cum_write<SRC, false><<<g2, RP_THREADS, 0, stream>>>(c, ws.statusB); }
    } else if (mode == RESIDUAL) k_cum_scan<SRC, true><<<g, SCAN_THREADS, 0, stream>>>(c);
    else k_cum_scan<SRC, false><<<g, SCAN_THREADS, 0, stream>>>(c);
    SMCB_CUDA(cudaGetLastError());
    DrawArgs d{};
    d.cum = dw.cum; d.n = n; d.enable_ptr = enable_ptr; d.floor_total = mode == RESIDUAL ? dw.floor_total : nullptr;
    d.uniforms = uniforms; d.step_ptr = step_ptr; d.seed = seed; d.seed_ptr = seed_ptr; d.count = dw.count;
    d.bucket = dw.bucket; d.log_nb = draw_log_buckets(n);
    long long need = ((2LL << d.log_nb) + 2 + 255) / 256;
    k_draw_buckets<<<(int)(need < grid_draw ? need : grid_draw), 256, 0, stream>>>(d);
    SMCB_CUDA(cudaGetLastError());
    need = (n + 255) / 256;
    k_multinomial_draw<<<(int)(need < grid_draw ? need : grid_draw), 256, 0, stream>>>(d);
    SMCB_CUDA(cudaGetLastError());
    if (dw.cslot) {
        // counts -> ancestor map as three streaming passes (tile totals / scan + zero-slot bitmap / k_rs_map): 85 us at 2^24 slots
        // against 326 us for the single-pass look-back kernel below.  statusA is free again (the integer scan is finished),
        // statusB / statusC are still zero.
        CountPassArgs cp{dw.count, n, enable_ptr, ws, am, dw.cslot};
        const int gp = resample_pass_grid(n);
        k_cnt_tiles<<<resample_pass_grid(n, 8), RP_THREADS, 0, stream>>>(cp);
        SMCB_CUDA(cudaGetLastError());
        k_cnt_scan<<<gp, RP_THREADS, 0, stream>>>(cp);
        SMCB_CUDA(cudaGetLastError());
        ScanArgs ma{};
        ma.n = n; ma.enable_ptr = enable_ptr; ma.step_ptr = step_ptr; ma.ws = ws; ma.am = am; ma.cslot = dw.cslot; ma.tail_info = tail_info;
        k_rs_map<false><<<resample_pass_grid(n, SMCB_RS_MAP_MINB), RP_THREADS, 0, stream>>>(ma);
        SMCB_CUDA(cudaGetLastError());
        return;
    }
    CountMapArgs m{dw.count, n, enable_ptr, dw.ticket2, ws, am, tail_info};
    g = (int)(tiles < grid_map ? tiles : grid_map);
    k_count_map<<<g, SCAN_THREADS, 0, stream>>>(m);
    SMCB_CUDA(cudaGetLastError());
}

// Materialise uRSIndices (sampler.h:845-857 layout) from the ancestor map.
static __global__ void k_decode_indices(AncestorMap am, long long n, uint32_t* idx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = decode_ancestor(am, i);
}

static __global__ void k_zero_u64(unsigned long long* p, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0ull;
}

}  // namespace smcb
