// Whole filter in ONE launch for small populations (N <= 2^16 by default, SMCB200_SMALL_MAX): the reference's loop Initialise(); IterateUntil(T-1)
// (inst/include/sampler.h:481-550, :701-764; e.g. the inner filter of src/nonLinPMMH.cpp:78-113, run once per MH iteration).
//
// At 10^3..10^5 particles a step's kernels finish faster than they can be launched (measured: 20 us per step with four
// launches replayed from a CUDA graph at 2^16 particles, 2 % of the HBM roofline).  Here a cooperative grid of at most one
// block per SM keeps the whole filter on the device: block b owns a contiguous range of slots, thread t a contiguous run of K
// slots inside it, the population lives in L2, and the steps are separated by EXCHANGES instead of kernel boundaries:
// every block posts a few values (its partial sums / masses / zero-slot counts) into a table, counts itself in on one monotone
// counter, and reads everybody else's once all have arrived; it then derives the step's scalars itself, in the same fixed order
// as every other block (so all blocks hold bit-identical values).  Per step: one exchange when the population is
// kept, four when it is resampled (SYSTEMATIC / STRATIFIED), six for the draw-based schemes.
//
// Same definitions as the large-population kernels (k_step, k_rs_*, k_cum_scan, k_multinomial_draw, k_count_map): Philox
// counters, weight grid, children counts and the slot order of the ancestor map are identical, so a filter gives the same
// ancestors whichever path runs it (tests/test_gpu_small.py).
#pragma once
// (included by engine.cuh after Ctrl / Traces)

namespace smcb {

constexpr int SF_THREADS = 256;
constexpr int SF_MAXB = 160;            // blocks (<= SMs of the device)
constexpr int SF_NV = 12;               // values per block and exchange
constexpr int SF_KINDS = 8;             // exchange points of a step (double-buffered by step parity)
constexpr int SF_HEAVY = 32;            // surplus runs longer than this are written by the whole block
constexpr int SF_QCAP = 128;            // queue of such runs per block and step

struct SfSync {
    unsigned long long* table;     // [SF_KINDS][2][SF_MAXB][SF_NV] values posted by the blocks
    unsigned long long* counter;   // arrivals since the launch (zeroed by the host before every launch)
};

template <class Model>
struct SmallArgs {
    double* xa[Model::D];
    double* xb[Model::D];
    double* lw;
    double* wt;
    long long n;
    int* enc;                   // [n] rank among the zero-count slots, or -1 for a slot that keeps its own particle
    uint32_t* surplus;          // [n] surplus[m] = parent of the m-th zero-count slot
    unsigned long long* cum;    // [n] inclusive integer scan of the weights (MULTINOMIAL / RESIDUAL)
    int* count;                 // [n] children per parent
    SfSync sync;
    Ctrl* ctrl;
    Traces tr;
    typename Model::Params params;
    const double* znoise;       // injected draws, as in StepArgs / ScanArgs / DrawArgs
    const double* ures;
    const double* ustrat;
    const double* udraw;
    long long tcap;
    long long T;                // steps to run, the first being the initialisation
    int mode;
    int k;                      // slots per thread (even)
    int resid_shift;
    int observe_final;
    int dbg;
};

struct SfShared {               // step scalars, identical in every block
    double lse, path, ess, inv_s, cs, cbase, u_res;
    double shift[MAX_SHIFT];
    int res;
};


// Exchange: thread 0 holds this block's `nv` values (raw 64-bit patterns) in mine[]; afterwards out[b][v] holds every block's.
// Doubles as a grid-wide barrier with release / acquire semantics for everything the blocks wrote before it.
// The values go into a table that is double-buffered by the parity of the step (a block may post its values of step t+1 while
// a slower block still reads those of step t, but not those of step t+2: it cannot get past step t+1 before everybody has
// posted there); arrival is counted on ONE monotone counter (`target` = blocks x exchanges so far, the same in every block), so
// only one thread per block spins, on one address, and the table is read once, without flags.
__device__ __forceinline__ void sf_exchange(const SfSync& sy, int kind, int par, unsigned long long& target, const unsigned long long* mine, int nv,
                                            unsigned long long (*out)[SF_NV]) {
    const int G = (int)gridDim.x, tid = (int)threadIdx.x;
    if (G == 1) {   // a population that fits one block: the exchange is a block barrier
        __syncthreads();
        if (tid == 0) for (int v = 0; v < nv; ++v) out[0][v] = mine[v];
        __syncthreads();
        return;
    }
    unsigned long long* tab = sy.table + ((size_t)(kind * 2 + par) * SF_MAXB) * SF_NV;
    target += (unsigned long long)G;
    __syncthreads();
    if (tid == 0) {
        for (int v = 0; v < nv; ++v) __stcg(tab + (size_t)blockIdx.x * SF_NV + v, mine[v]);
        __threadfence();
        atomicAdd(sy.counter, 1ull);
        while (*reinterpret_cast<volatile unsigned long long*>(sy.counter) < target) { }
        __threadfence();
    }
    __syncthreads();
    for (int e = tid; e < G * nv; e += SF_THREADS) {
        const int b = e / nv, v = e - b * nv;
        out[b][v] = __ldcg(tab + (size_t)b * SF_NV + v);
    }
    __syncthreads();
}

// Block-wide exclusive scan of one value per thread (fixed order); returns the exclusive prefix, `total` = block sum.
template <class T>
__device__ __forceinline__ T sf_block_scan(T v, T* s_warp /*[SF_THREADS/32 + 1]*/, T& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T t = __shfl_up_sync(SMCB_FULL, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();   // s_warp may still be read from the previous scan
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    T off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SF_THREADS / 32; ++w) { const T x = s_warp[w]; if (w < warp) off += x; tot += x; }
    total = tot;
    return off + inc - v;
}

template <class Model, bool INJECT>
__global__ void __launch_bounds__(SF_THREADS, 1) k_filter_small(SmallArgs<Model> a) {
    constexpr int D = Model::D, G = Model::NSHIFT, NOBS = Model::NOBS;
    constexpr int NF = 4 + G + NOBS;   // mx, s, q, obs0, g[G], obs[NOBS]
    static_assert(NF <= SF_NV, "exchange entry too small for this model");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = (int)gridDim.x;
    const int n = (int)a.n;
    const int K = a.k;
    const int slot0 = ((int)blockIdx.x * SF_THREADS + tid) * K;     // this thread's first slot
    const double nd = (double)n, inv_n = 1.0 / nd;
    const bool pow2 = (n & (n - 1)) == 0;

    __shared__ unsigned long long s_out[SF_MAXB][SF_NV];
    __shared__ double s_red[SF_THREADS / 32][SF_NV];
    __shared__ double s_scan_d[SF_THREADS / 32 + 1];
    __shared__ unsigned long long s_scan_u[SF_THREADS / 32 + 1];
    __shared__ int s_scan_i[SF_THREADS / 32 + 1];
    __shared__ SfShared S;
    __shared__ int s_qn;
    __shared__ int s_qE[SF_QCAP], s_qe[SF_QCAP];
    __shared__ uint32_t s_qid[SF_QCAP];

    const Ctrl c0 = *a.ctrl;
    const unsigned long long seed = c0.seed;
    const double log_n = c0.log_n, thr = c0.thr;
    typename Model::Params prm = a.params;
    Model::load_dynamic(prm, c0.mparam);
    if (tid == 0) {
        S.lse = 0.0; S.path = 0.0; S.ess = 0.0; S.inv_s = 1.0; S.cs = -log_n; S.cbase = -log_n; S.u_res = 0.0; S.res = 0;
        for (int j = 0; j < MAX_SHIFT; ++j) S.shift[j] = 0.0;
    }
    __syncthreads();
    int parity = 1;                 // the initialisation writes buffer A (destination of parity 1), as in Sampler::fresh_ctrl
    ScanArgs sa{};                  // what children_upto reads
    sa.ustrat = a.ustrat; sa.seed = seed; sa.seed_ptr = nullptr;

    unsigned long long sync_target = 0ull;   // arrivals the next exchange waits for
    // profiling switch (SMCB200_DBG_STEP = 64): cycles spent per phase by thread 0 of block 0, printed at the end
    long long pc[12];
    for (int i = 0; i < 12; ++i) pc[i] = 0;
    long long pt = clock64();
    auto tick = [&](int i) { if ((a.dbg & 64) && blockIdx.x == 0 && tid == 0) { const long long t = clock64(); pc[i] += t - pt; pt = t; } };
    // ancestor of slot s in the population the last step left behind
    auto ancestor = [&](int s, bool resampled) -> int {
        if (!resampled) return s;
        const int e = a.enc[s];
        return e < 0 ? s : (int)a.surplus[e];
    };

    for (long long step = 0; step <= a.T; ++step) {
        const bool init = step == 0, observe_only = step == a.T;
        if (observe_only && !(a.observe_final && NOBS > 0)) break;
        const long long tcur = step;                         // lTime handed to the model (sampler.h:733,776)
        const bool resampled = !init && S.res != 0;
        const double lse_prev = S.lse, cs = S.cs + ((a.dbg & 32) ? 1000.0 : 0.0), cbase = S.cbase;
        double shift[G > 0 ? G : 1];
#pragma unroll
        for (int j = 0; j < G; ++j) shift[j] = S.shift[j];
        double* const* src = parity ? a.xb : a.xa;
        double* const* dst = parity ? a.xa : a.xb;

        // ---- phase A: gather + pfMove / pfInitialise + log-weight + partial sums (one pair of slots at a time)
        WAcc<G> acc;
        acc.clear();
        double obs0 = 0.0, obs[NOBS > 0 ? NOBS : 1];
#pragma unroll
        for (int j = 0; j < NOBS; ++j) obs[j] = 0.0;
        constexpr int NZI = Model::NZ_INIT, NZM = Model::NZ_MOVE;
        for (int kk = 0; kk < K; kk += 2) {
            const int s = slot0 + kk;
            if (s >= n) break;
            const bool has1 = s + 1 < n;
            double x[2][D], lwv[2];
            if (!init) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int sl = (h == 0 || has1) ? s + h : s;
                    const int p = ancestor(sl, resampled);
#pragma unroll
                    for (int d = 0; d < D; ++d) x[h][d] = src[d][p];
                    lwv[h] = resampled ? -log_n : a.lw[sl] - lse_prev;
                }
                if (NOBS > 0) {  // moments of the population as the reference's Integrate sees it (sampler.h:559-571)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double f[NOBS > 0 ? NOBS : 1];
                        Model::observe(x[h], shift, f);
                        double wh = resampled ? 1.0 : fm_exp(lwv[h]);
                        wh = (h == 0 || has1) ? wh : 0.0;
                        obs0 += wh;
#pragma unroll
                        for (int j = 0; j < NOBS; ++j) obs[j] = fma(wh, f[j], obs[j]);
                    }
                }
            }
            if (!observe_only) {
                double z[2][(NZI > NZM ? NZI : NZM) > 0 ? (NZI > NZM ? NZI : NZM) : 1];
                const int nz = init ? NZI : NZM;
                if (INJECT) {
                    const long long off = init ? 0 : (long long)n * NZI + (tcur - 1) * (long long)n * NZM;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        for (int q = 0; q < nz; ++q) z[h][q] = (h == 0 || has1) ? a.znoise[off + (long long)(s + h) * nz + q] : 0.0;
                } else {
                    for (int q = 0; q < nz; ++q)
                        philox_normal2(seed, init ? STREAM_INIT : STREAM_MOVE, (uint32_t)tcur, (uint64_t)(s >> 1), q, z[0][q], z[1][q]);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double lw;
                    if (init) {
                        Model::init(x[h], lw, prm, z[h]);     // moveset.h:133-138
                        lw = lw - log_n;                      // sampler.h:495
                    } else {
                        lw = lwv[h];
                        Model::move(tcur, x[h], lw, prm, z[h]);  // moveset.h:162-168
                    }
                    double hs[G > 0 ? G : 1];
                    Model::shift_stat(x[h], hs);
                    const bool has = h == 0 || has1;
                    const double w = acc.add(has ? lw : -INFINITY, cs, hs);   // the step's only exponential per particle
                    if (has) {
#pragma unroll
                        for (int d = 0; d < D; ++d) dst[d][s + h] = x[h][d];
                        a.lw[s + h] = lw;
                        a.wt[s + h] = w;
                    }
                }
            }
        }
        // block partial sums in a fixed order
        acc.warp_reduce();
        obs0 = warp_sum(obs0);
#pragma unroll
        for (int j = 0; j < NOBS; ++j) obs[j] = warp_sum(obs[j]);
        __syncthreads();
        if (lane == 0) {
            s_red[warp][0] = acc.mx; s_red[warp][1] = acc.s; s_red[warp][2] = acc.q; s_red[warp][3] = obs0;
#pragma unroll
            for (int j = 0; j < G; ++j) s_red[warp][4 + j] = acc.g[j];
#pragma unroll
            for (int j = 0; j < NOBS; ++j) s_red[warp][4 + G + j] = obs[j];
        }
        __syncthreads();
        unsigned long long mine[SF_NV];
        if (tid == 0) {
            double v[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) v[f] = s_red[0][f];
            for (int w = 1; w < SF_THREADS / 32; ++w) {
                v[0] = fmax(v[0], s_red[w][0]);
#pragma unroll
                for (int f = 1; f < NF; ++f) v[f] += s_red[w][f];
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) mine[f] = (unsigned long long)__double_as_longlong(v[f]);
        }
        tick(0);
        sf_exchange(a.sync, 0, (int)(step & 1), sync_target, mine, NF, s_out);
        tick(1);

        // ---- phase B: every block closes the step for itself (same order everywhere => same bits everywhere)
        // every warp merges some of the NF values over the blocks (value f: lane-strided partial sums, then the shuffle tree --
        // the same order of additions in every block); the step's resampling uniform is drawn by another thread meanwhile
        __shared__ double s_fin[SF_NV], s_tmp[SF_NV];
        __shared__ int s_bad;
        auto merged = [&](double* out) {
            for (int f = warp; f < NF; f += SF_THREADS / 32) {
                double v = f == 0 ? -INFINITY : 0.0;
                for (int b = lane; b < nb; b += 32) {
                    const double x = __longlong_as_double((long long)s_out[b][f]);
                    v = f == 0 ? fmax(v, x) : v + x;
                }
                v = f == 0 ? warp_max(v) : warp_sum(v);
                if (lane == 0) out[f] = v;
            }
        };
        merged(s_fin);
        if (tid == SF_THREADS - 1 && !observe_only)
            S.u_res = a.ures ? a.ures[tcur] : u64_to_unit(philox_draw(seed, STREAM_RESAMPLE, (uint32_t)tcur, ~0ull, 1).a);
        __syncthreads();
        if (tid == 0) {
            const double d = s_fin[0] - cs;
            s_bad = (!observe_only && (d <= -250.0 || d >= 250.0) && s_fin[0] != -INFINITY) ? 1 : 0;
        }
        __syncthreads();
        double cs_used = cs;
        if (s_bad) {
            // the predicted shift of the stored weights was unusable (the weight level jumped by more than e^250): redo the sums
            // against the true maximum and rewrite the shifted weights
            const double M = s_fin[0];
            WAcc<G> r;
            r.clear();
            for (int kk = 0; kk < K; ++kk) {
                const int s = slot0 + kk;
                if (s >= n) break;
                double xi[D], hs[G > 0 ? G : 1];
#pragma unroll
                for (int d = 0; d < D; ++d) xi[d] = dst[d][s];
                Model::shift_stat(xi, hs);
                a.wt[s] = r.add(a.lw[s], M, hs);
            }
            r.warp_reduce();
            __syncthreads();
            if (lane == 0) {
                s_red[warp][1] = r.s; s_red[warp][2] = r.q;
#pragma unroll
                for (int j = 0; j < G; ++j) s_red[warp][4 + j] = r.g[j];
            }
            __syncthreads();
            if (tid == 0) {
                double v[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) v[f] = 0.0;
                for (int w = 0; w < SF_THREADS / 32; ++w) {
                    v[1] += s_red[w][1]; v[2] += s_red[w][2];
#pragma unroll
                    for (int j = 0; j < G; ++j) v[4 + j] += s_red[w][4 + j];
                }
#pragma unroll
                for (int f = 0; f < NF; ++f) mine[f] = (unsigned long long)__double_as_longlong(v[f]);
            }
            sf_exchange(a.sync, 1, (int)(step & 1), sync_target, mine, NF, s_out);
            merged(s_tmp);
            __syncthreads();
            if (tid == 0) {
                s_fin[1] = s_tmp[1]; s_fin[2] = s_tmp[2];
#pragma unroll
                for (int j = 0; j < G; ++j) s_fin[4 + j] = s_tmp[4 + j];
            }
            __syncthreads();
            cs_used = M;
        }
        if (tid == 0) {
            const double mx = s_fin[0], ssum = s_fin[1], qsum = s_fin[2], o0 = s_fin[3];
            const long long Tprev = step - 1;
            if (!init && NOBS > 0 && Tprev < a.tcap && blockIdx.x == 0) {
                double v[NOBS > 0 ? NOBS : 1], out[NOBS > 0 ? NOBS : 1];
#pragma unroll
                for (int j = 0; j < NOBS; ++j) v[j] = s_fin[4 + G + j] / o0;
                Model::finish_obs(v, shift, out);
#pragma unroll
                for (int j = 0; j < NOBS; ++j) a.tr.obs[Tprev * NOBS + j] = out[j];
            }
            if (!observe_only) {
                const double lse = cs_used + log(ssum);                // stableLogSumWeights, population.h:46-50
                const double path = (init ? 0.0 : S.path) + lse;       // sampler.h:498-499, :710-711
                const double ess = (ssum * ssum) / qsum;               // sampler.h:413-417
                const int res = ess < thr ? 1 : 0;                     // sampler.h:507, :719
                S.lse = lse; S.path = path; S.ess = ess; S.res = res; S.inv_s = 1.0 / ssum;
#pragma unroll
                for (int j = 0; j < G; ++j) S.shift[j] = s_fin[4 + j] / ssum;
                const double base_next = res ? -log_n : mx - lse;
                double c_next = base_next + (mx - cbase);
                if (!(fabs(c_next) < 1e300)) c_next = base_next;
                S.cbase = base_next; S.cs = c_next;
                if (blockIdx.x == 0 && tcur < a.tcap) { a.tr.lognc[tcur] = path; a.tr.ess[tcur] = ess; a.tr.resampled[tcur] = res; }
            }
        }
        __syncthreads();
        tick(2);
        if (observe_only) break;
        parity ^= 1;
        if (!S.res) continue;

        // ---- resampling: weights -> children counts (count[]) -> ancestor map (enc[], surplus[]); sampler.h:779-868
        const double lse = S.lse, inv_s = S.inv_s;
        int c_start = 0;                  // children handed out before this thread's first slot
        int zc = 0;                       // this thread's zero-count slots
        int ctot_sys = 0;                 // SYSTEMATIC / STRATIFIED: children handed out in total
        if (tid == 0) mine[2] = 0ull;
        if (a.mode == SYSTEMATIC || a.mode == STRATIFIED) {
            // exact cumulative weights on the 2^-52 grid (any summation order gives the same bits)
            double run = 0.0;
            for (int kk = 0; kk < K; ++kk) { const int s = slot0 + kk; if (s < n) run += rs_quantise<SRC_WTILDE>(a.wt[s], 0.0, inv_s); }
            double btot;
            const double texcl = sf_block_scan<double>(run, s_scan_d, btot);
            if (tid == 0) mine[0] = (unsigned long long)__double_as_longlong(btot);
            tick(3);
            sf_exchange(a.sync, 2, (int)(step & 1), sync_target, mine, 1, s_out);
            tick(4);
            double P = 0.0, Wtot = 0.0;
            for (int b = 0; b < nb; ++b) { const double m = __longlong_as_double((long long)s_out[b][0]); if (b < (int)blockIdx.x) P += m; Wtot += m; }
            double W = P + texcl;
            const double dRand = a.mode == SYSTEMATIC ? (0.0 + (inv_n - 0.0) * S.u_res) : 0.0;
            auto upto = [&](double w) -> int {
                if (a.mode == SYSTEMATIC) return pow2 ? children_upto<SYSTEMATIC, true>(w, sa, n, nd, inv_n, dRand, tcur) : children_upto<SYSTEMATIC, false>(w, sa, n, nd, inv_n, dRand, tcur);
                return children_upto<STRATIFIED, false>(w, sa, n, nd, inv_n, dRand, tcur);
            };
            c_start = upto(W);
            ctot_sys = upto(Wtot);
            int cprev = c_start;
            for (int kk = 0; kk < K; ++kk) {
                const int s = slot0 + kk;
                if (s >= n) break;
                W += rs_quantise<SRC_WTILDE>(a.wt[s], 0.0, inv_s);
                const int ck = upto(W);
                a.count[s] = ck - cprev;
                zc += (ck == cprev) ? 1 : 0;
                cprev = ck;
            }
        } else {
            // integer scan of the weights (RESIDUAL: of the fractions, with the floor counts handed out up front)
            const bool resid = a.mode == RESIDUAL;
            unsigned long long run = 0ull, fsum = 0ull;
            auto quant = [&](int s, unsigned long long& fl) -> unsigned long long {
                const double w = fm_exp(a.lw[s] - lse);
                const unsigned long long q = (unsigned long long)__double2ll_rn(w * TWO52);
                fl = 0ull;
                if (!resid) return q;
                const unsigned long long hi = __umul64hi(q, (unsigned long long)n), lo = q * (unsigned long long)n;
                fl = (hi << 12) | (lo >> 52);                                   // floor(n q / 2^52)
                return (lo & ((1ull << 52) - 1ull)) >> a.resid_shift;           // 52-bit fraction, right-shifted
            };
            for (int kk = 0; kk < K; ++kk) {
                const int s = slot0 + kk;
                if (s >= n) break;
                unsigned long long fl;
                run += quant(s, fl);
                fsum += fl;
            }
            unsigned long long btot, ftot;
            const unsigned long long texcl = sf_block_scan<unsigned long long>(run, s_scan_u, btot);
            (void)sf_block_scan<unsigned long long>(fsum, s_scan_u, ftot);
            if (tid == 0) { mine[0] = btot; mine[1] = ftot; }
            sf_exchange(a.sync, 2, (int)(step & 1), sync_target, mine, 2, s_out);
            unsigned long long P = 0ull, Q = 0ull, floor_total = 0ull;
            for (int b = 0; b < nb; ++b) { if (b < (int)blockIdx.x) P += s_out[b][0]; Q += s_out[b][0]; floor_total += s_out[b][1]; }
            unsigned long long cumv = P + texcl;
            for (int kk = 0; kk < K; ++kk) {
                const int s = slot0 + kk;
                if (s >= n) break;
                unsigned long long fl;
                cumv += quant(s, fl);
                a.cum[s] = cumv;
                a.count[s] = (int)fl;
            }
            sf_exchange(a.sync, 3, (int)(step & 1), sync_target, mine, 1, s_out);      // cum[] and the floor counts are complete
            // iid draws through the exact integer inverse CDF (same arithmetic as k_multinomial_draw)
            const long long ndraws = resid ? (long long)n - (long long)floor_total : (long long)n;
            for (long long j = (long long)blockIdx.x * SF_THREADS + tid; j < ndraws; j += (long long)nb * SF_THREADS) {
                unsigned long long m;
                if (a.udraw) m = (unsigned long long)(a.udraw[tcur * (long long)n + j] * 9007199254740992.0);
                else m = philox_draw(seed, STREAM_RESAMPLE, (uint32_t)tcur, (uint64_t)j, 2).a >> 11;
                const unsigned long long hi = __umul64hi(m, Q), lo = m * Q;
                const unsigned long long t = (hi << 11) | (lo >> 53);
                int l = 0, h = n - 1;
                while (l < h) {
                    const int mid = (l + h) >> 1;
                    if (__ldcg(a.cum + mid) > t) h = mid; else l = mid + 1;
                }
                atomicAdd(a.count + l, 1);
            }
            sf_exchange(a.sync, 4, (int)(step & 1), sync_target, mine, 1, s_out);      // the children counts are complete
            int csum = 0;
            for (int kk = 0; kk < K; ++kk) {
                const int s = slot0 + kk;
                if (s >= n) break;
                const int c = __ldcg(a.count + s);
                csum += c;
                zc += c == 0 ? 1 : 0;
            }
            int cbt;
            c_start = sf_block_scan<int>(csum, s_scan_i, cbt);
            if (tid == 0) mine[2] = (unsigned long long)(long long)cbt;     // posted with the zero-slot totals below
        }
        // zero-count slots in front of every thread, children in front of every block
        int zbt;
        const int zexcl = sf_block_scan<int>(zc, s_scan_i, zbt);
        if (tid == 0) { mine[0] = (unsigned long long)(long long)zbt; mine[1] = 0ull; }
        tick(5);
        sf_exchange(a.sync, 5, (int)(step & 1), sync_target, mine, 3, s_out);
        tick(6);
        int Zb = 0, ztotal = 0, Cb = 0, ctot_draws = 0;
        for (int b = 0; b < nb; ++b) {
            const int zb = (int)(long long)s_out[b][0], cb = (int)(long long)s_out[b][2];
            if (b < (int)blockIdx.x) { Zb += zb; Cb += cb; }
            ztotal += zb; ctot_draws += cb;
        }
        const bool draws = a.mode == MULTINOMIAL || a.mode == RESIDUAL;
        if (draws) c_start += Cb;
        if (tid == 0) s_qn = 0;
        __syncthreads();
        {
            int Zk = Zb + zexcl, cprev = c_start;
            for (int kk = 0; kk < K; ++kk) {
                const int s = slot0 + kk;
                if (s >= n) break;
                const int cnt = __ldcg(a.count + s);
                if (cnt == 0) { a.enc[s] = Zk; ++Zk; }
                else {
                    a.enc[s] = -1;
                    const int e = cnt - 1;
                    if (e > 0) {
                        const int E = cprev - s + Zk;          // first surplus child of parent s (the reference's in-place order, sampler.h:846-857)
                        int q = -1;
                        if (e > SF_HEAVY) { q = atomicAdd(&s_qn, 1); if (q >= SF_QCAP) q = -1; }
                        if (q >= 0) { s_qE[q] = E; s_qe[q] = e; s_qid[q] = (uint32_t)s; }
                        else for (int r = 0; r < e; ++r) a.surplus[E + r] = (uint32_t)s;
                    }
                }
                cprev += cnt;
            }
        }
        __syncthreads();
        {
            const int qn = min(s_qn, SF_QCAP);
            for (int q = 0; q < qn; ++q) {
                const int E = s_qE[q], e = s_qe[q];
                const uint32_t id = s_qid[q];
                for (int r = tid; r < e; r += SF_THREADS) a.surplus[E + r] = id;
            }
        }
        if (blockIdx.x == 0) {
            // zero-count slots beyond the children handed out copy particle 0 (uRSIndices is zero-initialised, sampler.h:845)
            const int ctot = draws ? ctot_draws : ctot_sys;
            const int etot = ctot - n + ztotal;
            for (int m = max(etot, 0) + tid; m < ztotal; m += SF_THREADS) a.surplus[m] = 0u;
        }
        tick(7);
        sf_exchange(a.sync, 6, (int)(step & 1), sync_target, mine, 1, s_out);          // the ancestor map is complete
        tick(8);
    }
    if ((a.dbg & 64) && blockIdx.x == 0 && tid == 0)
        printf("[k_filter_small] n=%d blocks=%d T=%lld cycles/step: A %lld | x0 %lld | B %lld | quantise+scan %lld | x2 %lld | counts %lld | x5 %lld | map %lld | x6 %lld\n",
               n, nb, a.T, pc[0] / a.T, pc[1] / a.T, pc[2] / a.T, pc[3] / a.T, pc[4] / a.T, pc[5] / a.T, pc[6] / a.T, pc[7] / a.T, pc[8] / a.T);
    if (blockIdx.x == 0 && tid == 0) {
        Ctrl* c = a.ctrl;
        c->lse = S.lse; c->lognc_path = S.path; c->ess = S.ess; c->resampled = S.res; c->inv_s = S.inv_s;
        c->T = a.T - 1; c->parity = parity; c->u_res = S.u_res; c->cshift = S.cs; c->cbase = S.cbase;
        for (int j = 0; j < MAX_SHIFT; ++j) c->shift[j] = S.shift[j];
    }
}

// uRSIndices after a run of k_filter_small (identity when the last step kept its population)
static __global__ void k_decode_small(const int* enc, const uint32_t* surplus, long long n, uint32_t* idx) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int e = enc[s];
    idx[s] = e < 0 ? (uint32_t)s : surplus[e];
}

}  // namespace smcb
