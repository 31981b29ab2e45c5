// blockpfGaussianOpt on device: block-sampling particle filter whose particles are whole paths.
// Replaces blockpfGaussianOpt_impl (reference src/blockpfgaussianopt.cpp:41-71; pfInitialise :81-86, pfMove :94-132) driven
// through sampler::Initialise / IterateEss (sampler.h:481-550, :701-764) with SYSTEMATIC resampling at ESS < N/2.
//
// Layout.  The reference stores one arma::vec of length T per particle and deep-copies whole paths at every resampling
// (sampler.h:860-867), O(T N) per resample.  Here a path is split by age:
//   live window   live[2][lag][N]  the trailing rows a future move can still read (time row tau lives at ring slot tau % lag);
//                                  double-buffered by step parity, every live row is rewritten at every step, so a step
//                                  needs only ONE old value per particle: row t-lag of its ancestor.
//   frozen rows   frozen[T][N]     row tau is written once, by the step that reads it as the window head (t = tau + lag),
//                                  indexed by the slot numbering of the population after step tau + lag - 1.
//   ancestors     anc[T][N]        explicit ancestor index of each slot for the steps that resampled.
// The paths the reference returns are rebuilt once at the end by walking the ancestor rows backwards (k_bpf_collect):
// O(lag N) bytes per step + O(T N) at the end instead of O(T N) per resampling.
#include <cmath>
#include <vector>

#include "cabi_common.cuh"
#include "la_engine.cuh"

using namespace smcb;
using namespace smcb_abi;

namespace {

struct BpfCtrl {
    long long T;          // sampler::T
    double lse, path, ess, thr, log_n, u_res;
    int resampled;
    unsigned int done, ticket;
};

struct BpfArgs {
    const double* y;
    long long t_total, n;
    int lag_max;
    long long zstride;    // lag the caller addressed its injected noise with (>= lag_max)
    const double* sig;    // sigma(k), k = 0..lag_max   (blockpfgaussianopt.cpp:111-119; particle independent)
    const double* sigh;   // sigmah(k)
    double* live[2];
    double* frozen;
    double* lw;
    uint32_t* anc;
    int* resflag;
    double* ess_tr;
    BpfCtrl* ctrl;
    double* partial;
    unsigned long long seed;
    const double* zin;    // injected normals z[(t * lag_max + k) * N + i], or null
    const double* uin;    // injected systematic base uniforms, one per step, or null
};

__device__ __forceinline__ double bpf_normal(const BpfArgs& a, long long t, int k, long long i, double& spare) {
    if (a.zin) return a.zin[((size_t)t * (size_t)a.zstride + (size_t)k) * (size_t)a.n + (size_t)i];
    if (k & 1) return spare;
    double z0;
    philox_normal2(a.seed, t == 0 ? STREAM_INIT : STREAM_MOVE, (uint32_t)t, (uint64_t)i, (uint32_t)(k >> 1), z0, spare);
    return z0;
}

// One kernel per time step: gather the window head through the previous step's ancestors, freeze it, run the forward
// filter / backward sampler of the block proposal, update the log-weight, and reduce (max, sum e, sum e^2) over the grid;
// the last block turns them into logNC increment, ESS and the resampling decision (sampler.h:707-724).
template <bool INIT>
__global__ void __launch_bounds__(256) k_bpf_step(BpfArgs a) {
    BpfCtrl* c = a.ctrl;
    const long long n = a.n;
    const long long t = INIT ? 0 : c->T + 1;
    const int L = a.lag_max;
    const int resampled = INIT ? 0 : c->resampled;
    const double lse_prev = c->lse, log_n = c->log_n;
    const int lag = (int)(t < L ? t : L);
    const double* src = a.live[(t + 1) & 1];
    double* dst = a.live[t & 1];
    const uint32_t* A = INIT ? nullptr : a.anc + (size_t)(t - 1) * (size_t)n;
    const size_t head_slot = INIT ? 0 : (size_t)((t - lag) % L);
    const bool freeze = !INIT && lag == L;
    const double inv_sqrt2 = 1.0 / sqrt(2.0);
    double sd_last = 0.0, sd_back = 0.0, sigh_lag = 0.0;
    if (!INIT && t >= 2) {
        const double sl = a.sig[lag];
        sd_last = sqrt(sl); sd_back = sqrt(sl / (sl + 1)); sigh_lag = a.sigh[lag];
    }
    LseAcc<0> acc; acc.clear();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        double spare = 0.0, lw;
        if (INIT) {
            dst[i] = 0.5 * a.y[0] + inv_sqrt2 * bpf_normal(a, 0, 0, i, spare);
            lw = 1.0 - log_n;   // logweight = 1.0 (blockpfgaussianopt.cpp:85), then the 1/N scaling (sampler.h:495)
        } else {
            const size_t par = resampled ? (size_t)A[i] : (size_t)i;
            const double head = src[head_slot * (size_t)n + par];
            lw = resampled ? -log_n : a.lw[i] - lse_prev;
            if (freeze) a.frozen[(size_t)(t - lag) * (size_t)n + (size_t)i] = head;
            else dst[i] = head;   // t < lag_max: row 0 stays in the window (ring slot 0)
            if (t == 1) {
                dst[(size_t)(1 % L) * (size_t)n + i] = (head + a.y[1]) / 2.0 + (0.0 + inv_sqrt2 * bpf_normal(a, 1, 0, i, spare));
                lw += -0.25 * (a.y[1] - head) * (a.y[1] - head);
            } else {
                // forward filtering: mu(k) parked in the row it will be overwritten by (same thread reads it back)
                double mu = head, mu_lm1 = head;
                for (int k = 1; k <= lag; ++k) {
                    const double sh = a.sigh[k];
                    mu_lm1 = mu;
                    mu = (sh * a.y[t - lag + k] + mu) / (sh + 1);
                    if (k < lag) dst[(size_t)((t - lag + k) % L) * (size_t)n + i] = mu;
                }
                // backward sampling
                double v = mu + sd_last * bpf_normal(a, t, 0, i, spare);
                dst[(size_t)(t % L) * (size_t)n + i] = v;
                int kd = 1;
                for (int k = lag - 1; k; --k, ++kd) {
                    double* slot = dst + (size_t)((t - lag + k) % L) * (size_t)n + i;
                    const double sk = a.sig[k];
                    const double mub = (sk * v + *slot) / (sk + 1);
                    v = mub + sd_back * bpf_normal(a, t, kd, i, spare);
                    *slot = v;
                }
                const double d = a.y[t] - mu_lm1;
                lw += -0.5 * (d * d) / (sigh_lag + 1);
            }
        }
        a.lw[i] = lw;
        acc.add(lw, nullptr);
    }
    if (la_lse_to_grid<0>(acc, a.partial, &c->done)) {
        const double lse = acc.m + log(acc.s);
        c->lse = lse;
        c->path = INIT ? lse : c->path + lse;
        c->ess = (acc.s * acc.s) / acc.q;
        c->resampled = c->ess < c->thr ? 1 : 0;
        if (c->resampled) c->u_res = a.uin ? a.uin[t] : u64_to_unit(philox_draw(a.seed, STREAM_RESAMPLE, (uint32_t)t, ~0ull, 1).a);
        a.resflag[t] = c->resampled;
        a.ess_tr[t] = c->ess;
        c->T = t;
        c->ticket = 0u;
        c->done = 0u;
    }
}

__global__ void k_bpf_decode(BpfArgs a, AncestorMap am) {
    const BpfCtrl* c = a.ctrl;
    if (!c->resampled) return;
    uint32_t* row = a.anc + (size_t)c->T * (size_t)a.n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x)
        row[i] = decode_ancestor(am, i);
}

// Rebuild GetParticleWeight() and the [N x T] path matrix (blockpfgaussianopt.cpp:61-66) from window, frozen rows and ancestors.
__global__ void __launch_bounds__(256) k_bpf_collect(BpfArgs a, double* weight, double* values) {
    const BpfCtrl* c = a.ctrl;
    const long long n = a.n, tl = a.t_total - 1;
    const int L = a.lag_max;
    const int res_last = a.resflag[tl];
    const double* cur = a.live[tl & 1];
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        if (weight) weight[i] = res_last ? exp(-c->log_n) : exp(a.lw[i] - c->lse);
        if (!values) continue;
        size_t M = res_last ? (size_t)a.anc[(size_t)tl * (size_t)n + (size_t)i] : (size_t)i;
        const long long first_live = tl - L + 1 > 0 ? tl - L + 1 : 0;
        for (long long tau = tl; tau >= first_live; --tau) values[(size_t)tau * (size_t)n + i] = cur[(size_t)(tau % L) * (size_t)n + M];
        long long s = tl - 1;   // frozen row tau is numbered like the population after step tau + lag - 1
        for (long long tau = tl - L; tau >= 0; --tau, --s) {
            values[(size_t)tau * (size_t)n + i] = a.frozen[(size_t)tau * (size_t)n + M];
            if (tau > 0 && a.resflag[s]) M = (size_t)a.anc[(size_t)s * (size_t)n + M];
        }
    }
}

void run_blockpf(const double* y_host, long T, long particles, long lag, uint64_t seed, const double* noise_host,
                 const double* unif_host, double* weight, double* values, double* lognc, double* ess_out, int* res_out,
                 uint32_t* anc_out, double* values_dev = nullptr) {
    if (!y_host || T < 1 || particles < 1 || lag < 1) throw std::invalid_argument("blockpfGaussianOpt: need data, particles >= 1, lag >= 1");
    if (particles >= (1LL << 31)) throw std::invalid_argument("blockpfGaussianOpt: particles must be < 2^31 (sampler.h:91-95 index width)");
    require_device();
    const long long n = particles;
    const int L = (int)(lag < T ? lag : (T > 1 ? T : 1));   // a window longer than the series never fills
    cudaStream_t st = 0;
    std::vector<double> sig(L + 1), sigh(L + 1);
    sig[0] = 0; sigh[0] = 0;
    for (int k = 1; k <= L; ++k) { sigh[k] = sig[k - 1] + 1; sig[k] = sigh[k] / (sigh[k] + 1); }
    DevBuf d_y(sizeof(double) * T), d_sig(sizeof(double) * (L + 1)), d_sigh(sizeof(double) * (L + 1));
    SMCB_CUDA(cudaMemcpy(d_y.p, y_host, sizeof(double) * T, cudaMemcpyHostToDevice));
    SMCB_CUDA(cudaMemcpy(d_sig.p, sig.data(), sizeof(double) * (L + 1), cudaMemcpyHostToDevice));
    SMCB_CUDA(cudaMemcpy(d_sigh.p, sigh.data(), sizeof(double) * (L + 1), cudaMemcpyHostToDevice));
    // The large buffers (the frozen rows and ancestor rows are T x N) are kept per calling thread and reused while the shape
    // matches: allocating and freeing 1.3 GB per call cost as much as the filter itself at 2^20 paths x 100 steps.
    struct Workspace {
        long T = -1; long long n = -1; int L = -1;
        DevBuf live, frozen, lw, anc, resflag, ess_tr, ctrl, zmask, zbase, surplus, status, cslot;
    };
    thread_local Workspace ws;
    const long long ntiles = scan_tiles(n);
    if (ws.T != T || ws.n != n || ws.L != L) {
        ws = Workspace{};
        ws.live = DevBuf(sizeof(double) * 2 * (size_t)L * n); ws.frozen = DevBuf(sizeof(double) * (size_t)T * n); ws.lw = DevBuf(sizeof(double) * n);
        ws.anc = DevBuf(sizeof(uint32_t) * (size_t)T * n); ws.resflag = DevBuf(sizeof(int) * T); ws.ess_tr = DevBuf(sizeof(double) * T);
        ws.ctrl = DevBuf(sizeof(BpfCtrl));
        ws.zmask = DevBuf(sizeof(uint32_t) * ((n + 31) / 32 + 64)); ws.zbase = DevBuf(sizeof(uint32_t) * ((n + 31) / 32 + 64));
        ws.surplus = DevBuf(sizeof(uint32_t) * n); ws.status = DevBuf(sizeof(unsigned long long) * 3 * ntiles);
        ws.cslot = DevBuf(sizeof(uint32_t) * resample_cslot_entries(n));
        ws.T = T; ws.n = n; ws.L = L;
    }
    DevBuf &live = ws.live, &frozen = ws.frozen, &lw = ws.lw, &anc = ws.anc, &resflag = ws.resflag, &ess_tr = ws.ess_tr, &ctrl = ws.ctrl;
    DevBuf &zmask = ws.zmask, &zbase = ws.zbase, &surplus = ws.surplus, &status = ws.status;
    int dev = 0, sms = 0;
    SMCB_CUDA(cudaGetDevice(&dev));
    SMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = (n + 255) / 256;
    const int grid = (int)(want < 4LL * sms ? want : 4LL * sms);
    DevBuf partial(sizeof(double) * 4 * (size_t)grid);
    DevBuf zin, uin;
    if (noise_host) {
        zin = DevBuf(sizeof(double) * (size_t)T * (size_t)lag * n);
        SMCB_CUDA(cudaMemcpy(zin.p, noise_host, sizeof(double) * (size_t)T * (size_t)lag * n, cudaMemcpyHostToDevice));
    }
    if (unif_host) {
        uin = DevBuf(sizeof(double) * T);
        SMCB_CUDA(cudaMemcpy(uin.p, unif_host, sizeof(double) * T, cudaMemcpyHostToDevice));
    }
    BpfArgs a{};
    a.y = d_y.as<double>(); a.t_total = T; a.n = n; a.lag_max = L;
    a.sig = d_sig.as<double>(); a.sigh = d_sigh.as<double>();
    a.live[0] = live.as<double>(); a.live[1] = live.as<double>() + (size_t)L * n;
    a.frozen = frozen.as<double>(); a.lw = lw.as<double>(); a.anc = anc.as<uint32_t>();
    a.resflag = resflag.as<int>(); a.ess_tr = ess_tr.as<double>(); a.ctrl = ctrl.as<BpfCtrl>(); a.partial = partial.as<double>();
    a.seed = seed; a.zin = zin.as<double>(); a.uin = uin.as<double>();
    a.zstride = lag;
    BpfCtrl h{};
    h.thr = 0.5 * (double)n; h.log_n = std::log((double)n);   // SetResampleParams(SYSTEMATIC, 0.5), sampler.h:885-893
    SMCB_CUDA(cudaMemcpy(ctrl.p, &h, sizeof h, cudaMemcpyHostToDevice));
    AncestorMap am{zmask.as<uint32_t>(), zbase.as<uint32_t>(), surplus.as<uint32_t>()};
    ScanArgs sc{};
    sc.in = a.lw; sc.n = n; sc.lse_ptr = &a.ctrl->lse; sc.enable_ptr = &a.ctrl->resampled; sc.u_ptr = &a.ctrl->u_res;
    sc.step_ptr = &a.ctrl->T; sc.seed = seed; sc.ticket = &a.ctrl->ticket;
    sc.ws = ScanWs{status.as<unsigned long long>(), status.as<unsigned long long>() + ntiles, status.as<unsigned long long>() + 2 * ntiles};
    sc.am = am; sc.n_glob = n;
    sc.cslot = ws.cslot.as<uint32_t>();
    for (long t = 0; t < T; ++t) {
        if (t == 0) k_bpf_step<true><<<grid, 256, 0, st>>>(a);
        else k_bpf_step<false><<<grid, 256, 0, st>>>(a);
        SMCB_CUDA(cudaMemsetAsync(status.p, 0, sizeof(unsigned long long) * 3 * ntiles, st));
        launch_resample_passes<SRC_LOGW, SYSTEMATIC>(sc, st);
        k_bpf_decode<<<grid, 256, 0, st>>>(a, am);
        SMCB_CUDA(cudaGetLastError());
    }
    // the path matrix goes to the caller's DEVICE buffer when one is given (nothing of size N x T crosses PCIe), else to a
    // staging buffer that is copied to the host
    DevBuf d_w(weight ? sizeof(double) * n : 0), d_v(values && !values_dev ? sizeof(double) * (size_t)T * n : 0);
    double* vdst = values_dev ? values_dev : (values ? d_v.as<double>() : nullptr);
    k_bpf_collect<<<grid, 256, 0, st>>>(a, weight ? d_w.as<double>() : nullptr, vdst);
    SMCB_CUDA(cudaGetLastError());
    SMCB_CUDA(cudaStreamSynchronize(st));
    SMCB_CUDA(cudaMemcpy(&h, ctrl.p, sizeof h, cudaMemcpyDeviceToHost));
    if (weight) SMCB_CUDA(cudaMemcpy(weight, d_w.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (values && !values_dev) SMCB_CUDA(cudaMemcpy(values, d_v.p, sizeof(double) * (size_t)T * n, cudaMemcpyDeviceToHost));
    if (lognc) *lognc = h.path;
    if (ess_out) SMCB_CUDA(cudaMemcpy(ess_out, ess_tr.p, sizeof(double) * T, cudaMemcpyDeviceToHost));
    std::vector<int> rs(T);
    SMCB_CUDA(cudaMemcpy(rs.data(), resflag.p, sizeof(int) * T, cudaMemcpyDeviceToHost));
    if (res_out) for (long t = 0; t < T; ++t) res_out[t] = rs[t];
    if (anc_out) {
        for (long t = 0; t < T; ++t) {
            uint32_t* row = anc_out + (size_t)t * n;
            if (rs[t]) SMCB_CUDA(cudaMemcpy(row, anc.as<uint32_t>() + (size_t)t * n, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
            else for (long long i = 0; i < n; ++i) row[i] = (uint32_t)i;
        }
    }
}

}  // namespace

extern "C" int smcb200_blockpfGaussianOpt(const double* data_host, long n_data, long particles, long lag, uint64_t seed,
                                          const double* noise_host, const double* uniforms_host, double* weight_host,
                                          double* values_host, double* lognc_host, double* ess_host, int* resampled_host,
                                          uint32_t* ancestors_host) {
    return guarded([&]() -> int {
        run_blockpf(data_host, n_data, particles, lag, seed, noise_host, uniforms_host, weight_host, values_host, lognc_host,
                    ess_host, resampled_host, ancestors_host);
        return SMCB200_OK;
    });
}

// Same filter with the path matrix left on the device: values_dev is a DEVICE pointer to particles x n_data doubles (column-major,
// as values_host above).  For callers that keep working on the GPU (smoothing functionals, further kernels): the reference's
// List(weight, values, logNC) costs 8 N T bytes over PCIe, which at 2^20 paths x 100 steps is twice the time of the filter itself.
extern "C" int smcb200_blockpfGaussianOpt_device(const double* data_host, long n_data, long particles, long lag, uint64_t seed,
                                                 double* weight_host, double* values_dev, double* lognc_host, double* ess_host,
                                                 int* resampled_host) {
    return guarded([&]() -> int {
        if (!values_dev) throw std::invalid_argument("smcb200_blockpfGaussianOpt_device: values_dev (device pointer) required");
        run_blockpf(data_host, n_data, particles, lag, seed, nullptr, nullptr, weight_host, nullptr, lognc_host, ess_host, resampled_host,
                    nullptr, values_dev);
        return SMCB200_OK;
    });
}
