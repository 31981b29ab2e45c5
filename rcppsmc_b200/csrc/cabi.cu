// C ABI of libsmcb200.so (include/smcb200.h).  Thin: argument checks, host<->device staging, launches.
#include "../../include/smcb200.h"

#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "cabi_common.cuh"
#include "models.cuh"

using namespace smcb;

namespace smcb_abi {
std::string& last_error() { thread_local std::string e; return e; }
}  // namespace smcb_abi
using namespace smcb_abi;

namespace {

// ---- log-sum-exp / ESS reduction (population.h:46-50, sampler.h:413-417) --------------------------------
__global__ void __launch_bounds__(256) k_lse(const double* lw, long long n, double* partial, unsigned int* done, double* out3) {
    LseAcc<0> acc;
    acc.clear();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) acc.add(lw[i], nullptr);
    __shared__ double s_red[8][3];
    __shared__ bool s_last;
    acc.warp_reduce();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[warp][0] = acc.m; s_red[warp][1] = acc.s; s_red[warp][2] = acc.q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        LseAcc<0> b; b.clear();
        for (int w = 0; w < 8; ++w) { LseAcc<0> t; t.m = s_red[w][0]; t.s = s_red[w][1]; t.q = s_red[w][2]; b.merge(t); }
        const int nb = gridDim.x;
        partial[blockIdx.x] = b.m; partial[nb + blockIdx.x] = b.s; partial[2 * nb + blockIdx.x] = b.q;
        __threadfence();
        s_last = atomicAdd(done, 1u) == (unsigned)nb - 1u;
    }
    __syncthreads();
    if (!s_last || warp != 0) return;
    __threadfence();
    const int nb = gridDim.x;
    const volatile double* P = partial;
    LseAcc<0> b; b.clear();
    for (int k = lane; k < nb; k += 32) { LseAcc<0> t; t.m = P[k]; t.s = P[nb + k]; t.q = P[2 * nb + k]; b.merge(t); }
    b.warp_reduce();
    if (lane == 0) { out3[0] = b.m + log(b.s); out3[1] = (b.s * b.s) / b.q; out3[2] = b.m; *done = 0u; }
}

__global__ void k_philox_normals(unsigned long long seed, uint32_t stream, uint32_t step, long long n, double* out) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * p >= n) return;
    double z0, z1;
    philox_normal2(seed, stream, step, (uint64_t)p, 0, z0, z1);
    out[2 * p] = z0;
    if (2 * p + 1 < n) out[2 * p + 1] = z1;
}

// Philox4x32-10 evaluated ON THE DEVICE (the __umulhi path of philox.cuh) for known-answer tests
__global__ void k_philox_raw(const uint32_t* ctr_key, int n, uint32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o[4];
    Philox::rand4(ctr_key[6 * i], ctr_key[6 * i + 1], ctr_key[6 * i + 2], ctr_key[6 * i + 3], ctr_key[6 * i + 4], ctr_key[6 * i + 5], o);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = o[k];
}

__global__ void k_fastmath(int fn, const double* in, long long n, double* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = in[i], s, c;
    switch (fn) {
        case 0: out[i] = fm_exp(x); break;
        case 1: out[i] = fm_neg2log(x); break;
        case 2: fm_sincos2pi(x, s, c); out[i] = s; break;
        case 3: fm_sincos2pi(x, s, c); out[i] = c; break;
        default: out[i] = fm_sqrt(x); break;
    }
}

}  // namespace

// ---- pfNonlinBS handle --------------------------------------------------------------------------------
struct smcb200_pf {
    long long n, tcap, T = 0;
    unsigned long long seed;
    cudaStream_t stream;
    std::unique_ptr<Sampler<NonlinBS>> s;
    DevBuf y, cterm, noise, unif;
    bool has_noise = false, has_unif = false, ran = false;
    int mode = SMCB200_MULTINOMIAL;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long launches = 0;
    std::vector<double> obs;
    bool profile = false;
    std::vector<cudaEvent_t> kev;  // per-launch events when profiling
    size_t kev_used = 0;
};

extern "C" {

const char* smcb200_last_error(void) { return last_error().c_str(); }
int smcb200_version(void) { return 100; }
int smcb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int smcb200_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    if (!ctr || !key || !out) return fail(SMCB200_ERR_INVALID, "null argument");
    uint32_t o[4];
    Philox::rand4(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], o);
    for (int i = 0; i < 4; ++i) out[i] = o[i];
    return SMCB200_OK;
}

int smcb200_philox4x32_device(const uint32_t* ctr_key_host, long n, uint32_t* out_host) {
    return guarded([&]() -> int {
        if (n < 1 || !ctr_key_host || !out_host) throw std::invalid_argument("smcb200_philox4x32_device: n >= 1, inputs and out required");
        require_device();
        DevBuf in(sizeof(uint32_t) * 6 * n), out(sizeof(uint32_t) * 4 * n);
        SMCB_CUDA(cudaMemcpy(in.p, ctr_key_host, sizeof(uint32_t) * 6 * n, cudaMemcpyHostToDevice));
        k_philox_raw<<<(unsigned)((n + 127) / 128), 128>>>(in.as<uint32_t>(), (int)n, out.as<uint32_t>());
        SMCB_CUDA(cudaGetLastError());
        SMCB_CUDA(cudaMemcpy(out_host, out.p, sizeof(uint32_t) * 4 * n, cudaMemcpyDeviceToHost));
        return SMCB200_OK;
    });
}

int smcb200_philox_normals(uint64_t seed, uint32_t stream, uint32_t step, long n, double* out_host) {
    return guarded([&]() -> int {
        if (n < 1 || !out_host) throw std::invalid_argument("smcb200_philox_normals: n >= 1 and out required");
        require_device();
        DevBuf d(sizeof(double) * n);
        long long pairs = (n + 1) / 2;
        k_philox_normals<<<(unsigned)((pairs + 255) / 256), 256>>>(seed, stream, step, n, d.as<double>());
        SMCB_CUDA(cudaGetLastError());
        SMCB_CUDA(cudaMemcpy(out_host, d.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
        return SMCB200_OK;
    });
}

int smcb200_fastmath_eval(int fn, const double* in_host, long n, double* out_host) {
    return guarded([&]() -> int {
        if (n < 1 || !in_host || !out_host || fn < 0 || fn > 4) throw std::invalid_argument("smcb200_fastmath_eval: bad argument");
        require_device();
        DevBuf in(sizeof(double) * n), out(sizeof(double) * n);
        SMCB_CUDA(cudaMemcpy(in.p, in_host, sizeof(double) * n, cudaMemcpyHostToDevice));
        k_fastmath<<<(unsigned)((n + 255) / 256), 256>>>(fn, in.as<double>(), n, out.as<double>());
        SMCB_CUDA(cudaGetLastError());
        SMCB_CUDA(cudaMemcpy(out_host, out.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
        return SMCB200_OK;
    });
}

int smcb200_lse(const double* logw_host, long n, double* out3_host) {
    return guarded([&]() -> int {
        if (n < 1 || !logw_host || !out3_host) throw std::invalid_argument("smcb200_lse: n >= 1, logw and out3 required");
        require_device();
        DevBuf lw(sizeof(double) * n), out(sizeof(double) * 3), done(sizeof(unsigned));
        int grid = persistent_grid(k_lse, 256, n, 256);
        DevBuf partial(sizeof(double) * 3 * grid);
        SMCB_CUDA(cudaMemcpy(lw.p, logw_host, sizeof(double) * n, cudaMemcpyHostToDevice));
        SMCB_CUDA(cudaMemset(done.p, 0, sizeof(unsigned)));
        k_lse<<<grid, 256>>>(lw.as<double>(), n, partial.as<double>(), done.as<unsigned>(), out.as<double>());
        SMCB_CUDA(cudaGetLastError());
        SMCB_CUDA(cudaMemcpy(out3_host, out.p, sizeof(double) * 3, cudaMemcpyDeviceToHost));
        return SMCB200_OK;
    });
}

int smcb200_resample(int mode, const double* weights_host, long n, const double* uniforms_host, long n_uniforms,
                     const int* counts_in_host, unsigned int* idx_out_host, int* counts_out_host) {
    return guarded([&]() -> int {
        if (n < 1 || n > 2147483647L) throw std::invalid_argument("smcb200_resample: n must be in [1, 2^31)");
        if (!idx_out_host) throw std::invalid_argument("smcb200_resample: idx_out required");
        if (mode < 0 || mode > 3) throw std::invalid_argument("smcb200_resample: unknown resampling mode");
        require_device();
        const long long ntiles = scan_tiles(n);
        DevBuf zmask(sizeof(uint32_t) * ((n + 31) / 32 + 64)), zbase(sizeof(uint32_t) * ((n + 31) / 32 + 64));
        DevBuf surplus(sizeof(uint32_t) * n), status(sizeof(unsigned long long) * 3 * ntiles);
        DevBuf misc(sizeof(unsigned) * 4 + sizeof(double) * 2), idx(sizeof(uint32_t) * n), counts(sizeof(int) * n);
        SMCB_CUDA(cudaMemset(status.p, 0, sizeof(unsigned long long) * 3 * ntiles));
        SMCB_CUDA(cudaMemset(misc.p, 0, sizeof(unsigned) * 4 + sizeof(double) * 2));
        AncestorMap am{zmask.as<uint32_t>(), zbase.as<uint32_t>(), surplus.as<uint32_t>()};
        ScanWs ws{status.as<unsigned long long>(), status.as<unsigned long long>() + ntiles, status.as<unsigned long long>() + 2 * ntiles};
        unsigned* ticket = misc.as<unsigned>();
        unsigned* tail = misc.as<unsigned>() + 1;
        double* u_dev = (double*)(misc.as<unsigned>() + 4);
        if (counts_in_host) {
            long long tot = 0;
            for (long i = 0; i < n; ++i) {
                if (counts_in_host[i] < 0) throw std::invalid_argument("smcb200_resample: negative count");
                tot += counts_in_host[i];
            }
            if (tot > n) throw std::invalid_argument("smcb200_resample: counts sum to more than n");
            SMCB_CUDA(cudaMemcpy(counts.p, counts_in_host, sizeof(int) * n, cudaMemcpyHostToDevice));
            CountMapArgs a{counts.as<int>(), n, nullptr, ticket, ws, am, tail};
            int grid = persistent_grid(k_count_map, SCAN_THREADS, n, SCAN_TILE);
            k_count_map<<<grid, SCAN_THREADS>>>(a);
            SMCB_CUDA(cudaGetLastError());
        } else {
            if (!weights_host) throw std::invalid_argument("smcb200_resample: weights required");
            DevBuf w(sizeof(double) * n);
            SMCB_CUDA(cudaMemcpy(w.p, weights_host, sizeof(double) * n, cudaMemcpyHostToDevice));
            ScanArgs a{};
            a.in = w.as<double>(); a.n = n; a.n_glob = n; a.u_ptr = u_dev; a.ticket = ticket; a.ws = ws; a.am = am;
            a.count_out = counts.as<int>(); a.tail_info = tail;
            DevBuf cslot(sizeof(uint32_t) * resample_cslot_entries(n));
            a.cslot = cslot.as<uint32_t>();
            const char* impl_env = getenv("SMCB200_SCAN_IMPL");
            const bool old_impl = impl_env && atoi(impl_env) == 0;   // block-tiled kernel, kept for A/B measurements
            DevBuf us;
            if (mode == SMCB200_SYSTEMATIC) {
                if (!uniforms_host || n_uniforms < 1) throw std::invalid_argument("smcb200_resample: SYSTEMATIC needs 1 uniform");
                SMCB_CUDA(cudaMemcpy(u_dev, uniforms_host, sizeof(double), cudaMemcpyHostToDevice));
                if (old_impl) launch_resample_scan<SRC_WEIGHTS, SYSTEMATIC>(a, 0);
                else launch_resample_passes<SRC_WEIGHTS, SYSTEMATIC>(a, 0);
            } else if (mode == SMCB200_STRATIFIED) {
                if (!uniforms_host || n_uniforms < n + 1) throw std::invalid_argument("smcb200_resample: STRATIFIED needs n+1 uniforms");
                us = DevBuf(sizeof(double) * (n + 1));
                SMCB_CUDA(cudaMemcpy(us.p, uniforms_host, sizeof(double) * (n + 1), cudaMemcpyHostToDevice));
                a.ustrat = us.as<double>();
                if (old_impl) launch_resample_scan<SRC_WEIGHTS, STRATIFIED>(a, 0);
                else launch_resample_passes<SRC_WEIGHTS, STRATIFIED>(a, 0);
            } else {
                if (!uniforms_host || n_uniforms < n) throw std::invalid_argument("smcb200_resample: MULTINOMIAL / RESIDUAL need n uniforms");
                us = DevBuf(sizeof(double) * n);
                SMCB_CUDA(cudaMemcpy(us.p, uniforms_host, sizeof(double) * n, cudaMemcpyHostToDevice));
                DevBuf cum(sizeof(unsigned long long) * n), extra(sizeof(unsigned long long) * 2);
                SMCB_CUDA(cudaMemset(extra.p, 0, sizeof(unsigned long long) * 2));
                DevBuf bucket(sizeof(uint32_t) * draw_bucket_entries(n));
                DrawWs dw{cum.as<unsigned long long>(), counts.as<int>(), extra.as<unsigned long long>(), (unsigned int*)(extra.as<unsigned long long>() + 1), bucket.as<uint32_t>(),
                          cslot.as<uint32_t>()};   // the streaming passes of the engine path (integer scan, sorted draws, counts -> map)
                launch_resample_draws<SRC_WEIGHTS>(mode, w.as<double>(), n, nullptr, nullptr, us.as<double>(), nullptr, 0, ticket, ws, dw, am, tail, 0);
                SMCB_CUDA(cudaDeviceSynchronize());
            }
            SMCB_CUDA(cudaGetLastError());
        }
        k_decode_indices<<<(unsigned)((n + 255) / 256), 256>>>(am, n, idx.as<uint32_t>());
        SMCB_CUDA(cudaGetLastError());
        SMCB_CUDA(cudaMemcpy(idx_out_host, idx.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
        if (counts_out_host) SMCB_CUDA(cudaMemcpy(counts_out_host, counts.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
        return SMCB200_OK;
    });
}

// ---- pfNonlinBS -----------------------------------------------------------------------------------------
int smcb200_pfNonlinBS_create(long particles, long tcap, uint64_t seed, void* stream, smcb200_pf** out) {
    return guarded([&]() -> int {
        if (!out) throw std::invalid_argument("smcb200_pfNonlinBS_create: out required");
        if (particles < 1 || tcap < 1) throw std::invalid_argument("smcb200_pfNonlinBS_create: particles >= 1 and tcap >= 1 required");
        require_device();
        std::unique_ptr<smcb200_pf> h(new smcb200_pf);
        h->n = particles; h->tcap = tcap; h->seed = seed; h->stream = (cudaStream_t)stream;
        h->s.reset(new Sampler<NonlinBS>(particles, tcap, seed, h->stream));
        h->y = DevBuf(sizeof(double) * tcap);
        h->cterm = DevBuf(sizeof(double) * tcap);
        SMCB_CUDA(cudaEventCreate(&h->ev0));
        SMCB_CUDA(cudaEventCreate(&h->ev1));
        *out = h.release();
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_create_island(long particles_local, long tcap, uint64_t seed, void* stream, int rank, int world, smcb200_pf** out) {
    return guarded([&]() -> int {
        if (!out) throw std::invalid_argument("smcb200_pfNonlinBS_create_island: out required");
        if (particles_local < 2 || tcap < 1) throw std::invalid_argument("smcb200_pfNonlinBS_create_island: particles_local >= 2 and tcap >= 1 required");
        require_device();
        std::unique_ptr<smcb200_pf> h(new smcb200_pf);
        h->n = particles_local; h->tcap = tcap; h->seed = seed; h->stream = (cudaStream_t)stream;
        h->s.reset(new Sampler<NonlinBS>(particles_local, tcap, seed, h->stream, rank, world));
        h->y = DevBuf(sizeof(double) * tcap);
        h->cterm = DevBuf(sizeof(double) * tcap);
        SMCB_CUDA(cudaEventCreate(&h->ev0));
        SMCB_CUDA(cudaEventCreate(&h->ev1));
        *out = h.release();
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_ipc_export(smcb200_pf* h, void* handle64_out) {
    return guarded([&]() -> int {
        if (!h || !handle64_out) throw std::invalid_argument("null argument");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
        h->s->ExportIpc((cudaIpcMemHandle_t*)handle64_out);
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_ipc_import(smcb200_pf* h, const void* handles_world_x64) {
    return guarded([&]() -> int {
        if (!h || !handles_world_x64) throw std::invalid_argument("null argument");
        h->s->ImportPeers((const cudaIpcMemHandle_t*)handles_world_x64);
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_set_data(smcb200_pf* h, const double* data_host, long T) {
    return guarded([&]() -> int {
        if (!h || !data_host) throw std::invalid_argument("smcb200_pfNonlinBS_set_data: null argument");
        if (T < 1 || T > h->tcap) throw std::invalid_argument("smcb200_pfNonlinBS_set_data: T must be in [1, tcap]");
        std::vector<double> c(T);
        for (long t = 0; t < T; ++t) c[t] = 8.0 * std::cos(1.2 * (double)t);  // src/pfnonlinbs.cpp:107, particle-independent
        SMCB_CUDA(cudaMemcpyAsync(h->y.p, data_host, sizeof(double) * T, cudaMemcpyHostToDevice, h->stream));
        SMCB_CUDA(cudaMemcpyAsync(h->cterm.p, c.data(), sizeof(double) * T, cudaMemcpyHostToDevice, h->stream));
        SMCB_CUDA(cudaStreamSynchronize(h->stream));
        h->T = T;
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_set_noise(smcb200_pf* h, const double* noise_host, const double* uniforms_host, long n_uniforms) {
    return guarded([&]() -> int {
        if (!h) throw std::invalid_argument("smcb200_pfNonlinBS_set_noise: null handle");
        if (h->T < 1) throw std::runtime_error("smcb200_pfNonlinBS_set_noise: set_data first");
        h->has_noise = noise_host != nullptr;
        h->has_unif = uniforms_host != nullptr;
        if (noise_host) {
            size_t bytes = sizeof(double) * (size_t)h->n * (size_t)h->T;
            h->noise = DevBuf(bytes);
            SMCB_CUDA(cudaMemcpy(h->noise.p, noise_host, bytes, cudaMemcpyHostToDevice));
        }
        if (uniforms_host) {
            if (n_uniforms < h->T) throw std::invalid_argument("smcb200_pfNonlinBS_set_noise: need at least T uniforms");
            h->unif = DevBuf(sizeof(double) * n_uniforms);
            SMCB_CUDA(cudaMemcpy(h->unif.p, uniforms_host, sizeof(double) * n_uniforms, cudaMemcpyHostToDevice));
        }
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_run(smcb200_pf* h, int resample_mode, double threshold) {
    return guarded([&]() -> int {
        if (!h) throw std::invalid_argument("smcb200_pfNonlinBS_run: null handle");
        if (h->T < 1) throw std::runtime_error("smcb200_pfNonlinBS_run: set_data first");
        if (resample_mode < 0 || resample_mode > 3) throw std::invalid_argument("smcb200_pfNonlinBS_run: unknown resampling mode");
        Sampler<NonlinBS>& s = *h->s;
        NonlinBS::Params p{h->y.as<double>(), h->cterm.as<double>()};
        s.SetParams(p);
        s.SetSeed(h->seed);
        s.SetResampleParams(resample_mode, threshold);
        const double* un = h->has_unif ? h->unif.as<double>() : nullptr;
        s.SetInjectedNoise(h->has_noise ? h->noise.as<double>() : nullptr,
                           resample_mode == SMCB200_SYSTEMATIC ? un : nullptr,
                           resample_mode == SMCB200_STRATIFIED ? un : nullptr,
                           (resample_mode == SMCB200_MULTINOMIAL || resample_mode == SMCB200_RESIDUAL) ? un : nullptr);
        h->mode = resample_mode;
        h->kev_used = 0;
        if (h->profile) {
            while (h->kev.size() < (size_t)(3 * h->T + 2)) { cudaEvent_t e; SMCB_CUDA(cudaEventCreate(&e)); h->kev.push_back(e); }
            smcb200_pf* hp = h;
            s.SetAfterLaunchHook([hp]() { if (hp->kev_used < hp->kev.size()) cudaEventRecord(hp->kev[hp->kev_used++], hp->stream); });
            SMCB_CUDA(cudaEventRecord(h->kev[h->kev_used++], h->stream));
        } else {
            s.SetAfterLaunchHook(nullptr);
        }
        SMCB_CUDA(cudaEventRecord(h->ev0, h->stream));
        if (s.SmallEligible()) {   // small populations: the whole filter is one cooperative launch (csrc/small.cuh)
            s.RunFilterSmall(h->T, true);
            h->launches = 1;
        } else {
            s.Initialise();
            for (long t = 1; t < h->T; ++t) s.Iterate();
            s.ObserveFinal();
            // per step: k_step + the three resampling passes (mass / count / map), or + the four kernels of the draw-based schemes
            h->launches = h->T * ((resample_mode == SMCB200_SYSTEMATIC || resample_mode == SMCB200_STRATIFIED) ? 4 : 5) + 1;
        }
        SMCB_CUDA(cudaEventRecord(h->ev1, h->stream));
        h->ran = true;
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_sync(smcb200_pf* h) {
    return guarded([&]() -> int {
        if (!h) throw std::invalid_argument("null handle");
        SMCB_CUDA(cudaStreamSynchronize(h->stream));
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_elapsed_ms(smcb200_pf* h, float* ms) {
    return guarded([&]() -> int {
        if (!h || !ms) throw std::invalid_argument("null argument");
        if (!h->ran) throw std::runtime_error("smcb200_pfNonlinBS_elapsed_ms: run first");
        SMCB_CUDA(cudaEventSynchronize(h->ev1));
        SMCB_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_read(smcb200_pf* h, double* mean_host, double* sd_host, double* lognc_host, double* ess_host,
                            int* resampled_host) {
    return guarded([&]() -> int {
        if (!h) throw std::invalid_argument("null handle");
        if (!h->ran) throw std::runtime_error("smcb200_pfNonlinBS_read: run first");
        const long T = h->T;
        h->obs.resize(2 * (size_t)T);
        h->s->ReadTraces(T, lognc_host, ess_host, resampled_host, (mean_host || sd_host) ? h->obs.data() : nullptr);
        for (long t = 0; t < T; ++t) {
            if (mean_host) mean_host[t] = h->obs[2 * t];
            if (sd_host) sd_host[t] = std::sqrt(h->obs[2 * t + 1]);  // src/pfnonlinbs.cpp:72
        }
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_read_ancestors(smcb200_pf* h, unsigned int* idx_host) {
    return guarded([&]() -> int {
        if (!h || !idx_host) throw std::invalid_argument("null argument");
        if (!h->ran) throw std::runtime_error("smcb200_pfNonlinBS_read_ancestors: run first");
        h->s->ReadAncestors(idx_host);
        return SMCB200_OK;
    });
}

long smcb200_pfNonlinBS_launches(smcb200_pf* h) { return h ? h->launches : 0; }

#ifdef SMCB_PROFILE_LOOKBACK
void smcb200_dbg_lookback_stats(unsigned long long* out8, int reset) {
    cudaMemcpyFromSymbol(out8, smcb::g_lb_stats, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(smcb::g_lb_stats, z, sizeof z); }
}
#endif

int smcb200_pfNonlinBS_set_profile(smcb200_pf* h, int on) {
    if (!h) return fail(SMCB200_ERR_INVALID, "null handle");
    h->profile = on != 0;
    return SMCB200_OK;
}

int smcb200_pfNonlinBS_kernel_ms(smcb200_pf* h, float* move_ms, long* move_launches, float* scan_ms, long* scan_launches) {
    return guarded([&]() -> int {
        if (!h || !move_ms || !move_launches || !scan_ms || !scan_launches) throw std::invalid_argument("null argument");
        if (!h->ran || !h->profile || h->kev_used < 3) throw std::runtime_error("smcb200_pfNonlinBS_kernel_ms: run with profiling enabled first");
        if (h->mode != SMCB200_SYSTEMATIC && h->mode != SMCB200_STRATIFIED) throw std::runtime_error("smcb200_pfNonlinBS_kernel_ms: per-kernel split is defined for SYSTEMATIC / STRATIFIED runs");
        SMCB_CUDA(cudaEventSynchronize(h->kev[h->kev_used - 1]));
        // event order: init, passes, (move, passes) x (T-1), observe; interval i lies between events i and i+1 ("passes" = the
        // three resampling kernels of a step, timed together)
        double mv = 0, sc = 0; long nm = 0, ns = 0;
        const size_t nint = h->kev_used - 1;
        for (size_t i = 0; i < nint; ++i) {
            float ms; SMCB_CUDA(cudaEventElapsedTime(&ms, h->kev[i], h->kev[i + 1]));
            if (i == 0 || i == nint - 1) continue;           // init and the final moment pass are not counted
            if (i % 2 == 1) { sc += ms; ++ns; } else { mv += ms; ++nm; }
        }
        if (h->s->island() && getenv("SMCB200_PROF_PRINT"))
            fprintf(stderr, "[smcb200] island rank %d: k_step %.1f us, resampling passes %.1f us per step\n", h->s->rank(),
                    1e3 * mv / (nm ? nm : 1), 1e3 * sc / (ns ? ns : 1));
        *move_ms = (float)mv; *move_launches = nm; *scan_ms = (float)sc; *scan_launches = ns;
        return SMCB200_OK;
    });
}

int smcb200_pfNonlinBS_destroy(smcb200_pf* h) {
    if (!h) return SMCB200_OK;
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (cudaEvent_t e : h->kev) cudaEventDestroy(e);
    delete h;
    return SMCB200_OK;
}

int smcb200_pfNonlinBS(const double* data_host, long T, long particles, int resample_mode, double threshold,
                       uint64_t seed, const double* noise_host, const double* uniforms_host, double* mean_host,
                       double* sd_host, double* lognc_host, double* ess_host, int* resampled_host) {
    // One engine is kept per calling thread and reused while the shape matches (avoids cudaMalloc per call).
    thread_local smcb200_pf* cache = nullptr;
    if (!data_host || T < 1 || particles < 1) return fail(SMCB200_ERR_INVALID, "smcb200_pfNonlinBS: data, T >= 1 and particles >= 1 required");
    int rc;
    if (cache && (cache->n != particles || cache->tcap < T)) { smcb200_pfNonlinBS_destroy(cache); cache = nullptr; }
    if (!cache) { rc = smcb200_pfNonlinBS_create(particles, T, seed, nullptr, &cache); if (rc) return rc; }
    cache->seed = seed;
    if ((rc = smcb200_pfNonlinBS_set_data(cache, data_host, T))) return rc;
    long nu = resample_mode == SMCB200_STRATIFIED ? T * (particles + 1) : (resample_mode == SMCB200_SYSTEMATIC ? T : T * particles);
    if ((rc = smcb200_pfNonlinBS_set_noise(cache, noise_host, uniforms_host, nu))) return rc;
    if ((rc = smcb200_pfNonlinBS_run(cache, resample_mode, threshold))) return rc;
    return smcb200_pfNonlinBS_read(cache, mean_host, sd_host, lognc_host, ess_host, resampled_host);
}

}  // extern "C"
