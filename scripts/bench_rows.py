"""Measurement of the widened rows on one B200, with the reference's own CPU implementation
(oracle/_ref, one host core) timed beside each on a bounded sample.  Not the bench.py contract: one JSON object per row,
written to the path given as argv[1] (default: stdout only).

Each GPU figure is the public C-ABI call with host buffers (host<->device copies inside the timed region), best of 3 after
one warm-up call.  Algorithmic bytes follow DESIGN.md: 16 d + 40 per particle-step for the filters."""
import ctypes
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import helpers as H  # noqa: E402
import rcppsmc_b200 as smc  # noqa: E402

PEAK = 6458.4
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


def best_of(f, reps=3):
    f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return min(ts)


def cpu_time(f):
    t0 = time.perf_counter(); f(); return time.perf_counter() - t0


rows = []
have_ref = H.ref_available()
R = H.ref() if have_ref else None
dp, L, D = H.dp, H.L, H.D

# ---- pfLineartBS: 2^20 particles x 250 steps, d = 4 (BASELINE configs[1]) --------------------------------------------
T, N = 250, 1 << 20
data = H.sim_lineart(T, 1)
for mode, name in ((smc.STRATIFIED, "stratified"), (smc.MULTINOMIAL, "multinomial"), (smc.RESIDUAL, "residual (reference default)")):
    t = best_of(lambda: smc.pfLineartBS(data, N, resample_mode=mode, threshold=0.5, seed=3))
    row = {"row": "pfLineartBS", "config": f"2^20 particles x 250 steps, d=4, {name}, ESS < N/2", "gpu_seconds": t,
           "particle_steps_per_s": N * T / t, "alg_bytes_per_particle_step": 104, "roofline_frac": N * T * 104 / t / 1e9 / PEAK}
    if have_ref and mode == smc.STRATIFIED:
        Nc = 1 << 14
        d = H._colmajor(data)
        out = [np.empty(T) for _ in range(6)]; rs = np.empty(T, dtype=np.int32)
        H.r_clear(); R.smcref_seed(ctypes.c_uint64(1))
        tc = cpu_time(lambda: R.smcref_pflineart(dp(d), L(T), L(Nc), mode, D(0.5), *[dp(o) for o in out], rs.ctypes.data_as(H.IP)))
        row["cpu_reference"] = {"particle_steps_per_s": Nc * T / tc, "cores": 1, "sample": f"2^14 particles x 250 steps, {tc:.1f} s"}
    rows.append(row)

# ---- LinRegLA_adapt: 2^18 particles, adaptive temperatures + MCMC (BASELINE configs[2]) --------------------------------
Data = H.radiata()[:, :2]
N = 1 << 18
res = {}
t = best_of(lambda: res.update(smc.LinRegLA_adapt(Data, N, 0.5, 0.9, seed=5)), reps=2)
Lg = res["Temps"].size
moves = float(np.sum(res["mcmcRepeats"])) * N
row = {"row": "LinRegLA_adapt", "config": "radiata (42 obs), 2^18 particles, resampTol 0.5, tempTol 0.9", "gpu_seconds": t, "temperatures": int(Lg),
       "mcmc_moves_per_s": moves / t, "logNC_standard": float(res["logNC_standard"]),
       "note": "compute-bound on the 42-term likelihood (2 per MH move); bytes are negligible next to it"}
if have_ref:
    Nc, maxT = 1 << 12, 512
    flat = np.ascontiguousarray(Data.T.reshape(-1))
    th, ll, lp, w = np.empty(maxT * 3 * Nc), np.empty(maxT * Nc), np.empty(maxT * Nc), np.empty(maxT * Nc)
    ess, tt, reps = np.empty(maxT), np.empty(maxT), np.empty(maxT)
    nc = np.empty(4)
    R.smcref_linreg_la_adapt.restype = L
    H.r_clear(); R.smcref_seed(ctypes.c_uint64(2))
    box = {}
    tc = cpu_time(lambda: box.update(L=R.smcref_linreg_la_adapt(dp(flat), L(42), L(Nc), D(0.5), D(0.9), L(maxT), dp(th), dp(ll), dp(lp), dp(w),
                                                                 dp(ess), dp(tt), dp(reps), dp(nc))))
    row["cpu_reference"] = {"mcmc_moves_per_s": float(np.sum(reps[:box["L"]])) * Nc / tc, "cores": 1, "sample": f"2^12 particles, {box['L']} temperatures, {tc:.1f} s"}
rows.append(row)

# ---- nonLinPMMH: 2^16-particle filters over 500 observations (BASELINE configs[3]) -----------------------------------------
T, N, M = 500, 1 << 16, 200
y = H.sim_nonlin(T, 7, var_init=5.0, cos_offset=0.0)[1]
t = best_of(lambda: smc.nonLinPMMH(y, N, M, seed=9), reps=1)
row = {"row": "nonLinPMMH", "config": "2^16-particle filters x 500 observations, 200 MH iterations (201 filters), multinomial / 0.5, one chain", "gpu_seconds": t,
       "filters_per_s": (M + 1) / t, "particle_steps_per_s": (M + 1) * N * T / t,
       "note": "each filter is ONE launch of k_filter_small (csrc/small.cuh); chains shard one per GPU"}
if have_ref:
    tc = cpu_time(lambda: H.r_pmmh(y, 1 << 13, 2, 3))
    row["cpu_reference"] = {"particle_steps_per_s": 3 * (1 << 13) * T / tc, "cores": 1, "sample": f"2^13-particle filters x 500 obs, 3 filters, {tc:.1f} s"}
rows.append(row)

# ---- blockpfGaussianOpt: 2^20 paths x 100 steps, lag 5 ---------------------------------------------------------------------
T, N, lag = 100, 1 << 20, 5
y = H.sim_lg(T, 5, 10.0)
out = {}
t = best_of(lambda: out.update(smc.blockpfGaussianOpt(y, N, lag, seed=3, diagnostics=True)), reps=2)
alg = 8 * (lag + 1) + 16 + 28 + 8
row = {"row": "blockpfGaussianOpt", "config": "2^20 path-valued particles x 100 steps, lag 5, systematic, ESS < N/2",
       "gpu_seconds": t, "particle_steps_per_s": N * T / t, "resampled_steps": int(out["resampled"].sum()),
       "alg_bytes_per_particle_step": alg, "roofline_frac": N * T * alg / t / 1e9 / PEAK,
       "note": "timed call returns the full [N x T] path matrix to the host (0.84 GB D2H inside the timed region)"}
if have_ref:
    Nc = 1 << 13
    w, v, nc1 = np.zeros(Nc), np.zeros(Nc * T), np.zeros(1)
    H.r_clear(); R.smcref_seed(ctypes.c_uint64(4))
    tc = cpu_time(lambda: R.smcref_blockpf(dp(H.f64(y)), L(T), L(Nc), L(lag), dp(w), dp(v), dp(nc1)))
    row["cpu_reference"] = {"particle_steps_per_s": Nc * T / tc, "cores": 1, "sample": f"2^13 particles x 100 steps, {tc:.1f} s"}
rows.append(row)
try:
    import torch
    buf = torch.empty((T, N), dtype=torch.float64, device="cuda")
    t = best_of(lambda: smc.blockpfGaussianOpt_device(y, N, lag, buf.data_ptr(), seed=3), reps=2)
    rows.append({"row": "blockpfGaussianOpt (paths left on the device)", "config": "2^20 path-valued particles x 100 steps, lag 5, systematic, ESS < N/2",
                 "gpu_seconds": t, "particle_steps_per_s": N * T / t, "alg_bytes_per_particle_step": alg, "roofline_frac": N * T * alg / t / 1e9 / PEAK,
                 "note": "smcb200_blockpfGaussianOpt_device: weights, logNC, ESS to the host, the [N x T] path matrix stays in HBM"})
    del buf
except ImportError:
    pass

# ---- compareNCestimates: 1000 particles, 100 simulations x (4 plain + 4 conditional) filters --------------------------------
T, N, S = 50, 1000, 100
x, y = H.sim_lgss(T, 4)
ref = x[:, None] + 0.3 * np.random.default_rng(1).standard_normal((T, S))
par = (0.9, 1.0, 0.0, 1.0)
for ind in (False, True):
    t = best_of(lambda: smc.compareNCestimates(y, N, S, par, ref, seed=7, independent_sims=ind), reps=2)
    row = {"row": "compareNCestimates", "config": f"1000 particles x 50 steps, simNum 100 (800 filters), independent_sims={ind}", "gpu_seconds": t,
           "filters_per_s": 8 * S / t}
    if have_ref and not ind:
        a, b = np.zeros(4 * S), np.zeros(4 * S)
        refc = np.ascontiguousarray(ref.T)
        f = R.smcref_compare_nc
        f.argtypes = [H.DP if hasattr(H, "DP") else ctypes.POINTER(ctypes.c_double), ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                      ctypes.c_double, ctypes.c_double] + [ctypes.POINTER(ctypes.c_double)] * 3
        H.r_clear(); R.smcref_seed(ctypes.c_uint64(6))
        tc = cpu_time(lambda: f(dp(H.f64(y)), T, N, S, 0.9, 1.0, 0.0, 1.0, dp(refc), dp(a), dp(b)))
        row["cpu_reference"] = {"filters_per_s": 8 * S / tc, "cores": 1, "sample": f"same configuration, {tc:.1f} s"}
    rows.append(row)

# ---- ancestral lines: [T][N] u32 ancestor rows -> lines -----------------------------------------------------------------------
T, N = 64, 1 << 22
rng = np.random.default_rng(2)
anc = np.sort(rng.integers(0, N, size=(T, N)), axis=1).astype(np.uint32)
t = best_of(lambda: smc.ancestor_lines(anc), reps=2)
rows.append({"row": "ancestor_lines (GetALineInd for every particle)", "config": "64 steps x 2^22 particles", "gpu_seconds": t,
             "lines_per_s": N / t, "note": "host-buffer call: 1 GiB H2D + 1 GiB D2H dominate; kernel traffic 8 B per particle-step"})

for r in rows:
    print(json.dumps(r))
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
