"""ctypes access to the oracle (oracle/liboracle.so) and, when built, the compiled reference
(oracle/_ref/libsmcref.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
DP = ctypes.POINTER(ctypes.c_double)
IP = ctypes.POINTER(ctypes.c_int)
UP = ctypes.POINTER(ctypes.c_uint)
L = ctypes.c_long
D = ctypes.c_double

MULTINOMIAL, RESIDUAL, STRATIFIED, SYSTEMATIC = 0, 1, 2, 3


def dp(a):
    return None if a is None else a.ctypes.data_as(DP)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        o = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        o.smco_lse.restype = D
        o.smco_ess.restype = D
        o.smco_integrate.restype = D
        o.smco_time_pfnonlinbs.restype = D
        for f in ("smco_counts_systematic", "smco_counts_stratified", "smco_counts_multinomial_iid", "smco_residual_floor"):
            getattr(o, f).restype = L
        _oracle = o
    return _oracle


def ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsmcref.so"))


def ref():
    global _ref
    if _ref is None:
        r = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsmcref.so"))
        r.smcref_lse.restype = D
        r.smcref_ess.restype = D
        r.smcref_time_pfnonlinbs.restype = D
        for f in ("smcref_pending_uniforms", "smcref_pending_normals", "smcref_consumed_uniforms", "smcref_consumed_normals"):
            getattr(r, f).restype = L
        _ref = r
    return _ref


# ---- oracle wrappers -----------------------------------------------------------------------------------
def o_lse(lw):
    lw = f64(lw)
    return oracle().smco_lse(dp(lw), L(lw.size))


def o_ess(lw):
    lw = f64(lw)
    return oracle().smco_ess(dp(lw), L(lw.size))


def o_quantise(w):
    w = f64(w)
    out = np.empty_like(w)
    oracle().smco_quantise_weights(dp(w), L(w.size), dp(out))
    return out


def o_normalised(lw):
    lw = f64(lw)
    out = np.empty_like(lw)
    oracle().smco_normalised_weights(dp(lw), L(lw.size), dp(out))
    return out


def o_resample_w(mode, w, U=None, counts=None):
    """(idx, counts) from normalised weights (oracle restatement of sampler.h:786-857)."""
    n = (w if w is not None else counts).size
    idx = np.empty(n, dtype=np.uint32)
    cout = np.empty(n, dtype=np.int32)
    Uu = None if U is None else f64(np.atleast_1d(U))
    c = None if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
    rc = oracle().smco_resample_from_weights(mode, dp(None if w is None else f64(w)), L(n), dp(Uu), L(0 if Uu is None else Uu.size),
                                             None if c is None else c.ctypes.data_as(IP), idx.ctypes.data_as(UP),
                                             cout.ctypes.data_as(IP))
    assert rc == 0
    return idx, cout


def o_resample_lw(mode, lw, U=None, counts=None):
    lw = f64(lw)
    n = lw.size
    idx = np.empty(n, dtype=np.uint32)
    cout = np.empty(n, dtype=np.int32)
    Uu = None if U is None else f64(np.atleast_1d(U))
    c = None if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
    rc = oracle().smco_resample(mode, dp(lw), L(n), dp(Uu), L(0 if Uu is None else Uu.size),
                                None if c is None else c.ctypes.data_as(IP), idx.ctypes.data_as(UP), cout.ctypes.data_as(IP))
    assert rc == 0
    return idx, cout


def sim_nonlin(T, seed, var_init=10.0, var_evol=10.0, var_obs=1.0, cos_offset=-1.0):
    """Restated R/simNonlin.R (oracle/smc_oracle.c:smco_sim_nonlin) on numpy normals."""
    rng = np.random.default_rng(seed)
    zs, zo = rng.standard_normal(T), rng.standard_normal(T)
    x, y = np.empty(T), np.empty(T)
    oracle().smco_sim_nonlin(L(T), dp(zs), dp(zo), D(var_init), D(var_evol), D(var_obs), D(cos_offset), dp(x), dp(y))
    return x, y


def o_pfnonlinbs(y, N, mode, threshold, z=None, U=None, seed=0, want_idx=False, by_step=False, grid=False):
    oracle().smco_set_uniforms_by_step(1 if by_step else 0)
    oracle().smco_set_grid_weights(1 if grid else 0)
    y = f64(y)
    T = y.size
    mean, sd, nc, ess = np.empty(T), np.empty(T), np.empty(T), np.empty(T)
    rs = np.empty(T, dtype=np.int32)
    idx = np.empty((T, N), dtype=np.uint32) if want_idx else None
    z = None if z is None else f64(z)
    U = None if U is None else f64(U)
    rc = oracle().smco_pfnonlinbs(dp(y), L(T), L(N), mode, D(threshold), dp(z), dp(U), ctypes.c_ulonglong(seed), dp(mean), dp(sd),
                                  dp(nc), dp(ess), rs.ctypes.data_as(IP), None if idx is None else idx.ctypes.data_as(UP))
    assert rc == 0
    out = {"mean": mean, "sd": sd, "logNC": nc, "ESS": ess, "resampled": rs}
    if want_idx:
        out["idx"] = idx
    return out


# ---- compiled-reference wrappers ---------------------------------------------------------------------------
def r_clear():
    ref().smcref_clear_queues()


def r_push(normals=None, uniforms=None, multinomial=None):
    if normals is not None:
        z = f64(normals)
        ref().smcref_push_normals(dp(z), L(z.size))
    if uniforms is not None:
        u = f64(np.atleast_1d(uniforms))
        ref().smcref_push_uniforms(dp(u), L(u.size))
    if multinomial is not None:
        c = np.ascontiguousarray(multinomial, dtype=np.int32)
        ref().smcref_push_multinomial(c.ctypes.data_as(IP), L(c.size))


def r_lse(lw):
    lw = f64(lw)
    return ref().smcref_lse(dp(lw), L(lw.size))


def r_ess(lw):
    lw = f64(lw)
    return ref().smcref_ess(dp(lw), L(lw.size))


def r_resample(mode, lw, uniforms=None, multinomial=None, weights_override=None):
    """uRSIndices from the compiled reference's sampler::Resample."""
    lw = f64(lw)
    n = lw.size
    r_clear()
    r_push(uniforms=uniforms, multinomial=multinomial)
    keep = None
    if weights_override is not None:
        keep = f64(weights_override)
        ref().smcref_override_exp(dp(keep), L(n), 1)
    idx = np.empty(n, dtype=np.uint32)
    val = np.empty(n)
    ref().smcref_resample(mode, dp(lw), L(n), idx.ctypes.data_as(UP), dp(val))
    return idx, val


def r_pfnonlinbs(y, N, mode, threshold, z, U):
    y = f64(y)
    T = y.size
    r_clear()
    r_push(normals=z, uniforms=U)
    mean, sd, nc, ess = np.empty(T), np.empty(T), np.empty(T), np.empty(T)
    rs = np.empty(T, dtype=np.int32)
    ref().smcref_pfnonlinbs(dp(y), L(T), L(N), mode, D(threshold), dp(mean), dp(sd), dp(nc), dp(ess), rs.ctypes.data_as(IP))
    return {"mean": mean, "sd": sd, "logNC": nc, "ESS": ess, "resampled": rs}


def random_weights(rng, n, spread=3.0, grid=True):
    """Normalised weights from Gaussian log-weights; optionally on the engine's 2^-52 grid."""
    lw = spread * rng.standard_normal(n)
    w = np.exp(lw - lw.max())
    w /= w.sum()
    return o_quantise(w) if grid else w


# ---- pfLineartBS ------------------------------------------------------------------------------------------
def sim_lineart(T, seed):
    """Restated R/simLineart.R:1-15: two Gaussian random walks observed with Student-t(4) noise.  Returns [T x 2]."""
    rng = np.random.default_rng(seed)
    sx, sy = np.cumsum(rng.standard_normal(T)), np.cumsum(rng.standard_normal(T))
    return np.column_stack([sx + rng.standard_t(4, T), sy + rng.standard_t(4, T)])


def _colmajor(data):
    return np.ascontiguousarray(np.asarray(data, dtype=np.float64).T.reshape(-1))


def o_pflineart(data, N, mode, threshold, z=None, U=None, seed=0, by_step=False):
    oracle().smco_set_uniforms_by_step(1 if by_step else 0)
    d = _colmajor(data)
    T = d.size // 2
    out = {k: np.empty(T) for k in ("Xm", "Xv", "Ym", "Yv", "logNC", "ESS")}
    rs = np.empty(T, dtype=np.int32)
    z = None if z is None else f64(z)
    U = None if U is None else f64(U)
    rc = oracle().smco_pflineart(dp(d), L(T), L(N), mode, D(threshold), dp(z), dp(U), ctypes.c_ulonglong(seed), dp(out["Xm"]),
                                 dp(out["Xv"]), dp(out["Ym"]), dp(out["Yv"]), dp(out["logNC"]), dp(out["ESS"]), rs.ctypes.data_as(IP))
    assert rc == 0
    out["resampled"] = rs
    return out


def r_pflineart(data, N, mode, threshold, z, U):
    d = _colmajor(data)
    T = d.size // 2
    r_clear()
    r_push(normals=z, uniforms=U)
    out = {k: np.empty(T) for k in ("Xm", "Xv", "Ym", "Yv", "logNC", "ESS")}
    rs = np.empty(T, dtype=np.int32)
    ref().smcref_pflineart(dp(d), L(T), L(N), mode, D(threshold), dp(out["Xm"]), dp(out["Xv"]), dp(out["Ym"]), dp(out["Yv"]),
                           dp(out["logNC"]), dp(out["ESS"]), rs.ctypes.data_as(IP))
    out["resampled"] = rs
    return out


# ---- LinRegLA / LinRegLA_adapt --------------------------------------------------------------------------------
def radiata():
    """The 42-observation radiata pine data (y, x1, x2): fixture extracted from the reference's bundled data file by
    tests/golden/make_golden.py (reference inst/rawdata/radiata-data.csv)."""
    return np.load(os.path.join(GOLDEN, "radiata.npz"))["data"]


def o_linreg_la(Data, N, adaptive=True, temps=None, resampTol=0.5, tempTol=0.9, z=None, U=None, seed=0, max_temps=512):
    d = np.asarray(Data, dtype=np.float64)
    flat = np.ascontiguousarray(d.T.reshape(-1)); n_obs = d.shape[0]
    th = np.empty(max_temps * 3 * N); ll = np.empty(max_temps * N); lp = np.empty(max_temps * N); w = np.empty(max_temps * N)
    ess, tt, reps, acc = (np.empty(max_temps) for _ in range(4))
    nc = np.empty(4)
    z = None if z is None else f64(z); U = None if U is None else f64(U)
    tin = None if temps is None else f64(temps)
    o = oracle(); o.smco_linreg_la.restype = L
    Lr = o.smco_linreg_la(dp(flat), L(n_obs), L(N), 1 if adaptive else 0, dp(tin), L(0 if tin is None else tin.size), D(resampTol), D(tempTol),
                          dp(z), dp(U), ctypes.c_ulonglong(seed), L(max_temps), dp(th), dp(ll), dp(lp), dp(w), dp(ess), dp(tt), dp(reps), dp(nc), dp(acc))
    assert Lr > 0
    return {"theta": th[:Lr * 3 * N].reshape(Lr, 3, N).transpose(2, 1, 0), "loglike": ll[:Lr * N].reshape(Lr, N).T,
            "logprior": lp[:Lr * N].reshape(Lr, N).T, "Weights": w[:Lr * N].reshape(Lr, N).T, "ESS": ess[:Lr].copy(), "Temps": tt[:Lr].copy(),
            "mcmcRepeats": reps[:Lr].copy(), "logNC_standard": nc[0], "logNC_ps_rect": nc[1], "logNC_ps_trap": nc[2], "logNC_ps_trap2": nc[3],
            "acceptProb": acc[:Lr].copy()}


# ---- nonLinPMMH -----------------------------------------------------------------------------------------------------
def o_pmmh_filter(y, N, sigv, sigw, mode, threshold, z=None, U=None, seed=0):
    y = f64(y)
    o = oracle(); o.smco_pmmh_filter.restype = D
    return o.smco_pmmh_filter(dp(y), L(y.size), L(N), D(sigv), D(sigw), mode, D(threshold), dp(None if z is None else f64(z)),
                              dp(None if U is None else f64(U)), ctypes.c_ulonglong(seed))


def o_pmmh(y, N, M, mode, threshold, innov, filter_noise=None, filter_unif=None, seed=0):
    y = f64(y)
    out = [np.empty(M + 1) for _ in range(5)]
    iv, iw, uu = (f64(a) for a in innov)
    rc = oracle().smco_nonlin_pmmh(dp(y), L(y.size), L(N), L(M), mode, D(threshold), dp(iv), dp(iw), dp(uu),
                                   dp(None if filter_noise is None else f64(filter_noise)), dp(None if filter_unif is None else f64(filter_unif)),
                                   ctypes.c_ulonglong(seed), *[dp(o) for o in out])
    assert rc == 0
    return dict(zip(("samples_sigv", "samples_sigw", "loglike", "logprior", "MH_acceptance_rate"), out))


def r_pmmh(y, N, M, seed):
    y = f64(y)
    r_clear(); ref().smcref_seed(ctypes.c_uint64(seed))
    out = [np.empty(M + 1) for _ in range(5)]
    rc = ref().smcref_nonlin_pmmh(dp(y), L(y.size), L(N), L(M), *[dp(o) for o in out])
    assert rc == 0
    return dict(zip(("samples_sigv", "samples_sigw", "loglike", "logprior", "MH_acceptance_rate"), out))


def o_linreg(Data, N, z=None, U=None, Udraw=None, seed=0):
    d = np.asarray(Data, dtype=np.float64)
    flat = np.ascontiguousarray(d.T.reshape(-1)); n_obs = d.shape[0]
    th, w, nc = np.empty(3 * N), np.empty(N), np.empty(1)
    rc = oracle().smco_linreg(dp(flat), L(n_obs), L(N), dp(None if z is None else f64(z)), dp(None if U is None else f64(U)),
                              dp(None if Udraw is None else f64(Udraw)), ctypes.c_ulonglong(seed), dp(th), dp(w), dp(nc))
    assert rc == 0
    return {"theta": th.reshape(3, N).T, "weights": w, "logNC": float(nc[0])}


def sim_lg(T, seed, obs_sd=1.0):
    """Random walk + noise series for the block filter (R/blockpfGaussianOpt.R simGaussian: x_t = x_{t-1} + e, y = x + obs_sd e')."""
    rng = np.random.default_rng(seed)
    return np.cumsum(rng.standard_normal(T)) + obs_sd * rng.standard_normal(T)


def blockpf_draw_count(t, lag):
    return 1 if t < 2 else min(t, lag)


def o_blockpf(y, N, lag, z=None, U=None, seed=0):
    """Oracle port of blockpfGaussianOpt_impl.  z [T][lag][N] addressed normals, U [T] uniforms by step."""
    y = f64(y); T = y.size
    w, v, nc, ess = np.zeros(N), np.zeros(N * T), np.zeros(1), np.zeros(T)
    rs, anc = np.zeros(T, dtype=np.int32), np.zeros((T, N), dtype=np.uint32)
    zz = None if z is None else f64(z)
    uu = None if U is None else f64(U)
    oracle().smco_blockpf(dp(y), L(T), L(N), L(lag), dp(zz) if zz is not None else None, dp(uu) if uu is not None else None,
                          ctypes.c_ulonglong(seed), dp(w), dp(v), dp(nc), dp(ess), rs.ctypes.data_as(IP), anc.ctypes.data_as(UP))
    return {"weight": w, "values": v.reshape(T, N).T.copy(), "logNC": float(nc[0]), "ESS": ess, "resampled": rs, "ancestors": anc}


def r_blockpf(y, N, lag, z, U_consumed):
    """The compiled reference on the same draws: normals re-ordered into its consumption order (particle-major inside a step),
    uniforms only for the steps that resample."""
    y = f64(y); T = y.size
    z = np.asarray(z).reshape(T, lag, N)
    q = [z[t, :blockpf_draw_count(t, lag), :].T.reshape(-1) for t in range(T)]
    r_clear(); r_push(normals=np.concatenate(q), uniforms=U_consumed)
    w, v, nc = np.zeros(N), np.zeros(N * T), np.zeros(1)
    rc = ref().smcref_blockpf(dp(y), L(T), L(N), L(lag), dp(w), dp(v), dp(nc))
    assert rc == 0
    assert ref().smcref_pending_normals() == 0 and ref().smcref_pending_uniforms() == 0
    return {"weight": w, "values": v.reshape(T, N).T.copy(), "logNC": float(nc[0])}


# ---- conditional SMC on the linear-Gaussian model (compareNCestimates) -------------------------------------------------
LP = ctypes.POINTER(ctypes.c_long)


def sim_lgss(T, seed, phi=0.9, var_evol=1.0, var_obs=1.0, x0=0.0):
    rng = np.random.default_rng(seed)
    x = np.zeros(T)
    prev = x0
    for t in range(T):
        x[t] = phi * prev + np.sqrt(var_evol) * rng.standard_normal()
        prev = x[t]
    return x, x + np.sqrt(var_obs) * rng.standard_normal(T)


def o_csmc_run(y, N, mode, par, ref_traj, refidx, z=None, U=None, addressed=True, seed=0):
    """One conditional run of the oracle port; `refidx` (uint32 [T]) is the persistent in/out reference-index state."""
    y = f64(y); T = y.size
    nc = np.zeros(1)
    anc, val, lw = np.zeros((T, N), dtype=np.uint32), np.zeros((T, N)), np.zeros((T, N))
    rs, kt, ns = np.zeros(T, dtype=np.int32), np.zeros(T, dtype=np.int64), np.zeros(T, dtype=np.int64)
    zz = None if z is None else f64(z)
    uu = None if U is None else f64(U)
    rt = f64(ref_traj)
    oracle().smco_csmc_lgss_run(dp(y), L(T), L(N), mode, D(par[0]), D(par[1]), D(par[2]), D(par[3]), dp(rt), refidx.ctypes.data_as(UP),
                                dp(zz) if zz is not None else None, dp(uu) if uu is not None else None, 1 if addressed else 0,
                                ctypes.c_ulonglong(seed), dp(nc), anc.ctypes.data_as(UP), dp(val), dp(lw), rs.ctypes.data_as(IP),
                                kt.ctypes.data_as(LP), ns.ctypes.data_as(LP))
    return {"logNC": float(nc[0]), "ancestors": anc, "values": val, "logw": lw, "resampled": rs, "Kt": kt, "nstoch": ns,
            "refidx": refidx.copy()}


def csmc_uniform_queue(mode, U, run, N):
    """Addressed uniforms U [T][N+2] -> the sequence the reference consumes, given the port's trace of the same run."""
    U = np.asarray(U).reshape(-1, N + 2)
    q = []
    for t in range(U.shape[0]):
        if not run["resampled"][t]:
            continue
        kt = int(run["Kt"][t])
        if mode == MULTINOMIAL:
            q.extend(U[t, 2 + 1:2 + N])
        elif mode == STRATIFIED:
            q.append(U[t, 0]); q.extend(U[t, 2 + i] for i in range(N) if i != kt)
        elif mode == SYSTEMATIC:
            q.append(U[t, 0]); q.append(U[t, 1])
        else:
            q.append(U[t, 0]); q.extend(U[t, 2 + k] for k in range(int(run["nstoch"][t]), N) if k != kt)
    return np.array(q, dtype=np.float64)


def r_csmc_open(y, N, par):
    y = f64(y)
    ref().smcref_csmc_open(dp(y), L(y.size), L(N), D(par[0]), D(par[1]), D(par[2]), D(par[3]))


def r_csmc_run(mode, ref_traj, T, N, z, uq):
    r_clear(); r_push(normals=f64(z).reshape(-1), uniforms=uq if uq.size else None)
    nc = np.zeros(1)
    ri, rs = np.zeros(T, dtype=np.uint32), np.zeros(T, dtype=np.int32)
    anc, lines = np.zeros((T, N), dtype=np.uint32), np.zeros((T, N), dtype=np.uint32)
    val, lw, lv = np.zeros((T, N)), np.zeros((T, N)), np.zeros((T, N))
    rt = f64(ref_traj)
    rc = ref().smcref_csmc_run(mode, dp(rt), dp(nc), ri.ctypes.data_as(UP), anc.ctypes.data_as(UP), dp(val), dp(lw), rs.ctypes.data_as(IP),
                               lines.ctypes.data_as(UP), dp(lv))
    assert rc == 0
    assert ref().smcref_pending_normals() == 0 and ref().smcref_pending_uniforms() == 0
    return {"logNC": float(nc[0]), "ancestors": anc, "values": val, "logw": lw, "resampled": rs, "refidx": ri, "lines": lines,
            "line_values": lv}


def o_ancestor_lines(anc, values=None):
    anc = np.ascontiguousarray(anc, dtype=np.uint32)
    T, N = anc.shape
    lines, lv = np.zeros((T, N), dtype=np.uint32), np.zeros((T, N))
    v = None if values is None else f64(values)
    oracle().smco_ancestor_lines(anc.ctypes.data_as(UP), dp(v) if v is not None else None, L(T), L(N), lines.ctypes.data_as(UP), dp(lv))
    return lines, lv


def o_bpf_lgss(y, N, mode, par, z=None, U=None, seed=0, want_flags=False):
    y = f64(y)
    o = oracle(); o.smco_bpf_lgss.restype = D
    zz = None if z is None else f64(z)
    uu = None if U is None else f64(U)
    rs = np.zeros(y.size, dtype=np.int32)
    v = o.smco_bpf_lgss(dp(y), L(y.size), L(N), mode, D(par[0]), D(par[1]), D(par[2]), D(par[3]), dp(zz) if zz is not None else None,
                        dp(uu) if uu is not None else None, ctypes.c_ulonglong(seed), rs.ctypes.data_as(IP))
    return (v, rs) if want_flags else v


def r_bpf_lgss(mode, z, uq):
    """Plain bootstrap filter of the compiled reference on the model opened by r_csmc_open."""
    r_clear(); r_push(normals=f64(z).reshape(-1), uniforms=uq if uq.size else None)
    nc = np.zeros(1)
    assert ref().smcref_bpf_lgss_run(mode, dp(nc)) == 0
    assert ref().smcref_pending_normals() == 0 and ref().smcref_pending_uniforms() == 0
    return float(nc[0])
