"""ctypes binding of include/smcb200.h.  Fails loudly when the CUDA extension is absent."""
import ctypes
import os

import numpy as np

from . import _build

MULTINOMIAL, RESIDUAL, STRATIFIED, SYSTEMATIC = 0, 1, 2, 3  # ResampleType::Enum (reference sampler.h:45-51)

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_uint_p = ctypes.POINTER(ctypes.c_uint)
_c_u32_p = ctypes.POINTER(ctypes.c_uint32)


class SmcError(RuntimeError):
    """Raised for any non-zero status of the C ABI (the analogue of smc::exception / Rcpp::stop)."""


_lib = None


def lib_path():
    return os.environ.get("SMCB200_LIB", _build.LIB)  # override only for kernel experiments


def lib():
    """Loads libsmcb200.so (never builds implicitly on import; __graft_entry__.build() or _build.build() does)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise SmcError(
            f"{path} is missing: the CUDA extension is not built (run `python -m rcppsmc_b200._build`); "
            "rcppsmc_b200 has no CPU fallback")
    L = ctypes.CDLL(path)
    L.smcb200_last_error.restype = ctypes.c_char_p
    L.smcb200_version.restype = ctypes.c_int
    L.smcb200_device_count.restype = ctypes.c_int
    L.smcb200_philox4x32.argtypes = [_c_u32_p, _c_u32_p, _c_u32_p]
    L.smcb200_philox4x32_device.argtypes = [_c_u32_p, ctypes.c_long, _c_u32_p]
    L.smcb200_philox_normals.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_long, _c_double_p]
    L.smcb200_lse.argtypes = [_c_double_p, ctypes.c_long, _c_double_p]
    L.smcb200_fastmath_eval.argtypes = [ctypes.c_int, _c_double_p, ctypes.c_long, _c_double_p]
    L.smcb200_resample.argtypes = [ctypes.c_int, _c_double_p, ctypes.c_long, _c_double_p, ctypes.c_long, _c_int_p,
                                   _c_uint_p, _c_int_p]
    L.smcb200_pfNonlinBS.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_double,
                                     ctypes.c_uint64, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                     _c_double_p, _c_double_p, _c_int_p]
    L.smcb200_pfLineartBS.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_uint64,
                                      _c_double_p, _c_double_p, ctypes.c_void_p, ctypes.c_void_p, _c_double_p, _c_double_p,
                                      _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_int_p]
    L.smcb200_LinRegLA_adapt.restype = ctypes.c_long
    L.smcb200_LinRegLA_adapt.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_uint64,
                                         _c_double_p, ctypes.c_long, _c_double_p, ctypes.c_long, ctypes.c_long] + [_c_double_p] * 9
    L.smcb200_LinRegLA.argtypes = [_c_double_p, ctypes.c_long, _c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_uint64,
                                   _c_double_p, ctypes.c_long, _c_double_p, ctypes.c_long] + [_c_double_p] * 6
    L.smcb200_nonLinPMMH.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_uint64, ctypes.c_int, ctypes.c_double,
                                     _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p, ctypes.c_void_p, ctypes.c_long,
                                     ctypes.c_void_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p]
    L.smcb200_LinReg.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_uint64, _c_double_p, ctypes.c_long, _c_double_p, ctypes.c_long,
                                 _c_double_p, _c_double_p, _c_double_p, _c_double_p]
    L.smcb200_blockpfGaussianOpt_device.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_uint64, _c_double_p, ctypes.c_void_p,
                                                    _c_double_p, _c_double_p, _c_int_p]
    L.smcb200_blockpfGaussianOpt.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_uint64, _c_double_p, _c_double_p,
                                             _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_int_p, _c_uint_p]
    L.smcb200_conditionalBPF.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_long, ctypes.c_long, _c_int_p, _c_double_p, ctypes.c_uint64, _c_double_p, _c_double_p, _c_double_p,
                                         _c_uint_p, _c_uint_p, _c_double_p, _c_double_p, _c_int_p]
    L.smcb200_lgssBPF.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                  ctypes.c_int, ctypes.c_uint64, _c_double_p, _c_double_p, _c_double_p]
    L.smcb200_compareNCestimates.argtypes = [_c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                             ctypes.c_double, _c_double_p, ctypes.c_uint64, ctypes.c_int, _c_double_p, _c_double_p]
    L.smcb200_ancestor_lines.argtypes = [_c_uint_p, ctypes.c_long, ctypes.c_long, _c_double_p, _c_uint_p, _c_double_p]
    L.smcb200_pfNonlinBS_create.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.c_uint64, ctypes.c_void_p,
                                            ctypes.POINTER(ctypes.c_void_p)]
    L.smcb200_pfNonlinBS_create_island.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int,
                                                   ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.smcb200_pfNonlinBS_ipc_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.smcb200_pfNonlinBS_ipc_import.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.smcb200_pfNonlinBS_set_data.argtypes = [ctypes.c_void_p, _c_double_p, ctypes.c_long]
    L.smcb200_pfNonlinBS_set_noise.argtypes = [ctypes.c_void_p, _c_double_p, _c_double_p, ctypes.c_long]
    L.smcb200_pfNonlinBS_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
    L.smcb200_pfNonlinBS_sync.argtypes = [ctypes.c_void_p]
    L.smcb200_pfNonlinBS_elapsed_ms.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
    L.smcb200_pfNonlinBS_read.argtypes = [ctypes.c_void_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_int_p]
    L.smcb200_pfNonlinBS_read_ancestors.argtypes = [ctypes.c_void_p, _c_uint_p]
    L.smcb200_pfNonlinBS_launches.argtypes = [ctypes.c_void_p]
    L.smcb200_pfNonlinBS_launches.restype = ctypes.c_long
    L.smcb200_pfNonlinBS_destroy.argtypes = [ctypes.c_void_p]
    L.smcb200_pfNonlinBS_set_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.smcb200_pfNonlinBS_kernel_ms.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_long),
                                               ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_long)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise SmcError(f"smcb200 status {rc}: {lib().smcb200_last_error().decode()}")


def _dp(a):
    return None if a is None else a.ctypes.data_as(_c_double_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    return lib().smcb200_device_count()


def philox4x32(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    _check(lib().smcb200_philox4x32(c.ctypes.data_as(_c_u32_p), k.ctypes.data_as(_c_u32_p), out.ctypes.data_as(_c_u32_p)))
    return out


def philox4x32_device(ctr_key):
    """Philox4x32-10 evaluated by a kernel: rows of (ctr0..3, key0..1) -> rows of 4 output words."""
    ck = np.ascontiguousarray(ctr_key, dtype=np.uint32).reshape(-1, 6)
    out = np.zeros((ck.shape[0], 4), dtype=np.uint32)
    _check(lib().smcb200_philox4x32_device(ck.ctypes.data_as(_c_u32_p), ck.shape[0], out.ctypes.data_as(_c_u32_p)))
    return out


def philox_normals(seed, stream, step, n):
    out = np.empty(n, dtype=np.float64)
    _check(lib().smcb200_philox_normals(seed, stream, step, n, _dp(out)))
    return out


def fastmath_eval(fn, x):
    x = _f64(x)
    out = np.empty_like(x)
    _check(lib().smcb200_fastmath_eval(fn, _dp(x), x.size, _dp(out)))
    return out


def log_sum_exp(logw):
    """(log sum exp(logw), ESS, max) -- smc::stableLogSumWeights (population.h:46-50) + GetESS (sampler.h:413-417)."""
    lw = _f64(logw)
    out = np.zeros(3)
    _check(lib().smcb200_lse(_dp(lw), lw.size, _dp(out)))
    return float(out[0]), float(out[1]), float(out[2])


def resample(mode, weights=None, uniforms=None, counts=None, return_counts=False):
    """Ancestor indices (uRSIndices) for normalised ``weights`` -- sampler::Resample (sampler.h:786-857)."""
    w = _f64(weights)
    u = _f64(np.atleast_1d(uniforms)) if uniforms is not None else None
    c = None if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
    n = w.size if w is not None else c.size
    idx = np.empty(n, dtype=np.uint32)
    cout = np.empty(n, dtype=np.int32) if return_counts else None
    _check(lib().smcb200_resample(mode, _dp(w), n, _dp(u), 0 if u is None else u.size,
                                  None if c is None else c.ctypes.data_as(_c_int_p), idx.ctypes.data_as(_c_uint_p),
                                  None if cout is None else cout.ctypes.data_as(_c_int_p)))
    return (idx, cout) if return_counts else idx


def pfNonlinBS(data, particles, resample_mode=MULTINOMIAL, threshold=None, seed=1, noise=None, uniforms=None,
               full=False):
    """Mirror of ``pfNonlinBS_impl(data, part)`` (reference src/pfnonlinbs.cpp:52-79): returns dict(mean, sd).

    The two reference arguments come first; the trailing ones are the extensions listed in include/smcb200.h
    (the reference hard-codes MULTINOMIAL with threshold 1.01 N).  ``full=True`` adds logNC / ESS / resampled traces.
    """
    y = _f64(data)
    T = y.size
    thr = 1.01 * particles if threshold is None else threshold
    mean, sd = np.empty(T), np.empty(T)
    lognc, ess, res = np.empty(T), np.empty(T), np.empty(T, dtype=np.int32)
    z = _f64(noise)
    u = _f64(uniforms)
    _check(lib().smcb200_pfNonlinBS(_dp(y), T, particles, resample_mode, thr, seed, _dp(z), _dp(u), _dp(mean), _dp(sd),
                                    _dp(lognc), _dp(ess), res.ctypes.data_as(_c_int_p)))
    out = {"mean": mean, "sd": sd}
    if full:
        out.update(logNC=lognc, ESS=ess, resampled=res)
    return out


_STEP_CB = ctypes.CFUNCTYPE(None, _c_double_p, _c_double_p, ctypes.c_long, ctypes.c_void_p)


def pfLineartBS(data, particles, usef=False, fun=None, resample_mode=RESIDUAL, threshold=0.5, seed=1, noise=None,
                uniforms=None, full=False):
    """Mirror of ``pfLineartBS_impl(data, part, usef, fun)`` (reference src/pflineart.cpp:53-99): dict(Xm, Xv, Ym, Yv).

    ``data`` is a [T x 2] array (x, y observation columns).  ``fun(Xm, Ym)`` is called after every step when ``usef``.
    Trailing keyword arguments are the ABI extensions (the reference hard-codes RESIDUAL, 0.5)."""
    d = np.asfortranarray(np.asarray(data, dtype=np.float64))
    if d.ndim != 2 or d.shape[1] != 2:
        raise ValueError("data must be [T x 2]")
    T = d.shape[0]
    flat = np.ascontiguousarray(d.T.reshape(-1))  # column-major: x column then y column
    Xm, Xv, Ym, Yv, lognc, ess = (np.empty(T) for _ in range(6))
    res = np.empty(T, dtype=np.int32)
    z, u = _f64(noise), _f64(uniforms)
    cb = None
    if usef and fun is not None:
        def _cb(xm, ym, n, _user):
            fun(np.ctypeslib.as_array(xm, shape=(n,)).copy(), np.ctypeslib.as_array(ym, shape=(n,)).copy())
        cb = _STEP_CB(_cb)
    _check(lib().smcb200_pfLineartBS(_dp(flat), T, particles, resample_mode, threshold, seed, _dp(z), _dp(u),
                                     ctypes.cast(cb, ctypes.c_void_p) if cb else None, None, _dp(Xm), _dp(Xv), _dp(Ym), _dp(Yv),
                                     _dp(lognc), _dp(ess), res.ctypes.data_as(_c_int_p)))
    out = {"Xm": Xm, "Xv": Xv, "Ym": Ym, "Yv": Yv}
    if full:
        out.update(logNC=lognc, ESS=ess, resampled=res)
    return out


def _la_data(Data):
    d = np.asarray(Data, dtype=np.float64)
    if d.ndim != 2 or d.shape[1] != 2:
        raise ValueError("Data must be [n_obs x 2] (y, x)")
    return np.ascontiguousarray(d.T.reshape(-1)), d.shape[0]   # column-major: y column then x column


def LinReg(Data, lNumber, seed=1, noise=None, uniforms=None, resample_uniforms=None):
    """Mirror of ``LinReg_impl(Data, lNumber)`` (reference src/LinReg.cpp:46-84): dict(theta [N,3], weights [N], logNC)."""
    flat, n_obs = _la_data(Data)
    N = int(lNumber)
    th, w, nc = np.empty(3 * N), np.empty(N), np.empty(1)
    z, u, ud = _f64(noise), _f64(uniforms), _f64(resample_uniforms)
    _check(lib().smcb200_LinReg(_dp(flat), n_obs, N, seed, _dp(z), 0 if z is None else z.size, _dp(u), 0 if u is None else u.size, _dp(ud),
                                _dp(th), _dp(w), _dp(nc)))
    return {"theta": th.reshape(3, N).T, "weights": w, "logNC": float(nc[0])}


def LinRegLA_adapt(Data, lNumber, resampTol=0.5, tempTol=0.9, seed=1, noise=None, uniforms=None, max_temps=512):
    """Mirror of ``LinRegLA_adapt_impl(Data, lNumber, resampTol, tempTol)`` (reference src/LinReg_LA_adapt.cpp:36-106).
    Returns the reference's list: theta [N,3,L], loglike/logprior/Weights [N,L], ESS, Temps, mcmcRepeats [L] and the four
    log-evidence estimates."""
    flat, n_obs = _la_data(Data)
    N = int(lNumber)
    z, u = _f64(noise), _f64(uniforms)
    # The history (6 doubles per particle and temperature, on the device and here) is sized by the temperature capacity, while a
    # run typically needs 20-40 temperatures: try a small capacity first and repeat with the caller's only if the schedule turns
    # out longer (the draws are counter-based, so the repeat reproduces the same run).
    for cap in ([64, max_temps] if max_temps > 64 else [max_temps]):
        th = np.empty(cap * 3 * N); ll = np.empty(cap * N); lp = np.empty(cap * N); w = np.empty(cap * N)
        ess, temps, reps, acc = (np.empty(cap) for _ in range(4))
        nc = np.empty(4)
        L = lib().smcb200_LinRegLA_adapt(_dp(flat), n_obs, N, resampTol, tempTol, seed, _dp(z), 0 if z is None else z.size, _dp(u),
                                         0 if u is None else u.size, cap, _dp(th), _dp(ll), _dp(lp), _dp(w), _dp(ess), _dp(temps),
                                         _dp(reps), _dp(nc), _dp(acc))
        if L > 0 or cap == max_temps or b"more temperatures than max_temps" not in lib().smcb200_last_error():
            break
    if L <= 0:
        raise SmcError(f"smcb200 status {-L}: {lib().smcb200_last_error().decode()}")
    return {"theta": th[:L * 3 * N].reshape(L, 3, N).transpose(2, 1, 0), "loglike": ll[:L * N].reshape(L, N).T,
            "logprior": lp[:L * N].reshape(L, N).T, "Weights": w[:L * N].reshape(L, N).T, "ESS": ess[:L].copy(), "Temps": temps[:L].copy(),
            "mcmcRepeats": reps[:L].copy(), "logNC_standard": nc[0], "logNC_ps_rect": nc[1], "logNC_ps_trap": nc[2],
            "logNC_ps_trap2": nc[3], "acceptProb": acc[:L].copy()}


def LinRegLA(Data, intemps, lNumber, seed=1, noise=None, uniforms=None):
    """Mirror of ``LinRegLA_impl(Data, intemps, lNumber)`` (reference src/LinReg_LA.cpp:45-111)."""
    flat, n_obs = _la_data(Data)
    N = int(lNumber)
    t = _f64(intemps)
    L = t.size
    th = np.empty(L * 3 * N); ll = np.empty(L * N); lp = np.empty(L * N); w = np.empty(L * N)
    ess, nc = np.empty(L), np.empty(4)
    z, u = _f64(noise), _f64(uniforms)
    _check(lib().smcb200_LinRegLA(_dp(flat), n_obs, _dp(t), L, N, seed, _dp(z), 0 if z is None else z.size, _dp(u), 0 if u is None else u.size,
                                  _dp(th), _dp(ll), _dp(lp), _dp(w), _dp(ess), _dp(nc)))
    return {"theta": th.reshape(L, 3, N).transpose(2, 1, 0), "loglike": ll.reshape(L, N).T, "logprior": lp.reshape(L, N).T,
            "Weights": w.reshape(L, N).T, "ESS": ess, "logNC_standard": nc[0], "logNC_ps_rect": nc[1], "logNC_ps_trap": nc[2],
            "logNC_ps_trap2": nc[3]}


_PMMH_CB = ctypes.CFUNCTYPE(None, ctypes.c_long, ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                            ctypes.c_double, ctypes.c_void_p)


def nonLinPMMH(data, lNumber, lMCMCits, verbose=False, msg_freq=100, resample_mode=MULTINOMIAL, threshold=0.5, seed=1, innovations=None,
               filter_noise=None, filter_uniforms=None, progress=None):
    """Mirror of ``nonLinPMMH_impl(data, lNumber, lMCMCits, verbose, msg_freq)`` (reference src/nonLinPMMH.cpp:37-153): dict with
    samples_sigv, samples_sigw, loglike, logprior, MH_acceptance_rate (lMCMCits + 1 entries each).
    ``innovations`` = (innov_v, innov_w, unif) pre-drawn outer-chain randomness; filter_* inject the inner filters' draws."""
    y = _f64(np.asarray(data).reshape(-1))
    M = int(lMCMCits)
    out = [np.empty(M + 1) for _ in range(5)]
    iv = iw = uu = None
    if innovations is not None:
        iv, iw, uu = (_f64(a) for a in innovations)
    zn, un = _f64(filter_noise), _f64(filter_uniforms)
    cb = None
    if progress is not None or verbose:
        def _cb(i, m, mv, mw, ll, lp, ar, _user):
            if progress is not None:
                progress(i, m, mv, mw, ll, lp, ar)
            else:
                print(f"Percentage completed: {round(100.0 * i / m)}%. mean sigv {mv:.4f} mean sigw {mw:.4f} loglike {ll:.3f} acceptance {round(100 * ar)}%")
        cb = _PMMH_CB(_cb)
    _check(lib().smcb200_nonLinPMMH(_dp(y), y.size, int(lNumber), M, seed, resample_mode, threshold, _dp(iv), _dp(iw), _dp(uu), _dp(zn), _dp(un),
                                    ctypes.cast(cb, ctypes.c_void_p) if cb else None, msg_freq, None, *[_dp(o) for o in out]))
    return dict(zip(("samples_sigv", "samples_sigw", "loglike", "logprior", "MH_acceptance_rate"), out))


def blockpfGaussianOpt(data, particles, lag, seed=1, noise=None, uniforms=None, diagnostics=False):
    """Mirror of ``blockpfGaussianOpt_impl(data, part, lag)`` (reference src/blockpfgaussianopt.cpp:41-71): dict with weight [N],
    values [N x T] and logNC.  ``noise`` [T][lag][N] / ``uniforms`` [T] inject the draws; ``diagnostics`` adds the per-step ESS,
    resampling flags and ancestor rows [T][N]."""
    y = _f64(np.asarray(data).reshape(-1))
    T, N = y.size, int(particles)
    w, v, nc = np.empty(N), np.empty((T, N)), np.empty(1)
    zn, un = _f64(noise), _f64(uniforms)
    if zn is not None and zn.size != T * int(lag) * N:
        raise ValueError("noise must have T * lag * particles entries")
    ess = rs = anc = None
    if diagnostics:
        ess, rs, anc = np.empty(T), np.empty(T, dtype=np.int32), np.empty((T, N), dtype=np.uint32)
    _check(lib().smcb200_blockpfGaussianOpt(_dp(y), T, N, int(lag), seed, _dp(zn), _dp(un), _dp(w), _dp(v), _dp(nc), _dp(ess),
                                            rs.ctypes.data_as(_c_int_p) if rs is not None else None,
                                            anc.ctypes.data_as(_c_uint_p) if anc is not None else None))
    out = {"weight": w, "values": v.T, "logNC": float(nc[0])}
    if diagnostics:
        out.update(ESS=ess, resampled=rs, ancestors=anc)
    return out


def blockpfGaussianOpt_device(data, particles, lag, values_dev, seed=1):
    """``blockpfGaussianOpt`` with the [N x T] path matrix left on the GPU: ``values_dev`` is a device pointer (int) to
    particles x T doubles, column-major (e.g. ``torch.empty((T, N), dtype=torch.float64, device="cuda").data_ptr()``).
    Returns weight [N], logNC, ESS [T], resampled [T] in host arrays."""
    y = _f64(np.asarray(data).reshape(-1))
    T, N = y.size, int(particles)
    w, nc, ess, rs = np.empty(N), np.empty(1), np.empty(T), np.empty(T, dtype=np.int32)
    _check(lib().smcb200_blockpfGaussianOpt_device(_dp(y), T, N, int(lag), seed, _dp(w), ctypes.c_void_p(int(values_dev)), _dp(nc), _dp(ess),
                                                   rs.ctypes.data_as(_c_int_p)))
    return {"weight": w, "logNC": float(nc[0]), "ESS": ess, "resampled": rs}


def _lgss_par(parInits):
    if isinstance(parInits, dict):
        return tuple(float(parInits[k]) for k in ("phi", "varStateEvol", "stateInit", "varObs"))
    return tuple(float(v) for v in parInits)


def compareNCestimates(data, lParticleNum, simNum, parInits, referenceTraj, seed=1, independent_sims=False):
    """Mirror of ``compareNCestimates_imp(data, lParticleNum, simNum, parInits, referenceTraj)`` (reference src/cSMCexamples.cpp:36-139):
    dict with smcOut and csmcOut [simNum x 4] (columns multinomial, residual, stratified, systematic).  ``parInits`` is the
    reference's list {phi, varStateEvol, stateInit, varObs}; ``referenceTraj`` is [T x simNum]."""
    y = _f64(np.asarray(data).reshape(-1))
    ref = np.asarray(referenceTraj, dtype=np.float64)
    if ref.shape != (y.size, int(simNum)):
        raise ValueError("referenceTraj must be [T x simNum]")
    refc = np.ascontiguousarray(ref.T)   # column-major [T x simNum]
    a, b = np.empty((4, int(simNum))), np.empty((4, int(simNum)))
    phi, ve, x0, vo = _lgss_par(parInits)
    _check(lib().smcb200_compareNCestimates(_dp(y), y.size, int(lParticleNum), int(simNum), phi, ve, x0, vo, _dp(refc), seed,
                                            1 if independent_sims else 0, _dp(a), _dp(b)))
    return {"smcOut": a.T, "csmcOut": b.T}


def conditionalBPF(data, particles, parInits, modes, reference, runs_per_chain=None, seed=1, noise=None, uniforms=None, history=False):
    """Chains of conditional bootstrap-filter runs (``smc::conditionalSampler`` on the LGSS example, HistoryType::AL).  ``modes`` [runs],
    ``reference`` [runs][T]; runs are grouped into chains of ``runs_per_chain`` (default: one chain) that share the sampler's
    reference-index state like runs on one reference sampler object.  Returns logNC [runs], refidx [runs][T] and, with
    ``history``, the AL history (ancestors, values, logw [runs][T][N], resampled [runs][T])."""
    y = _f64(np.asarray(data).reshape(-1))
    T, N = y.size, int(particles)
    modes = np.ascontiguousarray(modes, dtype=np.int32).reshape(-1)
    runs = modes.size
    rpc = runs if runs_per_chain is None else int(runs_per_chain)
    if runs % rpc:
        raise ValueError("runs must be a multiple of runs_per_chain")
    ref = _f64(np.asarray(reference).reshape(runs, T))
    zn, un = _f64(noise), _f64(uniforms)
    nc, ridx = np.empty(runs), np.empty((runs, T), dtype=np.uint32)
    anc = val = lw = rs = None
    if history:
        anc, val, lw = np.empty((runs, T, N), dtype=np.uint32), np.empty((runs, T, N)), np.empty((runs, T, N))
        rs = np.empty((runs, T), dtype=np.int32)
    phi, ve, x0, vo = _lgss_par(parInits)
    _check(lib().smcb200_conditionalBPF(_dp(y), T, N, phi, ve, x0, vo, runs // rpc, rpc, modes.ctypes.data_as(_c_int_p), _dp(ref), seed,
                                        _dp(zn), _dp(un), _dp(nc), ridx.ctypes.data_as(_c_uint_p),
                                        anc.ctypes.data_as(_c_uint_p) if history else None, _dp(val), _dp(lw),
                                        rs.ctypes.data_as(_c_int_p) if history else None))
    out = {"logNC": nc, "refidx": ridx}
    if history:
        out.update(ancestors=anc, values=val, logw=lw, resampled=rs)
    return out


def lgssBPF(data, particles, parInits, resample_mode, seed=1, noise=None, uniforms=None):
    """Log normalising constant of one plain bootstrap filter on the LGSS example (SamplerBPF of cSMCexamples.cpp:66-99)."""
    y = _f64(np.asarray(data).reshape(-1))
    nc = np.empty(1)
    phi, ve, x0, vo = _lgss_par(parInits)
    _check(lib().smcb200_lgssBPF(_dp(y), y.size, int(particles), phi, ve, x0, vo, int(resample_mode), seed, _dp(_f64(noise)), _dp(_f64(uniforms)), _dp(nc)))
    return float(nc[0])


def ancestor_lines(ancestors, values=None):
    """``GetALineInd(n)`` for every particle n (columns of the returned [T][N] array) and, with ``values`` [T][N], ``GetALineSpace``
    (reference sampler.h:445-476) from the per-step ancestor rows of an AL history."""
    anc = np.ascontiguousarray(ancestors, dtype=np.uint32)
    T, N = anc.shape
    lines = np.empty((T, N), dtype=np.uint32)
    v = _f64(values)
    lv = np.empty((T, N)) if v is not None else None
    _check(lib().smcb200_ancestor_lines(anc.ctypes.data_as(_c_uint_p), T, N, _dp(v), lines.ctypes.data_as(_c_uint_p), _dp(lv)))
    return (lines, lv) if v is not None else lines


class PfNonlinBS:
    """Handle-based filter: device buffers persist across runs (benchmarking, repeated filtering)."""

    def __init__(self, particles, tcap, seed=1, stream=None, island=None):
        """``island=(rank, world)``: this process owns ``particles`` particles of one population of ``world * particles``
        with global resampling (one process per GPU); exchange ``ipc_handle()`` between ranks and ``connect()``."""
        self._h = ctypes.c_void_p()
        self.particles, self.tcap = particles, tcap
        self.world = 1 if island is None else island[1]
        if island is None:
            _check(lib().smcb200_pfNonlinBS_create(particles, tcap, seed, stream, ctypes.byref(self._h)))
        else:
            _check(lib().smcb200_pfNonlinBS_create_island(particles, tcap, seed, stream, island[0], island[1], ctypes.byref(self._h)))
        self.T = 0

    def ipc_handle(self):
        buf = ctypes.create_string_buffer(64)
        _check(lib().smcb200_pfNonlinBS_ipc_export(self._h, buf))
        return buf.raw

    def connect(self, handles):
        blob = ctypes.create_string_buffer(b"".join(handles), 64 * len(handles))
        _check(lib().smcb200_pfNonlinBS_ipc_import(self._h, blob))

    def set_data(self, data):
        y = _f64(data)
        self.T = y.size
        _check(lib().smcb200_pfNonlinBS_set_data(self._h, _dp(y), y.size))

    def set_noise(self, noise=None, uniforms=None):
        z, u = _f64(noise), _f64(uniforms)
        _check(lib().smcb200_pfNonlinBS_set_noise(self._h, _dp(z), _dp(u), 0 if u is None else u.size))

    def run(self, resample_mode=SYSTEMATIC, threshold=None):
        thr = 1.01 * self.particles * self.world if threshold is None else threshold  # absolute thresholds refer to the whole population
        _check(lib().smcb200_pfNonlinBS_run(self._h, resample_mode, thr))

    def sync(self):
        _check(lib().smcb200_pfNonlinBS_sync(self._h))

    def elapsed_ms(self):
        ms = ctypes.c_float()
        _check(lib().smcb200_pfNonlinBS_elapsed_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def read(self):
        T = self.T
        mean, sd, lognc, ess = np.empty(T), np.empty(T), np.empty(T), np.empty(T)
        res = np.empty(T, dtype=np.int32)
        _check(lib().smcb200_pfNonlinBS_read(self._h, _dp(mean), _dp(sd), _dp(lognc), _dp(ess), res.ctypes.data_as(_c_int_p)))
        return {"mean": mean, "sd": sd, "logNC": lognc, "ESS": ess, "resampled": res}

    def read_ancestors(self):
        idx = np.empty(self.particles, dtype=np.uint32)
        _check(lib().smcb200_pfNonlinBS_read_ancestors(self._h, idx.ctypes.data_as(_c_uint_p)))
        return idx

    def set_profile(self, on=True):
        _check(lib().smcb200_pfNonlinBS_set_profile(self._h, 1 if on else 0))

    def kernel_ms(self):
        """(move_ms, move_launches, scan_ms, scan_launches) of the last profiled run."""
        mv, sc = ctypes.c_float(), ctypes.c_float()
        nm, ns = ctypes.c_long(), ctypes.c_long()
        _check(lib().smcb200_pfNonlinBS_kernel_ms(self._h, ctypes.byref(mv), ctypes.byref(nm), ctypes.byref(sc), ctypes.byref(ns)))
        return mv.value, nm.value, sc.value, ns.value

    def launches(self):
        return lib().smcb200_pfNonlinBS_launches(self._h)

    def close(self):
        if self._h:
            lib().smcb200_pfNonlinBS_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
