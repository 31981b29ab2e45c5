"""GPU parity, unit level: LSE/ESS, Philox normals, resampling indices -- CUDA path vs the oracle."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def smc():
    import rcppsmc_b200 as s
    assert s.device_count() >= 1, "GPU tests need a CUDA device"
    return s


def test_lse_known_answer(smc):
    # inst/skeleton/rcppsmc_hello_world.Rd:29-32 evaluated by the compiled reference
    lse, ess, mx = smc.log_sum_exp(-np.sqrt(np.arange(1.0, 5.0)))
    assert abs(lse - (-0.07985233853818718)) < 1e-14
    assert mx == -1.0


@pytest.mark.parametrize("n", [1, 2, 31, 257, 1000, 4099, 1 << 16, (1 << 20) + 3])
def test_lse_ess_vs_oracle(smc, n):
    rng = np.random.default_rng(n)
    lw = 5.0 * rng.standard_normal(n) - 700.0  # far from 0: needs the max shift
    lse, ess, mx = smc.log_sum_exp(lw)
    assert abs(lse - H.o_lse(lw)) <= 1e-12 * abs(H.o_lse(lw))
    assert abs(ess - H.o_ess(lw)) <= 1e-10 * H.o_ess(lw)
    assert mx == lw.max()


def test_lse_with_weightless_particles(smc):
    lw = np.array([-np.inf, -3.0, -np.inf, -1.0, -np.inf])
    lse, ess, _ = smc.log_sum_exp(lw)
    assert abs(lse - np.log(np.exp(-3.0) + np.exp(-1.0))) < 1e-14
    assert 1.0 < ess < 2.0


def test_philox_normals_moments(smc):
    z = smc.philox_normals(seed=7, stream=2, step=3, n=1 << 20)
    n = z.size
    assert abs(z.mean()) < 4.0 / np.sqrt(n)
    assert abs(z.var() - 1.0) < 4.0 * np.sqrt(2.0 / n)
    assert abs((z ** 3).mean()) < 4.0 * np.sqrt(15.0 / n)
    assert abs((z ** 4).mean() - 3.0) < 4.0 * np.sqrt(96.0 / n)
    # pairs come from one Box-Muller: the two halves must be uncorrelated
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 4.0 / np.sqrt(n / 2)
    # determinism and stream separation
    assert np.array_equal(z, smc.philox_normals(seed=7, stream=2, step=3, n=1 << 20))
    assert not np.array_equal(z[:64], smc.philox_normals(seed=7, stream=2, step=4, n=64))


def test_appendix_b_vectors(smc):
    # generated from the compiled reference, SYSTEMATIC / STRATIFIED / count mapping hand-verified
    w = H.o_quantise(np.array([0.05, 0.30, 0.02, 0.20, 0.01, 0.25, 0.10, 0.07]))
    assert np.array_equal(smc.resample(smc.SYSTEMATIC, w, [0.37]), [0, 1, 1, 3, 3, 5, 6, 5])
    U = [0.37, 0.91, 0.05, 0.50, 0.66, 0.12, 0.99, 0.25, 0.40]
    assert np.array_equal(smc.resample(smc.STRATIFIED, w, U), [0, 1, 1, 3, 5, 5, 6, 6])
    assert np.array_equal(smc.resample(smc.MULTINOMIAL, counts=[0, 3, 0, 2, 0, 2, 1, 0]), [1, 1, 1, 3, 3, 5, 6, 5])
    # counts summing below N: unfilled slot copies particle 0 (sampler.h:845)
    assert np.array_equal(smc.resample(smc.RESIDUAL, counts=[1, 2, 0, 2, 0, 1, 1, 0]), [0, 1, 1, 3, 3, 5, 6, 0])


SIZES = [1, 2, 3, 8, 33, 255, 2047, 2048, 2049, 5000, 65536, 100003, 1 << 20]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("mode", [H.SYSTEMATIC, H.STRATIFIED])
def test_resample_bit_exact_vs_oracle(smc, n, mode):
    rng = np.random.default_rng(1000 + n)
    for spread in (0.5, 3.0, 12.0):
        w = H.random_weights(rng, n, spread)
        U = rng.random(1 if mode == H.SYSTEMATIC else n + 1)
        idx, cnt = smc.resample(mode, w, U, return_counts=True)
        oidx, ocnt = H.o_resample_w(mode, w, U)
        assert np.array_equal(cnt, ocnt), f"counts differ (n={n}, spread={spread})"
        assert np.array_equal(idx, oidx), f"indices differ (n={n}, spread={spread})"


@pytest.mark.parametrize("n", [8, 1000, 4096, 70001])
def test_count_map_bit_exact_vs_oracle(smc, n):
    rng = np.random.default_rng(n)
    for total in (n, n - 1, max(1, n // 2)):
        p = rng.dirichlet(np.full(n, 0.3))
        counts = rng.multinomial(total, p).astype(np.int32)
        idx = smc.resample(smc.MULTINOMIAL, counts=counts)
        oidx, _ = H.o_resample_w(H.MULTINOMIAL, None, counts=counts)
        assert np.array_equal(idx, oidx)


@pytest.mark.parametrize("n", [1, 2, 9, 2047, 2049, 30000, 1 << 18])
@pytest.mark.parametrize("mode", [H.MULTINOMIAL, H.RESIDUAL])
def test_native_multinomial_residual_bit_exact_vs_oracle(smc, n, mode):
    # the engine's native definitions (iid uniforms through the exact integer inverse CDF; integer residual split)
    rng = np.random.default_rng(50 + n + mode)
    for spread in (0.7, 5.0):
        w = H.random_weights(rng, n, spread)
        U = rng.random(n)
        idx, cnt = smc.resample(mode, w, U, return_counts=True)
        oidx, ocnt = H.o_resample_w(mode, w, U)
        assert cnt.sum() == n
        assert np.array_equal(cnt, ocnt)
        assert np.array_equal(idx, oidx)


@pytest.mark.parametrize("n", [1 << 20, (1 << 21) + 7])
def test_residual_large_fraction_mass_stays_inside_the_status_payload(smc, n):
    # almost every slot has n w = 0.9: the residual fractions sum to ~0.9 n 2^(52 - shift), which must stay below the 62-bit
    # payload of the look-back status words (it once reached 2^62 at n = 2^20 and corrupted the flags of the last tiles)
    rng = np.random.default_rng(n)
    w = np.full(n, 0.9 / n)
    heavy = rng.choice(n, size=64, replace=False)
    w[heavy] += (1.0 - w.sum()) / 64
    w = H.o_quantise(w / w.sum())
    U = rng.random(n)
    idx, cnt = smc.resample(H.RESIDUAL, w, U, return_counts=True)
    oidx, ocnt = H.o_resample_w(H.RESIDUAL, w, U)
    assert cnt.sum() == n and np.array_equal(cnt, ocnt) and np.array_equal(idx, oidx)
    again = smc.resample(H.RESIDUAL, w, U)
    assert np.array_equal(idx, again)


def test_resample_degenerate_weights(smc):
    # one particle carries (almost) everything: exercises the heavy-parent paths of the surplus writer
    for n in (5000, 1 << 18):
        w = np.full(n, 1e-9 / n)
        w[n // 3] = 1.0 - w.sum() + w[n // 3]
        w = H.o_quantise(w)
        for mode, U in ((H.SYSTEMATIC, [0.123]), (H.STRATIFIED, np.random.default_rng(5).random(n + 1))):
            idx = smc.resample(mode, w, U)
            oidx, _ = H.o_resample_w(mode, w, U)
            assert np.array_equal(idx, oidx)
    # a few heavy parents of very different sizes
    n = 300000
    w = np.full(n, 0.2 / n)
    for k, m in ((10, 0.4), (150000, 0.3), (299999, 0.1)):
        w[k] += m
    w = H.o_quantise(w / w.sum())
    idx = smc.resample(H.SYSTEMATIC, w, [0.9])
    assert np.array_equal(idx, H.o_resample_w(H.SYSTEMATIC, w, [0.9])[0])


def test_resample_raw_weights_mismatch_accounting(smc):
    # Off-grid weights: the engine quantises to the 2^-52 grid, the reference sums raw doubles sequentially.
    # Any disagreement must be a boundary flip to an adjacent parent and must be rare.
    rng = np.random.default_rng(77)
    n = 1 << 18
    w = H.random_weights(rng, n, 2.0, grid=False)
    U = [rng.random()]
    _, cnt = smc.resample(H.SYSTEMATIC, w, U, return_counts=True)
    _, ocnt = H.o_resample_w(H.SYSTEMATIC, w, U)
    diff = np.flatnonzero(cnt != ocnt)
    assert diff.size <= 8
    assert int(cnt.sum()) == int(ocnt.sum()) or abs(int(cnt.sum()) - int(ocnt.sum())) <= 1
    assert np.abs(np.cumsum(cnt - ocnt)).max() <= 1


def test_invalid_arguments(smc):
    with pytest.raises(smc.SmcError):
        smc.resample(smc.SYSTEMATIC, np.ones(4) / 4, uniforms=None)
    with pytest.raises(smc.SmcError):
        smc.resample(smc.STRATIFIED, np.ones(4) / 4, uniforms=[0.1, 0.2])
    with pytest.raises(smc.SmcError):
        smc.resample(smc.MULTINOMIAL, counts=[3, 3, 0, 0])


def _ulp_err(got, exp):
    return np.abs(got - exp) / np.spacing(np.abs(exp))


def test_fastmath_accuracy(smc):
    # the engine's branch-free elementary functions against numpy (glibc), in units in the last place
    rng = np.random.default_rng(12)
    x = np.concatenate([-rng.random(200000) * 50.0, -rng.random(100000) * 700.0, [0.0, -1e-300, -708.0, -709.5, -800.0, -np.inf]])
    got = smc.fastmath_eval(0, x)
    exp = np.exp(x)
    ok = x >= -708.0
    assert _ulp_err(got[ok], exp[ok]).max() <= 2.0
    assert np.all(got[~ok] == 0.0)
    u = np.concatenate([rng.random(300000), 2.0 ** -rng.integers(1, 53, 1000), [1.0, 0.5, 2.0 ** -53, 1.0 - 2.0 ** -53]])
    u = u[u > 0]
    got = smc.fastmath_eval(1, u)
    exp = -2.0 * np.log(u)
    assert np.all(np.abs(got - exp) <= 4.0 * np.spacing(np.maximum(np.abs(exp), 1e-300)) + 1e-300)
    v = np.concatenate([rng.random(300000), [0.0, 0.25, 0.5, 0.75, 0.125, 1.0 - 2.0 ** -53]])
    s, c = smc.fastmath_eval(2, v), smc.fastmath_eval(3, v)
    assert np.abs(s - np.sin(2 * np.pi * v)).max() <= 1e-15 and np.abs(c - np.cos(2 * np.pi * v)).max() <= 1e-15
    assert np.abs(s * s + c * c - 1.0).max() <= 1e-15
    y = np.concatenate([rng.random(200000) * 80.0, 10.0 ** rng.uniform(-16, 2, 100000)])
    assert _ulp_err(smc.fastmath_eval(4, y), np.sqrt(y)).max() <= 1.0
