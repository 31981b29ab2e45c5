"""GPU: the filter-level oracle comparisons once more with the single-launch path switched off (SMCB200_SMALL_MAX=0), so that
the kernel-per-step path stays covered at the small sizes the oracle can afford -- all four resampling schemes, injected draws."""
import os

import numpy as np
import pytest

import helpers as H
import rcppsmc_b200 as smc

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture
def kernel_per_step_path():
    old = os.environ.get("SMCB200_SMALL_MAX")
    os.environ["SMCB200_SMALL_MAX"] = "0"
    yield
    if old is None:
        os.environ.pop("SMCB200_SMALL_MAX", None)
    else:
        os.environ["SMCB200_SMALL_MAX"] = old


@pytest.mark.parametrize("N,T", [(1000, 50), (37, 20), (4099, 30), (1 << 15, 12)])
@pytest.mark.parametrize("mode", [H.SYSTEMATIC, H.STRATIFIED, H.MULTINOMIAL, H.RESIDUAL])
def test_kernel_per_step_path_matches_oracle(kernel_per_step_path, N, T, mode):
    rng = np.random.default_rng(N + T + mode)
    _, y = H.sim_nonlin(T, N + T + mode)
    z = rng.standard_normal(N * T)
    U = rng.random({H.SYSTEMATIC: T, H.STRATIFIED: T * (N + 1)}.get(mode, T * N))
    got = smc.pfNonlinBS(y, N, resample_mode=mode, threshold=1.01 * N, noise=z, uniforms=U, full=True)
    exp = H.o_pfnonlinbs(y, N, mode, 1.01 * N, z=z, U=U)
    assert np.array_equal(got["resampled"], exp["resampled"])
    np.testing.assert_allclose(got["logNC"], exp["logNC"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(got["ESS"], exp["ESS"], rtol=RTOL)
    np.testing.assert_allclose(got["mean"], exp["mean"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(got["sd"], exp["sd"], rtol=RTOL, atol=1e-9)


@pytest.mark.parametrize("mode", [H.MULTINOMIAL, H.RESIDUAL])
def test_draw_schemes_same_ancestors_on_both_paths(mode):
    """counts -> ancestor map: streaming passes (kernel-per-step path) against the single-launch path, device draws"""
    T, N = 12, 60000
    _, y = H.sim_nonlin(T, 9)
    res = []
    for cut in ("0", str(1 << 18)):
        os.environ["SMCB200_SMALL_MAX"] = cut
        try:
            f = smc.PfNonlinBS(N, T, seed=5)
            f.set_data(y)
            f.run(mode, 1.01 * N)
            f.sync()
            res.append((f.read(), f.read_ancestors()))
            f.close()
        finally:
            os.environ.pop("SMCB200_SMALL_MAX", None)
    (a, anc_a), (b, anc_b) = res
    np.testing.assert_allclose(a["logNC"], b["logNC"], rtol=1e-9)
    assert np.mean(anc_a != anc_b) < 1e-3
