import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, helpers as H, rcppsmc_b200 as smc

def philox_u(seed, stream, step, ids, slot):
    """numpy Philox4x32-10: returns the `a` lane (u64) of philox_draw for each id"""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    ids = ids.astype(np.uint64)
    c0 = (ids & np.uint64(0xFFFFFFFF)); c1 = (ids >> np.uint64(32))
    c2 = np.full_like(c0, step); c3 = np.full_like(c0, (stream << 24) | slot)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for r in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c1 ^ np.uint64(k0); n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF; k1 = (k1 + W1) & 0xFFFFFFFF
    return (c1 << np.uint64(32)) | c0

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
T = 3
N = 1 << logn
seed = 77
_, y = H.sim_nonlin(T, 3)
nat = smc.pfNonlinBS(y, N, resample_mode=smc.STRATIFIED, seed=seed, full=True)
# rebuild the same draws on the host
z = np.concatenate([smc.philox_normals(seed, 1, 0, N)] + [smc.philox_normals(seed, 2, t, N) for t in range(1, T)])
U = np.empty(T * (N + 1))
for t in range(T):
    a = philox_u(seed, 4, t, np.arange(N + 1), 0)
    U[t * (N + 1):(t + 1) * (N + 1)] = (a >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
inj = smc.pfNonlinBS(y, N, resample_mode=smc.STRATIFIED, seed=seed, noise=z, uniforms=U, full=True)
print("native vs injected logNC", nat["logNC"], inj["logNC"], "mean", nat["mean"], inj["mean"], flush=True)
t0 = time.time()
o = H.o_pfnonlinbs(y, N, H.STRATIFIED, 1.01 * N, z=z, U=U)
print("oracle(raw) logNC", o["logNC"], "mean", o["mean"], f"({time.time()-t0:.0f}s)")
o = H.o_pfnonlinbs(y, N, H.STRATIFIED, 1.01 * N, z=z, U=U, grid=True)
print("oracle(grid) logNC", o["logNC"], "mean", o["mean"])
np.set_printoptions(precision=17)
print("diff native-oracle(grid) logNC", nat["logNC"] - o["logNC"], "mean", nat["mean"] - o["mean"])
