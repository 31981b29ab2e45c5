"""rcppsmc_b200 -- B200-native SMC engine behind the RcppSMC plug-in surface.

The product is ``libsmcb200.so`` (hand-written CUDA for sm_100a behind the C ABI in ``include/smcb200.h``).
This package is the thin host-side mirror of the reference's exported entry points
(reference ``R/RcppExports.R:4-34``): same names, same argument meaning, numpy arrays in and out.
There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but every compute
call raises ``SmcError`` when the CUDA extension or a device is missing.
"""
from ._native import (  # noqa: F401
    MULTINOMIAL, RESIDUAL, STRATIFIED, SYSTEMATIC, SmcError, PfNonlinBS, lib, lib_path, device_count,
    log_sum_exp, fastmath_eval, resample, philox4x32, philox4x32_device, philox_normals, pfNonlinBS, pfLineartBS, LinReg, LinRegLA, LinRegLA_adapt, nonLinPMMH, blockpfGaussianOpt, blockpfGaussianOpt_device, compareNCestimates, conditionalBPF, lgssBPF, ancestor_lines,
)

__all__ = [
    "MULTINOMIAL", "RESIDUAL", "STRATIFIED", "SYSTEMATIC", "SmcError", "PfNonlinBS", "lib", "lib_path",
    "device_count", "log_sum_exp", "fastmath_eval", "resample", "philox4x32", "philox4x32_device", "philox_normals", "pfNonlinBS", "pfLineartBS", "LinReg", "LinRegLA", "LinRegLA_adapt", "nonLinPMMH", "blockpfGaussianOpt", "blockpfGaussianOpt_device", "compareNCestimates", "conditionalBPF", "lgssBPF", "ancestor_lines",
]
