"""Build recipe for libsmcb200.so (hand-written CUDA for sm_100a) -- in-tree, so the .so travels with the repo."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libsmcb200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("cabi.cu", "cabi_lineart.cu", "cabi_linreg.cu", "cabi_pmmh.cu", "cabi_blockpf.cu", "cabi_csmc.cu")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in ("common.cuh", "fastmath.cuh", "philox.cuh", "resample.cuh", "engine.cuh", "models.cuh", "cabi_common.cuh", "island.cuh", "la_engine.cuh")]
HEADERS.append(os.path.join(ROOT, "include", "smcb200.h"))
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # model arithmetic keeps the reference's operation order; fma() is explicit where wanted
    "-Xcompiler", "-fPIC", "-shared",
]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libsmcb200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


def build_variant(name, defines):
    """Experiment builds (kernel A/B measurements): build/libsmcb200_<name>.so with extra -D switches; loaded via SMCB200_LIB."""
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libsmcb200_{name}.so")
    cmd = [os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed building {out}")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
