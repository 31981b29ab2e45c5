"""Build recipe for libsmcb200.so (hand-written CUDA for sm_100a) -- in-tree, so the .so travels with the repo."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libsmcb200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("cabi.cu", "cabi_lineart.cu", "cabi_linreg.cu", "cabi_pmmh.cu", "cabi_blockpf.cu", "cabi_csmc.cu")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in ("csmc.cuh", "small.cuh", "generic.cuh", "common.cuh", "fastmath.cuh", "philox.cuh", "resample.cuh", "engine.cuh", "models.cuh", "cabi_common.cuh", "island.cuh", "la_engine.cuh")]
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


def _compile_objects(flags, obj_dir, sources, force, verbose):
    """One nvcc per translation unit, all at once (the units are independent: no relocatable device code), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    newest_header = max(os.path.getmtime(f) for f in HEADERS)
    jobs = []
    for src in sources:
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append((src, obj))
    def run(job):
        cmd = [nvcc] + [f for f in flags if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", job[1], job[0]]
        return job, subprocess.run(cmd, capture_output=True, text=True)
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for job, r in ex.map(run, jobs):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed compiling " + job[0])
            if verbose:
                sys.stderr.write(r.stderr)
    return [os.path.join(obj_dir, os.path.basename(src) + ".o") for src in sources]


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = _compile_objects(NVCC_FLAGS, os.path.join(ROOT, ".objcache", "main"), SOURCES, force, verbose)
    r = subprocess.run([os.environ.get("NVCC", "nvcc"), "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libsmcb200.so")
    return LIB


def build_variant(name, defines):
    """Experiment builds (kernel A/B measurements): build/libsmcb200_<name>.so with extra -D switches; loaded via SMCB200_LIB."""
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libsmcb200_{name}.so")
    objs = _compile_objects(NVCC_FLAGS + [f"-D{d}" for d in defines], os.path.join(ROOT, ".objcache", name), SOURCES, True, False)
    r = subprocess.run([os.environ.get("NVCC", "nvcc"), "-shared", "-o", out] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed linking {out}")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
