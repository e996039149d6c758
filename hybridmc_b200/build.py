"""In-tree build of the CUDA library (sm_100a only).

`build_library()` compiles hybridmc_b200/csrc/*.cu with nvcc into
hybridmc_b200/libhybridmc_b200.so.  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   Blackwell B200 only, no PTX fallback
  -fmad=false                               no FMA contraction: the event times must
                                            be bit-identical to the reference's
                                            non-fused x86-64 arithmetic
  -lineinfo                                 ncu source page maps to the .cu
  -l:libstdc++.so.6                         this image would otherwise link libstdc++
                                            statically, which clashes with torch's
                                            shared copy inside one Python process
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhybridmc_b200.so")
SOURCES = ["hmc_kernels.cu", "hmc_api.cu", "hmc_host.cu", "hmc_program.cu", "hmc_h5.cu", "hmc_comm.cu"]
HEADERS = [os.path.join(CSRC, "hmc_internal.h"), os.path.join(CSRC, "host", "h5lite.h"),
           os.path.join(os.path.dirname(HERE), "include", "hybridmc_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the engine cannot be built (there is no CPU fallback)")


def nvcc_flags(extra=()):
    return ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
            "-std=c++17", "-Xcompiler", "-fPIC", *extra]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, out=None, defines=()):
    """out/defines: build an experimental variant next to the default library."""
    if out is None and not force and not is_stale():
        return LIB
    cmd = [nvcc_path(), *nvcc_flags(["-D" + d for d in defines]), "-shared", "-o", out or LIB,
           *[os.path.join(CSRC, s) for s in SOURCES],
           "-cudart", "static", "-Xlinker", "-l:libstdc++.so.6", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out or LIB


# Checking builds of the same sources, loaded through HMC_LIB by tests/test_gpu_debug_builds.py:
#   verify: every pair the single-precision pre-filter drops is predicted exactly anyway and
#           error bit 0x40000000 is raised if one of them has an event
#   bounds: every index into the shared-memory replica image is range-checked
VARIANTS = {"verify": ("HMC_VERIFY_PREFILTER", os.path.join(HERE, "libhybridmc_b200_verify.so")),
            "bounds": ("HMC_DEBUG_BOUNDS", os.path.join(HERE, "libhybridmc_b200_bounds.so"))}


def build_variants(force=False):
    """Both checking builds, compiled concurrently."""
    from concurrent.futures import ThreadPoolExecutor
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    todo = [(d, out) for d, out in VARIANTS.values()
            if force or not os.path.exists(out)
            or any(os.path.getmtime(x) > os.path.getmtime(out) for x in deps)]
    with ThreadPoolExecutor(max_workers=2) as ex:
        list(ex.map(lambda t: build_library(force=True, out=t[1], defines=(t[0],)), todo))
    return {k: v[1] for k, v in VARIANTS.items()}


EXE = os.path.join(HERE, "hybridmc")


def build_executable(force=False):
    """The C++17 host executable (hybridmc_b200/hybridmc), linked against the library."""
    src = os.path.join(CSRC, "host", "hybridmc_main.cpp")
    deps = [src, os.path.join(CSRC, "host", "mini_json.h"), LIB] + HEADERS
    if not force and os.path.exists(EXE) and all(
            os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps if os.path.exists(d)):
        return EXE
    # zlib: --input-file may be a stock reference output (chunked + deflate datasets)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-DH5LITE_WITH_ZLIB", "-o", EXE, src, LIB,
           "-Wl,-rpath,$ORIGIN", "-l:libstdc++.so.6", "-lz"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return EXE


if __name__ == "__main__":
    import sys
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv,
                        out=outs[0] if outs else None, defines=defs))
    if not outs:
        print(build_executable(force="--force" in sys.argv))
    if "--variants" in sys.argv:
        print(build_variants(force="--force" in sys.argv))
