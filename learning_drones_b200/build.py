"""Build libldrones.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m learning_drones_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libldrones.so")
# the same library with the episode kernel's own bounds checks compiled in (-DLD_CHECKED); tests/test_gpu_soak.py runs it
LIB_CHECKED = os.path.join(HERE, "libldrones_checked.so")
SOURCES = [os.path.join(CSRC, "ldrones.cu")]
DEPS = SOURCES + [os.path.join(CSRC, h) for h in ("ld_step.cuh", "ld_math.cuh", "ld_f32.cuh")] + [os.path.join(ROOT, "include", "ldrones.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # python-float expressions of the reference are unfused; fused products are explicit fma()
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-I", os.path.join(ROOT, "include"),
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libldrones.so cannot be built")


def is_stale() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(LIB_CHECKED)):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(LIB_CHECKED))
    return any(os.path.getmtime(f) > t for f in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    checked = subprocess.Popen([nvcc_path()] + NVCC_FLAGS + ["-DLD_CHECKED", "-o", LIB_CHECKED] + SOURCES, stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True)  # in parallel with the product build
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-Xptxas", "-v", "-o", LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    checked_out, _ = checked.communicate()
    if checked.returncode != 0:
        sys.stderr.write(checked_out)
        raise RuntimeError("nvcc failed building libldrones_checked.so")
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libldrones.so")
    spilled = episode_kernel_spills(r.stdout + r.stderr)
    if spilled:
        os.remove(LIB)
        raise RuntimeError(f"the fp64 episode kernel spills registers ({spilled}): its hot loop must stay in registers")
    return LIB


def episode_kernel_spills(ptxas_log: str):
    """ptxas -v lines of the fly_mlp_f64_kernel instantiations that report spill stores / loads (a spilling hot loop
    costs 30 % — measured)."""
    bad, current = [], None
    for line in ptxas_log.splitlines():
        if "Function properties for" in line or "Compiling entry function" in line:
            current = line
        elif "spill stores" in line and current and "fly_mlp_f64_kernel" in current and " 0 bytes spill stores, 0 bytes spill loads" not in line:
            bad.append(line.strip())
    return bad


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
