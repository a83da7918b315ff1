"""In-tree build of libfastkv_b200.so with nvcc for sm_100a (no torch headers, no JIT cache).

The shared library lands next to this file so it travels with the repo snapshot to the GPU box.
"""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfastkv_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "--threads", "8",
    "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "fastkv_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ into libfastkv_b200.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = list(NVCC_FLAGS)
    if os.environ.get("FKV_BUILD_TRACE") == "1":      # compile the %globaltimer phase timelines into the scoring kernel
        flags.append("-DFKV_TC_TRACING")
    if os.environ.get("FKV_BUILD_XTRACE") == "1":     # %globaltimer stamps of the peer-memory exchange (tools/exchange_trace.py)
        flags.append("-DFKV_X_TRACING")
    cmd = [nvcc] + flags + ["-o", LIB] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "csrc", "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-8000:])
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
