"""Builds the two native libraries in-tree (they travel to the GPU box with the snapshot):

  raytracingdenoiser_b200/libnrd_b200.so   product: scheduler (C++) + CUDA executor + sm_100a kernels (nvcc cross-compiles here)
  oracle/liboracle.so                      test infrastructure: CPU restatement of the reference shaders (g++, OpenMP)

Usage: python -m raytracingdenoiser_b200.build [--force]
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "raytracingdenoiser_b200")
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libnrd_b200.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
# the same oracle sources with FMA contraction allowed: a second, equally IEEE-legal evaluation of the same math.  Tests use the
# pair to measure the rounding-noise floor of the (chaotic) temporal chains: how far two legal CPU evaluations drift apart.
ORACLE_FMA_LIB = os.path.join(ORACLE_DIR, "liboracle_fma.so")  # noise-floor variant: FMA contraction allowed
ORACLE_SRC_LIB = os.path.join(ORACLE_DIR, "liboracle_src.so")  # the reference's association order where the oracle's differs (oracle/relax.cpp)
ORACLE_UV_LIB = os.path.join(ORACLE_DIR, "liboracle_uv.so")    # noise-floor variant: bilinear fetches one float ulp further (oracle/hlsl.h)

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-prec-div=false", "-prec-sqrt=false", "-ftz=true"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra"]
# -ffp-contract=off: the oracle evaluates a*b+c with two roundings, like the "pinned" arithmetic of the kernels
ORACLE_FLAGS = ["-std=c++17", "-O3", "-mavx2", "-mfma", "-mf16c", "-ffp-contract=off", "-fopenmp", "-fPIC", "-Wall", "-Wextra", "-shared"]


def _sources():
    out = []
    for d in (CSRC, os.path.join(CSRC, "device")):
        for f in sorted(os.listdir(d)):
            if f.endswith(".cpp") or f.endswith(".cu"):
                out.append(os.path.join(d, f))
    return out


def _headers_digest():
    h = hashlib.sha1()
    for d in (CSRC, os.path.join(CSRC, "device"), os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".h", ".cuh")):
                h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def _compile(src, obj, stamp, extra):
    if src.endswith(".cu"):
        cmd = [NVCC] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
    else:
        cmd = ["g++"] + CXX_FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compile failed: %s\n%s%s" % (" ".join(cmd), r.stdout, r.stderr))
    open(stamp, "w").write("ok")
    return r.stderr


def build_product(force=False):
    os.makedirs(OBJ, exist_ok=True)
    digest = _headers_digest()
    objs, jobs = [], []
    for src in _sources():
        # kernel translation units are built twice: strip (multi-GPU) addressing and plain one-GPU addressing (device/common.cuh)
        is_kernel_tu = os.path.basename(os.path.dirname(src)) == "device" and src.endswith(".cu")
        for suffix, extra in ((("", []), ("_single", ["-DNRD_B200_NO_STRIPS"])) if is_kernel_tu else (("", []),)):
            key = hashlib.sha1((digest + open(src, "rb").read().decode("utf-8", "replace") + " ".join(NVCC_FLAGS + CXX_FLAGS + extra)).encode()).hexdigest()[:16]
            base = os.path.basename(src).replace(".", "_") + suffix
            obj = os.path.join(OBJ, base + ".o")
            stamp = os.path.join(OBJ, base + "." + key + ".stamp")
            objs.append(obj)
            if force or not (os.path.exists(obj) and os.path.exists(stamp)):
                for f in os.listdir(OBJ):
                    if f.startswith(base + ".") and f.endswith(".stamp"):
                        os.remove(os.path.join(OBJ, f))
                jobs.append((src, obj, stamp, extra))
    for f in os.listdir(OBJ):  # objects of sources that no longer exist must not be linked by accident
        if f.endswith(".o") and os.path.join(OBJ, f) not in objs:
            os.remove(os.path.join(OBJ, f))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for fut in [ex.submit(_compile, *j) for j in jobs]:
                fut.result()
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s%s" % (" ".join(cmd), r.stdout, r.stderr))
    return LIB


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in sorted(os.listdir(ORACLE_DIR)) if f.endswith(".cpp")]
    deps = srcs + [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith(".h")]
    libs = ((ORACLE_LIB, ORACLE_FLAGS), (ORACLE_FMA_LIB, [f if f != "-ffp-contract=off" else "-ffp-contract=fast" for f in ORACLE_FLAGS]),
            (ORACLE_UV_LIB, ORACLE_FLAGS + ["-DORACLE_NUDGE_UV"]), (ORACLE_SRC_LIB, ORACLE_FLAGS + ["-DORACLE_REFERENCE_ASSOCIATION"]))
    if not force and all(os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps) for lib, _ in libs):
        return ORACLE_LIB
    for lib, flags in libs:
        r = subprocess.run(["g++"] + flags + ["-o", lib] + srcs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n%s%s" % (r.stdout, r.stderr))
    return ORACLE_LIB


def build_reference_host(force=False):
    """oracle/_ref/libnrd_ref.so: the reference's own host code (pass scheduling) compiled from /root/reference by
    oracle/Makefile.ref -- test infrastructure, only buildable where the reference tree is mounted (this container); the GPU box
    uses the prebuilt file that travelled with the snapshot."""
    if not os.path.isdir("/root/reference/Source"):
        return None
    if force:
        import shutil
        shutil.rmtree(os.path.join(ORACLE_DIR, "_ref"), ignore_errors=True)
    r = subprocess.run(["make", "-s", "-f", "oracle/Makefile.ref"], cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference host build failed:\n%s%s" % (r.stdout, r.stderr))
    return os.path.join(ORACLE_DIR, "_ref", "libnrd_ref.so")


def build_reference_shaders():
    """The reference's own pass shaders compiled for the CPU (oracle/build_refshaders.py -> oracle/_ref/shaders/*.so) -- test
    infrastructure that pins the oracle; only buildable where the reference tree is mounted, the prebuilt files travel."""
    if not os.path.isdir("/root/reference/Shaders/Source"):
        return
    sys.path.insert(0, ORACLE_DIR)
    try:
        import build_refshaders
        build_refshaders.build_all()
    finally:
        sys.path.remove(ORACLE_DIR)


def build_all(force=False):
    """Builds whatever is out of date.  Serialised across processes with a file lock: every rank of a torchrun job calls this."""
    import fcntl
    with open(os.path.join(PKG, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            build_reference_host(force)
            libs = build_product(force), build_oracle(force)
            build_reference_shaders()
            return libs
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build_all("--force" in sys.argv))
