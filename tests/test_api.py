"""Boundary tests (no GPU): the C-ABI library loads, exports every declared symbol, and its structs match the header."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from raytracingdenoiser_b200 import nrd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nrd_b200.h")


def test_every_declared_symbol_is_exported():
    text = open(HEADER).read()
    declared = set(re.findall(r"^\s*NRD_API[^;(]*?\b(\w+)\s*\(", text, re.M)) - {"__attribute__"}
    assert len(declared) >= 19, declared
    lib = C.CDLL(nrd.library_path())
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in nrd_b200.h is not exported" % name
    assert set(nrd.EXPORTED_SYMBOLS) <= declared


def test_struct_layouts_match_the_header():
    """Compile a probe against include/nrd_b200.h and compare sizeof/offsetof with the ctypes mirror."""
    probe = r'''
#include "nrd_b200.h"
#include <cstdio>
#include <cstddef>
#define S(T) printf(#T " %zu\n", sizeof(nrd::T))
#define O(T, m) printf(#T "." #m " %zu\n", offsetof(nrd::T, m))
int main() {
  S(CommonSettings); S(ReblurSettings); S(RelaxSettings); S(SigmaSettings); S(DispatchDesc); S(InstanceDesc); S(PipelineDesc); S(LibraryDesc);
  S(ResourceDesc); S(TextureDesc); S(InstanceCreationDesc);
  O(CommonSettings, motionVectorScale); O(CommonSettings, resourceSize); O(CommonSettings, viewZScale); O(CommonSettings, printfAt); O(CommonSettings, rectOrigin);
  O(CommonSettings, frameIndex); O(CommonSettings, accumulationMode); O(CommonSettings, enableValidation);
  O(ReblurSettings, maxAccumulatedFrameNum); O(ReblurSettings, checkerboardMode); O(ReblurSettings, minMaterialForDiffuse); O(ReblurSettings, usePrepassOnlyForSpecularMotionEstimation);
  O(RelaxSettings, atrousIterationNum); O(RelaxSettings, checkerboardMode); O(RelaxSettings, minMaterialForSpecular);
  O(DispatchDesc, constantBufferData); O(DispatchDesc, pipelineIndex); O(DispatchDesc, gridHeight);
  printf("NrdCudaContextDesc %zu\nNrdCudaTextureInfo %zu\n", sizeof(NrdCudaContextDesc), sizeof(NrdCudaTextureInfo));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "probe.cpp"), os.path.join(d, "probe")
        open(src, "w").write(probe)
        subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split("\n")
    for line in out:
        if not line.strip():
            continue
        key, val = line.split()
        if "." in key:
            t, m = key.split(".")
            assert getattr(getattr(nrd, t), m).offset == int(val), key
        else:
            assert C.sizeof(getattr(nrd, key)) == int(val), key


def test_library_desc():
    d = nrd.get_library_desc()
    assert (d["versionMajor"], d["versionMinor"], d["versionBuild"]) == (4, 14, 0)
    assert d["normalEncoding"] == 2 and d["roughnessEncoding"] == 1          # R10_G10_B10_A2_UNORM, LINEAR (reference CMakeLists.txt:28-29)
    assert d["spirvBindingOffsets"] == (100, 200, 300, 400)
    assert nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR in d["supportedDenoisers"]


def test_name_tables():
    assert nrd.get_resource_type_string(nrd.ResourceType.IN_MV) == "IN_MV"
    assert nrd.get_resource_type_string(nrd.ResourceType.PERMANENT_POOL) == "PERMANENT_POOL"
    assert nrd.get_resource_type_string(33) is None
    assert nrd.get_denoiser_string(nrd.Denoiser.SIGMA_SHADOW) == "SIGMA_SHADOW"
    assert nrd.get_denoiser_string(19) is None


def test_create_instance_errors():
    with pytest.raises(nrd.NrdError) as e:
        nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE_SH)])       # not implemented -> UNSUPPORTED, like the reference for an unknown denoiser
    assert e.value.result == nrd.Result.UNSUPPORTED
    with pytest.raises(nrd.NrdError) as e:
        nrd.Instance([(5, nrd.Denoiser.REBLUR_DIFFUSE), (5, nrd.Denoiser.SIGMA_SHADOW)])
    assert e.value.result == nrd.Result.NON_UNIQUE_IDENTIFIER


def test_user_allocator_is_honoured():
    """Every allocation of an instance goes through AllocationCallbacks and is returned on DestroyInstance."""
    lib = C.CDLL(nrd.library_path())
    libc = C.CDLL(None)
    libc.aligned_alloc.restype = C.c_void_p
    libc.aligned_alloc.argtypes = [C.c_size_t, C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    live = {}
    ALLOC = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t)
    REALLOC = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t)
    FREE = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)

    def alloc(_, size, align):
        align = max(align, 16)
        p = libc.aligned_alloc(align, (size + align - 1) // align * align)
        live[p] = size
        return p

    def free(_, p):
        if p:
            del live[p]
            libc.free(p)

    cbs = (ALLOC(alloc), REALLOC(lambda u, p, s, a: None), FREE(free))
    desc = nrd.InstanceCreationDesc()
    desc.allocationCallbacks.Allocate = C.cast(cbs[0], C.c_void_p)
    desc.allocationCallbacks.Reallocate = C.cast(cbs[1], C.c_void_p)
    desc.allocationCallbacks.Free = C.cast(cbs[2], C.c_void_p)
    arr = (nrd.DenoiserDesc * 1)()
    arr[0].identifier, arr[0].denoiser = 1, int(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR)
    desc.denoisers, desc.denoisersNum = arr, 1
    handle = C.c_void_p()
    lib.CreateInstance.argtypes = [C.c_void_p, C.c_void_p]
    assert lib.CreateInstance(C.byref(desc), C.byref(handle)) == 0
    assert len(live) > 5
    lib.DestroyInstance.argtypes = [C.c_void_p]
    lib.DestroyInstance(handle)
    assert not live, "leaked %d allocations" % len(live)


def test_cuda_context_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    inst = nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE)])
    with pytest.raises(nrd.NrdError):
        nrd.CudaContext(inst, 64, 64)


def test_public_struct_layout_matches_reference_headers():
    """sizeof / alignof / default-initialised bytes of every public struct, enum extents and the version, probed by
    tests/layout_probe.cpp: include/nrd_b200.h against the fixture generated from the reference's Include/NRD.h
    (tests/golden/nrd_layout.json, generator tests/golden/make_layout_golden.py); regenerated live when the reference is here."""
    import json
    import sys
    golden_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden_dir)
    import make_layout_golden as m
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    golden = json.load(open(os.path.join(golden_dir, "nrd_layout.json")))
    mine = json.loads(m.run_probe(os.path.join(root, "include"), "nrd_b200.h"))
    assert mine == golden, [k for k in golden if golden[k] != mine.get(k)]
    assert len(golden) == 23 and golden["CommonSettings"]["sizeof"] > 300
    if os.path.isdir("/root/reference/Include"):
        assert json.loads(m.run_probe("/root/reference/Include", "NRD.h")) == golden
