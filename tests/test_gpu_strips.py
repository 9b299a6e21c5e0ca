"""Strip mode (the multi-GPU data path) on ONE GPU: the frame is cut into strips that are separate contexts of this
process, connected through their arena pointers; every kernel then takes the ghost-row / owner-lookup path for rows of
other strips and the passes are separated by the flag barrier.  The result must be bit-identical to the single-context run
of the same (strip) build of the kernels, and within the parity tolerance of the one-GPU build (the two builds are
compiled separately and differ in the last bits).
(The cross-process CUDA-IPC variant of the same check is tests/multi_gpu_check.py, run under torchrun on 2+ GPUs.)"""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def _run(denoiser_name, w, h, world, frames, halo, whole_frame_call, weighted=False, settings=None, frame_fn=None):
    import os
    os.environ["NRD_B200_FORCE_STRIP_KERNELS"] = "1"  # the full-frame reference runs the strip build of the kernels
    try:
        _run_inner(denoiser_name, w, h, world, frames, halo, whole_frame_call, weighted, settings, frame_fn)
    finally:
        del os.environ["NRD_B200_FORCE_STRIP_KERNELS"]


def _run_inner(denoiser_name, w, h, world, frames, halo, whole_frame_call, weighted, settings=None, frame_fn=None):
    import torch
    from raytracingdenoiser_b200 import harness, nrd, scene, strips
    den = getattr(nrd.Denoiser, denoiser_name)
    mode = harness.radiance_mode(den)
    full = harness.GpuDenoiser(den, w, h, settings=settings)
    partition = None
    if weighted:  # cost-balanced, non-uniform strips from the sky mask of the first frame
        cost = strips.tile_row_cost_from_viewz(scene.Scene(w, h).frame(0, mode)["IN_VIEWZ"])
        partition = strips.partition_rows_weighted(h, world, cost, min_rows=max(halo, 16))
        assert len(set(y1 - y0 for y0, y1 in partition[1])) > 1, partition
    parts = [strips.StripDenoiser(den, w, h, r, world, halo_rows=halo, partition=partition, settings=settings) for r in range(world)]
    try:  # a failing case must not leak its strip-mode contexts (there are only kMaxPeerSlots per process)
        for p in parts:
            p.connect_local(parts)
        streams = [torch.cuda.Stream() for _ in parts]
        sc = scene.Scene(w, h)
        for f in range(frames):
            fr = sc.frame(f, mode)
            if frame_fn:
                fr = frame_fn(fr, f)
            cs = harness.make_common_settings(fr, w, h, f)
            full.set_inputs(fr)
            full.denoise(cs)
            torch.cuda.synchronize()
            if whole_frame_call:
                # nrdCudaDenoise per rank (with the ghost look-ahead): all launches are asynchronous, the ranks meet on the device
                for p, st in zip(parts, streams):
                    p.set_inputs(fr, st)
                    p.denoise(cs, st)
            else:
                lists = []
                for p, st in zip(parts, streams):
                    p.set_inputs(fr, st)
                    lists.append(p.dispatches(cs))
                    p.ctx.barrier(st.cuda_stream)
                n = lists[0][1]
                assert all(m == n for _, m in lists)
                # application-driven dispatch loop, ranks interleaved pass by pass
                for i in range(n):
                    for p, st, (raw, _) in zip(parts, streams, lists):
                        p.ctx.execute_raw(C.byref(raw[i]), st.cuda_stream)
            for p, st in zip(parts, streams):
                p.synchronize(st)
        ref = full.outputs()
        outs = [p.read_outputs(stream=st) for p, st in zip(parts, streams)]
        torch.cuda.synchronize()
        for name, t in ref.items():
            got = torch.cat([o[name] for o in outs], dim=0)
            same = (got.view(torch.uint8) == t.view(torch.uint8))
            assert bool(same.all()), (name, float(same.float().mean()))
    finally:
        for p in parts:
            p.destroy()
        full.destroy()


@pytest.mark.parametrize("denoiser,w,h,world,frames,halo,whole_frame_call", [
    ("REBLUR_DIFFUSE_SPECULAR", 320, 192, 2, 4, 32, True),     # ghost rows + direct peer loads beyond them
    ("REBLUR_DIFFUSE_SPECULAR", 320, 192, 2, 3, 0, False),     # minimum halo (one tile): most foreign taps are direct peer loads
    ("REBLUR_DIFFUSE_SPECULAR", 250, 141, 3, 3, 96, True),     # halo clamped to the strip height (48 rows)
    ("RELAX_DIFFUSE_SPECULAR", 320, 180, 2, 4, 16, True),
    ("RELAX_DIFFUSE_SPECULAR", 320, 180, 3, 3, 64, False),
    ("SIGMA_SHADOW", 320, 180, 4, 4, 16, True),
])
def test_strips_bit_identical_to_full_frame(denoiser, w, h, world, frames, halo, whole_frame_call):
    _run(denoiser, w, h, world, frames, halo, whole_frame_call)


@pytest.mark.parametrize("denoiser,w,h,world,halo", [("REBLUR_DIFFUSE_SPECULAR", 320, 192, 3, 32), ("RELAX_DIFFUSE_SPECULAR", 320, 180, 2, 16),
                                                    ("SIGMA_SHADOW", 320, 180, 3, 16)])
def test_cost_balanced_strips_bit_identical_to_full_frame(denoiser, w, h, world, halo):
    _run(denoiser, w, h, world, 3, halo, True, weighted=True)


def test_strip_context_rejects_foreign_user_pointers_and_bad_geometry():
    import torch
    from raytracingdenoiser_b200 import nrd
    inst = nrd.Instance([(0, nrd.Denoiser.SIGMA_SHADOW)])
    with pytest.raises(nrd.NrdError):
        nrd.CudaContext(inst, 256, 128, strip=(0, 60), strip_height=60)      # not whole tiles
    with pytest.raises(nrd.NrdError):
        nrd.CudaContext(inst, 256, 128, strip=(24, 88), strip_height=64)     # y0 is not a multiple of 16
    with pytest.raises(nrd.NrdError):
        nrd.CudaContext(inst, 256, 128, strip=(0, 96), strip_height=64)      # taller than the reserved strip height
    ctx = nrd.CudaContext(inst, 256, 128, strip=(64, 128), strip_height=64)
    t = torch.zeros((64, 256), dtype=torch.float32, device="cuda")
    with pytest.raises(nrd.NrdError):
        ctx.set_user_texture(nrd.ResourceType.IN_VIEWZ, t.data_ptr(), 1024, nrd.Format.R32_SFLOAT)
    info = ctx.get_texture(nrd.ResourceType.IN_VIEWZ)
    assert (info.firstRow, info.rowsNum, info.width, info.height) == (64, 64, 256, 128)
    ctx.destroy()
    inst.destroy()


def test_strip_build_and_one_gpu_build_agree_within_parity_tolerance():
    """The two builds of the kernels (device/common.cuh) are separate compilations: same source, last-bit differences."""
    import os
    import numpy as np
    import torch
    from raytracingdenoiser_b200 import harness, nrd, scene
    den, w, h = nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, 320, 180
    outs = []
    for force in (False, True):
        if force:
            os.environ["NRD_B200_FORCE_STRIP_KERNELS"] = "1"
        try:
            gpu = harness.GpuDenoiser(den, w, h)
            sc = scene.Scene(w, h)
            for f in range(6):
                fr = sc.frame(f)
                gpu.set_inputs(fr)
                gpu.denoise(harness.make_common_settings(fr, w, h, f))
            torch.cuda.synchronize()
            outs.append({k: v.float().cpu().numpy() for k, v in gpu.outputs().items()})
            gpu.destroy()
        finally:
            os.environ.pop("NRD_B200_FORCE_STRIP_KERNELS", None)
    for k in outs[0]:
        a, b = outs[0][k], outs[1][k]
        ok = np.abs(a - b) <= 1e-3 * np.maximum(np.abs(a), np.abs(b)) + 1e-4
        assert ok.all(axis=-1).mean() >= 0.99, (k, ok.all(axis=-1).mean())   # the sequence gate of tests/parity.py: two builds diverge like GPU vs oracle
