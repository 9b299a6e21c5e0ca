"""Size-independent properties at the benchmark's full size (3840x2160), where the oracle is too slow to be the checker:
a constant signal on a flat surface is a fixed point of every chain, results are deterministic, SIGMA passes fully lit /
fully shadowed frames through unchanged."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 3840, 2160


def _flat_frame(torch, scene, dev, radiance_mode, value):
    """Wall facing the camera at z = 10, static camera, zero motion, constant radiance / hit distance."""
    z = torch.full((H, W), 10.0, device=dev)
    n = torch.zeros((H, W, 3), device=dev)
    n[..., 2] = -1.0
    rough = torch.full((H, W), 0.5, device=dev)
    mat = torch.zeros((H, W), device=dev)
    rad = torch.tensor(value, device=dev).expand(H, W, 3)
    hit = torch.full((H, W), 3.0, device=dev)
    sky = torch.zeros((H, W), dtype=torch.bool, device=dev)
    fr = {"IN_VIEWZ": z.contiguous(), "IN_NORMAL_ROUGHNESS": scene.pack_normal_roughness(n, rough, mat),
          "IN_MV": torch.zeros((H, W, 4), dtype=torch.float16, device=dev)}
    if radiance_mode == "reblur":
        fr["IN_DIFF_RADIANCE_HITDIST"] = scene.pack_reblur(rad, hit, z, torch.ones_like(rough), sky)
        fr["IN_SPEC_RADIANCE_HITDIST"] = scene.pack_reblur(rad, hit, z, rough, sky)
    else:
        fr["IN_DIFF_RADIANCE_HITDIST"] = scene.pack_relax(rad, hit, sky)
        fr["IN_SPEC_RADIANCE_HITDIST"] = scene.pack_relax(rad, hit, sky)
    proj = scene.perspective_lh(60.0, W / float(H))
    view = np.eye(4, dtype=np.float32)
    fr.update({"viewToClip": proj, "worldToView": view, "worldToViewPrev": view})
    return fr


@pytest.mark.parametrize("denoiser", ["REBLUR_DIFFUSE_SPECULAR", "RELAX_DIFFUSE_SPECULAR"])
def test_constant_signal_is_a_fixed_point_at_4k(denoiser):
    import torch
    from raytracingdenoiser_b200 import harness, nrd, scene
    den = getattr(nrd.Denoiser, denoiser)
    mode = harness.radiance_mode(den)
    dev = torch.device("cuda", 0)
    fr = _flat_frame(torch, scene, dev, mode, (1.0, 0.5, 0.25))
    gpu = harness.GpuDenoiser(den, W, H)
    outs = []
    for f in range(4):
        gpu.set_inputs(fr)
        gpu.denoise(harness.make_common_settings(fr, W, H, f))
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in gpu.outputs().items()})
    for sig in ("DIFF", "SPEC"):
        exp = fr["IN_%s_RADIANCE_HITDIST" % sig].float()
        got = outs[-1]["OUT_%s_RADIANCE_HITDIST" % sig].float()
        err = (got[..., :3] - exp[..., :3]).abs().max().item()
        assert err <= 2e-3, (denoiser, sig, err)          # radiance (REBLUR: YCoCg) unchanged up to FP16 rounding
        assert torch.isfinite(got).all()
    gpu.destroy()


def test_reblur_is_deterministic_at_4k():
    import torch
    from raytracingdenoiser_b200 import harness, nrd, scene
    den = nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR
    sc = scene.Scene(W, H, device="cuda:0")
    frames = [sc.frame(f) for f in range(3)]
    results = []
    for run in range(2):
        gpu = harness.GpuDenoiser(den, W, H)
        for f, fr in enumerate(frames):
            gpu.set_inputs(fr)
            gpu.denoise(harness.make_common_settings(fr, W, H, f))
        torch.cuda.synchronize()
        results.append({k: v.clone() for k, v in gpu.outputs().items()})
        gpu.destroy()
    for k in results[0]:
        assert torch.equal(results[0][k].view(torch.int16), results[1][k].view(torch.int16)), k
        # and something was denoised: the output is not the input
        assert not torch.equal(results[0][k].view(torch.int16), frames[-1]["IN" + k[3:]].view(torch.int16))


@pytest.mark.parametrize("penumbra,expected", [(65504.0, 255), (0.0, 0)])
def test_sigma_uniform_visibility_passes_through_at_4k(penumbra, expected):
    """IN_PENUMBRA = 65504 (NRD_FP16_MAX) means fully lit, 0 fully shadowed (NRD.hlsli SIGMA front end)."""
    import torch
    from raytracingdenoiser_b200 import harness, nrd, scene
    dev = torch.device("cuda", 0)
    fr = _flat_frame(torch, scene, dev, "reblur", (1.0, 1.0, 1.0))
    fr["IN_PENUMBRA"] = torch.full((H, W), penumbra, dtype=torch.float16, device=dev)
    gpu = harness.GpuDenoiser(nrd.Denoiser.SIGMA_SHADOW, W, H)
    for f in range(3):
        gpu.set_inputs(fr)
        gpu.denoise(harness.make_common_settings(fr, W, H, f))
    torch.cuda.synchronize()
    out = gpu.outputs()["OUT_SHADOW_TRANSLUCENCY"]
    assert int(out.min()) == expected and int(out.max()) == expected, (int(out.min()), int(out.max()))
    gpu.destroy()
