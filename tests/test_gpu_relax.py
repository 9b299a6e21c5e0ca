"""GPU parity of the RELAX_DIFFUSE_SPECULAR kernels against the oracle (oracle/relax.cpp)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def _dump(name, report):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, name), "w") as f:
        json.dump(report, f, indent=1)


@pytest.mark.parametrize("width,height,frames", [(640, 360, 4), (250, 141, 6)])
def test_relax_per_pass_parity(width, height, frames):
    import parity
    from raytracingdenoiser_b200 import nrd
    sbs = parity.SideBySide(nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, width, height)
    report = sbs.run_per_pass(frames)
    _dump("parity_RELAX_%dx%d.json" % (width, height), report)
    assert not sbs.failures(), sbs.describe_failures()


def test_relax_settings_variants_per_pass():
    """Non-default settings: 3 A-trous iterations (odd/even binding variants), anti-lag off, roughness edge stopping off."""
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.RelaxSettings()
    s.atrousIterationNum = 3
    s.enableRoughnessEdgeStopping = False
    s.historyFixFrameNum = 2
    s.diffusePrepassBlurRadius = 0.0
    s.enableAntiFirefly = False
    sbs = parity.SideBySide(nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, 320, 180, settings=s)
    report = sbs.run_per_pass(4)
    _dump("parity_RELAX_variant.json", report)
    assert not sbs.failures(), sbs.describe_failures()


def test_relax_sequence_parity():
    import parity
    from raytracingdenoiser_b200 import nrd
    res = parity.run_sequence(nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, 320, 180, 12)
    _dump("sequence_relax.json", res)
    for name, (frac, psnr) in res.items():
        assert frac >= 0.99 and psnr >= 60.0, (name, frac, psnr)


def test_relax_against_golden_vector():
    """Committed fixture (oracle output, tests/golden/make_golden.py): the GPU path alone must reproduce it."""
    import numpy as np
    import torch
    import make_golden_path  # noqa: F401
    import make_golden
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    den, w, h, frames = make_golden.CASES["relax_diffuse_specular_96x64"]
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "relax_diffuse_specular_96x64.npz"))
    d = getattr(nrd.Denoiser, den)
    gpu = harness.GpuDenoiser(d, w, h)
    sc = scene.Scene(w, h)
    for f in range(frames):
        fr = sc.frame(f, harness.radiance_mode(d))
        gpu.set_inputs(fr)
        gpu.denoise(harness.make_common_settings(fr, w, h, f))
    torch.cuda.synchronize()
    for name, t in gpu.outputs().items():
        got = t.cpu().numpy().view(ref[name].dtype).reshape(ref[name].shape)
        frac, _ = orr.compare(ref[name], got, nrd.Format.RGBA16_SFLOAT, 1e-3, 1e-4)
        assert frac >= 0.99, (name, frac)


def test_relax_optional_inputs_per_pass():
    """History confidence (temporal accumulation caps + confidence-driven relaxation of the A-trous edge stopping,
    RELAX_TemporalAccumulation.hlsli:585-632, RELAX_Atrous.hlsli:55-106) and the disocclusion-threshold mix (:483-484)."""
    import parity
    from raytracingdenoiser_b200 import nrd
    common = {"isHistoryConfidenceAvailable": True, "isDisocclusionThresholdMixAvailable": True}
    s = nrd.RelaxSettings()
    s.confidenceDrivenRelaxationMultiplier = 0.7
    s.confidenceDrivenLuminanceEdgeStoppingRelaxation = 0.4
    s.confidenceDrivenNormalEdgeStoppingRelaxation = 0.5
    sbs = parity.SideBySide(nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, 250, 141, settings=s, common=common, noise_floor=True)
    report = sbs.run_per_pass(4)
    _dump("parity_RELAX_optional_inputs.json", report)
    assert not sbs.failures(), sbs.describe_failures()


def test_relax_anti_firefly_per_pass():
    """RelaxSettings::enableAntiFirefly: the Copy pass and the 3x3 rank-conditioned rank-selection (RELAX_Copy.hlsli, RELAX_AntiFirefly.hlsli)."""
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.RelaxSettings()
    s.enableAntiFirefly = True
    sbs = parity.SideBySide(nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, 250, 141, settings=s, noise_floor=True)
    report = sbs.run_per_pass(4)
    names = {r["shader"] for r in report}
    assert "RELAX_DiffuseSpecular_AntiFirefly.cs" in names and "RELAX_DiffuseSpecular_Copy.cs" in names
    _dump("parity_RELAX_antifirefly.json", report)
    assert not sbs.failures(), sbs.describe_failures()


@pytest.mark.parametrize("denoiser_name", ["RELAX_DIFFUSE", "RELAX_SPECULAR"])
def test_relax_one_signal_per_pass_parity(denoiser_name):
    """RELAX_DIFFUSE / RELAX_SPECULAR: the same kernels compiled without the other signal (Source/Denoisers/Relax_Diffuse.hpp,
    Relax_Specular.hpp), anti-firefly on so that Copy / AntiFirefly are covered too."""
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.RelaxSettings()
    s.enableAntiFirefly = True
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), 250, 141, settings=s, noise_floor=True)
    report = sbs.run_per_pass(4)
    _dump("parity_%s.json" % denoiser_name, report)
    assert not sbs.failures(), sbs.describe_failures()


@pytest.mark.parametrize("denoiser_name", ["RELAX_DIFFUSE", "RELAX_SPECULAR"])
def test_relax_one_signal_sequence_parity(denoiser_name):
    import parity
    from raytracingdenoiser_b200 import nrd
    res = parity.run_sequence(getattr(nrd.Denoiser, denoiser_name), 320, 180, 8)
    _dump("sequence_%s.json" % denoiser_name.lower(), res)
    for name, (frac, psnr) in res.items():
        assert frac >= 0.99 and psnr >= 60.0, (name, frac, psnr)


@pytest.mark.parametrize("denoiser_name,mode", [("RELAX_DIFFUSE_SPECULAR", "AREA_3X3"), ("RELAX_DIFFUSE_SPECULAR", "AREA_5X5"), ("RELAX_SPECULAR", "AREA_5X5"), ("RELAX_DIFFUSE", "AREA_3X3")])
def test_relax_hit_distance_reconstruction_per_pass(denoiser_name, mode):
    """RelaxSettings::hitDistanceReconstructionMode: the extra 3x3 / 5x5 pass (RELAX_HitDistReconstruction.hlsli) and the chain behind it."""
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.RelaxSettings()
    s.hitDistanceReconstructionMode = int(getattr(nrd.HitDistanceReconstructionMode, mode))
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), 250, 141, settings=s, noise_floor=True)
    report = sbs.run_per_pass(3)
    assert any("HitDistReconstruction" in r["shader"] for r in report)
    _dump("parity_hitdist_%s_%s.json" % (denoiser_name, mode), report)
    assert not sbs.failures(), sbs.describe_failures()
