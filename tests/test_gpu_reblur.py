"""GPU parity of the REBLUR kernels against the oracle, through the C-ABI (nrdCudaExecuteDispatch / nrdCudaDenoise)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def _dump(name, report):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, name), "w") as f:
        json.dump(report, f, indent=1)


@pytest.mark.parametrize("denoiser_name,width,height,frames", [
    ("REBLUR_DIFFUSE_SPECULAR", 250, 141, 6),   # ragged size: partial tiles and groups on both axes
    ("REBLUR_DIFFUSE", 256, 144, 4),
    ("REBLUR_SPECULAR", 256, 144, 4),
])
def test_reblur_per_pass_parity(denoiser_name, width, height, frames):
    import parity
    from raytracingdenoiser_b200 import nrd
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), width, height)
    report = sbs.run_per_pass(frames)
    _dump("parity_%s.json" % denoiser_name, report)
    assert not sbs.failures(), sbs.describe_failures()


def test_reblur_spatial_only_config():
    """BASELINE config 2: REBLUR_DIFFUSE with temporal accumulation off (Blur + PostBlur_NoTemporalStabilization)."""
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.ReblurSettings(maxAccumulatedFrameNum=0, maxFastAccumulatedFrameNum=0, maxStabilizedFrameNum=0, historyFixFrameNum=0, diffusePrepassBlurRadius=0.0)
    sbs = parity.SideBySide(nrd.Denoiser.REBLUR_DIFFUSE, 480, 270, settings=s)
    report = sbs.run_per_pass(2)
    names = {r["shader"] for r in report}
    assert "REBLUR_Diffuse_PostBlur_NoTemporalStabilization.cs" in names and "REBLUR_Diffuse_PrePass.cs" not in names
    _dump("parity_spatial_only.json", report)
    assert not sbs.failures(), sbs.describe_failures()


def test_reblur_sequence_parity():
    """Statistical gate: 12 independent frames end to end; >= 99 % of texels within tolerance, PSNR >= 60 dB."""
    import parity
    from raytracingdenoiser_b200 import nrd
    res = parity.run_sequence(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, 320, 180, 12)
    _dump("sequence_reblur.json", res)
    for name, (frac, psnr) in res.items():
        assert frac >= 0.99 and psnr >= 60.0, (name, frac, psnr)


@pytest.mark.parametrize("denoiser_name,mode,width,height", [
    ("REBLUR_DIFFUSE_SPECULAR", "AREA_3X3", 250, 141),
    ("REBLUR_DIFFUSE_SPECULAR", "AREA_5X5", 640, 360),
    ("REBLUR_DIFFUSE", "AREA_5X5", 256, 144),
    ("REBLUR_SPECULAR", "AREA_3X3", 256, 144),
])
def test_reblur_hit_distance_reconstruction_per_pass(denoiser_name, mode, width, height):
    """ReblurSettings::hitDistanceReconstructionMode: the extra 3x3 / 5x5 pass (shared-memory tile staged by TMA) and the chain behind it."""
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.ReblurSettings(hitDistanceReconstructionMode=int(getattr(nrd.HitDistanceReconstructionMode, mode)))
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), width, height, settings=s)
    report = sbs.run_per_pass(3)
    assert any("HitDistReconstruction" in r["shader"] for r in report)
    _dump("parity_hitdist_%s_%s.json" % (denoiser_name, mode), report)
    assert not sbs.failures(), sbs.describe_failures()


@pytest.mark.parametrize("denoiser_name", ["REBLUR_DIFFUSE_SPECULAR", "REBLUR_SPECULAR"])
def test_reblur_optional_inputs_per_pass(denoiser_name):
    """CommonSettings::isHistoryConfidenceAvailable / isDisocclusionThresholdMixAvailable: temporal accumulation reads IN_DIFF_CONFIDENCE /
    IN_SPEC_CONFIDENCE / IN_DISOCCLUSION_THRESHOLD_MIX (REBLUR_TemporalAccumulation.hlsli:220-221, :327-328, :830-831)."""
    import parity
    from raytracingdenoiser_b200 import nrd
    common = {"isHistoryConfidenceAvailable": True, "isDisocclusionThresholdMixAvailable": True}
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), 250, 141, common=common)
    report = sbs.run_per_pass(4)
    bound = {res for r in report for res in [r["resource"]]}
    assert any("TemporalAccumulation" in r["shader"] for r in report), bound
    _dump("parity_optional_inputs_%s.json" % denoiser_name, report)
    assert not sbs.failures(), sbs.describe_failures()


@pytest.mark.parametrize("denoiser_name,hitdist", [("REBLUR_DIFFUSE_SPECULAR", "OFF"), ("REBLUR_DIFFUSE_SPECULAR", "AREA_3X3"), ("REBLUR_DIFFUSE", "OFF"), ("REBLUR_SPECULAR", "AREA_5X5")])
def test_reblur_performance_mode_per_pass(denoiser_name, hitdist):
    """ReblurSettings::enablePerformanceMode: the REBLUR_Perf_* permutations (6 taps of g_Special6, screen-space sampling for both
    signals, bilinear instead of CatRom history filters, no rank clamp in temporal stabilization, anti-firefly radius 3;
    REBLUR_Config.hlsli:196-238, Source/Reblur.cpp:106-118)."""
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.ReblurSettings(enablePerformanceMode=True, enableAntiFirefly=True, hitDistanceReconstructionMode=int(getattr(nrd.HitDistanceReconstructionMode, hitdist)))
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), 250, 141, settings=s)
    report = sbs.run_per_pass(5)
    shaders = {r["shader"] for r in report if not r["shader"].startswith("Clear_")}
    assert all(x.startswith("REBLUR_Perf_") or x == "REBLUR_ClassifyTiles.cs" for x in shaders), shaders
    _dump("parity_perf_%s_%s.json" % (denoiser_name, hitdist), report)
    assert not sbs.failures(), sbs.describe_failures()


def test_reblur_performance_mode_sequence_parity():
    import parity
    from raytracingdenoiser_b200 import nrd
    res = parity.run_sequence(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, 320, 180, 10, settings=nrd.ReblurSettings(enablePerformanceMode=True))
    _dump("sequence_reblur_perf.json", res)
    for name, (frac, psnr) in res.items():
        assert frac >= 0.99 and psnr >= 60.0, (name, frac, psnr)


def test_reblur_base_color_motion_vector_patch_per_pass():
    """CommonSettings::isBaseColorMetalnessAvailable: temporal stabilization rewrites IN_MV of specular-dominant pixels with the motion of
    their reflection (REBLUR_TemporalStabilization.hlsli:250-285; BRDF helpers restated from MathLib, oracle/mathlib.h)."""
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.ReblurSettings()
    s.specularProbabilityThresholdsForMvModification[0], s.specularProbabilityThresholdsForMvModification[1] = 0.2, 0.6
    sbs = parity.SideBySide(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, 250, 141, settings=s, common={"isBaseColorMetalnessAvailable": True})
    report = sbs.run_per_pass(4)
    mv = [r for r in report if "TemporalStabilization" in r["shader"] and r["resource"].startswith("IN_MV")]
    assert mv, "temporal stabilization must bind IN_MV as an output"
    _dump("parity_basecolor_mv_patch.json", report)
    assert not sbs.failures(), sbs.describe_failures()
