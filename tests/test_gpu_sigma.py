"""GPU parity of the SIGMA_SHADOW kernels against the oracle (BASELINE config 1 is the 1280x720 single frame)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def _dump(name, report):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, name), "w") as f:
        json.dump(report, f, indent=1)


@pytest.mark.parametrize("width,height,frames", [(1280, 720, 1), (250, 141, 5)])
def test_sigma_per_pass_parity(width, height, frames):
    import parity
    from raytracingdenoiser_b200 import nrd
    sbs = parity.SideBySide(nrd.Denoiser.SIGMA_SHADOW, width, height)
    report = sbs.run_per_pass(frames)
    _dump("parity_SIGMA_%dx%d.json" % (width, height), report)
    assert not sbs.failures(), sbs.describe_failures()


@pytest.mark.parametrize("width,height,frames", [(640, 360, 3), (250, 141, 5)])
def test_sigma_translucency_per_pass_parity(width, height, frames):
    """SIGMA_SHADOW_TRANSLUCENCY: float4 signal in RGBA8 textures, IN_TRANSLUCENCY feeds ClassifyTiles and the first blur."""
    import parity
    from raytracingdenoiser_b200 import nrd
    sbs = parity.SideBySide(nrd.Denoiser.SIGMA_SHADOW_TRANSLUCENCY, width, height)
    report = sbs.run_per_pass(frames)
    assert {"SIGMA_ShadowTranslucency_ClassifyTiles.cs", "SIGMA_ShadowTranslucency_Blur.cs", "SIGMA_ShadowTranslucency_PostBlur.cs",
            "SIGMA_ShadowTranslucency_TemporalStabilization.cs", "SIGMA_Copy.cs"} <= {r["shader"] for r in report}
    _dump("parity_SIGMA_TRANSLUCENCY_%dx%d.json" % (width, height), report)
    assert not sbs.failures(), sbs.describe_failures()


def test_sigma_translucency_sequence_parity():
    import parity
    from raytracingdenoiser_b200 import nrd
    res = parity.run_sequence(nrd.Denoiser.SIGMA_SHADOW_TRANSLUCENCY, 320, 180, 10)
    _dump("sequence_sigma_translucency.json", res)
    for name, (frac, psnr) in res.items():
        assert frac >= 0.99 and psnr >= 45.0, (name, frac, psnr)   # RGBA8 output: 1 LSB = 1/255


def test_sigma_sequence_parity():
    import parity
    from raytracingdenoiser_b200 import nrd
    res = parity.run_sequence(nrd.Denoiser.SIGMA_SHADOW, 320, 180, 10)
    _dump("sequence_sigma.json", res)
    for name, (frac, psnr) in res.items():
        assert frac >= 0.99 and psnr >= 45.0, (name, frac, psnr)   # R8 output: 1 LSB = 1/255


def test_sigma_against_golden_vector():
    """Committed fixture (oracle output, tests/golden/make_golden.py): the GPU path alone must reproduce it."""
    import numpy as np
    import torch
    import make_golden_path  # noqa: F401
    import make_golden
    from raytracingdenoiser_b200 import harness, nrd, scene
    den, w, h, frames = make_golden.CASES["sigma_shadow_96x64"]
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sigma_shadow_96x64.npz"))
    gpu = harness.GpuDenoiser(getattr(nrd.Denoiser, den), w, h)
    sc = scene.Scene(w, h)
    for f in range(frames):
        fr = sc.frame(f)
        gpu.set_inputs(fr)
        gpu.denoise(harness.make_common_settings(fr, w, h, f))
    torch.cuda.synchronize()
    got = gpu.outputs()["OUT_SHADOW_TRANSLUCENCY"].cpu().numpy()
    d = np.abs(got.astype(np.int32) - ref["OUT_SHADOW_TRANSLUCENCY"].astype(np.int32))
    assert (d <= 1).mean() >= 0.99, (d <= 1).mean()
