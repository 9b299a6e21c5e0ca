"""GPU parity of the auxiliary passes against the oracle: split-screen passes of the three families (CommonSettings::splitScreen)
and the REFERENCE denoiser."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def _dump(name, report):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, name), "w") as f:
        json.dump(report, f, indent=1)


@pytest.mark.parametrize("denoiser_name,split", [
    ("REBLUR_DIFFUSE_SPECULAR", 0.4), ("REBLUR_SPECULAR", 1.0), ("RELAX_DIFFUSE_SPECULAR", 0.4), ("RELAX_DIFFUSE", 0.6),
    ("SIGMA_SHADOW", 0.5), ("SIGMA_SHADOW_TRANSLUCENCY", 0.5),
])
def test_split_screen_per_pass(denoiser_name, split):
    """splitScreen in (0, 1): the whole chain plus the split-screen pass; splitScreen >= 1: the split-screen pass alone
    (Source/Reblur.cpp:120-128, Relax.cpp:196-204, Sigma.cpp:38-46)."""
    import parity
    from raytracingdenoiser_b200 import nrd
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), 250, 141, common={"splitScreen": split})
    report = sbs.run_per_pass(3)
    shaders = [r["shader"] for r in report]
    assert any("SplitScreen" in s for s in shaders)
    if split >= 1.0:
        assert all("SplitScreen" in s or s.startswith("Clear_") for s in shaders), shaders
    _dump("parity_split_%s.json" % denoiser_name, report)
    assert not sbs.failures(), sbs.describe_failures()


def test_reference_denoiser_per_pass_and_sequence():
    """Denoiser::REFERENCE: running average of IN_SIGNAL in an RGBA32F history (Source/Denoisers/Reference.hpp)."""
    import numpy as np
    import parity
    from raytracingdenoiser_b200 import nrd
    sbs = parity.SideBySide(nrd.Denoiser.REFERENCE, 250, 141)
    report = sbs.run_per_pass(4)
    assert {r["shader"] for r in report} >= {"REFERENCE_TemporalAccumulation.cs", "REFERENCE_Copy.cs"}
    _dump("parity_REFERENCE.json", report)
    assert not sbs.failures(), sbs.describe_failures()
    res = parity.run_sequence(nrd.Denoiser.REFERENCE, 320, 180, 6)
    for name, (frac, psnr) in res.items():
        assert frac >= 0.999 and psnr >= 60.0, (name, frac, psnr)


def _rects(f):
    """Dynamic resolution: the rect changes size after two frames and sits at a non-zero origin of the G-buffer inputs."""
    return (16, 8, 250, 141) if f < 2 else (24, 16, 280, 158)


@pytest.mark.parametrize("denoiser_name", ["REBLUR_DIFFUSE_SPECULAR", "RELAX_DIFFUSE_SPECULAR", "SIGMA_SHADOW"])
def test_dynamic_resolution_per_pass(denoiser_name):
    """rectSize < resourceSize, rectOrigin != 0, rectSizePrev != rectSize (Source/InstanceImpl.cpp:834-856, Common.hlsli:200-222):
    textures are 320x192, the passes run over a 250x141 rect and then over a 280x158 one."""
    import parity
    from raytracingdenoiser_b200 import nrd
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), 320, 192, noise_floor=denoiser_name.startswith("RELAX"))
    report = sbs.run_per_pass(4, rect_fn=_rects)
    _dump("parity_dynres_%s.json" % denoiser_name, report)
    assert not sbs.failures(), sbs.describe_failures()
