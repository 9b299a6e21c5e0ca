"""Checkerboarded inputs (checkerboardMode of ReblurSettings / RelaxSettings): each signal arrives at half rate, packed into the left
half of its texture; the pre-pass resolves it (REBLUR_PrePass.hlsli:43-100, RELAX_PrePass.hlsli:28-110), temporal accumulation
slows down on resolved pixels (REBLUR_TemporalAccumulation.hlsli:731-735, :880, :915; RELAX_TemporalAccumulation.hlsli:597-606,
:854-887), the split-screen passes stretch it (*_SplitScreen.hlsli).  Hit-distance reconstruction is switched off by the host code
in this mode (Reblur.cpp, Relax.cpp)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dump(name, obj):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", name), "w") as f:
        json.dump(obj, f, indent=1, default=str)


def _modes(mode):
    return (0, 1) if mode == "BLACK" else (1, 0)  # (diffuse, specular) parity with data: Source/Reblur.cpp


def _frame_fn(mode):
    from raytracingdenoiser_b200 import scene
    d, s = _modes(mode)
    return lambda fr, f: scene.checkerboard_frame(fr, f, d, s)


@pytest.mark.parametrize("denoiser_name,mode,prepass,perf", [
    ("REBLUR_DIFFUSE_SPECULAR", "BLACK", True, False),
    ("REBLUR_DIFFUSE_SPECULAR", "WHITE", False, False),   # radii 0: the pre-pass only resolves
    ("REBLUR_DIFFUSE_SPECULAR", "WHITE", True, True),     # performance-mode permutations
    ("REBLUR_DIFFUSE", "WHITE", True, False),
    ("REBLUR_SPECULAR", "BLACK", True, False),
])
def test_reblur_checkerboard_per_pass(denoiser_name, mode, prepass, perf):
    import parity
    from raytracingdenoiser_b200 import nrd
    kw = {} if prepass else dict(diffusePrepassBlurRadius=0.0, specularPrepassBlurRadius=0.0)
    s = nrd.ReblurSettings(checkerboardMode=int(getattr(nrd.CheckerboardMode, mode)), enablePerformanceMode=perf, enableAntiFirefly=True, **kw)
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), 252, 142, settings=s, frame_fn=_frame_fn(mode))
    report = sbs.run_per_pass(4)
    assert any("PrePass" in r["shader"] for r in report)
    _dump("parity_checkerboard_%s_%s_%d%d.json" % (denoiser_name, mode, prepass, perf), report)
    assert not sbs.failures(), sbs.describe_failures()


def test_reblur_checkerboard_split_screen():
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.ReblurSettings(checkerboardMode=int(nrd.CheckerboardMode.BLACK))
    sbs = parity.SideBySide(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, 252, 142, settings=s, frame_fn=_frame_fn("BLACK"), common={"splitScreen": 0.4})
    report = sbs.run_per_pass(2)
    assert any("SplitScreen" in r["shader"] for r in report)
    assert not sbs.failures(), sbs.describe_failures()


def test_reblur_checkerboard_strips_bit_identical_to_full_frame():
    import test_gpu_strips as tgs
    from raytracingdenoiser_b200 import nrd
    s = nrd.ReblurSettings(checkerboardMode=int(nrd.CheckerboardMode.BLACK))
    tgs._run("REBLUR_DIFFUSE_SPECULAR", 320, 192, 2, 3, 32, True, settings=s, frame_fn=_frame_fn("BLACK"))


@pytest.mark.parametrize("denoiser_name,mode,prepass", [
    ("RELAX_DIFFUSE_SPECULAR", "BLACK", True),
    ("RELAX_DIFFUSE_SPECULAR", "WHITE", False),   # radii 0: the pre-pass only resolves
    ("RELAX_DIFFUSE", "WHITE", True),
    ("RELAX_SPECULAR", "BLACK", True),
])
def test_relax_checkerboard_per_pass(denoiser_name, mode, prepass):
    import parity
    from raytracingdenoiser_b200 import nrd
    kw = {} if prepass else dict(diffusePrepassBlurRadius=0.0, specularPrepassBlurRadius=0.0)
    s = nrd.RelaxSettings(checkerboardMode=int(getattr(nrd.CheckerboardMode, mode)), enableAntiFirefly=True, **kw)
    sbs = parity.SideBySide(getattr(nrd.Denoiser, denoiser_name), 252, 142, settings=s, frame_fn=_frame_fn(mode))
    report = sbs.run_per_pass(4)
    assert any("PrePass" in r["shader"] for r in report)
    _dump("parity_checkerboard_%s_%s_%d.json" % (denoiser_name, mode, prepass), report)
    assert not sbs.failures(), sbs.describe_failures()


def test_relax_checkerboard_split_screen():
    import parity
    from raytracingdenoiser_b200 import nrd
    s = nrd.RelaxSettings(checkerboardMode=int(nrd.CheckerboardMode.WHITE))
    sbs = parity.SideBySide(nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, 252, 142, settings=s, frame_fn=_frame_fn("WHITE"), common={"splitScreen": 0.4})
    report = sbs.run_per_pass(2)
    assert any("SplitScreen" in r["shader"] for r in report)
    assert not sbs.failures(), sbs.describe_failures()


def test_relax_checkerboard_strips_bit_identical_to_full_frame():
    import test_gpu_strips as tgs
    from raytracingdenoiser_b200 import nrd
    s = nrd.RelaxSettings(checkerboardMode=int(nrd.CheckerboardMode.WHITE))
    tgs._run("RELAX_DIFFUSE_SPECULAR", 320, 180, 2, 3, 16, True, settings=s, frame_fn=_frame_fn("WHITE"))
