"""Oracle tests (no GPU): the CPU restatement against its committed golden vectors and against properties the filter
chain must have (sky untouched, energy preserved, variance reduced, history length grows, determinism)."""
import os
import sys

import numpy as np
import pytest

import oracle_runner as orr
from raytracingdenoiser_b200 import harness, nrd, scene

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)


def _cases():
    import make_golden
    return make_golden.CASES


@pytest.mark.parametrize("name", ["reblur_diffuse_specular_96x64", "sigma_shadow_96x64", "relax_diffuse_specular_96x64"])
def test_oracle_reproduces_golden_vectors(name):
    import make_golden
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden vector not generated yet")
    den, w, h, frames = make_golden.CASES[name]
    out = make_golden.run_case(den, w, h, frames)
    ref = np.load(path)
    for k in ref.files:
        # the oracle is plain IEEE arithmetic + libm; allow 1 ulp of the storage format for libm differences between hosts
        frac, worst = orr.compare(ref[k], out[k], orr.CpuDenoiser(getattr(nrd.Denoiser, den), 16, 16).user_fmt[k], rel=1e-3, abs_tol=1e-4)
        assert frac >= 0.9995, (k, frac, worst)


def test_reblur_chain_properties():
    W, H, N = 160, 90, 8
    sc = scene.Scene(W, H)
    cpu = orr.CpuDenoiser(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H)
    for f in range(N):
        fr = sc.frame(f)
        cpu.set_inputs(fr)
        cpu.denoise(harness.make_common_settings(fr, W, H, f))
        if f == 0:
            cpu.set_inputs(fr)
    z = cpu.user["IN_VIEWZ"]
    sky = z > 5e5
    out = cpu.user["OUT_DIFF_RADIANCE_HITDIST"].astype(np.float32)
    noisy = cpu.user["IN_DIFF_RADIANCE_HITDIST"].astype(np.float32)
    assert sky.any() and (~sky).any()
    assert not np.isnan(out).any()
    # luminance (Y of YCoCg) is preserved on average and its pixel-to-pixel variation is reduced
    m_in, m_out = noisy[~sky][:, 0].mean(), out[~sky][:, 0].mean()
    assert abs(m_in - m_out) / m_in < 0.08, (m_in, m_out)
    hf = lambda a: np.abs(np.diff(a[..., 0], axis=1))[~sky[:, 1:] & ~sky[:, :-1]].mean()
    assert hf(out) < 0.35 * hf(noisy)
    # history length: internal data (6 bits per signal) grows by about one frame per frame on static-ish surfaces
    internal = cpu.permanent[2]
    diff_frames = (internal & 63)[~sky]
    assert np.median(diff_frames) >= N - 2
    # PREV_VIEWZ (written by Blur) is the current viewZ wherever the tile is not all-sky
    prev_z = cpu.permanent[0]
    assert np.array_equal(prev_z[~sky], z[~sky])


def test_oracle_is_deterministic():
    W, H = 64, 48
    sc = scene.Scene(W, H)
    outs = []
    for _ in range(2):
        cpu = orr.CpuDenoiser(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H)
        for f in range(3):
            fr = sc.frame(f)
            cpu.set_inputs(fr)
            cpu.denoise(harness.make_common_settings(fr, W, H, f))
            if f == 0:
                cpu.set_inputs(fr)
        outs.append(cpu.user["OUT_SPEC_RADIANCE_HITDIST"].copy())
    assert np.array_equal(outs[0].view(np.uint16), outs[1].view(np.uint16))


def test_scene_encodings():
    sc = scene.Scene(64, 48)
    fr = sc.frame(2)
    nr = fr["IN_NORMAL_ROUGHNESS"].numpy().view(np.uint32)
    z = fr["IN_VIEWZ"].numpy()
    assert fr["IN_MV"].dtype.__str__() == "torch.float16" and fr["IN_MV"].shape == (48, 64, 4)
    assert ((nr >> 30) <= 3).all() and (z > 0).all()
    # decode the oct-packed normals of surface pixels: unit length after normalisation, facing the camera half-space mostly
    p = np.stack([(nr & 1023) / 1023.0, ((nr >> 10) & 1023) / 1023.0], -1) * 2 - 1
    n = np.concatenate([p, 1 - np.abs(p).sum(-1, keepdims=True)], -1)
    t = np.clip(-n[..., 2], 0, 1)
    n[..., 0] -= t * np.where(n[..., 0] >= 0, 1, -1)
    n[..., 1] -= t * np.where(n[..., 1] >= 0, 1, -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    ground = (z < 5e5) & (np.abs(n[..., 1] - 1.0) < 0.01)
    assert ground.mean() > 0.1


def _flat_frame(w, h, radiance_mode, value=(1.0, 0.5, 0.25)):
    """Wall facing a static camera at z = 10, zero motion, constant radiance and hit distance (same as tests/test_gpu_properties.py)."""
    import torch
    z = torch.full((h, w), 10.0)
    n = torch.zeros((h, w, 3))
    n[..., 2] = -1.0
    rough = torch.full((h, w), 0.5)
    rad = torch.tensor(value).expand(h, w, 3)
    hit = torch.full((h, w), 3.0)
    sky = torch.zeros((h, w), dtype=torch.bool)
    fr = {"IN_VIEWZ": z, "IN_NORMAL_ROUGHNESS": scene.pack_normal_roughness(n, rough, torch.zeros((h, w))), "IN_MV": torch.zeros((h, w, 4), dtype=torch.float16),
          "IN_PENUMBRA": torch.full((h, w), 65504.0, dtype=torch.float16)}
    if radiance_mode == "reblur":
        fr["IN_DIFF_RADIANCE_HITDIST"] = scene.pack_reblur(rad, hit, z, torch.ones_like(rough), sky)
        fr["IN_SPEC_RADIANCE_HITDIST"] = scene.pack_reblur(rad, hit, z, rough, sky)
    else:
        fr["IN_DIFF_RADIANCE_HITDIST"] = scene.pack_relax(rad, hit, sky)
        fr["IN_SPEC_RADIANCE_HITDIST"] = scene.pack_relax(rad, hit, sky)
    view = np.eye(4, dtype=np.float32)
    fr.update({"viewToClip": scene.perspective_lh(60.0, w / float(h)), "worldToView": view, "worldToViewPrev": view})
    return fr


def _run_oracle(den, w, h, frames_fn, n):
    cpu = orr.CpuDenoiser(den, w, h)
    for f in range(n):
        fr = frames_fn(f)
        cpu.set_inputs(fr)
        cpu.denoise(harness.make_common_settings(fr, w, h, f))
        if f == 0:
            cpu.set_inputs(fr)
    return cpu, fr


@pytest.mark.parametrize("denoiser", ["REBLUR_DIFFUSE_SPECULAR", "RELAX_DIFFUSE_SPECULAR"])
def test_constant_signal_is_a_fixed_point_of_the_oracle(denoiser):
    """Every weight of every pass is a convex combination: a constant signal on a flat surface must come out unchanged."""
    den = getattr(nrd.Denoiser, denoiser)
    w, h = 96, 64
    fr0 = _flat_frame(w, h, harness.radiance_mode(den))
    cpu, fr = _run_oracle(den, w, h, lambda f: fr0, 4)
    for sig in ("DIFF", "SPEC"):
        exp = fr["IN_%s_RADIANCE_HITDIST" % sig].float().numpy()
        got = cpu.user["OUT_%s_RADIANCE_HITDIST" % sig].astype(np.float32)
        assert np.isfinite(got).all()
        assert np.abs(got[..., :3] - exp[..., :3]).max() <= 1e-3, (denoiser, sig)


@pytest.mark.parametrize("penumbra,expected", [(65504.0, 255), (0.0, 0)])
def test_sigma_oracle_passes_uniform_visibility_through(penumbra, expected):
    import torch
    w, h = 96, 64
    fr0 = _flat_frame(w, h, "reblur")
    fr0["IN_PENUMBRA"] = torch.full((h, w), penumbra, dtype=torch.float16)
    cpu, _ = _run_oracle(nrd.Denoiser.SIGMA_SHADOW, w, h, lambda f: fr0, 3)
    out = cpu.user["OUT_SHADOW_TRANSLUCENCY"]
    assert int(out.min()) == expected and int(out.max()) == expected


def test_oracle_result_does_not_depend_on_the_thread_count():
    """The OpenMP loops write disjoint pixels: one thread and all threads must agree bit for bit (no races, no reductions)."""
    lib = orr.oracle_lib()
    threads = lib.oracle_num_threads()
    w, h = 80, 48
    outs = []
    try:
        for n in (1, max(threads, 2)):
            lib.oracle_set_num_threads(n)
            sc = scene.Scene(w, h)
            cpu, _ = _run_oracle(nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, w, h, lambda f: sc.frame(f, "relax"), 3)
            outs.append({k: v.copy() for k, v in cpu.user.items() if k.startswith("OUT_")})
    finally:
        lib.oracle_set_num_threads(threads)
    for k in outs[0]:
        assert np.array_equal(outs[0][k].view(np.uint16), outs[1][k].view(np.uint16)), k


def test_relax_chain_properties():
    W, H, N = 160, 90, 6
    sc = scene.Scene(W, H)
    cpu, fr = _run_oracle(nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, W, H, lambda f: sc.frame(f, "relax"), N)
    z = cpu.user["IN_VIEWZ"]
    sky = z > 5e5
    out = cpu.user["OUT_DIFF_RADIANCE_HITDIST"].astype(np.float32)
    noisy = cpu.user["IN_DIFF_RADIANCE_HITDIST"].astype(np.float32)
    assert not np.isnan(out).any()
    lum = lambda a: 0.2126 * a[..., 0] + 0.7152 * a[..., 1] + 0.0722 * a[..., 2]
    m_in, m_out = lum(noisy)[~sky].mean(), lum(out)[~sky].mean()
    assert abs(m_in - m_out) / m_in < 0.08, (m_in, m_out)
    hf = lambda a: np.abs(np.diff(lum(a), axis=1))[~sky[:, 1:] & ~sky[:, :-1]].mean()
    assert hf(out) < 0.4 * hf(noisy)
    # PREV_VIEWZ written by the first A-trous pass is the current viewZ everywhere (also on sky pixels)
    assert np.array_equal(cpu.permanent[9], z)


def test_rounding_noise_floor_of_the_reblur_chain():
    """Two IEEE-legal CPU evaluations of the same oracle sources (with / without FMA contraction) drift apart frame by frame:
    REBLUR's temporal feedback is chaotic at the 1e-3 level.  This is the yardstick the long-sequence GPU gate is read against
    (tests/test_gpu_baseline_configs.py::test_config3...)."""
    import numpy as np
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    den, w, h = nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, 256, 144
    a, b = orr.CpuDenoiser(den, w, h), orr.CpuDenoiser(den, w, h, variant="fma")
    sc = scene.Scene(w, h)
    fractions = []
    for f in range(24):
        fr = sc.frame(f)
        cs = harness.make_common_settings(fr, w, h, f)
        for c in (a, b):
            c.set_inputs(fr)
            c.denoise(cs)
            if f == 0:
                c.set_inputs(fr)
        fractions.append(min(orr.compare(a.user[n], b.user[n], a.user_fmt[n])[0] for n in ("OUT_DIFF_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST")))
    assert fractions[0] > 0.999                       # one frame: the two evaluations agree almost everywhere
    assert fractions[-1] < fractions[0] and fractions[-1] < 0.999   # ... and drift apart through the history feedback
    assert fractions[-1] > 0.9 and np.isfinite(a.user["OUT_DIFF_RADIANCE_HITDIST"].astype(np.float32)).all()


def test_subtexel_noise_floor_of_relax_temporal_accumulation():
    """The third build of the oracle moves the uv of every bilinear fetch by one float ulp (oracle/hlsl.h, ORACLE_NUDGE_UV): the
    CatRom history of RELAX's temporal accumulation reacts to it -- per pass, on identical inputs -- because the sub-texel position
    a shader hands to the sampler is only known to ulp(uv) * size texels and the second moment has contrast edges.  The kernels
    merge the CatRom's bilinear taps with exact weights, so this floor (not the plain 99.9 % gate) is what their temporal
    accumulation is read against at 4K (tests/test_gpu_baseline_configs.py::test_config4...)."""
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    den, w, h = nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, 512, 288
    a = orr.CpuDenoiser(den, w, h)
    b = orr.CpuDenoiser(den, w, h, instance=a.instance, variant="uv")
    sc = scene.Scene(w, h)
    worst = {}
    for f in range(5):
        fr = sc.frame(f)
        a.set_inputs(fr)
        cs = harness.make_common_settings(fr, w, h, f)
        a.instance.set_common_settings(cs)
        r, raw, n = a.instance.get_compute_dispatches_raw([0])
        assert r == nrd.Result.SUCCESS
        pipelines = a.instance.get_instance_desc()["pipelines"]
        for i in range(n):
            d = nrd.Dispatch(raw[i], pipelines)
            for _, rtype, index in d.resources:  # same inputs for both builds, pass by pass
                b.resolve(rtype, index)[0][...] = a.resolve(rtype, index)[0]
            a.run_dispatch(d)
            b.run_dispatch(d)
            for dtype_, rtype, index in d.resources:
                if dtype_ != nrd.DescriptorType.STORAGE_TEXTURE:
                    continue
                ref, fmt = a.resolve(rtype, index)
                frac = orr.compare(ref, b.resolve(rtype, index)[0], fmt)[0]
                worst[d.shaderFileName] = min(worst.get(d.shaderFileName, 1.0), frac)
        if f == 0:
            a.set_inputs(fr)
    ta = worst["RELAX_DiffuseSpecular_TemporalAccumulation.cs"]
    assert ta < 1.0, worst               # the pass does react to a one-ulp move of its fetches ...
    assert ta > 0.99, worst              # ... mildly (the effect grows with the frame size: ulp(uv) * size)
    # passes that only point-sample are untouched
    assert worst["RELAX_DiffuseSpecular_HistoryClamping.cs"] == 1.0, worst


def test_optional_inputs_reach_the_oracle():
    """History confidence and the disocclusion-threshold mix are bound and read: switching them on changes the temporal result."""
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    w, h = 96, 64
    sc = scene.Scene(w, h)
    outs = []
    for common in (None, {"isHistoryConfidenceAvailable": True}, {"isDisocclusionThresholdMixAvailable": True}):
        cpu = orr.CpuDenoiser(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h, common=common)
        for f in range(3):
            fr = sc.frame(f)
            cpu.set_inputs(fr)
            cpu.denoise(harness.make_common_settings(fr, w, h, f, common=common))
            if f == 0:
                cpu.set_inputs(fr)
        outs.append(cpu.user["OUT_DIFF_RADIANCE_HITDIST"].copy())
    assert (outs[0] != outs[1]).any() and (outs[0] != outs[2]).any()


def test_relax_anti_firefly_removes_fireflies_in_the_oracle():
    """The synthetic scene plants 50x fireflies on 0.1 % of the pixels: with enableAntiFirefly the brightest output texel drops."""
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    w, h = 96, 64
    sc = scene.Scene(w, h)
    peak = []
    for af in (False, True):
        s = nrd.RelaxSettings()
        s.enableAntiFirefly = af
        cpu = orr.CpuDenoiser(nrd.Denoiser.RELAX_DIFFUSE_SPECULAR, w, h, settings=s)
        names = set()
        for f in range(3):
            fr = sc.frame(f, "relax")
            cpu.set_inputs(fr)
            names |= {d.shaderFileName for d in cpu.denoise(harness.make_common_settings(fr, w, h, f))}
            if f == 0:
                cpu.set_inputs(fr)
        assert ("RELAX_DiffuseSpecular_AntiFirefly.cs" in names) == af
        out = cpu.user["OUT_DIFF_RADIANCE_HITDIST"].view(np.float16).astype(np.float32)[..., :3]
        assert np.isfinite(out).all()
        peak.append(float(out.max()))
    assert peak[1] <= peak[0]


@pytest.mark.parametrize("single,signal", [("RELAX_DIFFUSE", "DIFF"), ("RELAX_SPECULAR", "SPEC")])
def test_relax_one_signal_denoisers_equal_the_two_signal_denoiser_per_signal(single, signal):
    """RELAX_DIFFUSE / RELAX_SPECULAR run the two-signal passes without the other signal's bindings: with equal accumulation caps
    (the defaults) each must reproduce its signal of RELAX_DIFFUSE_SPECULAR bit for bit -- checks the one-signal binding layouts."""
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    w, h = 96, 64
    sc = scene.Scene(w, h)
    s = nrd.RelaxSettings()
    s.enableAntiFirefly = True
    outs = []
    for den in (getattr(nrd.Denoiser, single), nrd.Denoiser.RELAX_DIFFUSE_SPECULAR):
        cpu = orr.CpuDenoiser(den, w, h, settings=s)
        for f in range(4):
            fr = sc.frame(f, "relax")
            cpu.set_inputs(fr)
            cpu.denoise(harness.make_common_settings(fr, w, h, f))
            if f == 0:
                cpu.set_inputs(fr)
        outs.append(cpu.user["OUT_%s_RADIANCE_HITDIST" % signal].copy())
    assert outs[0].any() and (outs[0] == outs[1]).all()


def test_reference_denoiser_accumulates_a_running_average_in_the_oracle():
    """Static camera: after n frames OUT_SIGNAL is the mean of the n inputs (accumSpeed = 1 / (1 + n), Reference.hpp:60-73)."""
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    w, h = 64, 48
    sc = scene.Scene(w, h)
    cpu = orr.CpuDenoiser(nrd.Denoiser.REFERENCE, w, h)
    fr0 = sc.frame(0)
    acc = np.zeros((h, w, 4), dtype=np.float64)
    for f in range(4):
        fr = dict(fr0)  # frame-0 camera for every frame: worldToClip == worldToClipPrev after the first one
        fr["IN_SIGNAL"] = sc.frame(f)["IN_SIGNAL"]
        fr["worldToViewPrev"] = fr["worldToView"]
        cpu.set_inputs(fr)
        cpu.denoise(harness.make_common_settings(fr, w, h, f))
        acc += cpu.user["IN_SIGNAL"].view(np.float16).astype(np.float64)
    out = cpu.user["OUT_SIGNAL"].view(np.float16).astype(np.float64)
    assert np.allclose(out, acc / 4.0, rtol=2e-3, atol=1e-3)


def test_split_screen_shows_the_input_left_of_the_split_in_the_oracle():
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    w, h = 96, 64
    sc = scene.Scene(w, h)
    common = {"splitScreen": 0.5}
    cpu = orr.CpuDenoiser(nrd.Denoiser.REBLUR_DIFFUSE, w, h, common=common)
    names = set()
    for f in range(2):
        fr = sc.frame(f)
        cpu.set_inputs(fr)
        names |= {d.shaderFileName for d in cpu.denoise(harness.make_common_settings(fr, w, h, f, common=common))}
    assert "REBLUR_Diffuse_SplitScreen.cs" in names
    out = cpu.user["OUT_DIFF_RADIANCE_HITDIST"].view(np.float16).astype(np.float32)
    inp = cpu.user["IN_DIFF_RADIANCE_HITDIST"].view(np.float16).astype(np.float32)
    z = np.abs(cpu.user["IN_VIEWZ"]) < 500000.0
    left = slice(0, w // 2 - 1)
    assert np.array_equal(out[:, left][z[:, left]], inp[:, left][z[:, left]])
    assert not np.array_equal(out[:, w // 2 + 1:], inp[:, w // 2 + 1:])


def test_reblur_performance_mode_runs_the_perf_permutations_in_the_oracle():
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    w, h = 96, 64
    sc = scene.Scene(w, h)
    outs = []
    for perf in (False, True):
        cpu = orr.CpuDenoiser(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h, settings=nrd.ReblurSettings(enablePerformanceMode=perf, enableAntiFirefly=True))
        names = set()
        for f in range(4):
            fr = sc.frame(f)
            cpu.set_inputs(fr)
            names |= {d.shaderFileName for d in cpu.denoise(harness.make_common_settings(fr, w, h, f))}
            if f == 0:
                cpu.set_inputs(fr)
        assert any(n.startswith("REBLUR_Perf_") for n in names) == perf
        out = cpu.user["OUT_SPEC_RADIANCE_HITDIST"].view(np.float16).astype(np.float32)
        assert np.isfinite(out).all() and out.any()
        outs.append(out)
    assert not np.array_equal(outs[0], outs[1])


def test_dynamic_resolution_rect_origin_does_not_change_the_result_in_the_oracle():
    """A 96x64 rect of 160x96 textures: the guide inputs (viewZ, normals, motion) sit at rectOrigin of their textures (WithRectOrigin,
    Common.hlsli:200-205), everything else at (0, 0).  Moving the origin -- with garbage around the rect -- must not change a texel."""
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    w, h, RW, RH = 96, 64, 160, 96
    sc = scene.Scene(w, h)
    outs = []
    for ox, oy in ((0, 0), (16, 8), (64, 32)):
        cpu = orr.CpuDenoiser(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, RW, RH)
        for f in range(3):
            fr = sc.frame(f)
            for name in harness.RECT_ORIGIN_INPUTS:
                if name in cpu.user:
                    cpu.user[name][...] = np.random.RandomState(f).randint(0, 200, cpu.user[name].shape).astype(cpu.user[name].dtype)  # garbage around the rect
            cpu.set_inputs(fr, rect_origin=(ox, oy))
            cpu.denoise(harness.make_common_settings(fr, w, h, f, common=dict(resourceSize=(RW, RH), resourceSizePrev=(RW, RH), rectOrigin=(ox, oy))))
        outs.append({k: v[:h, :w].copy() for k, v in cpu.user.items() if k.startswith("OUT_")})
    for k in outs[0]:
        assert outs[0][k].any()
        assert np.array_equal(outs[0][k].view(np.uint16), outs[1][k].view(np.uint16)) and np.array_equal(outs[0][k].view(np.uint16), outs[2][k].view(np.uint16)), k


@pytest.mark.parametrize("family,mode", [("REBLUR", "BLACK"), ("REBLUR", "WHITE"), ("RELAX", "BLACK"), ("RELAX", "WHITE")])
def test_checkerboarded_inputs_are_resolved_by_the_pre_pass_in_the_oracle(family, mode):
    """checkerboardMode of ReblurSettings / RelaxSettings: each signal arrives at half rate, packed into the left half of its texture
    (scene.checkerboard_frame).  With the pre-pass radii at 0 the pre-pass only resolves (REBLUR_PrePass.hlsli:43-100,
    RELAX_PrePass.hlsli:28-110): a pixel that has data comes out unchanged, a pixel without takes the average of its left / right
    neighbours on the same surface."""
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    w, h = 96, 64
    sc = scene.Scene(w, h)
    diff_mode, spec_mode = (0, 1) if mode == "BLACK" else (1, 0)  # host code of both families: BLACK = diffuse on the 0-pixels, specular on the 1-pixels
    Settings = nrd.ReblurSettings if family == "REBLUR" else nrd.RelaxSettings
    settings = Settings(checkerboardMode=int(getattr(nrd.CheckerboardMode, mode)), diffusePrepassBlurRadius=0.0, specularPrepassBlurRadius=0.0)
    den = getattr(nrd.Denoiser, family + "_DIFFUSE_SPECULAR")
    cpu = orr.CpuDenoiser(den, w, h, settings=settings)
    out_slots = {"diff": 5, "spec": 6} if family == "REBLUR" else {"spec": 5, "diff": 6}  # *_PrePass.resources.hlsli
    yy, xx = np.mgrid[0:h, 0:w]
    for f in range(3):
        full = sc.frame(f, harness.radiance_mode(den))
        fr = scene.checkerboard_frame(full, f, diff_mode, spec_mode)
        grabbed = {}

        def grab(i, d, when):
            if when == "after" and d.shaderFileName.endswith("_PrePass.cs"):
                for key, slot in out_slots.items():
                    grabbed[key] = cpu.resolve(d.resources[slot][1], d.resources[slot][2])[0].copy()
        cpu.set_inputs(fr)
        cpu.denoise(harness.make_common_settings(fr, w, h, f), on_dispatch=grab)
        if f == 0:
            cpu.set_inputs(fr)
        assert grabbed, "checkerboarded frames always run the pre-pass"
        z = np.abs(cpu.user["IN_VIEWZ"]) < 500000.0
        for key, name, m in (("diff", "IN_DIFF_RADIANCE_HITDIST", diff_mode), ("spec", "IN_SPEC_RADIANCE_HITDIST", spec_mode)):
            want = full[name].cpu().numpy().view(np.float16).reshape(h, w, 4)
            got = grabbed[key].view(np.float16).reshape(h, w, 4)
            has = (((xx ^ yy ^ f) & 1) == m) & z
            # (the RELAX pre-pass clamps the hit distance to the denoising range, the radiance passes through)
            assert np.array_equal(got[has][:, :3].view(np.uint16), want[has][:, :3].view(np.uint16)), (key, f)
            holes = ~(((xx ^ yy ^ f) & 1) == m) & z
            assert np.isfinite(got[holes].astype(np.float32)).all()
            # interior holes between two data pixels of the same surface are filled with something: not left black everywhere
            assert (got[holes].astype(np.float32)[:, :3].sum(-1) > 0).mean() > 0.5, (key, f)
    out = cpu.user["OUT_DIFF_RADIANCE_HITDIST"].view(np.float16).astype(np.float32)
    assert np.isfinite(out).all() and out.any()
