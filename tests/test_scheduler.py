"""Host-logic tests (no GPU): the DispatchDesc stream of the scheduler against tables derived by hand from the reference's
pass graphs (Source/Denoisers/*.hpp, Source/Reblur.cpp:104-210, Relax.cpp:182-295, Sigma.cpp:25-90) and an independent
numpy restatement of the constant-block arithmetic (Source/InstanceImpl.cpp:339-470, Reblur.cpp:297-406)."""
import math
import struct

import numpy as np
import pytest

from raytracingdenoiser_b200 import nrd, scene

W, H = 1920, 1080
RT = nrd.ResourceType


def common(frame_index=0, w=W, h=H, yaw=0.0, prev_yaw=None, eye=(0.0, 1.7, -4.0), prev_eye=None):
    cs = nrd.CommonSettings()
    P = scene.perspective_lh(60.0, w / float(h))
    V = scene.look_at_lh(eye, yaw, -0.1)
    Vp = scene.look_at_lh(prev_eye if prev_eye else eye, prev_yaw if prev_yaw is not None else yaw, -0.1)
    for k, m in (("viewToClipMatrix", P), ("viewToClipMatrixPrev", P), ("worldToViewMatrix", V), ("worldToViewMatrixPrev", Vp)):
        for i, v in enumerate(scene.colmajor(m)):
            getattr(cs, k)[i] = v
    for k in ("resourceSize", "resourceSizePrev", "rectSize", "rectSizePrev"):
        getattr(cs, k)[0], getattr(cs, k)[1] = w, h
    cs.motionVectorScale[0], cs.motionVectorScale[1], cs.motionVectorScale[2] = 1.0 / w, 1.0 / h, 1.0
    cs.timeDeltaBetweenFrames = 16.6667
    cs.frameIndex = frame_index
    return cs, P, V, Vp


def names(dispatches):
    return [d.name for d in dispatches]


def test_reblur_diffuse_specular_frame_schedule():
    inst = nrd.Instance([(3, nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR)])
    cs, _, _, _ = common(0)
    assert inst.set_common_settings(cs) == nrd.Result.SUCCESS
    d0 = inst.get_compute_dispatches([3])
    # first use forces CLEAR_AND_RESTART: one clear per storage-bound texture (13 permanent incl. ping-pong partners... see below)
    clears = [d for d in d0 if d.name.startswith("Clear")]
    chain = [d for d in d0 if not d.name.startswith("Clear")]
    cleared = {(r[1], r[2]) for d in clears for r in d.resources}
    assert (RT.IN_MV, 0) in cleared and (RT.OUT_DIFF_RADIANCE_HITDIST, 0) in cleared      # reference quirk: user textures bound as storage are cleared too
    assert all((RT.PERMANENT_POOL, i) in cleared for i in range(13)) and all((RT.TRANSIENT_POOL, i) in cleared for i in range(8))
    assert len(clears) == 13 + 8 + 3
    assert [c.shaderFileName for c in clears if (RT.PERMANENT_POOL, 2) in {(r[1], r[2]) for r in c.resources}] == ["Clear_Uint.cs"]   # R16_UINT
    P = "REBLUR_DiffuseSpecular - "
    assert names(chain) == [P + n for n in ("Classify tiles", "Pre-pass", "Temporal accumulation", "History fix", "Blur", "Post-blur", "Temporal stabilization")]
    assert [c.shaderFileName for c in chain] == ["REBLUR_ClassifyTiles.cs", "REBLUR_DiffuseSpecular_PrePass.cs", "REBLUR_DiffuseSpecular_TemporalAccumulation.cs",
                                                 "REBLUR_DiffuseSpecular_HistoryFix.cs", "REBLUR_DiffuseSpecular_Blur.cs", "REBLUR_DiffuseSpecular_PostBlur.cs",
                                                 "REBLUR_DiffuseSpecular_TemporalStabilization.cs"]
    # grid = ceil(rect / group): tiles 16x16, everything else 8x16
    assert (chain[0].gridWidth, chain[0].gridHeight) == (120, 68)
    assert all((c.gridWidth, c.gridHeight) == (240, 68) for c in chain[1:])
    assert all(len(c.constants) == 832 for c in chain)
    assert all(c.constantsMatchPrevious for c in chain[1:]) and not chain[0].constantsMatchPrevious
    # Blur bindings (Reblur_DiffuseSpecular.hpp:208-226): TILES, N/R, DATA1, TEMP1 diff/spec (= the outputs), VIEWZ -> TMP2 diff/spec, PREV_VIEWZ
    blur = chain[4]
    assert [(r[0].name, r[1].name, r[2]) for r in blur.resources] == [
        ("TEXTURE", "TRANSIENT_POOL", 7), ("TEXTURE", "IN_NORMAL_ROUGHNESS", 0), ("TEXTURE", "TRANSIENT_POOL", 0), ("TEXTURE", "OUT_DIFF_RADIANCE_HITDIST", 0),
        ("TEXTURE", "OUT_SPEC_RADIANCE_HITDIST", 0), ("TEXTURE", "IN_VIEWZ", 0), ("STORAGE_TEXTURE", "TRANSIENT_POOL", 3), ("STORAGE_TEXTURE", "TRANSIENT_POOL", 5),
        ("STORAGE_TEXTURE", "PERMANENT_POOL", 0)]
    # temporal accumulation has 18 inputs + 7 outputs
    assert len(chain[2].resources) == 25 and sum(1 for r in chain[2].resources if r[0] == nrd.DescriptorType.STORAGE_TEXTURE) == 7


def test_ping_pong_swaps_every_frame():
    inst = nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR)])
    seen = []
    for f in range(3):
        cs, _, _, _ = common(f)
        inst.set_common_settings(cs)
        ds = [d for d in inst.get_compute_dispatches([0]) if "Temporal accumulation" in d.name][0]
        seen.append((ds.resources[16][2], ds.resources[22][2]))   # prev hit-dist-for-tracking in, current out
    assert seen[0][0] != seen[0][1] and seen[0] == seen[2] and seen[1] == (seen[0][1], seen[0][0])


def test_settings_select_permutations():
    inst = nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE)])
    cs, _, _, _ = common(0)
    inst.set_common_settings(cs)
    s = nrd.ReblurSettings(maxAccumulatedFrameNum=0, maxFastAccumulatedFrameNum=0, maxStabilizedFrameNum=0, historyFixFrameNum=0, diffusePrepassBlurRadius=0.0)
    inst.set_denoiser_settings(0, s)
    chain = [d for d in inst.get_compute_dispatches([0]) if not d.name.startswith("Clear")]
    assert [c.shaderFileName for c in chain] == ["REBLUR_ClassifyTiles.cs", "REBLUR_Diffuse_TemporalAccumulation.cs", "REBLUR_Diffuse_HistoryFix.cs",
                                                 "REBLUR_Diffuse_Blur.cs", "REBLUR_Diffuse_PostBlur_NoTemporalStabilization.cs"]
    # without a pre-pass temporal accumulation reads the noisy input directly
    assert chain[1].resources[9][1] == RT.IN_DIFF_RADIANCE_HITDIST
    s2 = nrd.ReblurSettings(enablePerformanceMode=True)
    inst.set_denoiser_settings(0, s2)
    cs.frameIndex = 1
    inst.set_common_settings(cs)
    chain = inst.get_compute_dispatches([0])
    assert chain[-1].shaderFileName == "REBLUR_Perf_Diffuse_TemporalStabilization.cs"


def test_sigma_and_relax_schedules():
    inst = nrd.Instance([(1, nrd.Denoiser.SIGMA_SHADOW), (2, nrd.Denoiser.RELAX_DIFFUSE_SPECULAR)])
    cs, _, _, _ = common(0, 1280, 720)
    inst.set_common_settings(cs)
    sig = [d for d in inst.get_compute_dispatches([1]) if not d.name.startswith("Clear")]
    assert [d.shaderFileName for d in sig] == ["SIGMA_Shadow_ClassifyTiles.cs", "SIGMA_SmoothTiles.cs", "SIGMA_Copy.cs", "SIGMA_Shadow_Blur.cs",
                                               "SIGMA_Shadow_PostBlur.cs", "SIGMA_Shadow_TemporalStabilization.cs"]
    assert all(len(d.constants) == 516 for d in sig)   # sizeof() of the reference struct (528 once padded to registers)
    assert (sig[1].gridWidth, sig[1].gridHeight) == (5, 3)            # smooth tiles runs at 1/16 resolution in 16x16 groups
    assert sig[4].resources[-1][1] == RT.TRANSIENT_POOL                # post-blur writes TEMP_2 when stabilization is on
    inst.set_denoiser_settings(1, nrd.SigmaSettings(maxStabilizedFrameNum=0))
    cs.frameIndex = 1
    inst.set_common_settings(cs)
    sig = inst.get_compute_dispatches([1])
    assert [d.shaderFileName for d in sig][-1] == "SIGMA_Shadow_PostBlur.cs" and sig[-1].resources[-1][1] == RT.OUT_SHADOW_TRANSLUCENCY and len(sig) == 4

    rel = [d for d in inst.get_compute_dispatches([2]) if not d.name.startswith("Clear")]
    sh = [d.shaderFileName for d in rel]
    assert sh[:5] == ["RELAX_ClassifyTiles.cs", "RELAX_DiffuseSpecular_PrePass.cs", "RELAX_DiffuseSpecular_TemporalAccumulation.cs",
                      "RELAX_DiffuseSpecular_HistoryFix.cs", "RELAX_DiffuseSpecular_HistoryClamping.cs"]
    assert sh[5:] == ["RELAX_DiffuseSpecular_AtrousSmem.cs"] + ["RELAX_DiffuseSpecular_Atrous.cs"] * 4
    steps = [struct.unpack_from("2I", d.constants, 704) for d in rel[5:]]
    assert steps == [(1, 0), (2, 0), (4, 0), (8, 0), (16, 1)]
    assert len(rel[5].constants) == 712 and len(rel[1].constants) == 704
    # iterations alternate PING/PONG, the last one writes the outputs (Relax.cpp:263-276)
    assert rel[-1].resources[-2][1] == RT.OUT_SPEC_RADIANCE_HITDIST and rel[-1].resources[-1][1] == RT.OUT_DIFF_RADIANCE_HITDIST
    assert rel[6].resources[-2][2] != rel[7].resources[-2][2]


def test_invalid_arguments():
    inst = nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE)])
    cs, _, _, _ = common(0)
    cs.denoisingRange = -1.0
    assert inst.set_common_settings(cs, check=False) == nrd.Result.INVALID_ARGUMENT
    cs, _, _, _ = common(0)
    cs.rectSize[0] = 0
    assert inst.set_common_settings(cs, check=False) == nrd.Result.INVALID_ARGUMENT
    assert inst.set_denoiser_settings(77, nrd.ReblurSettings(), check=False) == nrd.Result.INVALID_ARGUMENT
    r, _, n = inst.get_compute_dispatches_raw([])
    assert r == nrd.Result.SUCCESS and n == 0
    cs, _, _, _ = common(0)
    inst.set_common_settings(cs)
    r, _, n = inst.get_compute_dispatches_raw([123])          # unknown identifier -> empty list -> INVALID_ARGUMENT (InstanceImpl.cpp:577)
    assert r == nrd.Result.INVALID_ARGUMENT and n == 0


def test_transient_pool_is_shared_between_denoisers():
    a = nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR)]).get_instance_desc()
    b = nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR), (1, nrd.Denoiser.REBLUR_DIFFUSE)]).get_instance_desc()
    assert len(b["permanentPool"]) == 13 + 7
    # REBLUR_DIFFUSE needs R8_UNORM, R8_UINT, RGBA16F, R16F, R8@16: the last three alias textures of the first denoiser
    assert len(b["transientPool"]) == len(a["transientPool"]) + 2


# ---------------------------------------------------------------------------------------------------------------------
# independent restatement of the constants
# ---------------------------------------------------------------------------------------------------------------------
def _weyl(p, n):
    v = np.float32(p) + np.float32((n * 10368889) & 0xFFFFFFFF) / np.float32(16777216.0)
    return float(v - np.floor(v))


def _rotator(a):
    return np.array([math.cos(a), math.sin(a), -math.sin(a), math.cos(a)])


def _combine(r1, r2):
    return np.array([r1[0] * r2[0] + r1[2] * r2[1], r1[1] * r2[0] + r1[3] * r2[1], r1[0] * r2[2] + r1[2] * r2[3], r1[1] * r2[2] + r1[3] * r2[3]])


def test_reblur_constants_against_numpy_restatement():
    inst = nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR)])
    inst.set_common_settings(common(0)[0])
    inst.get_compute_dispatches([0])
    f = 7
    cs, P, V, Vp = common(f, yaw=0.02, prev_yaw=0.0, eye=(0.1, 1.7, -3.9), prev_eye=(0.0, 1.7, -4.0))
    inst.set_common_settings(cs)
    c = inst.get_compute_dispatches([0])[4].constants
    fl = lambda off, n: np.array(struct.unpack_from("%df" % n, c, off), dtype=np.float64)
    mat = lambda off: fl(off, 16).reshape(4, 4).T                       # column-major -> [row, col]

    P64, V64, Vp64 = P.astype(np.float64), V.astype(np.float64), Vp.astype(np.float64)
    v2w, v2w_prev = np.linalg.inv(V64), np.linalg.inv(Vp64)
    delta = v2w_prev[:3, 3] - v2w[:3, 3]
    v2w_rel = v2w.copy(); v2w_rel[:3, 3] = 0
    v2w_prev_rel = v2w_prev.copy(); v2w_prev_rel[:3, 3] = delta
    w2v_rel, w2v_prev_rel = np.linalg.inv(v2w_rel), np.linalg.inv(v2w_prev_rel)
    np.testing.assert_allclose(mat(0), P64 @ w2v_rel, atol=2e-6)        # gWorldToClip (camera relative)
    np.testing.assert_allclose(mat(64), P64, atol=1e-7)                # gViewToClip
    np.testing.assert_allclose(mat(128), v2w_rel, atol=2e-6)           # gViewToWorld
    np.testing.assert_allclose(mat(192), w2v_prev_rel, atol=2e-6)      # gWorldToViewPrev
    np.testing.assert_allclose(mat(256), P64 @ w2v_prev_rel, atol=4e-6)  # gWorldToClipPrev
    np.testing.assert_allclose(mat(320), np.eye(4), atol=0)            # gWorldPrevToWorld
    # rotators (InstanceImpl.cpp:339-349)
    bayer = lambda n: (n & 15) / 16.0                                   # Bayer4x4 at pixel (0,0): matrix entry 0
    np.testing.assert_allclose(fl(384, 4), _rotator(_weyl(0.5, f) * math.pi / 2), atol=2e-6)
    np.testing.assert_allclose(fl(400, 4), _combine(_rotator(_weyl(0.0, 2 * f) * math.pi / 2), _rotator(bayer(2 * f) * 2 * math.pi)), atol=2e-6)
    np.testing.assert_allclose(fl(416, 4), _combine(_rotator(_weyl(0.0, 2 * f + 1) * math.pi / 2), _rotator(bayer(2 * f + 1) * 2 * math.pi)), atol=2e-6)
    # frustum: view.xy = (uv * zw + xy) * z must invert the projection
    fr = fl(432, 4)
    t = math.tan(math.radians(30.0))
    np.testing.assert_allclose(fr, [-t * W / H, t, 2 * t * W / H, -2 * t], rtol=1e-6)
    np.testing.assert_allclose(fl(464, 3), delta, atol=2e-6)           # gCameraDelta
    np.testing.assert_allclose(fl(480, 4), [3.0, 0.1, 20.0, -25.0])    # gHitDistParams
    np.testing.assert_allclose(fl(496, 3), -v2w_rel[:3, 2], atol=2e-6)   # gViewVectorWorld
    np.testing.assert_allclose(fl(528, 4), [1.0 / W, 1.0 / H, 1.0, 0.0], rtol=1e-6)   # gMvScale
    assert struct.unpack_from("2i", c, 656) == (W - 1, H - 1)          # gRectSizeMinusOne
    scal = fl(664, 35)
    unproject = 1.0 / (0.5 * H * (1.0 / t))
    np.testing.assert_allclose(scal[0], 0.01 + 1.0 / H, rtol=1e-6)     # gDisocclusionThreshold (+ (1 + jitterDelta) / rectH)
    np.testing.assert_allclose(scal[5], 63.0 / 64.0, rtol=1e-6)        # gStabilizationStrength
    np.testing.assert_allclose(scal[9], unproject, rtol=1e-6)          # gUnproject
    np.testing.assert_allclose(scal[12], 33.333 / 16.6667, rtol=1e-5)  # gFramerateScale
    np.testing.assert_allclose(scal[13:17], [1.0, 30.0, 30.0, 50.0])   # min / max blur radius, pre-pass radii
    np.testing.assert_allclose(scal[17:19], [30.0, 6.0])               # accumulated frame limits
    np.testing.assert_allclose(scal[20], 0.15 * 0.15, rtol=1e-6)       # gLobeAngleFraction is squared (Reblur.cpp:384)
    np.testing.assert_allclose(scal[25], H * unproject, rtol=1e-6)     # gMinRectDimMulUnproject
    assert struct.unpack_from("7I", c, 804) == (0, 0, 2, 2, f, 0, 0)

    # history reset: accumulation limits and stabilization collapse to 0
    cs.accumulationMode = int(nrd.AccumulationMode.RESTART)
    cs.frameIndex = f + 1
    inst.set_common_settings(cs)
    c = inst.get_compute_dispatches([0])[0].constants
    assert struct.unpack_from("f", c, 664 + 5 * 4)[0] == 0.0 and struct.unpack_from("2f", c, 664 + 17 * 4) == (0.0, 0.0)
    assert struct.unpack_from("7I", c, 804)[6] == 1


def test_right_handed_input_is_converted():
    """A right-handed camera must produce the same constants as its left-handed mirror (InstanceImpl.cpp:392-408)."""
    inst_l = nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE)])
    inst_r = nrd.Instance([(0, nrd.Denoiser.REBLUR_DIFFUSE)])
    cs, P, V, Vp = common(0, yaw=0.1)
    inst_l.set_common_settings(cs)
    flip = np.diag([1.0, 1.0, -1.0, 1.0]).astype(np.float32)
    csr, _, _, _ = common(0, yaw=0.1)
    for k, m in (("viewToClipMatrix", P @ flip), ("viewToClipMatrixPrev", P @ flip), ("worldToViewMatrix", flip @ V), ("worldToViewMatrixPrev", flip @ Vp)):
        for i, v in enumerate(scene.colmajor(m)):
            getattr(csr, k)[i] = v
    inst_r.set_common_settings(csr)
    a = inst_l.get_compute_dispatches([0])[-1].constants
    b = inst_r.get_compute_dispatches([0])[-1].constants
    np.testing.assert_allclose(np.frombuffer(a[:544], np.float32), np.frombuffer(b[:544], np.float32), atol=1e-6)


def test_relax_constants_against_numpy_restatement():
    """RELAX_SHARED_CONSTANTS (RELAX_Config.hlsli:21-99) as filled by Relax.cpp:60-180, restated independently."""
    inst = nrd.Instance([(0, nrd.Denoiser.RELAX_DIFFUSE_SPECULAR)])
    inst.set_common_settings(common(0)[0])
    inst.get_compute_dispatches([0])
    f = 5
    cs, P, V, Vp = common(f, yaw=0.03, prev_yaw=0.01, eye=(0.2, 1.7, -3.8), prev_eye=(0.1, 1.7, -4.0))
    inst.set_common_settings(cs)
    ds = inst.get_compute_dispatches([0])
    c = [d for d in ds if d.shaderFileName == "RELAX_DiffuseSpecular_TemporalAccumulation.cs"][0].constants
    assert len(c) == 704
    fl = lambda off, n: np.array(struct.unpack_from("%df" % n, c, off), dtype=np.float64)
    mat = lambda off: fl(off, 16).reshape(4, 4).T

    P64, V64, Vp64 = P.astype(np.float64), V.astype(np.float64), Vp.astype(np.float64)
    v2w, v2w_prev = np.linalg.inv(V64), np.linalg.inv(Vp64)
    delta = v2w_prev[:3, 3] - v2w[:3, 3]
    v2w_rel = v2w.copy(); v2w_rel[:3, 3] = 0
    v2w_prev_rel = v2w_prev.copy(); v2w_prev_rel[:3, 3] = delta
    w2v_rel, w2v_prev_rel = np.linalg.inv(v2w_rel), np.linalg.inv(v2w_prev_rel)
    np.testing.assert_allclose(mat(0), P64 @ w2v_rel, atol=2e-6)             # gWorldToClip
    np.testing.assert_allclose(mat(64), P64 @ w2v_prev_rel, atol=4e-6)       # gWorldToClipPrev
    np.testing.assert_allclose(mat(128), w2v_prev_rel, atol=2e-6)            # gWorldToViewPrev
    np.testing.assert_allclose(mat(192), np.eye(4), atol=0)                  # gWorldPrevToWorld
    np.testing.assert_allclose(fl(256, 4), _rotator(_weyl(0.5, f) * math.pi / 2), atol=2e-6)   # gRotatorPre
    # world-space frustum axes: pixel ray = forward + right * (2u - 1) - up * (2v - 1), forward.z (view) = 1
    t = math.tan(math.radians(30.0))
    np.testing.assert_allclose(fl(272, 4), np.append(w2v_rel[0, :3] * t * W / H, 0.0), atol=2e-6)        # gFrustumRight
    np.testing.assert_allclose(fl(288, 4), np.append(w2v_rel[1, :3] * t, 0.0), atol=2e-6)                # gFrustumUp
    np.testing.assert_allclose(fl(304, 4), np.append(v2w_rel[:3, 2], 0.0), atol=2e-6)                    # gFrustumForward (symmetric projection)
    np.testing.assert_allclose(fl(320, 4), np.append(w2v_prev_rel[0, :3] * t * W / H, 0.0), atol=2e-6)   # gPrevFrustumRight
    np.testing.assert_allclose(fl(336, 4), np.append(w2v_prev_rel[1, :3] * t, 0.0), atol=2e-6)           # gPrevFrustumUp
    np.testing.assert_allclose(fl(352, 4), np.append(v2w_prev_rel[:3, 2], 0.0), atol=2e-6)               # gPrevFrustumForward
    np.testing.assert_allclose(fl(368, 4), np.append(delta, 0.0), atol=2e-6)                             # gCameraDelta
    np.testing.assert_allclose(fl(384, 4), [1.0 / W, 1.0 / H, 1.0, 0.0], rtol=1e-6)                      # gMvScale
    np.testing.assert_allclose(fl(400, 14), [0, 0, 1, 1, 0, 0, 1.0 / W, 1.0 / H, W, H, 1.0 / W, 1.0 / H, W, H], rtol=1e-6)   # jitter .. gRectSizePrev
    np.testing.assert_allclose(fl(456, 2), [1.0 / W, 1.0 / H], rtol=1e-6)                                # gResourceSizeInvPrev
    assert struct.unpack_from("4I2i", c, 464) == (9999, 9999, 0, 0, W, H)                                    # gPrintfAt, gRectOrigin, gRectSize
    s = fl(488, 47)
    np.testing.assert_allclose(s[0:4], [30.0, 6.0, 30.0, 6.0])                                           # accumulated frame limits
    np.testing.assert_allclose(s[4:6], [0.01 + 1.0 / H, 0.05 + 1.0 / H], rtol=1e-6)                      # disocclusion thresholds (+ (1 + jitterDelta) / rectH)
    np.testing.assert_allclose(s[6:9], [999.0, 999.0, 80e-6], rtol=1e-6)                                  # material ids, strand thickness
    np.testing.assert_allclose(s[9:16], [0.15, 0.0, 0.0, 30.0, 50.0, 0.003, 0.5], rtol=1e-6)             # roughnessFraction .. gLobeAngleFraction
    np.testing.assert_allclose(s[16], math.radians(0.15), rtol=1e-6)                                     # gSpecLobeAngleSlack is converted to radians
    np.testing.assert_allclose(s[17:25], [8.0, 1.0, 0.3, 2.0, 0.3, 0.5, 4.5, 0.5], rtol=1e-6)            # edge stopping, colour box, anti-lag
    np.testing.assert_allclose(s[25:28], [500000.0, 1.0, 2.0], rtol=1e-6)                                # denoising range, phi luminance
    assert np.isinf(s[28]) and np.isinf(s[29])                                                           # -log(saturate(minLuminanceWeight = 0))
    np.testing.assert_allclose(s[30], 1.0)                                # gLuminanceEdgeStoppingRelaxation takes roughnessEdgeStoppingRelaxation (Relax.cpp:154)
    np.testing.assert_allclose(s[34:36], [0.0, 0.0])                                                     # gDebug, gOrthoMode
    np.testing.assert_allclose(s[36], t / (0.5 * H), rtol=1e-6)                                          # gUnproject
    np.testing.assert_allclose(s[37], 16.66 / 16.6667, rtol=1e-5)                                        # gFramerateScale
    np.testing.assert_allclose(s[39:47], [0.0, 4.0, 14.0, 3.0, 1.0, 0.2, 4.0, 4.0], rtol=1e-6)           # jitterDelta .. min materials
    assert struct.unpack_from("7I", c, 676) == (1, f, 2, 2, 0, 0, 0)
    # the A-trous passes append gStepSize / gIsLastPass: 1, 2, 4, 8, 16 with the last one flagged
    steps = [struct.unpack_from("2I", d.constants, 704) for d in ds if "Atrous" in d.shaderFileName]
    assert steps == [(1, 0), (2, 0), (4, 0), (8, 0), (16, 1)]


def test_sigma_constants_against_numpy_restatement():
    """SIGMA_SHARED_CONSTANTS as filled by Sigma.cpp:92-144, restated independently."""
    inst = nrd.Instance([(0, nrd.Denoiser.SIGMA_SHADOW)])
    s = nrd.SigmaSettings()
    light = np.array([0.3, 0.8, -0.5]) / np.linalg.norm([0.3, 0.8, -0.5])
    for i in range(3):
        s.lightDirection[i] = float(light[i])
    inst.set_denoiser_settings(0, s)
    inst.set_common_settings(common(0)[0])
    inst.get_compute_dispatches([0])
    f = 3
    cs, P, V, Vp = common(f, yaw=0.04, prev_yaw=0.02, eye=(0.3, 1.7, -3.7), prev_eye=(0.2, 1.7, -3.9))
    inst.set_common_settings(cs)
    ds = inst.get_compute_dispatches([0])
    c = [d for d in ds if d.shaderFileName == "SIGMA_Shadow_Blur.cs"][0].constants
    assert len(c) in (516, 528)
    fl = lambda off, n: np.array(struct.unpack_from("%df" % n, c, off), dtype=np.float64)
    mat = lambda off: fl(off, 16).reshape(4, 4).T
    P64, V64, Vp64 = P.astype(np.float64), V.astype(np.float64), Vp.astype(np.float64)
    v2w, v2w_prev = np.linalg.inv(V64), np.linalg.inv(Vp64)
    delta = v2w_prev[:3, 3] - v2w[:3, 3]
    v2w_rel = v2w.copy(); v2w_rel[:3, 3] = 0
    v2w_prev_rel = v2w_prev.copy(); v2w_prev_rel[:3, 3] = delta
    w2v_rel, w2v_prev_rel = np.linalg.inv(v2w_rel), np.linalg.inv(v2w_prev_rel)
    np.testing.assert_allclose(mat(0), w2v_rel, atol=2e-6)                    # gWorldToView (camera relative)
    np.testing.assert_allclose(mat(64), P64, atol=1e-7)                       # gViewToClip
    np.testing.assert_allclose(mat(128), P64 @ w2v_prev_rel, atol=4e-6)       # gWorldToClipPrev
    np.testing.assert_allclose(mat(192), w2v_prev_rel, atol=2e-6)             # gWorldToViewPrev
    bayer = lambda n: (n & 15) / 16.0
    np.testing.assert_allclose(fl(256, 4), _combine(_rotator(_weyl(0.0, 2 * f) * math.pi / 2), _rotator(bayer(2 * f) * 2 * math.pi)), atol=2e-6)       # gRotator
    np.testing.assert_allclose(fl(272, 4), _combine(_rotator(_weyl(0.0, 2 * f + 1) * math.pi / 2), _rotator(bayer(2 * f + 1) * 2 * math.pi)), atol=2e-6)  # gRotatorPost
    np.testing.assert_allclose(fl(288, 3), -v2w_rel[:3, 2], atol=2e-6)         # gViewVectorWorld
    np.testing.assert_allclose(fl(304, 4), np.append(w2v_rel[:3, :3] @ light, 0.0), atol=2e-6)   # gLightDirectionView
    t = math.tan(math.radians(30.0))
    np.testing.assert_allclose(fl(320, 4), [-t * W / H, t, 2 * t * W / H, -2 * t], rtol=1e-6)     # gFrustum
    np.testing.assert_allclose(fl(336, 4), [-t * W / H, t, 2 * t * W / H, -2 * t], rtol=1e-6)     # gFrustumPrev
    np.testing.assert_allclose(fl(352, 3), delta, atol=2e-6)                   # gCameraDelta
    np.testing.assert_allclose(fl(368, 4), [1.0 / W, 1.0 / H, 1.0, 0.0], rtol=1e-6)               # gMvScale
    np.testing.assert_allclose(fl(384, 14), [1.0 / W, 1.0 / H, 1.0 / W, 1.0 / H, W, H, 1.0 / W, 1.0 / H, W, H, 1, 1, 0, 0], rtol=1e-6)
    assert struct.unpack_from("4I4i", c, 440) == (9999, 9999, 0, 0, W - 1, H - 1, (W + 15) // 16 - 1, (H + 15) // 16 - 1)
    sc = fl(472, 9)
    unproject = t / (0.5 * H)
    np.testing.assert_allclose(sc, [0.0, unproject, 500000.0, 0.02, 5.0 / 6.0, 0.0, 0.0, 1.0, H * unproject], rtol=1e-6)
    assert struct.unpack_from("2I", c, 508) == (f, 0)
