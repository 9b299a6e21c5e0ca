"""Glue between the synthetic scene, torch device memory and the NRD-style API: what an application would write.

`make_common_settings` fills nrd.CommonSettings from a scene frame; `GpuDenoiser` owns the user textures (IN_*/OUT_*)
as torch CUDA tensors, binds them to a nrd.CudaContext and runs `nrdCudaDenoise` per frame.  No torch op is on the
denoising path: tensors are only storage.
"""
import numpy as np
import torch

from . import nrd, scene

USER_FORMATS = {
    "IN_MV": (nrd.Format.RGBA16_SFLOAT, torch.float16, 4),
    "IN_NORMAL_ROUGHNESS": (nrd.Format.R10_G10_B10_A2_UNORM, torch.int32, 1),
    "IN_VIEWZ": (nrd.Format.R32_SFLOAT, torch.float32, 1),
    "IN_DIFF_RADIANCE_HITDIST": (nrd.Format.RGBA16_SFLOAT, torch.float16, 4),
    "IN_SPEC_RADIANCE_HITDIST": (nrd.Format.RGBA16_SFLOAT, torch.float16, 4),
    "OUT_DIFF_RADIANCE_HITDIST": (nrd.Format.RGBA16_SFLOAT, torch.float16, 4),
    "OUT_SPEC_RADIANCE_HITDIST": (nrd.Format.RGBA16_SFLOAT, torch.float16, 4),
    "IN_PENUMBRA": (nrd.Format.R16_SFLOAT, torch.float16, 1),
    "IN_SIGNAL": (nrd.Format.RGBA16_SFLOAT, torch.float16, 4),
    "OUT_SIGNAL": (nrd.Format.RGBA16_SFLOAT, torch.float16, 4),
    "IN_DIFF_CONFIDENCE": (nrd.Format.R8_UNORM, torch.uint8, 1),
    "IN_SPEC_CONFIDENCE": (nrd.Format.R8_UNORM, torch.uint8, 1),
    "IN_DISOCCLUSION_THRESHOLD_MIX": (nrd.Format.R8_UNORM, torch.uint8, 1),
    "IN_BASECOLOR_METALNESS": (nrd.Format.RGBA8_UNORM, torch.uint8, 4),
    "OUT_SHADOW_TRANSLUCENCY": (nrd.Format.R8_UNORM, torch.uint8, 1),
    "IN_TRANSLUCENCY": (nrd.Format.RGBA8_UNORM, torch.uint8, 4),
    "OUT_SHADOW_TRANSLUCENCY#RGBA8": (nrd.Format.RGBA8_UNORM, torch.uint8, 4),   # SIGMA_SHADOW_TRANSLUCENCY writes float4
}


def user_format(denoiser, name):
    """(format, torch dtype, channels) of a user texture; OUT_SHADOW_TRANSLUCENCY is RGBA8 for the translucent SIGMA variant."""
    if name == "OUT_SHADOW_TRANSLUCENCY" and denoiser == nrd.Denoiser.SIGMA_SHADOW_TRANSLUCENCY:
        return USER_FORMATS["OUT_SHADOW_TRANSLUCENCY#RGBA8"]
    return USER_FORMATS[name]


DENOISER_RESOURCES = {
    nrd.Denoiser.REBLUR_DIFFUSE: ["IN_MV", "IN_NORMAL_ROUGHNESS", "IN_VIEWZ", "IN_DIFF_RADIANCE_HITDIST", "OUT_DIFF_RADIANCE_HITDIST"],
    nrd.Denoiser.REBLUR_SPECULAR: ["IN_MV", "IN_NORMAL_ROUGHNESS", "IN_VIEWZ", "IN_SPEC_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST"],
    nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR: ["IN_MV", "IN_NORMAL_ROUGHNESS", "IN_VIEWZ", "IN_DIFF_RADIANCE_HITDIST", "IN_SPEC_RADIANCE_HITDIST",
                                           "OUT_DIFF_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST"],
    nrd.Denoiser.RELAX_DIFFUSE_SPECULAR: ["IN_MV", "IN_NORMAL_ROUGHNESS", "IN_VIEWZ", "IN_DIFF_RADIANCE_HITDIST", "IN_SPEC_RADIANCE_HITDIST",
                                          "OUT_DIFF_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST"],
    nrd.Denoiser.RELAX_DIFFUSE: ["IN_MV", "IN_NORMAL_ROUGHNESS", "IN_VIEWZ", "IN_DIFF_RADIANCE_HITDIST", "OUT_DIFF_RADIANCE_HITDIST"],
    nrd.Denoiser.RELAX_SPECULAR: ["IN_MV", "IN_NORMAL_ROUGHNESS", "IN_VIEWZ", "IN_SPEC_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST"],
    nrd.Denoiser.REFERENCE: ["IN_SIGNAL", "OUT_SIGNAL"],
    nrd.Denoiser.SIGMA_SHADOW: ["IN_MV", "IN_NORMAL_ROUGHNESS", "IN_VIEWZ", "IN_PENUMBRA", "OUT_SHADOW_TRANSLUCENCY"],
    nrd.Denoiser.SIGMA_SHADOW_TRANSLUCENCY: ["IN_MV", "IN_NORMAL_ROUGHNESS", "IN_VIEWZ", "IN_PENUMBRA", "IN_TRANSLUCENCY", "OUT_SHADOW_TRANSLUCENCY"],
}


# inputs an application binds from its full-size G-buffer: read at rectOrigin + pixel (Common.hlsli:200-205 WithRectOrigin)
RECT_ORIGIN_INPUTS = ("IN_VIEWZ", "IN_NORMAL_ROUGHNESS", "IN_MV", "IN_DIFF_CONFIDENCE", "IN_SPEC_CONFIDENCE", "IN_DISOCCLUSION_THRESHOLD_MIX", "IN_BASECOLOR_METALNESS")

OPTIONAL_INPUTS = {  # CommonSettings flag -> the user textures it makes the passes read, per signal
    "isHistoryConfidenceAvailable": {"diff": "IN_DIFF_CONFIDENCE", "spec": "IN_SPEC_CONFIDENCE"},
    "isDisocclusionThresholdMixAvailable": {"any": "IN_DISOCCLUSION_THRESHOLD_MIX"},
    "isBaseColorMetalnessAvailable": {"spec": "IN_BASECOLOR_METALNESS"},
}


def denoiser_resources(denoiser, common=None):
    """User textures of a denoiser, plus the optional inputs the CommonSettings overrides in `common` switch on."""
    names = list(DENOISER_RESOURCES[denoiser])
    for flag, extra in OPTIONAL_INPUTS.items():
        if common and common.get(flag):
            if "any" in extra:
                names.append(extra["any"])
            if "diff" in extra and "IN_DIFF_RADIANCE_HITDIST" in names:
                names.append(extra["diff"])
            if "spec" in extra and "IN_SPEC_RADIANCE_HITDIST" in names:
                names.append(extra["spec"])
    return names


def radiance_mode(denoiser):
    return "relax" if nrd.Denoiser(denoiser).name.startswith("RELAX") else "reblur"


def make_common_settings(frame, width, height, frame_index, time_delta_ms=16.6667, common=None):
    """CommonSettings for one synthetic frame: 2.5D motion vectors in pixels, deterministic frame time.  `common` = dict of
    CommonSettings fields to override (optional inputs, accumulation mode, ...)."""
    cs = nrd.CommonSettings()
    for k, m in (("viewToClipMatrix", frame["viewToClip"]), ("viewToClipMatrixPrev", frame["viewToClip"]),
                 ("worldToViewMatrix", frame["worldToView"]), ("worldToViewMatrixPrev", frame["worldToViewPrev"])):
        arr = getattr(cs, k)
        for i, v in enumerate(scene.colmajor(m)):
            arr[i] = v
    cs.motionVectorScale[0], cs.motionVectorScale[1], cs.motionVectorScale[2] = 1.0 / width, 1.0 / height, 1.0
    for k in ("resourceSize", "resourceSizePrev", "rectSize", "rectSizePrev"):
        getattr(cs, k)[0], getattr(cs, k)[1] = width, height
    common = dict(common or {})
    for k in ("resourceSize", "resourceSizePrev", "rectSize", "rectSizePrev", "rectOrigin"):  # array fields (dynamic resolution)
        if k in common:
            v = common.pop(k)
            getattr(cs, k)[0], getattr(cs, k)[1] = int(v[0]), int(v[1])
    cs.timeDeltaBetweenFrames = time_delta_ms
    cs.frameIndex = frame_index
    for k, v in (common or {}).items():
        setattr(cs, k, v)
    return cs


class GpuDenoiser(object):
    """One denoiser instance + CUDA context + user textures on one GPU (optionally one strip of the frame)."""

    def __init__(self, denoiser, width, height, device=0, identifier=0, settings=None, common=None):
        self.denoiser, self.width, self.height, self.identifier, self.common = denoiser, width, height, identifier, common
        self.device = torch.device("cuda", device)
        self.instance = nrd.Instance([(identifier, denoiser)])
        self.ctx = nrd.CudaContext(self.instance, width, height, device=device)
        self.tex = {}
        for name in denoiser_resources(denoiser, common):
            fmt, dtype, ch = user_format(denoiser, name)
            shape = (height, width, ch) if ch > 1 else (height, width)
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self.tex[name] = t
            self.ctx.set_user_texture(getattr(nrd.ResourceType, name), t.data_ptr(), t.stride(0) * t.element_size(), fmt)
        if settings is not None:
            self.instance.set_denoiser_settings(identifier, settings)

    def set_inputs(self, frame, non_blocking=False):
        for name, t in self.tex.items():
            if name.startswith("IN_"):
                t.copy_(frame[name], non_blocking=non_blocking)

    def denoise(self, common_settings, stream=None):
        self.instance.set_common_settings(common_settings)
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        return self.ctx.denoise([self.identifier], stream=s)

    def outputs(self):
        return {k: v for k, v in self.tex.items() if k.startswith("OUT_")}

    def destroy(self):
        self.ctx.destroy()
        self.instance.destroy()
