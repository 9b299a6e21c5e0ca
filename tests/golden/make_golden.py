"""Generates the committed golden vectors: outputs of the ORACLE (not of the reference -- the reference ships none and
cannot run here: parity unpinned) on tiny seeded synthetic sequences.  They pin the oracle against accidental change and
give the GPU tests a fixture that does not need the oracle at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = {  # name: (denoiser, width, height, frames)
    "reblur_diffuse_specular_96x64": ("REBLUR_DIFFUSE_SPECULAR", 96, 64, 4),
    "sigma_shadow_96x64": ("SIGMA_SHADOW", 96, 64, 3),
    "relax_diffuse_specular_96x64": ("RELAX_DIFFUSE_SPECULAR", 96, 64, 4),
}


def run_case(denoiser_name, w, h, frames):
    import oracle_runner as orr
    from raytracingdenoiser_b200 import harness, nrd, scene
    den = getattr(nrd.Denoiser, denoiser_name)
    sc = scene.Scene(w, h)
    cpu = orr.CpuDenoiser(den, w, h)
    for f in range(frames):
        fr = sc.frame(f, harness.radiance_mode(den))
        cpu.set_inputs(fr)
        cpu.denoise(harness.make_common_settings(fr, w, h, f))
        if f == 0:
            cpu.set_inputs(fr)
    return {k: v.copy() for k, v in cpu.user.items() if k.startswith("OUT_")}


if __name__ == "__main__":
    only = sys.argv[1:] 
    for name, (den, w, h, frames) in CASES.items():
        if only and name not in only:
            continue
        try:
            out = run_case(den, w, h, frames)
        except RuntimeError as e:
            print("skip", name, e)
            continue
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name, {k: v.shape for k, v in out.items()})
