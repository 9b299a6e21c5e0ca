import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Both native libraries must exist before anything imports them (they are git-ignored build products)."""
    from raytracingdenoiser_b200 import build
    build.build_all()
    yield


# ---- time budget of the GPU session -------------------------------------------------------------------------------------------
# The BASELINE-size parity tests compare against the CPU oracle, whose speed depends on the host of the GPU box (1.7 - 10 Mpixels/s
# have been seen on the same box class: cgroup quotas differ).  On a fast host the whole GPU suite takes ~15 minutes; so that a
# slow host cannot turn it into hours, the heavy tests ask for their estimated cost first and are SKIPPED (loudly, with the
# measured oracle speed) once the session would exceed NRD_B200_GPU_TEST_BUDGET_S (default 2700 s).  Their last results from a
# fast host are kept under profiles/ (parity_*.json summaries).
import time

_SESSION_START = time.time()
_ORACLE_SPEED = {}


def _oracle_mpixels_per_s():
    if "v" not in _ORACLE_SPEED:
        import oracle_runner as orr
        from raytracingdenoiser_b200 import harness, nrd, scene
        w, h = 640, 360
        sc = scene.Scene(w, h)
        cpu = orr.CpuDenoiser(nrd.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h)
        fr = sc.frame(0)
        cpu.set_inputs(fr)
        cpu.denoise(harness.make_common_settings(fr, w, h, 0))
        t0 = time.time()
        cpu.set_inputs(fr)
        cpu.denoise(harness.make_common_settings(fr, w, h, 1))
        _ORACLE_SPEED["v"] = w * h / max(time.time() - t0, 1e-3) / 1e6
    return _ORACLE_SPEED["v"]


@pytest.fixture
def gpu_time_budget():
    """gpu_time_budget(cost_s): cost_s = seconds the test takes on a host whose oracle runs at 2 Mpixels/s."""
    def ask(cost_s):
        budget = float(os.environ.get("NRD_B200_GPU_TEST_BUDGET_S", "2700"))
        speed = _oracle_mpixels_per_s()
        need = cost_s * max(1.0, 2.0 / speed)
        elapsed = time.time() - _SESSION_START
        if elapsed + need > budget:
            pytest.skip("host oracle runs at %.2f Mpixels/s: this test needs ~%.0f s, %.0f s of the %.0f s GPU-session budget are used "
                        "(NRD_B200_GPU_TEST_BUDGET_S raises it)" % (speed, need, elapsed, budget))
    return ask
