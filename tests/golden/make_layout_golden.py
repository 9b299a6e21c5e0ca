"""Generates tests/golden/nrd_layout.json from the REFERENCE's own headers (/root/reference/Include/NRD.h): sizeof, alignof and
default-initialised bytes of every public struct, enum extents, version.  Run in the build container (the reference tree is not
on the GPU box):  python tests/golden/make_layout_golden.py"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
PROBE = os.path.join(os.path.dirname(HERE), "layout_probe.cpp")


def run_probe(include_dir, header):
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "probe")
        subprocess.run(["g++", "-std=c++17", "-DPROBE_HEADER=\"%s\"" % header, "-I", include_dir, PROBE, "-o", exe], check=True)
        return subprocess.run([exe], check=True, capture_output=True, text=True).stdout


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/Include"
    out = run_probe(ref, "NRD.h")
    open(os.path.join(HERE, "nrd_layout.json"), "w").write(out)
    print("wrote nrd_layout.json (%d bytes)" % len(out))
