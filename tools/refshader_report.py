#!/usr/bin/env python
"""Table of tests/test_reference_shaders.py: per case and pass, how the oracle's pass compares with the reference's own shader source
compiled for the CPU (oracle/build_refshaders.py).  usage: python tools/refshader_report.py > profiles/rN_reference_shader_pin.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import test_reference_shaders as t
    print("# oracle pass vs the reference's own shader source (96x64 synthetic sequence, every dispatch, every written texture)")
    print("# case | shader | outputs compared | bit-identical | min fraction of texels within tolerance | min fraction of equal bytes | worst excess")
    identical = total = 0
    for name in sorted(t.CASES):
        stats, missing = t.run_case(name)
        for shader, s in sorted(stats.items()):
            total += 1
            identical += s["min_bytes_equal"] == 1.0
            print("%s | %s | %d | %s | %.5f | %.5f | %.1f" % (name, shader, s["outputs"], "yes" if s["min_bytes_equal"] == 1.0 else "no", s["min_fraction"], s["min_bytes_equal"], s["worst"]))
        if missing:
            print("%s | (no compiled reference shader: %s)" % (name, ", ".join(missing)))
    print("# %d of %d (case, shader) pairs are bit-identical" % (identical, total))
    print()
    print("# REBLUR and RELAX with the oracle built in the reference's association order (liboracle_src.so: oracle/reblur.cpp curvature, oracle/relax.cpp world positions)")
    identical = total = 0
    for name in sorted(n for n in t.CASES if n.startswith(("relax", "reblur"))):
        stats, _ = t.run_case(name, variant="src")
        for shader, s in sorted(stats.items()):
            total += 1
            identical += s["min_bytes_equal"] == 1.0
            print("%s | %s | %d | %s | %.5f | %.5f | %.1f" % (name, shader, s["outputs"], "yes" if s["min_bytes_equal"] == 1.0 else "no", s["min_fraction"], s["min_bytes_equal"], s["worst"]))
    print("# %d of %d (case, shader) pairs are bit-identical" % (identical, total))


if __name__ == "__main__":
    main()
