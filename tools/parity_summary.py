#!/usr/bin/env python
"""Condenses the parity reports the GPU tests write to gpurun_out/ (parity_*.json: one record per frame x pass x output texture,
sequence_*.json: end-to-end fractions) into one table: per report and shader the worst fraction of texels within tolerance, the
largest number of outliers (texels beyond 10x tolerance) against its budget, and non-finite texels.
usage: python tools/parity_summary.py [gpurun_out] > profiles/rN_parity_summary.txt"""
import glob
import json
import os
import sys


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    print("# per-pass parity (tests/parity.py: REL 1e-3 + ABS 1e-4, >= 99.9 %% of texels, outliers beyond 10x tolerance budgeted at 5e-5)")
    print("# report | shader | records | min fraction | required | max outliers / budget | non-finite")
    for path in sorted(glob.glob(os.path.join(src, "parity_*.json"))):
        try:
            recs = json.load(open(path))
        except ValueError:
            continue
        if not isinstance(recs, list) or not recs or "shader" not in recs[0]:
            continue
        name = os.path.basename(path)[len("parity_"):-len(".json")]
        by = {}
        for r in recs:
            if r["shader"].startswith("Clear_"):
                continue
            b = by.setdefault(r["shader"], dict(n=0, frac=1.0, req=1.0, out=0, budget=0, nonfinite=0, texels=0))
            b["n"] += 1
            if r["fraction"] <= b["frac"]:
                b["frac"], b["req"] = r["fraction"], r.get("min_fraction", 0.999)
            if r.get("outliers", 0) >= b["out"]:
                b["out"], b["budget"] = r.get("outliers", 0), r.get("outlier_budget", 0)
            b["nonfinite"] += r.get("nonfinite", 0)
            b["texels"] = max(b["texels"], r.get("texels", 0))
        for shader, b in sorted(by.items()):
            print("%s | %s | %d | %.6f | %.4f | %d / %d | %d" % (name, shader, b["n"], b["frac"], b["req"], b["out"], b["budget"], b["nonfinite"]))
    print()
    print("# sequence parity (independent end-to-end runs; fraction of texels within tolerance, PSNR dB[, FMA-build noise floor])")
    for path in sorted(glob.glob(os.path.join(src, "sequence_*.json"))):
        try:
            res = json.load(open(path))
        except ValueError:
            continue
        print("%s | %s" % (os.path.basename(path)[len("sequence_"):-len(".json")], json.dumps(res)))


if __name__ == "__main__":
    main()
