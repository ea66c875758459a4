#!/usr/bin/env python
"""How often does a rounding-level change of operation order flip a step accept/reject decision?

The CPU oracle is compiled twice from the same source -- `-ffp-contract=off` (the checker) and `-ffp-contract=fast
-mfma` -- and both integrate the same 10 000 draws of every BASELINE workload.  The fraction of draws whose
(accepted, rejected) counts differ is the knife-edge rate to expect when the golden file from real diffrax (XLA
contracts multiplies and adds as it pleases) is compared with this repository's kernels: those draws can differ by
O(rtol) in the saved states without either side being wrong, all others agree to rounding.

    python scripts/knife_edge_rate.py [--draws 10000] > profiles/r2/knife_edge_rate.md
"""
import argparse
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--draws", type=int, default=10000)
    args = ap.parse_args()
    from oracle import oracle as orc
    from tests.cases import make_case
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "libdynode_oracle.so", "libdynode_oracle_fma.so"],
                   check=True)
    print("# Knife-edge rate: oracle with and without FMA contraction\n")
    print(f"{args.draws} draws per workload (tests/cases.py seeds), rtol=1e-5, atol=1e-6.\n")
    print("| workload | draws with different (accepted, rejected) | rate | max rel diff, same counts | max rel diff, "
          "different counts |")
    print("|---|---|---|---|---|")
    res = {}
    for variant in ("libdynode_oracle.so", "libdynode_oracle_fma.so"):
        orc._lib = None
        orc._LIB_PATH = os.path.join(ROOT, "oracle", variant)
        for name in ("sir_age2", "seirs_seasonal", "seirs_multi_a2s3", "seirs_multi_g6s3"):
            B = args.draws if name != "seirs_multi_g6s3" else max(1, args.draws // 4)
            case = make_case(name, B)
            fam, dims, theta, shared = case["oracle"]
            ys, _, st = orc.solve(fam, dims, case["y0"], theta, shared, t1=case["t1"])
            res[(variant, name)] = (ys, st)
    for name in ("sir_age2", "seirs_seasonal", "seirs_multi_a2s3", "seirs_multi_g6s3"):
        (y0, s0), (y1, s1) = res[("libdynode_oracle.so", name)], res[("libdynode_oracle_fma.so", name)]
        diff = np.any(s0[:, 1:3] != s1[:, 1:3], axis=1)
        scale = np.abs(y0).max(axis=(1, 2), keepdims=True)
        rel = (np.abs(y0 - y1) / (np.abs(y0) + 1e-9 * scale)).max(axis=(1, 2))
        same = rel[~diff].max() if (~diff).any() else 0.0
        other = rel[diff].max() if diff.any() else 0.0
        print(f"| {name} | {int(diff.sum())} / {diff.size} | {diff.mean():.2e} | {same:.1e} | {other:.1e} |")


if __name__ == "__main__":
    main()
