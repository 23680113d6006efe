"""Soft pin of the FGR oracle (oracle/fgr_oracle.c) against the reference's shipped poses.

Run in the BUILD container only (it reads /root/reference):

    python oracle/pin_fgr_against_goldens.py [--stride 15]

The reference ships, per consecutive NCLT pair, the pose its own registro_FGR produced (relative_poses_FGR/NCLT, %.10f)
and the pose after the M-GICP refinement (relative_poses_FGR_GICP/NCLT).  FGR's tuple test is random, so two FGR runs of
the same pair differ by centimetres; what can be checked is that the oracle's FGR is as good a coarse alignment as the
reference's: the distance of both to the refined pose (the best available stand-in for the truth) is recorded in
tests/golden/nclt_fgr_pin.json and summarised in DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "point-cloud-registration-with-global-refinement_b200"))
import pcd_io  # noqa: E402
import synthetic  # noqa: E402

REF = "/root/reference"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stride", type=int, default=15)
    a = ap.parse_args()
    pairs = sorted(set(range(0, 900, a.stride)) | {0, 17})
    rows = []
    t0 = time.time()
    for i in pairs:
        tgt = pcd_io.read_pcd_xyz(f"{REF}/nuvens/nuvens_pre_processadas/NCLT/s{i}.pcd").astype(np.float64)
        src = pcd_io.read_pcd_xyz(f"{REF}/nuvens/nuvens_pre_processadas/NCLT/s{i + 1}.pcd").astype(np.float64)
        T_fgr = pcd_io.read_pose(f"{REF}/relative_poses_FGR/NCLT/pose_{i + 1}_{i}.txt")
        T_ref = pcd_io.read_pose(f"{REF}/relative_poses_FGR_GICP/NCLT/pose_{i + 1}_{i}.txt")
        T, nc = oracle.registro_FGR(src, tgt, 0.1, seed=i)
        ro, to = synthetic.pose_error(T, T_ref)
        rs, ts = synthetic.pose_error(T_fgr, T_ref)
        rows.append({"pair": i, "n_src": len(src), "n_tgt": len(tgt), "n_corres": nc, "oracle_vs_refined_rad": ro, "oracle_vs_refined_m": to,
                     "shipped_vs_refined_rad": rs, "shipped_vs_refined_m": ts})
        print(f"pair {i + 1}->{i}: oracle FGR {to:.3f} m / {ro:.4f} rad from the refined pose; shipped FGR {ts:.3f} m / {rs:.4f} rad", flush=True)
    q = lambda k, p: float(np.percentile([r[k] for r in rows], p))
    summary = {"pairs": len(rows), "seconds": time.time() - t0,
               "oracle_m_p50": q("oracle_vs_refined_m", 50), "oracle_m_p90": q("oracle_vs_refined_m", 90),
               "shipped_m_p50": q("shipped_vs_refined_m", 50), "shipped_m_p90": q("shipped_vs_refined_m", 90),
               "oracle_rad_p50": q("oracle_vs_refined_rad", 50), "oracle_rad_p90": q("oracle_vs_refined_rad", 90),
               "shipped_rad_p50": q("shipped_vs_refined_rad", 50), "shipped_rad_p90": q("shipped_vs_refined_rad", 90)}
    print(json.dumps(summary, indent=1))
    json.dump({"summary": summary, "rows": rows}, open(os.path.join(ROOT, "tests", "golden", "nclt_fgr_pin.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
