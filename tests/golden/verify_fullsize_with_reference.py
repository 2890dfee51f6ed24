"""Checks the committed full-size digests (tests/golden/fullsize_*.json, generated from the oracle by make_fullsize.py) against the
COMPILED REFERENCE SOLVER (oracle/_ref/libps_ref_full.so): counts, SHA-256 of every label / DOF-index / reduced-index / weight field, of
the CSR patterns and values of G and D^T, and |b|.  Setup only (the reference's serial CG at these sizes takes tens of minutes; its solve
is compared on smaller scenes in tests/test_ref_full.py).  Run from the repo root where /root/reference exists:
    python tests/golden/verify_fullsize_with_reference.py [S1 S2 ...]   > profiles/<log>
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_fullsize as mf  # noqa: E402
from oracle.ref_full import RefFull  # noqa: E402


def main():
    names = sys.argv[1:] or ["S1", "S2", "S5_half_tile16", "S4_192"]
    ok_all = True
    for name in names:
        with open(os.path.join(ROOT, "tests", "golden", f"fullsize_{name}.json")) as f:
            g = json.load(f)
        sc = mf.CASES[name]()
        t0 = time.time()
        saved = os.dup(1); devnull = os.open(os.devnull, os.O_WRONLY); os.dup2(devnull, 1)
        try:
            R = RefFull(sc).setup()
        finally:
            os.dup2(saved, 1); os.close(saved); os.close(devnull)
        counts = {k: R.count(k) for k in mf.COUNTS if k not in ("nFaceX", "nFaceY", "nFaceZ", "nEdgeYZ", "nEdgeXZ", "nEdgeXY") or True}
        fields = mf.field_digests(R.index_field, R.weight_field)
        csr = mf.csr_digests(R.csr)
        b = R.vector("b")
        bad = [k for k in mf.COUNTS if counts[k] != g["counts"][k]]
        bad += [k for k in fields if fields[k] != g["fields"][k]]
        bad += [f"csr {m} {q}" for m in csr for q in ("shape", "nnz", "pattern", "values") if csr[m][q] != g["csr"][m][q]]
        bn = float(np.sqrt(np.dot(b, b)))
        if abs(bn - g["b_norm"]) > 1e-12 * g["b_norm"]:
            bad.append(f"b_norm {bn} vs {g['b_norm']}")
        ok_all &= not bad
        print(f"{name}: {sc.nx}x{sc.ny}x{sc.nz}, n={counts['nSystemSize']}, regions={counts['regionCount']}: compiled reference setup in {time.time() - t0:.0f} s -> "
              + ("ALL DIGESTS EQUAL (14 counts, 35 field hashes, G / D^T patterns + values, |b|)" if not bad else f"MISMATCH: {bad}"), flush=True)
        R.close()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
