"""Full-size oracle digests for the BASELINE.json configurations (run from the repo root:
    python tests/golden/make_fullsize.py [S1 S2 S3 S5a ...]).

The oracle needs minutes and several GB at these sizes, so the GPU tests cannot run it on the box; instead this script
runs it once here and commits a small JSON digest per scene (tests/golden/fullsize_<name>.json):
  * every count of the DOF numbering,
  * a SHA-256 of every label / active-index / reduced-index field (int8 / int32, x fastest) and of every weight field (eighths as uint8),
  * a SHA-256 of the CSR patterns of G and D^T (rowptr int64 + colidx int32) and of their values (bit-equal on the GPU),
  * |b|, the CG iteration count, the reference stop-test error, and a strided sample of the solved velocity fields.
tests/test_gpu_fullsize.py recomputes the same digests from the CUDA path.  The digests come from the oracle (fast, OpenMP);
tests/golden/verify_fullsize_with_reference.py re-derives the setup part (counts, field hashes, G / D^T, |b|) with the COMPILED
REFERENCE SOLVER (oracle/_ref/libps_ref_full.so) and finds them equal -- log in profiles/r01_fullsize_digests_vs_compiled_reference.log.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from polystokes_b200 import scenes  # noqa: E402

COUNTS = ["nCenter", "nFaceX", "nFaceY", "nFaceZ", "nEdgeYZ", "nEdgeXZ", "nEdgeXY", "regionCount", "nSystemSize",
          "nActiveVs", "nReducedVs", "nPressures", "nStresses", "nTotalDOFs"]
SAMPLE_STRIDE = 4099          # prime: the sample walks through every x / y / z phase

CASES = {
    # BASELINE.json configs[0..2] at their full sizes, configs[4] (tile 16) at half scale (the oracle's explicit JG / JD^T
    # at 512x256x256 do not fit this container), configs[3] at 192^3 for the same reason
    "S1": lambda: scenes.scene_s1(),
    "S2": lambda: scenes.scene_s2(128),
    "S3": lambda: scenes.scene_s3(256),
    "S4_192": lambda: scenes.scene_s4(192),
    "S5_half_tile16": lambda: scenes.scene_s5(0.5, tileSize=16),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def field_digests(index_field, weight_field):
    """index_field(kind, slot) / weight_field(liquid, slot) -> arrays; the dtypes are normalised before hashing."""
    out = {}
    for slot in range(7):
        out[f"labels{slot}"] = sha(index_field(0, slot).astype(np.int8))
        out[f"active{slot}"] = sha(index_field(1, slot).astype(np.int32))
        out[f"reduced{slot}"] = sha(index_field(2, slot).astype(np.int32))
        for liquid in (1, 0):
            w8 = np.rint(weight_field(liquid, slot).astype(np.float64) * 8.0).astype(np.uint8)
            out[f"weights{slot}_{'liquid' if liquid else 'fluid'}"] = sha(w8)
    return out


def csr_digests(csr):
    out = {}
    for m in ("G", "Dt"):
        shape, ptr, idx, val = csr(m)
        out[m] = {"shape": [int(shape[0]), int(shape[1])], "nnz": int(idx.size),
                  "pattern": sha(ptr.astype(np.int64)) + sha(idx.astype(np.int32)), "values": sha(val.astype(np.float64))}
    return out


def velocity_sample(vel, valid):
    out = {}
    for a in range(3):
        flat = np.ascontiguousarray(vel[a]).ravel()
        out[f"vel{a}_sample"] = [float(v) for v in flat[::SAMPLE_STRIDE]]
        out[f"vel{a}_absmax"] = float(np.abs(flat).max())
        out[f"valid{a}"] = sha(np.ascontiguousarray(valid[a]).astype(np.uint8))
    return out


def collect(sc):
    from oracle.oracle import Oracle
    t0 = time.time()
    o = Oracle(sc).setup()
    d = {"scene": sc.name, "res": [sc.nx, sc.ny, sc.nz], "params": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in sc.params.items()}}
    d["counts"] = {k: o.count(k) for k in COUNTS}
    d["fields"] = field_digests(o.index_field, o.weight_field)
    d["csr"] = csr_digests(o.csr)
    b = o.vector("b")
    d["b_norm"] = float(np.sqrt(np.dot(b, b)))
    d["b_absmax"] = float(np.abs(b).max()) if b.size else 0.0
    d["setup_seconds"] = round(time.time() - t0, 1)
    t0 = time.time()
    d["result"] = int(o.solve())
    d["iterations"] = o.count("iterations")
    d["solveError"] = o.real("solveError")
    vel, valid = o.writeback()
    d.update(velocity_sample(vel, valid))
    d["solve_seconds"] = round(time.time() - t0, 1)
    return d


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for name in names:
        t0 = time.time()
        d = collect(CASES[name]())
        with open(os.path.join(ROOT, "tests", "golden", f"fullsize_{name}.json"), "w") as f:
            json.dump(d, f)
        print(f"wrote fullsize_{name}.json: n={d['counts']['nSystemSize']} regions={d['counts']['regionCount']} iterations={d['iterations']} "
              f"({time.time() - t0:.0f} s)", flush=True)
