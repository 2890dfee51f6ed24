"""Regenerates tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference has no fixtures of its own and its classifier / assembly cannot be run here (HDK), so the
label / index / matrix vectors pin the *oracle's* behaviour at the time they were generated; the KATs in tests/test_oracle_kat.py
pin it analytically.  The `refcode_*` entries are different: they are OUTPUTS OF THE REFERENCE'S OWN CODE -- pcg.h and
ApplyPressureStressMatrix.h compiled from /root/reference (oracle/_ref/libps_ref_solve.so, `make -C oracle ref`) and run on the
matrices above: the operator applied to a fixed vector, and the iteration count / error / solution of pcg_external_matrix_A.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from polystokes_b200 import scenes  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from oracle import ref_solve, ref_classify  # noqa: E402

CASES = {
    "blob20_tile8_pad1": lambda: scenes.blob_scene(20, seed=21, tile=8, pad=1),
    "box16_uniform": lambda: scenes.box_scene(16, doReduced=0, tolerance=1e-6),
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


REF_BLOCKS = ("Mc", "McInv", "u", "uInv", "G", "Dt", "JG", "JDt")


def probe_vector(n):
    return np.random.default_rng(20261017).standard_normal(n)


def collect(sc):
    o = Oracle(sc, threads=1).setup()        # one thread: reproducible reductions
    out = {}
    for slot in range(7):
        out[f"labels{slot}"] = o.index_field(0, slot).astype(np.int8)
        out[f"active{slot}"] = o.index_field(1, slot).astype(np.int32)
        out[f"reduced{slot}"] = o.index_field(2, slot).astype(np.int32)
    for m in ("G", "Dt", "JG", "JDt"):
        shape, ptr, idx, val = o.csr(m)
        out[f"{m}_shape"] = np.array(shape)
        out[f"{m}_pattern_sha256"] = np.frombuffer(bytes.fromhex(digest(ptr) + digest(idx)), dtype=np.uint8)
        out[f"{m}_values"] = val
    out["b"] = o.vector("b")
    if ref_classify.available():
        # outputs of the reference's own classifier + constructMatrixBlocks (oracle/_ref/libps_ref_classify.so) on the oracle's weights
        R = ref_classify.RefClassifier(sc.nx, sc.ny, sc.nz, sc.dx, sc.dt, sc.params, o.weight_field, density=sc.density)
        fields, counts, valid = R.results()
        out["refcode_counts"] = np.array([counts[k] for k in ref_classify.COUNT_NAMES])
        out["refcode_fields_sha256"] = np.array([digest(np.ascontiguousarray(fields[kind][slot], dtype=np.int64)) for kind in range(3) for slot in range(7)])
        out["refcode_valid_sha256"] = np.array([digest(np.ascontiguousarray(v, dtype=np.float32)) for v in valid])
        R.construct_blocks(sc.vel, sc.colvel, sc.viscosity, o.vector("com").reshape(-1, 3), ref_classify.oracle_coeff_fn())
        for m in REF_BLOCKS:
            shape, ptr, idx, val = R.csr(m)
            out[f"refcode_{m}_pattern_sha256"] = np.array([digest(ptr.astype(np.int64)) + digest(idx.astype(np.int32))])
            out[f"refcode_{m}_values_sha256"] = np.array([digest(val.astype(np.float64))])
        R.close()
    if ref_solve.available():
        R = ref_solve.RefSolve(o.csr, sc.dt)
        out["refcode_probe"] = probe_vector(R.n)
        out["refcode_apply"] = R.apply(out["refcode_probe"])
        res, it, x, err, used = R.solve_spd(out["b"], sc.params["tolerance"], sc.params["maxIterations"])
        out["refcode_result"] = np.array([res, it, used])
        out["refcode_error"] = np.array([err])
        out["refcode_solution"] = x
    o.solve()
    out["oracle_solution"] = o.vector("solution")
    vel, valid = o.writeback()
    out["iterations"] = np.array([o.count("iterations")])
    for a in range(3):
        out[f"vel{a}"] = vel[a]
        out[f"valid{a}"] = valid[a].astype(np.uint8)
    return out


if __name__ == "__main__":
    for name, mk in CASES.items():
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **collect(mk()))
        print("wrote", name)
