"""BASELINE.json configurations at their full sizes on the GPU.

(1) Against the committed oracle digests (tests/golden/fullsize_*.json, made by tests/golden/make_fullsize.py): counts, the SHA-256 of
    every weight / label / DOF-index field and of the CSR of G and D^T (pattern AND values: these are bit-equal), |b|, the CG
    iteration count and a strided sample of the solved velocity.
(2) Size-independent properties of the solve where the oracle cannot run in this container (S4 384^3, S5 512x256x256):
    the DOF numbering is a bijection, the operator is symmetric and linear, the returned x satisfies the reference's
    stop test when the residual is recomputed from scratch through ps_apply, valid == "face is in the system",
    and a second step on the same handle reproduces the first bit for bit.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_fullsize  # noqa: E402
import parity  # noqa: E402

pytestmark = pytest.mark.gpu


def _digest_path(name):
    return os.path.join(ROOT, "tests", "golden", f"fullsize_{name}.json")


def _check_numbering(s, n_expected):
    """serialAssignFieldIndices (S_Cls:1738-1770): per slot the active indices are 0..count-1, each exactly once, increasing in
    voxel-tile order is checked by the digests; here: bijection + labels consistent with the indices."""
    names = ["nCenter", "nFaceX", "nFaceY", "nFaceZ", "nEdgeYZ", "nEdgeXZ", "nEdgeXY"]
    for slot in range(7):
        idx = s.index_field(1, slot).ravel()
        lab = s.index_field(0, slot).ravel()
        act = idx[idx >= 0]
        assert act.size == s.count(names[slot]), f"slot {slot}: {act.size} active indices vs count {s.count(names[slot])}"
        seen = np.zeros(act.size, dtype=np.uint8)
        seen[act] = 1
        assert seen.all(), f"slot {slot}: active indices are not a bijection onto 0..count-1"
        red = s.index_field(2, slot).ravel()
        assert not np.any((idx >= 0) & (red >= 0) & (slot < 4)), "a centre / face sample is both active and reduced"
        del idx, lab, act, seen, red


def _check_operator_properties(s, seed=3):
    n = s.count("nSystemSize")
    rng = np.random.default_rng(seed)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Ax, Ay = s.apply(x), s.apply(y)
    sym = abs(np.dot(y, Ax) - np.dot(x, Ay)) / max(abs(np.dot(y, Ax)), 1e-300)
    assert sym <= 1e-10, f"<y,Ax> vs <x,Ay>: {sym:.2e}"
    assert np.dot(x, Ax) * np.dot(y, Ay) > 0, "the system matrix handed to CG must be definite (same sign of every Rayleigh quotient)"
    lin = parity.rel(s.apply(2.0 * x - 0.5 * y), 2.0 * Ax - 0.5 * Ay)
    assert lin <= 1e-12, f"linearity: {lin:.2e}"
    return n


def _check_stop_test(s, tol, converged=True):
    """pcg.h:316-325 recomputed from scratch: r = b - A x through ps_apply, min(|r|^2, |r|^2/|x|^2) < tol^2 (with a little slack
    for the recursive residual the loop itself tests)."""
    x, b = s.vector("solution"), s.vector("b")
    r = b - s.apply(x)
    rr, xx = float(np.dot(r, r)), float(np.dot(x, x))
    rre = min(rr, rr / xx) if xx > 0 else rr
    if converged:
        assert rre < (1.05 * tol) ** 2, f"true residual test {np.sqrt(rre):.3e} vs tol {tol:.1e}"
    assert abs(np.sqrt(rre) - s.real("solveError")) <= 0.05 * s.real("solveError") + 1e-14, f"{np.sqrt(rre):.6e} vs reported {s.real('solveError'):.6e}"


def _step_on_device(s, sc):
    import torch
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ins = (d(sc.surface), d(sc.collision), d(sc.viscosity), [d(v) for v in sc.vel], [d(v) for v in sc.colvel])
    vout = [v.clone() for v in ins[3]]
    valid = [torch.zeros_like(v) for v in ins[3]]
    rc = s.step(*ins, vout, valid)
    torch.cuda.synchronize()
    return rc, ins, vout, valid


@pytest.mark.parametrize("name", list(make_fullsize.CASES))
def test_gpu_matches_fullsize_oracle_digest(built, name):
    if not os.path.exists(_digest_path(name)):
        pytest.skip(f"no committed digest for {name}")
    from polystokes_b200 import PolyStokesSolver
    g = json.load(open(_digest_path(name)))
    sc = make_fullsize.CASES[name]()
    assert [sc.nx, sc.ny, sc.nz] == g["res"]
    s = PolyStokesSolver.from_scene(sc)
    rc, ins, vout, valid = _step_on_device(s, sc)
    assert rc == g["result"]
    for k, v in g["counts"].items():
        assert s.count(k) == v, f"count {k}: oracle {v} vs {s.count(k)}"
    mine = make_fullsize.field_digests(s.index_field, s.weight_field)
    for k, v in g["fields"].items():
        assert mine[k] == v, f"{name}: field {k} differs from the oracle"
    mine = make_fullsize.csr_digests(s.csr)
    for m, v in g["csr"].items():
        assert mine[m]["shape"] == v["shape"] and mine[m]["nnz"] == v["nnz"], f"{m} shape / nnz"
        assert mine[m]["pattern"] == v["pattern"], f"{name}: {m} sparsity pattern differs from the oracle"
        assert mine[m]["values"] == v["values"], f"{name}: {m} values are not bit-equal to the oracle's"
    b = s.vector("b")
    assert abs(float(np.sqrt(np.dot(b, b))) - g["b_norm"]) <= 1e-10 * g["b_norm"]
    it = s.count("iterations")
    assert abs(it - g["iterations"]) <= max(2, int(0.01 * g["iterations"])), f"iterations oracle {g['iterations']} vs {it}"
    tol = max(10 * sc.params["tolerance"], 4e-7)
    got = make_fullsize.velocity_sample([v.cpu().numpy() for v in vout], [v.cpu().numpy() for v in valid])
    dist = lambda A, B, a: float(np.abs(np.array(A[f"vel{a}_sample"]) - np.array(B[f"vel{a}_sample"])).max()) / g[f"vel{a}_absmax"]
    direct = [dist(g, got, a) for a in range(3)]
    for a in range(3):
        assert got[f"valid{a}"] == g[f"valid{a}"], f"valid field axis {a}"
    _check_stop_test(s, sc.params["tolerance"])
    s.close()
    if max(direct) > tol:
        # Long, ill-conditioned solves (S4: mu = 2000, > 1000 iterations): two CG runs that both pass the reference's stop test
        # min(|r|^2, |r|^2/|x|^2) < tol^2 still differ by (condition number) x tol in the velocity.  The gate then is: the oracle's
        # field is as close to the CONVERGED field (same system solved here to tol / 1000) as our own tol-run is, within 3x.
        tight = PolyStokesSolver.from_scene(sc, tolerance=sc.params["tolerance"] * 1e-3)
        rc2, _, vout2, valid2 = _step_on_device(tight, sc)
        assert rc2 == 1
        ref = make_fullsize.velocity_sample([v.cpu().numpy() for v in vout2], [v.cpu().numpy() for v in valid2])
        tight.close()
        for a in range(3):
            mine, theirs = dist(got, ref, a), dist(g, ref, a)
            assert theirs <= 3.0 * max(mine, tol), f"velocity axis {a}: oracle is {theirs:.2e} from the converged field, this library {mine:.2e}"
            assert mine <= 100 * tol, f"velocity axis {a}: {mine:.2e} from the converged field"


FULL = {
    # name -> (scene factory, overrides): sizes the oracle cannot reach in this container; CG capped so the test stays short
    "S3_256": lambda: __import__("polystokes_b200").scenes.scene_s3(256),
    "S5_512x256x256_tile16": lambda: __import__("polystokes_b200").scenes.scene_s5(1.0, tileSize=16, maxIterations=400, keepNonConvergedResults=1),
    "S5_512x256x256_tile32": lambda: __import__("polystokes_b200").scenes.scene_s5(1.0, tileSize=32, maxIterations=200, keepNonConvergedResults=1),
    "S4_384": lambda: __import__("polystokes_b200").scenes.scene_s4(384, maxIterations=150, keepNonConvergedResults=1),
}


@pytest.mark.parametrize("name", list(FULL))
def test_gpu_fullsize_properties(built, name):
    import torch
    from polystokes_b200 import PolyStokesSolver
    sc = FULL[name]()
    s = PolyStokesSolver.from_scene(sc)
    rc, ins, vout, valid = _step_on_device(s, sc)
    assert rc in (0, 1)
    n = s.count("nSystemSize")
    assert n == s.count("nPressures") + s.count("nStresses") and n > 0
    _check_numbering(s, n)
    _check_operator_properties(s)
    converged = (rc == 1) and not s.count("usedBiCGStab")
    if not s.count("usedBiCGStab"):
        _check_stop_test(s, sc.params["tolerance"], converged=converged)
    # valid (S_Cls:4-54): 1 exactly where the face label is neither UNSOLVED nor UNASSIGNED, i.e. where the face has a DOF or is SOLID
    for a in range(3):
        v = valid[a].cpu().numpy()
        assert set(np.unique(v).tolist()) <= {0.0, 1.0}
        has_dof = (s.index_field(1, 1 + a) >= 0) | (s.index_field(2, 1 + a) >= 0)
        assert np.all(v[has_dof] == 1.0), f"axis {a}: a face with a DOF is not valid"
        assert np.isfinite(vout[a].cpu().numpy()).all()
    # a second step on the same handle (buffers reused, tickets healed) reproduces the first bit for bit
    it1 = s.count("iterations")
    vout2 = [v.clone() for v in ins[3]]
    rc2 = s.step(*ins, vout2, None)
    torch.cuda.synchronize()
    assert rc2 == rc and s.count("iterations") == it1
    for a in range(3):
        assert torch.equal(vout[a], vout2[a]), f"axis {a}: repeated step differs"
    s.close()
