"""Analytic known-answer tests of the CPU oracle.  (The reference ships no tests / golden vectors; its whole solver is compiled here from its
own sources -- tests/test_ref_full.py -- and the oracle is bit-equal to it, modulo the stand-ins' definitions of the HDK primitives.  These
tests check the properties that do not depend on any reference build.)  Each test checks an analytic property of the restated algorithm."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.oracle import Oracle
from polystokes_b200 import scenes


def test_basis_is_divergence_free(built):
    """D1 (S.cpp:2107-2149): every one of the 26 basis fields v_n = (c_x[n], c_y[n], c_z[n]) has zero divergence."""
    rng = np.random.default_rng(0)
    h = 1e-3
    for _ in range(5):
        p = rng.uniform(-1, 1, 3)
        div = np.zeros(26)
        for a in range(3):
            e = np.zeros(3); e[a] = h
            div += (orc.conversion_coefficients(p + e, a) - orc.conversion_coefficients(p - e, a)) / (2 * h)
        assert np.abs(div).max() < 1e-9


def test_basis_nonzero_pattern(built):
    c = [orc.conversion_coefficients([0.3, -0.2, 0.7], a) for a in range(3)]
    assert [int((x != 0).sum()) for x in c] == [10, 10, 14]
    assert c[0][0] == 1 and c[1][1] == 1 and c[2][2] == 1


def test_partial_piv_inverse_and_full_piv_solve(built):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((26, 26)); A = A @ A.T + 0.1 * np.eye(26)
    Ai = orc.inverse_partial_piv_lu(A)
    assert np.abs(Ai @ A - np.eye(26)).max() < 1e-9
    b = rng.standard_normal(26)
    x, rank = orc.solve_full_piv_lu(A, b)
    assert rank == 26 and np.abs(x - np.linalg.solve(A, b)).max() < 1e-9
    # rank deficient but consistent: Eigen's fullPivLu().solve returns *a* solution (free variables 0)
    V = rng.standard_normal((26, 20)); S = V @ V.T
    bs = S @ rng.standard_normal(26)
    xs, rank = orc.solve_full_piv_lu(S, bs)
    assert rank == 20 and np.abs(S @ xs - bs).max() < 1e-8 * np.abs(bs).max()


def test_weights_of_a_plane(built):
    """Shim of computeSDFWeightsSampled: a liquid half space x < 6.5 cells gives exact eighths."""
    sc = scenes.box_scene(12, doReduced=0)
    x = (np.arange(12) + 0.5).reshape(1, 1, 12) / 12.0
    sc.surface[...] = np.broadcast_to(x - 6.5 / 12.0, sc.surface.shape).astype(np.float32)
    o = Oracle(sc); o.build_weights()
    w = o.weight_field(1, 0)          # centre weights
    assert np.all(w[:, :, :6] == 1.0) and np.all(w[:, :, 6] == 0.5) and np.all(w[:, :, 7:] == 0.0)
    wx = o.weight_field(1, 1)         # face-x weights: face i sits at x = i cells
    assert np.all(wx[:, :, :7] == 1.0) and np.all(wx[:, :, 7:] == 0.0)


def test_active_indices_follow_tile_order(built):
    """C11 (S_Cls:1738-1770): running counter in UT_VoxelArray order = 16^3 tiles x->y->z, x fastest inside."""
    sc = scenes.blob_scene((36, 20, 40), seed=2, doReduced=0)
    o = Oracle(sc).setup()
    for slot in range(7):
        lab = o.index_field(0, slot); idx = o.index_field(1, slot)
        rz, ry, rx = lab.shape
        n = 0
        for tk in range(0, rz, 16):
            for tj in range(0, ry, 16):
                for ti in range(0, rx, 16):
                    blk_l = lab[tk:tk + 16, tj:tj + 16, ti:ti + 16]; blk_i = idx[tk:tk + 16, tj:tj + 16, ti:ti + 16]
                    act = (blk_l == -4) | (blk_l == -9)
                    cnt = int(act.sum())
                    assert np.array_equal(blk_i[act], np.arange(n, n + cnt))   # C-order of the block = x fastest, then y, then z
                    assert np.all(blk_i[~act] == -1)
                    n += cnt


@pytest.mark.parametrize("mk", [lambda: scenes.blob_scene(24, tile=8, pad=1), lambda: scenes.box_scene(24, doReduced=0)])
def test_operator_symmetric_negative_definite_and_matches_explicit_A(built, mk):
    import scipy.sparse as sp
    sc = mk()
    o = Oracle(sc).setup(); o.assemble_explicit_A()
    A = o.scipy_csr("A"); n = A.shape[0]
    rng = np.random.default_rng(3); x = rng.standard_normal(n); y = rng.standard_normal(n)
    Ax, Ay = o.apply(x), o.apply(y)
    assert abs(y @ Ax - x @ Ay) <= 1e-12 * abs(y @ Ax)            # <y, Ax> = <Ay, x>
    assert x @ Ax < 0                                               # negative definite as coded
    assert np.abs(A @ x - Ax).max() <= 1e-12 * np.abs(Ax).max()     # explicit A (S_AS:381-397) == factored apply (Apply.h:102-179)
    assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
    # scipy second opinion on the factored form
    G, Dt, JG, JDt = (o.scipy_csr(k) for k in ("G", "Dt", "JG", "JDt"))
    McInv, uInv, Bi = o.scipy_csr("McInv"), o.scipy_csr("uInv"), o.scipy_csr("BInv")
    K = sp.hstack([G, Dt]).tocsr(); J = sp.hstack([JG, JDt]).tocsr(); np_ = G.shape[1]
    ref = -sc.dt * (K.T @ (McInv @ (K @ x))) - J.T @ (Bi @ (J @ x)); ref[np_:] -= 0.5 * (uInv @ x[np_:])
    assert np.abs(ref - Ax).max() <= 1e-12 * np.abs(Ax).max()
    B = o.scipy_csr("B")
    if B.shape[0]:
        assert np.abs((B @ Bi).toarray() - np.eye(B.shape[0])).max() < 1e-8


def test_least_squares_reproduces_polynomial_velocity(built):
    """D3 (S.cpp:374-417): if u* is itself a member of the 26-dim space, the fit returns its coefficients."""
    sc = scenes.box_scene(40, tileSize=16, tilePadding=2, liquid_hi_frac=0.8)
    o = Oracle(sc); o.build_weights(); o.classify()
    R = o.count("regionCount")
    assert R >= 1
    o2 = Oracle(sc).setup()
    com = o2.vector("com").reshape(R, 3)
    q = np.random.default_rng(5).standard_normal(26) * 0.1
    # overwrite u* with the polynomial field of region 0's frame, re-run
    for a in range(3):
        shp = sc.vel[a].shape
        kk, jj, ii = np.meshgrid(np.arange(shp[0]), np.arange(shp[1]), np.arange(shp[2]), indexing="ij")
        pos = [ii.astype(float), jj.astype(float), kk.astype(float)]
        pos[a] = pos[a] - 0.5
        off = [pos[d] * sc.dx - com[0, d] for d in range(3)]
        # vectorised basis evaluation
        ox, oy, oz = off
        if a == 0:
            v = q[0] + q[3] * ox + q[4] * oy + q[5] * oz + q[6] * ox * ox + q[7] * ox * oy + q[8] * ox * oz + q[9] * oy * oy + q[10] * oy * oz + q[11] * oz * oz
        elif a == 1:
            v = q[1] + q[12] * ox + q[13] * oy + q[14] * oz + q[15] * ox * ox + q[16] * ox * oy + q[17] * ox * oz + q[18] * oy * oy + q[19] * oy * oz + q[20] * oz * oz
        else:
            v = (q[2] - q[3] * oz - 2 * q[6] * ox * oz - q[7] * oy * oz - 0.5 * q[8] * oz * oz - q[13] * oz - q[16] * ox * oz
                 - 2 * q[18] * oy * oz - 0.5 * q[19] * oz * oz + q[21] * ox + q[22] * oy + q[23] * ox * ox + q[24] * ox * oy + q[25] * oy * oy)
        sc.vel[a][...] = v.astype(np.float32)
    o3 = Oracle(sc).setup()
    fit = o3.vector("bestFit").reshape(R, 26)[0]
    # the fitted polynomial must reproduce the field on the region's surface faces: compare through M_r v (rhs_r), a linear image
    Mr = o3.vector("MrDense").reshape(R, 26, 26)[0]
    assert np.abs(Mr @ fit - Mr @ q).max() <= 2e-4 * np.abs(Mr @ q).max()   # fp32 velocity input limits the accuracy


def test_marketio_format(built, tmp_path):
    """Eigen saveMarket / saveMarketVector text format (MarketIO.h:311-372)."""
    sc = scenes.box_scene(12, doReduced=0)
    o = Oracle(sc).setup()
    p = str(tmp_path / "G.mtx"); o.save_csr("G", p)
    lines = open(p).read().splitlines()
    assert lines[0] == "%%MatrixMarket matrix coordinate  real general"
    shape, ptr, idx, val = o.csr("G")
    assert lines[1] == f"{shape[0]} {shape[1]} {len(val)}"
    r, c, v = lines[2].split()
    assert int(r) == 1 + int(np.searchsorted(ptr, 0, side="right") - 1) and int(c) == idx[0] + 1
    assert v == "%.17e" % val[0]
    p = str(tmp_path / "b.mtx"); o.save_vector("b", p)
    lines = open(p).read().splitlines()
    assert lines[0] == "%%MatrixMarket matrix array real general" and lines[1] == f"{o.count('nSystemSize')} 1"


def test_cg_solves_the_system(built):
    sc = scenes.blob_scene(24, tile=8, pad=1, tolerance=1e-8)
    o = Oracle(sc).setup()
    assert o.solve() == 1
    b, x = o.vector("b"), o.vector("solution")
    assert np.linalg.norm(o.apply(x) - b) <= 1e-6 * np.linalg.norm(b)
