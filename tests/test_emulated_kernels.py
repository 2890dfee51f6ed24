"""CPU-only check of the kernel logic: the product sources compiled with -DPS_EMULATE (every per-thread
kernel body executed serially by g++) against the oracle.  This is NOT a product path -- the python package
never loads libpolystokes_emul.so; it exists so index/stencil bugs are caught on machines without a GPU."""
import pytest

import parity


@pytest.mark.parametrize("case", list(parity.SCENE_CASES))
def test_emulated_step_matches_oracle(built, case):
    parity.run_case(case, lib_path=parity.EMUL_LIB)


def test_emulated_bicgstab_fallback(built):
    parity.check_bicgstab_fallback(lib_path=parity.EMUL_LIB)


@pytest.mark.parametrize("warm", [1, 0])
def test_emulated_eigen_cg(built, warm):
    parity.check_eigen_cg(lib_path=parity.EMUL_LIB, warm=warm)


def test_emulated_eigen_cg_uniform(built):
    from polystokes_b200 import scenes
    parity.check_eigen_cg(lib_path=parity.EMUL_LIB, scene=scenes.box_scene(20, doReduced=0, tolerance=1e-6))


@pytest.mark.parametrize("case", list(parity.EXPLICIT_CASES))
def test_emulated_explicit_A(built, case, tmp_path):
    parity.check_explicit_A(case, lib_path=parity.EMUL_LIB, tmpdir=tmp_path)
