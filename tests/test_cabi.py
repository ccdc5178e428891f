"""The C-ABI library loads and exports every symbol include/ipp.h declares; without a GPU the product path
fails loudly instead of falling back to the CPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, cuda_available, mesh_path, specs


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ipp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ipp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    from inspection_path_planning_b200 import build
    return build.build()


def test_header_symbols_are_exported(built):
    lib = ctypes.CDLL(built)
    names = _declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_prototypes_cover_the_header(built):
    from inspection_path_planning_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared_symbols()
    _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "inspection_path_planning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), os.path.join(dirpath, f)


@pytest.mark.skipif(cuda_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu(built):
    from inspection_path_planning_b200 import cpp_binding
    with pytest.raises(RuntimeError, match="no CUDA device"):
        cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), specs(64))


def test_module_surface_matches_swig_binding():
    # plan_optimizer/cpp_wrapper/cpp_binding.i:20-39
    from inspection_path_planning_b200 import cpp_binding as cb
    assert (cb.normal, cb.circulating, cb.image, cb.nothing) == (0, 1, 2, 3)
    for name in ("evaluatePlan", "getNumberOfBoxes", "updateMemoisedEdges", "getSimplePlans", "storePlanImage",
                 "getMaxAllowedEnergy", "interpretAndExportPlan", "evaluatePopulation"):
        assert callable(getattr(cb.PlanCoverageEstimator, name))
    assert callable(cb.viewMatrixSelector) and cb.Line and cb.Array and cb.StringVector
