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


def test_mesh_ingest_reproduces_the_fixtures_byte_for_byte():
    """csrc/mesh_io.cpp (the product's replacement of osgDB::readNodeFile, SceneKeeper.cpp:73-86) on the reference's real assets
    -- a binary STL of 26 707 facets, a Blender OBJ of triangles and the quad-faced OBJ sphere, vendored under
    tests/golden/assets/ -- against the .ippm fixtures every other test and bench.py load (written by the oracle's loader,
    tests/golden/make_fixtures.py): same triangles, same order, same bytes.  Host code only: runs without a GPU."""
    import numpy as np
    from inspection_path_planning_b200 import cpp_binding
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for asset, fixture, T in (("mockup_only.stl", "mockup_only.ippm", 26707), ("luis1.obj", "luis1.ippm", 13925), ("sphere.obj", "sphere.ippm", 960)):
        tris = cpp_binding.load_mesh(os.path.join(root, "tests", "golden", "assets", asset))
        raw = open(os.path.join(root, "tests", "golden", "meshes", fixture), "rb").read()
        assert raw[:4] == b"IPPM" and int(np.frombuffer(raw[4:8], dtype=np.uint32)[0]) == T == tris.shape[0]
        assert tris.tobytes() == raw[8:], asset
        # and the fixture itself comes back unchanged through the same entry point
        assert cpp_binding.load_mesh(os.path.join(root, "tests", "golden", "meshes", fixture)).tobytes() == raw[8:]
    import pytest
    with pytest.raises((ValueError, RuntimeError)):
        cpp_binding.load_mesh(os.path.join(root, "tests", "golden", "assets", "does_not_exist.stl"))
