"""Shared fixtures.  `-m "not gpu"` tests run on the CPU only (oracle, host logic, C-ABI symbol check);
`-m gpu` tests are the parity tests proper and call libipp.so through its C ABI."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MESHES = os.path.join(ROOT, "tests", "golden", "meshes")
DEFAULT_SPECS = [2.0, 46, 46, 1024, 0.1, 10, 1, 1]  # plan_optimizer/settings/Parameters.py:55


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


def mesh_path(name):
    return os.path.join(MESHES, name + ".ippm")


def specs(height=1024, interval=2.0, front=1, down=1, near=0.1, far=10.0, fov=46):
    return [interval, fov, fov, height, near, far, front, down]


def cuda_available():
    try:
        L = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return False
    n = ctypes.c_int(0)
    return L.cuInit(0) == 0 and L.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as om
    om.build()
    return om


def random_plans(rng, n_boxes, pop, length):
    """genes i.i.d. uniform on [0, N) with consecutive duplicates re-drawn (mirrors cleanUpIndividual,
    plan_optimizer/evolutionary_operators/mutationAndRecombination.py:64-87)."""
    plans = []
    for _ in range(pop):
        L = length if isinstance(length, int) else int(rng.integers(length[0], length[1] + 1))
        g = []
        while len(g) < L:
            x = int(rng.integers(0, n_boxes))
            if g and g[-1] == x:
                continue
            g.append(x)
        plans.append([[x] for x in g])
    return plans


def popcount(words):
    return int(np.unpackbits(np.ascontiguousarray(words, dtype=np.uint32).view(np.uint8)).sum())


def bit_ids(words):
    bits = np.unpackbits(np.ascontiguousarray(words, dtype=np.uint32).view(np.uint8), bitorder="little")
    return np.nonzero(bits)[0]
