"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gbp_b200.h declares, and fails loudly (no fallback) without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    entry.build()
    import global_body_planner_b200 as gbp
    return gbp.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gbp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gbp_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(L):
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"libgbp_b200.so lacks {missing}"


def test_struct_layout_matches_header():
    import global_body_planner_b200 as gbp
    assert gbp.PLAN_STATS_DTYPE.itemsize == 80
    assert ctypes.sizeof(gbp.PlanParams) == 80
    assert ctypes.sizeof(gbp.SvParams) == 152 and ctypes.sizeof(gbp.SvResult) == 64


def test_argument_validation_without_device(L):
    import global_body_planner_b200 as gbp
    h = ctypes.c_void_p()
    x = np.array([0.0, 0.0, 1.0]); z = np.zeros(9)
    rc = L.gbp_terrain_create(3, 3, x.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p),
                              z.ctypes.data_as(ctypes.c_void_p), None, None, None, ctypes.byref(h))
    assert rc == -1 and b"strictly increasing" in L.gbp_last_error()
    if gbp.device_count() == 0:
        with pytest.raises(gbp.GbpError, match="no CUDA device"):
            gbp.Terrain([0, 1], [0, 1], [[0, 0], [0, 0]])


def test_product_does_not_reference_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "global_body_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("import pyoracle", "libgbp_oracle", "libgbp_ref", "#include \"gbp_oracle", "dlopen"):
                    assert needle not in txt, f"{f} references the oracle ({needle})"
