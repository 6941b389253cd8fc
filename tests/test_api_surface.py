"""The drop-in headers (include/global_body_planner/*.h) declare every public member function / free function the
reference's headers declare for the path (the reference's "FFI" is its C++ class API), same names and
parameter counts.  tests/golden/api_surface.json holds the reference's surface as (scope, name, arity) triples, minted
here from /root/reference/include/global_body_planner (python tests/test_api_surface.py --mint); where the reference tree
is present the committed file is also checked against it."""
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import api_surface as A  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "include", "global_body_planner")
REFERENCE = "/root/reference/include/global_body_planner"
FIXTURE = os.path.join(ROOT, "tests", "golden", "api_surface.json")
HEADERS = ("fast_terrain_map.h", "planning_utils.h", "graph_class.h", "planner_class.h", "rrt.h", "rrt_connect.h",
           "rrt_star_connect.h", "global_body_planner.h")
# declared by the reference, deliberately not by the drop-in
NOT_BUILT = {("FastTerrainMap", "saveData", 1): "debug dump that throws after writing (fast_terrain_map.cpp:226-231)",
             ("FastTerrainMap", "saveDataToTxt", 2): "helper of saveData (fast_terrain_map.cpp:234-260)",
             ("GlobalBodyPlanner", "spin", 0): "the ROS loop around callPlanner (global_body_planner.cpp:381-397); the drop-in driver is ROS-free"}


def _surface(directory):
    return {h: sorted(A.surface(open(os.path.join(directory, h)).read())) for h in HEADERS}


def _fixture():
    return {h: {tuple(t) for t in v} for h, v in json.load(open(FIXTURE)).items()}


@pytest.mark.parametrize("header", HEADERS)
def test_dropin_declares_the_reference_surface(header):
    want = _fixture()[header]
    have = A.surface(open(os.path.join(DROPIN, header)).read())
    missing = {t for t in want - have if t not in NOT_BUILT}
    assert not missing, f"{header}: declared by the reference, absent from the drop-in: {sorted(missing)}"
    assert len(want) >= 3, "the fixture lost its content"


def test_exceptions_are_real():
    """every stated exception is in the reference's surface and is indeed absent (no stale entries)"""
    ref = set().union(*_fixture().values())
    have = set().union(*(A.surface(open(os.path.join(DROPIN, h)).read()) for h in HEADERS))
    for t in NOT_BUILT:
        assert t in ref and t not in have, t


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="no /root/reference on this box")
def test_fixture_is_the_reference_surface():
    live = _surface(REFERENCE)
    assert {h: [list(t) for t in v] for h, v in live.items()} == json.load(open(FIXTURE))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="no /root/reference on this box")
def test_reference_signatures_compile_against_the_dropin(tmp_path):
    """stronger than names and arities: for every public function the reference's headers declare, a pointer spelled with
    the REFERENCE's return and parameter types is initialised from the DROP-IN's function of that name (g++ -fsyntax-only
    over the drop-in headers), overloads included.  Nothing of the reference is compiled or stored: the check file is
    generated in a temporary directory from the declarations parsed above."""
    import subprocess
    body, skipped = [], []
    for h in HEADERS:
        text = open(os.path.join(REFERENCE, h)).read()
        decls = A.declarations(text)
        for (kind, scope, ret, fn, params, const, static), line in zip(decls, A.pointer_checks(text)):
            if (scope, fn, A._nargs(params)) in NOT_BUILT or any(ns in params + ret for ns in ("grid_map::", "ros::", "nav_msgs::")):
                skipped.append((scope, fn))
                continue
            body.append(f"\t// {h}: {scope}::{fn}\n\t{line}")
    assert len(body) >= 95 and len(skipped) <= 6, (len(body), skipped)
    src = tmp_path / "signatures.cpp"
    src.write_text("".join(f'#include "global_body_planner/{h}"\n' for h in HEADERS) + "void check() {\n" + "\n".join(body) + "\n}\n")
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


if __name__ == "__main__" and "--mint" in sys.argv:
    json.dump({h: [list(t) for t in v] for h, v in _surface(REFERENCE).items()}, open(FIXTURE, "w"), indent=0)
    print(FIXTURE, sum(len(v) for v in _surface(REFERENCE).values()), "declarations")
