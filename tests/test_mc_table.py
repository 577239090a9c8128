"""The marching-cubes table is validated structurally (skimage's tables are not available offline)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import validate_mc_table as V  # noqa: E402


def test_table_is_a_valid_marching_cubes_table():
    table = V.load_table()
    errors, info = V.validate(table)
    assert errors == []
    assert info["orientation_signs"] in ([-1], [1])          # one global winding
    assert len(info["ambiguous_face_rules"]) == 1            # one pairing rule on ambiguous faces: no cracks
    assert info["max_triangles"] == 5 and info["total_triangles"] == 820
    assert V.symmetry_classes(table) == []                   # rotation classes share their triangle counts


def test_known_rows():
    t = V.load_table()
    assert t[0][0] == -1 and t[255][0] == -1
    assert t[1][:4] == [0, 8, 3, -1] and t[254][:4] == [0, 3, 8, -1]
    assert t[3][:7] == [1, 8, 3, 9, 8, 1, -1]
    assert t[7][:10] == [2, 8, 3, 2, 10, 8, 10, 9, 8, -1]
    assert t[15][:7] == [9, 8, 10, 10, 8, 11, -1]


# ---- tilings of the ambiguous configurations (tools/gen_mc33_tables.py) ------------------------------------------------
def test_mc33_rows_are_valid_from_first_principles():
    import validate_mc33 as V33
    errors, info = V33.validate()
    assert errors == []
    # 1056 = 24*4 (case 3) + 48*4 (6) + 16*16 (7) + 6*8 (10) + 24*8 (12) + 2*128 (13) + 8*2 (4)
    assert info["rows"] == 1056 and info["tunnel_rows"] == 110
    # rows with an interior diagonal inside a cube face: unavoidable without Lewiner's 13th vertex (pinned)
    assert info["face_diagonal_rows"] == 286


def test_oracle_tables_are_its_own_copy_and_equal_the_products():
    import validate_mc33 as V33
    a, b = V33.load(V33.PRODUCT_H), V33.load(V33.ORACLE_H)
    assert a["rows"] == b["rows"] and a["base"] == b["base"] and a["ntri"] == b["ntri"]
    # the oracle's own classic table against csrc/mc_tables.h
    assert b["classic"] == V.load_table()
    # the oracle includes nothing from the product tree
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "oracle", "mc_ref.c")).read()
    assert '#include "mc33_tables_oracle.h"' in src and "mc_tables.h\"" not in src.replace("mc33_tables_oracle.h\"", "")
    assert "csrc" not in open(os.path.join(root, "oracle", "cpu_ref.py")).read().split('"""', 2)[2]


def test_mc33_decision_metadata_matches_the_oracles_derivation(oracle):
    """K / FACES / NEED / POL / SIGN of csrc/mc33_tables.h (generator) against oracle/mc_ref.c's own derivation, probed
    through resolve_cube: every (index, J) reachable with +-values gives the same row."""
    import itertools
    import numpy as np
    import validate_mc33 as V33
    t = V33.load(V33.PRODUCT_H)
    for idx in range(1, 255):
        if t["base"][idx] == 0xffff:
            tris, J, tube, row = oracle.resolve_cube([1.0 if (idx >> c) & 1 else -1.0 for c in range(8)])
            assert row == -1
            continue
        k = t["K"][idx]
        # magnitudes 2 on the positive corners: p1*p2 - n1*n2 > 0 on every ambiguous face -> all joined; 0.5: none joined
        for mag, want in ((2.0, (1 << k) - 1), (0.5, 0)):
            v = [mag if (idx >> c) & 1 else -1.0 for c in range(8)]
            tris, J, tube, row = oracle.resolve_cube(v)
            assert J == want and row == t["base"][idx] + (J | (tube << k))
            assert tube in (0, 1) and (tube == 0 or (t["NEED"][idx] >> J) & 1)
            assert tris.tolist() == [t["rows"][row][i:i + 3] for i in range(0, 3 * t["ntri"][row], 3)]
