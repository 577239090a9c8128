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
