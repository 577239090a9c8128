#!/usr/bin/env python3
"""Structural validator for csrc/mc_tables.h (the 256-row marching-cubes triangle table).

skimage (whose Lewiner tables the reference calls through surface_extractor.py:55) is not
installable offline, so the table cannot be diffed against skimage's.  Instead every row is
checked to be what a marching-cubes row has to be, independently of how it was obtained:

  1. only sign-changing cube edges are referenced, and every sign-changing edge is used;
  2. a mesh edge lying in a cube face is used by exactly one triangle (it is on the patch
     boundary), any other mesh edge by exactly two triangles with opposite directions
     (oriented 2-manifold interior);
  3. in every cube face each cut edge is the end point of exactly one boundary segment, and
     on faces with four cut edges (the ambiguous ones) the pairing rule is the same for all
     256 rows, so two cubes sharing a face always agree (no cracks);
  4. every triangle is wound the same way relative to the inside corners;
  5. the number of triangles equals sum(len(loop) - 2) over the boundary loops.

Run:  python tools/validate_mc_table.py        (exit code 0 = table valid)
Imported by tests/test_mc_table.py.
"""
import itertools
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "tomography_3d_reconstructor_b200", "csrc", "mc_tables.h")

# corner coordinates as (x, y, z), edge -> corner pairs (see mc_tables.h)
CORNERS = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0),
                    (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)], dtype=float)
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4),
         (0, 4), (1, 5), (2, 6), (3, 7)]
# faces as (axis, value)
FACES = [(a, v) for a in range(3) for v in (0, 1)]


def load_table(path=HEADER):
    txt = open(path).read()
    body = txt.split("#define T3D_TRI_TABLE_ROWS", 1)[1]
    rows = re.findall(r"\{([^}]*)\}", body)
    table = [[int(t) for t in r.split(",")] for r in rows]
    return table


def edge_faces(e):
    a, b = EDGES[e]
    out = []
    for f, (ax, v) in enumerate(FACES):
        if CORNERS[a][ax] == v and CORNERS[b][ax] == v:
            out.append(f)
    return out


EDGE_FACES = [set(edge_faces(e)) for e in range(12)]


def face_corners(f):
    ax, v = FACES[f]
    return [c for c in range(8) if CORNERS[c][ax] == v]


def validate(table):
    errors = []
    ambiguous_rule = {}  # (frozenset of inside corners on face) -> set of rules seen
    orient_signs = set()
    tri_counts = []
    if len(table) != 256:
        return ["table has %d rows" % len(table)], None
    for case, row in enumerate(table):
        if len(row) != 16:
            errors.append("case %d: row length %d" % (case, len(row)))
            continue
        n = row.index(-1) if -1 in row else 16
        if any(v != -1 for v in row[n:]) or n % 3:
            errors.append("case %d: malformed row" % case)
            continue
        tris = [tuple(row[i:i + 3]) for i in range(0, n, 3)]
        tri_counts.append(len(tris))
        inside = [(case >> c) & 1 for c in range(8)]
        cut = {e for e, (a, b) in enumerate(EDGES) if inside[a] != inside[b]}
        used = set(itertools.chain.from_iterable(tris))
        if not used <= cut:
            errors.append("case %d: uses non-cut edges %s" % (case, sorted(used - cut)))
            continue
        if used != cut:
            errors.append("case %d: cut edges never used %s" % (case, sorted(cut - used)))
            continue
        if any(len(set(t)) != 3 for t in tris):
            errors.append("case %d: degenerate triangle" % case)
            continue
        # directed mesh edges
        directed = {}
        for t in tris:
            for i in range(3):
                d = (t[i], t[(i + 1) % 3])
                directed[d] = directed.get(d, 0) + 1
        if any(v != 1 for v in directed.values()):
            errors.append("case %d: a directed mesh edge is used twice (inconsistent winding)" % case)
            continue
        boundary = []
        ok = True
        for (a, b) in directed:
            onface = EDGE_FACES[a] & EDGE_FACES[b]
            rev = (b, a) in directed
            if onface:
                if rev:
                    errors.append("case %d: mesh edge %d-%d lies in a cube face but is interior" % (case, a, b))
                    ok = False
                else:
                    boundary.append((a, b, next(iter(onface))))
            else:
                if not rev:
                    errors.append("case %d: interior mesh edge %d-%d has no twin" % (case, a, b))
                    ok = False
        if not ok:
            continue
        # per-face: each cut edge in the face is an end point of exactly one boundary segment
        for f in range(6):
            fc = face_corners(f)
            fcut = [e for e in cut if f in EDGE_FACES[e]]
            segs = [(a, b) for (a, b, ff) in boundary if ff == f]
            ends = list(itertools.chain.from_iterable(segs))
            if sorted(ends) != sorted(fcut):
                errors.append("case %d face %d: boundary segments %s do not match cut edges %s"
                              % (case, f, segs, sorted(fcut)))
                ok = False
                continue
            if len(fcut) == 4:
                # ambiguous face: does a segment cut off a single inside corner, or a single outside corner?
                ins = frozenset(c for c in fc if inside[c])
                rules = set()
                for (a, b) in segs:
                    shared = set(EDGES[a]) & set(EDGES[b])
                    if len(shared) != 1:
                        errors.append("case %d face %d: segment %d-%d joins opposite face edges" % (case, f, a, b))
                        ok = False
                        continue
                    c = next(iter(shared))
                    rules.add("separate_inside" if inside[c] else "separate_outside")
                if len(rules) != 1:
                    errors.append("case %d face %d: mixed pairing %s" % (case, f, rules))
                    ok = False
                else:
                    ambiguous_rule.setdefault((f, ins), set()).update(rules)
        if not ok:
            continue
        # boundary loops -> triangle count
        nxt = {}
        for (a, b, _f) in boundary:
            if a in nxt:
                errors.append("case %d: boundary vertex %d has two successors" % (case, a))
                ok = False
            nxt[a] = b
        if not ok:
            continue
        seen, expected = set(), 0
        for s in list(nxt):
            if s in seen:
                continue
            k, cur = 0, s
            while cur not in seen:
                seen.add(cur)
                cur = nxt.get(cur)
                k += 1
                if cur is None:
                    errors.append("case %d: open boundary" % case)
                    ok = False
                    break
            if ok:
                expected += k - 2
        if ok and expected != len(tris):
            # a patch with an interior tunnel/extra vertex would differ; the classic table has none
            errors.append("case %d: %d triangles, loops imply %d" % (case, len(tris), expected))
        # orientation: n . (outside - inside) along each vertex' own cube edge
        mid = np.array([(CORNERS[a] + CORNERS[b]) / 2 for a, b in EDGES])
        for t in tris:
            p = mid[list(t)]
            nrm = np.cross(p[1] - p[0], p[2] - p[0])
            tot = 0.0
            for e in t:
                a, b = EDGES[e]
                d = (CORNERS[b] - CORNERS[a]) * (1 if inside[a] else -1)  # inside -> outside
                tot += float(np.dot(nrm, d))
            # summed over the three vertices: a steep middle triangle of a hexagonal patch may have
            # one vertex whose own cube edge leans the other way, the sum never does
            if abs(tot) < 1e-12:
                errors.append("case %d: triangle %s has undetermined winding" % (case, t))
            else:
                orient_signs.add(1 if tot > 0 else -1)
    # global checks
    all_rules = set()
    for key, rules in ambiguous_rule.items():
        all_rules |= rules
    info = {
        "ambiguous_face_rules": sorted(all_rules),
        "orientation_signs": sorted(orient_signs),
        "total_triangles": int(sum(tri_counts)),
        "max_triangles": int(max(tri_counts)) if tri_counts else 0,
    }
    if len(orient_signs) != 1:
        errors.append("inconsistent winding across the table: %s" % sorted(orient_signs))
    if len(all_rules) > 1:
        # two neighbouring cubes see the same face with the same corner signs; a single global rule
        # keyed on the inside/outside state guarantees they pair the cut edges the same way
        per_pattern_conflict = [k for k, r in ambiguous_rule.items() if len(r) > 1]
        # the same geometric face pattern seen from the two sides is (f, ins) vs (f^1, mirrored ins)
        conflicts = 0
        for (f, ins), rules in ambiguous_rule.items():
            ax, v = FACES[f]
            g = FACES.index((ax, 1 - v))
            # mirror corners across the axis
            def mirror(c):
                p = CORNERS[c].copy()
                p[ax] = 1 - p[ax]
                return int(np.where((CORNERS == p).all(axis=1))[0][0])
            mins = frozenset(mirror(c) for c in ins)
            other = ambiguous_rule.get((g, mins))
            if other is not None and other != rules:
                conflicts += 1
        info["face_rule_conflicts"] = conflicts
        if per_pattern_conflict or conflicts:
            errors.append("ambiguous faces are not paired consistently: %d conflicts" % (conflicts + len(per_pattern_conflict)))
    return errors, info


def symmetry_classes(table):
    """Triangle counts must be invariant under the 24 cube rotations."""
    import itertools as it
    rots = []
    pts = CORNERS * 2 - 1
    for perm in it.permutations(range(3)):
        for signs in it.product((1, -1), repeat=3):
            m = np.zeros((3, 3))
            for i, p in enumerate(perm):
                m[i, p] = signs[i]
            if np.linalg.det(m) > 0:
                rots.append(m)
    bad = []
    ntri = [(r.index(-1) if -1 in r else 16) // 3 for r in table]
    for m in rots:
        mapped = pts @ m.T
        perm = [int(np.where((pts == q).all(axis=1))[0][0]) for q in mapped]
        for case in range(256):
            c2 = 0
            for c in range(8):
                if (case >> c) & 1:
                    c2 |= 1 << perm[c]
            if ntri[case] != ntri[c2]:
                bad.append((case, c2))
    return bad


def main():
    table = load_table()
    errors, info = validate(table)
    bad = symmetry_classes(table)
    print("info:", info)
    print("rotation-class triangle-count mismatches:", len(bad))
    for e in errors:
        print("ERROR:", e)
    return 1 if (errors or bad) else 0


if __name__ == "__main__":
    sys.exit(main())
