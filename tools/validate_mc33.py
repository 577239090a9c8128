#!/usr/bin/env python3
"""First-principles validator of the generated MC33 rows (csrc/mc33_tables.h, oracle/mc33_tables_oracle.h).

Written independently of tools/gen_mc33_tables.py (own geometry, own orientation test).  For every cube index with an
extended entry, every outcome J of the face tests and both values of the tunnel bit, the row must be

  1. built from all and only the sign-changing cube edges, no degenerate triangle;
  2. an oriented 2-manifold patch: every directed mesh edge at most once, every edge without a twin lies in a cube face
     (it is a boundary edge), all others have their twin.  An INTERIOR edge lying in a cube face (a diagonal between two
     vertices of one face) cannot be avoided without Lewiner's 13th vertex in sub-cases 7.3, 7.4.2, 10.1.2 / 12.1.2, 10.2,
     13.3, 13.4: such rows are counted ("face_diagonal_rows", pinned by tests/test_mc_table.py) and must obey the ownership
     rule (owns_face_diagonal) that keeps two neighbouring cubes from ever using the same diagonal;
  3. bounded exactly by the face polylines (index, J) prescribe: on an ambiguous face whose positive corners are joined
     (bit of J set) the two boundary segments cut off the two NEGATIVE corners, otherwise the two POSITIVE corners -- this
     is what makes two cubes sharing a face agree (the face test only sees the four shared corner values);
  4. wound like the classic table (the positive corner is always on the same side of every boundary segment);
  5. of the right topology: disks only (Euler characteristic = number of boundary loops = number of components), or, in
     the rows where the tunnel bit is honoured, exactly one annulus among them (Euler characteristic = loops - 2,
     components = loops - 1).

Run: python tools/validate_mc33.py   (exit 0 = all rows valid).  Imported by tests/test_mc_table.py.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_H = os.path.join(HERE, "..", "tomography_3d_reconstructor_b200", "csrc", "mc33_tables.h")
ORACLE_H = os.path.join(HERE, "..", "oracle", "mc33_tables_oracle.h")

P = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)], dtype=float)  # x,y,z
E = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
FACE = [(a, v) for a in range(3) for v in (0, 1)]     # x=0, x=1, y=0, y=1, z=0, z=1


def macro(txt, name):
    m = re.search(r"#define %s (.*?)(?=\n#define|\n\n|\n/\*|\Z)" % name, txt, re.S)
    if not m:
        raise KeyError(name)
    return m.group(1).replace("\\\n", " ")


def load(path):
    txt = open(path).read()
    width = int(macro(txt, "T3D_MC33_ROW"))
    base = [int(v) for v in macro(txt, "T3D_MC33_BASE_VALUES").split(",")]
    ntri = [int(v) for v in macro(txt, "T3D_MC33_NTRI_VALUES").split(",")]
    rows = [[int(v) for v in r.split(",")] for r in re.findall(r"\{([^}]*)\}", macro(txt, "T3D_MC33_TRI_ROWS"))]
    assert all(len(r) == width for r in rows) and len(rows) == len(ntri)
    out = {"width": width, "base": base, "ntri": ntri, "rows": rows}
    for key in ("K", "NEED", "POL", "SIGN"):
        try:
            out[key] = [int(v.replace("ull", ""), 0) for v in macro(txt, "T3D_MC33_%s_VALUES" % key).split(",")]
        except KeyError:
            pass
    try:
        out["FACES"] = [[int(v) for v in r.split(",")] for r in re.findall(r"\{([^}]*)\}", macro(txt, "T3D_MC33_FACES_VALUES"))]
    except KeyError:
        pass
    try:
        out["classic"] = [[int(v) for v in r.split(",")] for r in re.findall(r"\{([^}]*)\}", macro(txt, "T3D_ORACLE_CLASSIC_ROWS"))]
    except KeyError:
        pass
    return out


def face_of_edge_pair(a, b):
    for f, (ax, v) in enumerate(FACE):
        if all(P[c][ax] == v for c in E[a] + E[b]):
            return f
    return None


def owns_face_diagonal(a, b, f):
    """Ownership of a diagonal inside face f (written independently of the generator): classify the two cube edges by their
    direction inside the face; a diagonal between two PARALLEL (opposite) edges belongs to the cube that has the face on
    its high side (x=1 / y=1 / z=1), a diagonal between two PERPENDICULAR (adjacent) edges to the cube that has it on its
    low side."""
    ax, v = FACE[f]
    da = np.abs(P[E[a][0]] - P[E[a][1]])
    db = np.abs(P[E[b][0]] - P[E[b][1]])
    parallel = bool((da == db).all())
    return parallel == (v == 1)


def ambiguous_faces(idx):
    out = []
    for f, (ax, v) in enumerate(FACE):
        cs = [c for c in range(8) if P[c][ax] == v]
        pos = [c for c in cs if (idx >> c) & 1]
        if len(pos) == 2 and np.abs(P[pos[0]] - P[pos[1]]).sum() == 2:     # two positive corners, diagonal in the face
            out.append(f)
    return out


def check_row(idx, faces, Jbits, tunnel_expected, tris, ref_side):
    """Returns an error string or None.  ref_side: dict updated with the side (+1/-1) the positive corner lies on."""
    inside = [(idx >> c) & 1 for c in range(8)]
    cut = {e for e, (a, b) in enumerate(E) if inside[a] != inside[b]}
    used = {e for t in tris for e in t}
    if used != cut:
        return "uses %s, cut edges are %s" % (sorted(used), sorted(cut))
    if any(len(set(t)) != 3 for t in tris):
        return "degenerate triangle"
    directed = {}
    for t in tris:
        for i in range(3):
            d = (t[i], t[(i + 1) % 3])
            if d in directed:
                return "directed edge %s twice" % (d,)
            directed[d] = 1
    boundary = []
    for (a, b) in directed:
        f = face_of_edge_pair(a, b)
        twin = (b, a) in directed
        if f is not None and twin:
            # an interior edge inside a cube face: tolerated where unavoidable without a 13th vertex, but only if THIS cube owns
            # the diagonal (the neighbour across the face then cannot use it: no mesh edge with four triangles)
            if not owns_face_diagonal(a, b, f):
                return "edge %d-%d runs inside face %d, which this cube does not own for that diagonal" % (a, b, f)
            ref_side["face_diagonal"] = True
            continue
        if f is None and not twin:
            return "interior edge %d-%d has no twin" % (a, b)
        if f is not None:
            boundary.append((a, b, f))
    # 3. the polylines on every face
    for f, (ax, v) in enumerate(FACE):
        fcut = sorted(e for e in cut if all(P[c][ax] == v for c in E[e]))
        segs = [(a, b) for (a, b, ff) in boundary if ff == f]
        if sorted(x for s in segs for x in s) != fcut:
            return "face %d: segments %s vs cut edges %s" % (f, segs, fcut)
        if len(fcut) == 4:
            joined = (Jbits >> faces.index(f)) & 1
            for (a, b) in segs:
                shared = set(E[a]) & set(E[b])
                if len(shared) != 1:
                    return "face %d: segment %d-%d joins opposite edges" % (f, a, b)
                c = next(iter(shared))
                if inside[c] == joined:      # joined: cut-off corners must be negative; not joined: positive
                    return "face %d: J=%d but the segment cuts off a %s corner" % (f, joined, "positive" if inside[c] else "negative")
        # 4. orientation: side of the positive corners relative to each directed boundary segment, seen from outside
        n = np.zeros(3)
        n[ax] = 1 if v else -1
        mids = {e: (P[E[e][0]] + P[E[e][1]]) / 2 for e in fcut}
        for (a, b) in segs:
            d = mids[b] - mids[a]
            left = np.cross(n, d)
            # a positive corner of this face adjacent to one of the two cut edges
            pc = [c for c in E[a] + E[b] if inside[c]]
            side = np.sign(np.dot(left, P[pc[0]] - mids[a]))
            ref_side.setdefault("side", side)
            if side != ref_side["side"] or side == 0:
                return "face %d: segment %d->%d wound the other way" % (f, a, b)
    # 5. topology
    nxt = {a: b for (a, b, _f) in boundary}
    if len(nxt) != len(boundary):
        return "boundary vertex with two successors"
    seen, loops = set(), 0
    for s in nxt:
        if s in seen:
            continue
        loops += 1
        cur = s
        while cur not in seen:
            seen.add(cur)
            cur = nxt[cur]
    verts = sorted(used)
    und = {tuple(sorted(d)) for d in directed}
    chi = len(verts) - len(und) + len(tris)
    # components by union-find over triangles
    parent = {v: v for v in verts}

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for t in tris:
        parent[find(t[0])] = find(t[1])
        parent[find(t[1])] = find(t[2])
    comps = len({find(v) for v in verts})
    if tunnel_expected:
        if chi != loops - 2 or comps != loops - 1:
            return "tunnel row: chi %d, components %d, loops %d" % (chi, comps, loops)
    elif chi != loops or comps != loops:
        return "disk row: chi %d, components %d, loops %d" % (chi, comps, loops)
    return None


def validate(path=PRODUCT_H, meta_path=PRODUCT_H):
    t = load(path)
    m = load(meta_path)
    errors, n_rows, n_tunnel, n_diag = [], 0, 0, 0
    ref_side = {}
    # the side convention comes from the classic case-1 row {0, 8, 3}
    check_row(1, [], 0, False, [(0, 8, 3)], ref_side)
    for idx in range(256):
        faces = ambiguous_faces(idx)
        k = len(faces)
        diag = any(idx in ((1 << a) | (1 << b), 255 ^ ((1 << a) | (1 << b))) for a, b in ((0, 6), (1, 7), (2, 4), (3, 5)))
        extended = k > 0 or diag
        if (t["base"][idx] != 0xffff) != extended:
            errors.append("index %d: extended entry %s, expected %s" % (idx, t["base"][idx] != 0xffff, extended))
            continue
        if not extended:
            continue
        if m["K"][idx] != k or m["FACES"][idx][:k] != faces:
            errors.append("index %d: metadata K/FACES %s %s vs %d %s" % (idx, m["K"][idx], m["FACES"][idx], k, faces))
        for code in range(1 << (k + 1)):
            J, tb = code & ((1 << k) - 1), code >> k
            r = t["base"][idx] + code
            row = t["rows"][r]
            n = t["ntri"][r]
            if any(v != -1 for v in row[3 * n:]) or any(v < 0 for v in row[:3 * n]):
                errors.append("index %d code %d: malformed row" % (idx, code))
                continue
            tris = [tuple(row[i:i + 3]) for i in range(0, 3 * n, 3)]
            tunnel = bool(tb and (m["NEED"][idx] >> J) & 1)
            ref_side.pop("face_diagonal", None)
            err = check_row(idx, faces, J, tunnel, tris, ref_side)
            n_rows += 1
            n_tunnel += tunnel
            n_diag += bool(ref_side.pop("face_diagonal", False))
            if err:
                errors.append("index %d J=%s tube=%d: %s" % (idx, bin(J), tb, err))
    return errors, {"rows": n_rows, "tunnel_rows": n_tunnel, "face_diagonal_rows": n_diag}


def main():
    errors, info = validate()
    print("info:", info)
    for e in errors[:40]:
        print("ERROR:", e)
    a, b = load(PRODUCT_H), load(ORACLE_H)
    same = a["rows"] == b["rows"] and a["base"] == b["base"] and a["ntri"] == b["ntri"]
    print("product and oracle copies identical:", same)
    return 1 if (errors or not same) else 0


if __name__ == "__main__":
    sys.exit(main())
