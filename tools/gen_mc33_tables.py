#!/usr/bin/env python3
"""Generate the tilings of the AMBIGUOUS marching-cubes configurations (Chernyaev's MC33 / Lewiner et al. 2003).

skimage.measure.marching_cubes (the reference's call at surface_extractor.py:55) is Lewiner's implementation: for cubes
with an ambiguous face (Lewiner's cases 3, 6, 7, 10, 12, 13) it decides with the asymptotic decider how the face is
cut, for case 4 and for some sub-cases whether the interior carries a tunnel, and picks one of 33 topological tilings.
skimage's look-up tables are not available offline (no network, package absent), so the tilings are GENERATED here from
the topology they have to realise.  What is pinned and what is not is stated in DESIGN.md section 2.

For every cube index with an ambiguous face (or case 4) and every outcome of the tests, one row:

    row(index, J, tube),   J = bit i set <=> the POSITIVE corners are joined across the i-th ambiguous face of `index`
                           tube = 1 <=> the interior test asks for a tunnel

  * the polylines the surface has to follow on the six cube faces are fixed by (index, J): on an ambiguous face the two
    segments cut off the negative corners if J, the positive corners otherwise; they close into oriented loops;
  * tube = 0: every loop is closed by a disk.  If the classic 256-row table (csrc/mc_tables.h) already has exactly these
    boundary segments, its row is used verbatim (Lewiner's default sub-cases reuse the classic rows); otherwise each loop
    is triangulated without extra vertices (see disk());
  * tube = 1 (only where the decision tree of Lewiner's implementation runs the interior test: 4, 6.1, 7.4, 10.1, 12.1,
    13.5): the two loops bounding the tunnel are joined by a triangle strip (m + n triangles, see tube()), other loops
    get disks.
    No additional (centre) vertex is used anywhere: Lewiner's 13th vertex (sub-cases 6.1.2, 7.3, 10.2, 12.2, 13.3, 13.4)
    is the known deviation -- same topology, different triangle count in those sub-cases.  Where a triangulation without
    a diagonal inside a cube face does not exist, only diagonals owned by this cube are used (diagonal_allowed), which keeps
    the mesh free of edges with more than two triangles.

Outputs (identical tables, separately stored so that the CPU oracle does not include product sources):
    tomography_3d_reconstructor_b200/csrc/mc33_tables.h     (+ per-index decision metadata for the kernels)
    oracle/mc33_tables_oracle.h                             (+ its own copy of the classic table)

    python tools/gen_mc33_tables.py          regenerate both files
    python tools/gen_mc33_tables.py --check  exit 1 if the files on disk differ from what would be generated
"""
import itertools
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CLASSIC_H = os.path.join(ROOT, "tomography_3d_reconstructor_b200", "csrc", "mc_tables.h")
OUT_PRODUCT = os.path.join(ROOT, "tomography_3d_reconstructor_b200", "csrc", "mc33_tables.h")
OUT_ORACLE = os.path.join(ROOT, "oracle", "mc33_tables_oracle.h")

# corner (x, y, z), edge -> corners: the conventions of mc_tables.h
CORN = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
FACES = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)]      # (axis, value): x=0, x=1, y=0, y=1, z=0, z=1
ROW = 32                                                      # edge ids per extended row (at most 10 triangles occur)


def face_cycle(f):
    """The four corners of face f in cyclic order."""
    ax, val = FACES[f]
    u, w = [a for a in range(3) if a != ax]
    out = []
    for (cu, cw) in ((0, 0), (1, 0), (1, 1), (0, 1)):
        for c, p in enumerate(CORN):
            if p[ax] == val and p[u] == cu and p[w] == cw:
                out.append(c)
    return out


FACE_CYCLE = [face_cycle(f) for f in range(6)]
FACE_EDGES = [[e for e, (a, b) in enumerate(EDGES) if a in FACE_CYCLE[f] and b in FACE_CYCLE[f]] for f in range(6)]
DIAGS = [(0, 6), (1, 7), (2, 4), (3, 5)]


def load_classic():
    body = open(CLASSIC_H).read().split("#define T3D_TRI_TABLE_ROWS", 1)[1]
    rows = [[int(t) for t in r.split(",")] for r in re.findall(r"\{([^}]*)\}", body)]
    assert len(rows) == 256 and all(len(r) == 16 for r in rows)
    return [[tuple(r[i:i + 3]) for i in range(0, r.index(-1) if -1 in r else 16, 3)] for r in rows]


def signs(idx):
    return [(idx >> c) & 1 for c in range(8)]


def ambiguous_faces(idx):
    s = signs(idx)
    return [f for f in range(6) if s[FACE_CYCLE[f][0]] == s[FACE_CYCLE[f][2]] != s[FACE_CYCLE[f][1]] == s[FACE_CYCLE[f][3]]]


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def mid2(e):
    """Twice the midpoint of cube edge e (integers)."""
    a, b = EDGES[e]
    return tuple(CORN[a][i] + CORN[b][i] for i in range(3))


def directed_segment(f, e1, e2, side_corners, side_sign, positive_left):
    """Segment between cut edges e1, e2 of face f with the corners `side_corners` (sign `side_sign`) on one side.
    Returned directed so that, seen from outside the cube, the positive side is on the left (or right)."""
    ax, val = FACES[f]
    n = [0, 0, 0]
    n[ax] = 1 if val else -1
    p1, p2 = mid2(e1), mid2(e2)
    d = tuple(b - a for a, b in zip(p1, p2))
    left = cross(n, d)
    cen = [sum(2 * CORN[c][i] for c in side_corners) / len(side_corners) for i in range(3)]
    m = [(a + b) / 2 for a, b in zip(p1, p2)]
    s = dot(left, [c - q for c, q in zip(cen, m)])
    assert s != 0
    side_is_left = s > 0
    positive_is_left = side_is_left == bool(side_sign)
    return (e1, e2) if positive_is_left == positive_left else (e2, e1)


def required_segments(idx, J, positive_left):
    """Directed boundary segments {(e_from, e_to)} for cube index idx, J = {face: positives joined?} for ambiguous faces."""
    s = signs(idx)
    segs = set()
    for f in range(6):
        cyc = FACE_CYCLE[f]
        cut = [e for e in FACE_EDGES[f] if s[EDGES[e][0]] != s[EDGES[e][1]]]
        if len(cut) == 2:
            pos = [c for c in cyc if s[c]]
            segs.add(directed_segment(f, cut[0], cut[1], pos, 1, positive_left))
        elif len(cut) == 4:
            cut_sign = 0 if J[f] else 1          # positives joined: the negative corners are cut off
            for c in cyc:
                if s[c] == cut_sign:
                    ee = [e for e in FACE_EDGES[f] if c in EDGES[e]]
                    segs.add(directed_segment(f, ee[0], ee[1], [c], cut_sign, positive_left))
    return segs


def boundary_of(tris):
    """Directed mesh edges without a twin."""
    d = set()
    for t in tris:
        for i in range(3):
            d.add((t[i], t[(i + 1) % 3]))
    return {e for e in d if (e[1], e[0]) not in d}


def loops_of(segs):
    nxt = dict(segs)
    assert len(nxt) == len(segs)
    left = set(nxt)
    loops = []
    while left:
        start = min(left)
        loop, e = [], start
        while True:
            loop.append(e)
            left.discard(e)
            e = nxt[e]
            if e == start:
                break
        loops.append(loop)
    return loops


def same_face(a, b):
    """Do cube edges a and b lie in a common cube face?"""
    return any(all(c in FACE_CYCLE[f] for c in EDGES[a] + EDGES[b]) for f in range(6))


def face_label(f, e):
    """Face-local label of cube edge e lying in face f, the same for the two cubes that share the face: 0 / 1 = the edge runs
    along the first in-face axis at the low / high end of the second, 2 / 3 = along the second axis at the low / high end of
    the first."""
    ax, _val = FACES[f]
    u, w = [a for a in range(3) if a != ax]
    pa, pb = CORN[EDGES[e][0]], CORN[EDGES[e][1]]
    return pa[w] if pa[u] != pb[u] else 2 + pa[u]


def diagonal_allowed(a, b):
    """A mesh edge between two vertices of one cube face that is not a boundary segment runs INSIDE that face.  Two cubes
    share the face; if both used the same such diagonal the mesh edge would carry four triangles.  Ownership rule: a
    diagonal joining OPPOSITE edges of the face belongs to the cube on the LOW side of the face (the one that sees it as its
    x=1 / y=1 / z=1 face), a diagonal joining ADJACENT edges to the cube on the high side.  Of the 64 possible assignments
    of the six kinds of diagonal, four let every loop and every tunnel of every sub-case be triangulated with owned
    diagonals only (exhaustive search; this is one of them), so no mesh edge is ever used by more than two triangles --
    without Lewiner's 13th vertex.  The generator asserts it, tools/validate_mc33.py re-checks it independently."""
    for f in range(6):
        if all(c in FACE_CYCLE[f] for c in EDGES[a] + EDGES[b]):
            low_side_owns = (face_label(f, a) < 2) == (face_label(f, b) < 2)
            if low_side_owns != (FACES[f][1] == 1):
                return False
    return True


def dist2(a, b):
    return sum((p - q) ** 2 for p, q in zip(mid2(a), mid2(b)))


def polygon_triangulations(n):
    """All triangulations of the convex n-gon 0..n-1 in a fixed enumeration order (apex of edge (i, j) ascending)."""
    memo = {}

    def rec(i, j):
        if j - i < 2:
            return [[]]
        if (i, j) not in memo:
            memo[(i, j)] = [L + [(i, k, j)] + R for k in range(i + 1, j) for L in rec(i, k) for R in rec(k, j)]
        return memo[(i, j)]
    return rec(0, n - 1)


def disk(loop):
    """Triangulation of the loop (v0 = lowest edge id) without additional vertices.  A diagonal joining two loop vertices
    that lie in one cube face would run inside that face: such diagonals are avoided where a triangulation without them
    exists (it does not for the 9-loops of 7.3 / 13.3, the 8-loops of 10.2 and the 12-loops of 13.4 -- the sub-cases
    where Lewiner inserts a 13th vertex) and otherwise restricted to the diagonals this cube OWNS (diagonal_allowed: the
    neighbour across the face can never use the same one); cost = (diagonals not owned [always 0], face diagonals, summed
    squared diagonal length), first minimum in enumeration order.  Returns (triangles, number of face diagonals)."""
    n = len(loop)
    best = None
    for tr in polygon_triangulations(n):
        chords = {(min(p, q), max(p, q)) for t in tr for p, q in ((t[0], t[1]), (t[1], t[2]), (t[0], t[2]))
                  if (q - p) % n not in (1, n - 1)}
        cost = (sum(not diagonal_allowed(loop[p], loop[q]) for p, q in chords), sum(same_face(loop[p], loop[q]) for p, q in chords),
                sum(dist2(loop[p], loop[q]) for p, q in chords))
        if best is None or cost < best[0]:
            best = (cost, tr)
    (forbidden, bad, _), tr = best
    assert forbidden == 0, loop
    # every triangle (i, k, j) with i < k < j follows the loop direction
    return [(loop[i], loop[k], loop[j]) for (i, k, j) in tr], bad


def tube(A, B):
    """Triangulation of the annulus between loops A and B (both directed as boundary of the surface: they run in opposite
    senses around the tunnel) without additional vertices: m + n triangles.  Cutting the annulus along a cross edge
    (a_i, b_j) leaves the polygon a_i, a_i+1, ..., a_i (again), b_j, b_j+1, ..., b_j (again); every triangulation of that
    polygon which is a valid mesh (no degenerate triangle, no directed edge twice) is a candidate.  Cost = (diagonals inside
    a cube face, summed squared diagonal length); first minimum in (i, j, enumeration) order."""
    m, n = len(A), len(B)
    best = None
    tri_sets = polygon_triangulations(m + n + 2)
    for i in range(m):
        for j in range(n):
            poly = [A[(i + t) % m] for t in range(m + 1)] + [B[(j + t) % n] for t in range(n + 1)]
            for tr in tri_sets:
                tris = [(poly[p], poly[q], poly[r]) for (p, q, r) in tr]
                if any(len(set(t)) != 3 for t in tris):
                    continue
                directed = [(t[c], t[(c + 1) % 3]) for t in tris for c in range(3)]
                if len(set(directed)) != len(directed):
                    continue
                on_a = {(A[t], A[(t + 1) % m]) for t in range(m)}
                on_b = {(B[t], B[(t + 1) % n]) for t in range(n)}
                inner = {(min(d), max(d)) for d in directed if d not in on_a and d not in on_b}
                if any((d[1], d[0]) in on_a or (d[1], d[0]) in on_b for d in directed):
                    continue                   # a boundary edge used backwards
                cost = (sum(not diagonal_allowed(p, q) for p, q in inner), sum(same_face(p, q) for p, q in inner),
                        sum(dist2(p, q) for p, q in inner))
                if best is None or cost < best[0]:
                    best = (cost, tris)
    assert best[0][0] == 0, (A, B)
    return best[1], best[0][1]


def classify(idx):
    """Per-index decision metadata: (k, ambiguous faces, need, polarity, s).
    need: bit J set <=> pattern J of the face tests is followed by the interior test;
    polarity 1: tunnel iff the interior test returns true (case 7.4), 0: tunnel iff it returns false;
    s: +1 / -1 = sign argument of the interior test (the minority corners' sign; + for 4/4)."""
    sg = signs(idx)
    p = sum(sg)
    faces = ambiguous_faces(idx)
    k = len(faces)
    minority_positive = p <= 4
    mc = min(p, 8 - p)
    s = 1 if minority_positive else -1
    need, pol = 0, 0

    def T(Jbits, i):      # minority corners joined on the i-th ambiguous face
        j = (Jbits >> i) & 1
        return j if minority_positive else 1 - j

    if k == 0:
        if any(idx in ((1 << a) | (1 << b), 255 ^ ((1 << a) | (1 << b))) for a, b in DIAGS):
            need = 1                                   # case 4
    elif k == 1 and mc == 3:
        need = sum(1 << J for J in range(2) if not T(J, 0))                    # case 6.1
    elif k == 3:
        need = sum(1 << J for J in range(8) if all(T(J, i) for i in range(3)))  # case 7.4
        pol = 1
    elif k == 2:
        need = 1                                       # cases 10.1 / 12.1: positives joined on neither face
    elif k == 6:
        for q in range(8):                             # case 13.5: joined exactly on the three faces of a negative corner
            if not sg[q]:
                J = sum(1 << i for i, f in enumerate(faces) if q in FACE_CYCLE[f])
                need |= 1 << J
    return k, faces, need, pol, s


def build():
    classic = load_classic()
    # orientation convention of the classic table: try both, keep the one its case-1 row follows, check all rows
    positive_left = None
    for cand in (True, False):
        if boundary_of(classic[1]) == required_segments(1, {}, cand):
            positive_left = cand
    assert positive_left is not None
    for idx in range(1, 255):
        faces = ambiguous_faces(idx)
        b = boundary_of(classic[idx])
        assert any(b == required_segments(idx, dict(zip(faces, J)), positive_left)
                   for J in itertools.product((0, 1), repeat=len(faces))), idx

    base = [0xffff] * 256
    meta = [(0, [], 0, 0, 1)] * 256
    rows = []
    face_chord_rows = []
    for idx in range(1, 255):
        k, faces, need, pol, s = classify(idx)
        meta[idx] = (k, faces, need, pol, s)
        if k == 0 and not need:
            continue
        base[idx] = len(rows)
        sg = signs(idx)
        for code in range(1 << (k + 1)):
            Jbits, tb = code & ((1 << k) - 1), code >> k
            J = {f: (Jbits >> i) & 1 for i, f in enumerate(faces)}
            segs = required_segments(idx, J, positive_left)
            loops = loops_of(segs)
            if tb and (need >> Jbits) & 1:
                if k == 6:       # 13.5: hexagon + the triangle around the isolated positive corner
                    hexa = [l for l in loops if len(l) == 6]
                    tri = [l for l in loops if len(l) == 3 and any(sg[c] and all(c in EDGES[e] for e in l) for c in range(8))]
                    assert len(hexa) == 1 and len(tri) == 1 and len(loops) == 3, (idx, Jbits, loops)
                    pair = (tri[0], hexa[0]) if min(tri[0]) < min(hexa[0]) else (hexa[0], tri[0])
                else:
                    assert len(loops) == 2, (idx, Jbits, loops)
                    pair = (loops[0], loops[1])
                tris = []
                done = False
                for l in loops:
                    if l is pair[0] or l is pair[1]:
                        if not done:
                            tt, bad = tube(pair[0], pair[1])
                            if bad:
                                face_chord_rows.append((idx, code, -(len(pair[0]) + len(pair[1])), bad))
                            tris += tt
                            done = True
                    else:
                        tris += disk(l)[0]
            elif boundary_of(classic[idx]) == segs:
                tris = list(classic[idx])
            else:
                tris = []
                for l in loops:
                    tt, bad = disk(l)
                    tris += tt
                    if bad:
                        face_chord_rows.append((idx, code, len(l), bad))
            assert 3 * len(tris) <= ROW, (idx, code, len(tris))
            rows.append(tris)
    build.face_chord_rows = face_chord_rows
    return classic, base, meta, rows


def fmt_rows(rows, width):
    out = []
    for tris in rows:
        flat = [e for t in tris for e in t]
        flat += [-1] * (width - len(flat))
        out.append("{" + ",".join(str(v) for v in flat) + "}")
    return ",\\\n".join(out)


def render(classic, base, meta, rows, for_oracle):
    L = []
    L.append("/* GENERATED by tools/gen_mc33_tables.py -- do not edit.  Tilings of the ambiguous marching-cubes configurations")
    L.append(" * (MC33 topology; see the generator's docstring and DESIGN.md section 2).  row = BASE[index] + (J | tube << K[index]). */")
    guard = "T3D_MC33_TABLES_ORACLE_H" if for_oracle else "T3D_MC33_TABLES_H"
    L += ["#ifndef " + guard, "#define " + guard, ""]
    L.append("#define T3D_MC33_ROW %d" % ROW)
    L.append("#define T3D_MC33_NROWS %d" % len(rows))
    L.append("#define T3D_MC33_NONE 0xffff")
    L.append("#define T3D_MC33_BASE_VALUES " + ",".join(str(b) for b in base))
    L.append("#define T3D_MC33_NTRI_VALUES " + ",".join(str(len(r)) for r in rows))
    L.append("#define T3D_MC33_TRI_ROWS \\\n" + fmt_rows(rows, ROW))
    if for_oracle:
        L.append("")
        L.append("/* the oracle's own copy of the classic 256-row table (public-domain Lorensen/Cline/Bourke rows) */")
        L.append("#define T3D_ORACLE_CLASSIC_ROW 16")
        L.append("#define T3D_ORACLE_CLASSIC_ROWS \\\n" + fmt_rows(classic, 16))
    else:
        L.append("")
        L.append("/* per-index decision metadata: number of ambiguous faces, their ids (x=0,x=1,y=0,y=1,z=0,z=1), the J patterns")
        L.append(" * followed by the interior test (bit J of NEED), tunnel polarity, sign argument of the interior test */")
        L.append("#define T3D_MC33_K_VALUES " + ",".join(str(m[0]) for m in meta))
        L.append("#define T3D_MC33_FACES_VALUES " + ",".join("{" + ",".join(str(f) for f in (m[1] + [255] * 6)[:6]) + "}" for m in meta))
        L.append("#define T3D_MC33_NEED_VALUES " + ",".join("0x%xull" % m[2] for m in meta))
        L.append("#define T3D_MC33_POL_VALUES " + ",".join(str(m[3]) for m in meta))
        L.append("#define T3D_MC33_SIGN_VALUES " + ",".join(str(m[4]) for m in meta))
        L.append("/* the four corners of each face in cyclic order */")
        L.append("#define T3D_MC33_FACE_CYCLE_VALUES " + ",".join("{" + ",".join(str(c) for c in FACE_CYCLE[f]) + "}" for f in range(6)))
    L += ["", "#endif", ""]
    return "\n".join(L)


def main():
    classic, base, meta, rows = build()
    outs = {OUT_PRODUCT: render(classic, base, meta, rows, False), OUT_ORACLE: render(classic, base, meta, rows, True)}
    if "--check" in sys.argv:
        bad = [p for p, txt in outs.items() if not os.path.exists(p) or open(p).read() != txt]
        if bad:
            print("stale:", bad)
            sys.exit(1)
        print("mc33 tables up to date: %d rows" % len(rows))
        return
    for p, txt in outs.items():
        open(p, "w").write(txt)
    n_classic = sum(1 for idx in range(256) if base[idx] != 0xffff for c in range(1 << (meta[idx][0] + 1))
                    if rows[base[idx] + c] == classic[idx])
    print("wrote %d rows (%d identical to the classic row), max %d triangles" % (len(rows), n_classic, max(len(r) for r in rows)))
    import collections
    print("rows with a diagonal inside a cube face (loop length: rows):",
          dict(collections.Counter(l for _i, _c, l, _b in build.face_chord_rows)))


if __name__ == "__main__":
    main()
