/*
 * oracle/mc_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
 *
 * Scalar CPU restatement of skimage.measure.marching_cubes(volume, level) with its defaults, which
 * is what the reference calls at /root/reference/surface_extractor.py:55.  scikit-image (floor pin
 * `scikit-image>=0.18.0`, /root/reference/requirements.txt:5) is NOT installed in this image, so the
 * restatement follows the published algorithm (skimage/measure/_marching_cubes_lewiner_cy.pyx):
 *
 *   - volume is float32, C order (nz, ny, nx); cubes visited z-major, then y, x fastest;
 *   - corner order 0:(z,y,x) 1:(z,y,x+1) 2:(z,y+1,x+1) 3:(z,y+1,x) 4..7 same at z+1;
 *     case bit i = ((double)v_i - level) > 0;
 *   - one vertex per cut grid edge, created on first use (so vertex order = first-use order) and
 *     shared through per-layer index caches;
 *   - vertex on edge (a -> b, b = a + 1 along one axis):
 *         wa = 1/(FLT_EPSILON + |va|), wb = 1/(FLT_EPSILON + |vb|)   (v = value - level, double)
 *         coordinate = a + 1*wb/(wa + wb)  computed in double, stored float32
 *   - output vertices are (z, y, x); gradient_direction='descent' reverses every triangle.
 *
 * PARITY UNPINNED for the tiling of ambiguous cubes: triangles come from the classic 256-row table
 * (csrc/mc_tables.h, validated structurally by tools/validate_mc_table.py).  skimage's Lewiner tables
 * use the same rows for all cases without an ambiguous face; for cubes WITH an ambiguous face (or the
 * two-opposite-corner "case 4") Lewiner runs extra face/interior tests that may pick another tiling.
 * Such cubes are counted in *n_ambiguous so tests can assert the count is zero on the benchmark
 * phantoms.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mc_tables.h"

static const int8_t TRI[256][T3D_MC_ROW] = { T3D_TRI_TABLE_ROWS };

/* edge -> (corner a, corner b) */
static const int EDGE_CORNER[12][2] = {
    {0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
/* corner -> (dz, dy, dx) */
static const int CORNER_OFF[8][3] = {
    {0, 0, 0}, {0, 0, 1}, {0, 1, 1}, {0, 1, 0}, {1, 0, 0}, {1, 0, 1}, {1, 1, 1}, {1, 1, 0}};

/* face -> its 4 corners in cyclic order (for the ambiguity diagnostic) */
static const int FACE_CORNERS[6][4] = {
    {0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 3, 7, 4}, {1, 2, 6, 5}};

static int is_ambiguous(int idx)
{
    for (int f = 0; f < 6; ++f) {
        int a = (idx >> FACE_CORNERS[f][0]) & 1, b = (idx >> FACE_CORNERS[f][1]) & 1;
        int c = (idx >> FACE_CORNERS[f][2]) & 1, d = (idx >> FACE_CORNERS[f][3]) & 1;
        if (a == c && b == d && a != b) return 1; /* alternating corners on a face */
    }
    /* Lewiner case 4: exactly two inside (or outside) corners, diagonally opposite in the cube */
    static const int DIAG[4][2] = {{0, 6}, {1, 7}, {2, 4}, {3, 5}};
    for (int k = 0; k < 4; ++k) {
        int m = (1 << DIAG[k][0]) | (1 << DIAG[k][1]);
        if (idx == m || idx == (255 ^ m)) return 1;
    }
    return 0;
}

/*
 * Returns 0 on success.  verts (3 floats per vertex, z,y,x) and faces (3 int32 per face) may be NULL
 * for a counting pass; otherwise they must hold at least cap_v / cap_f entries (error 2 if exceeded).
 */
int t3d_oracle_marching_cubes(const float *vol, int nz, int ny, int nx, double level,
                              float *verts, int64_t cap_v, int32_t *faces, int64_t cap_f,
                              int64_t *n_verts, int64_t *n_faces, int64_t *n_ambiguous,
                              int64_t *case_hist /* 256 entries or NULL */)
{
    int64_t nv = 0, nf = 0, namb = 0;
    if (nz < 2 || ny < 2 || nx < 2) {
        *n_verts = 0; *n_faces = 0; if (n_ambiguous) *n_ambiguous = 0;
        return 0;
    }
    const size_t plane = (size_t)ny * nx;
    /* index caches: x- and y-edges for the two planes of the current cube layer, z-edges between */
    int32_t *xe[2], *ye[2], *ze;
    for (int k = 0; k < 2; ++k) {
        xe[k] = (int32_t *)malloc(plane * sizeof(int32_t));
        ye[k] = (int32_t *)malloc(plane * sizeof(int32_t));
        if (!xe[k] || !ye[k]) return 1;
        memset(xe[k], 0xff, plane * sizeof(int32_t));
        memset(ye[k], 0xff, plane * sizeof(int32_t));
    }
    ze = (int32_t *)malloc(plane * sizeof(int32_t));
    if (!ze) return 1;
    int rc = 0;

    for (int z = 0; z < nz - 1 && !rc; ++z) {
        /* rotate caches: upper plane becomes lower plane */
        if (z > 0) {
            int32_t *t = xe[0]; xe[0] = xe[1]; xe[1] = t;
            t = ye[0]; ye[0] = ye[1]; ye[1] = t;
            memset(xe[1], 0xff, plane * sizeof(int32_t));
            memset(ye[1], 0xff, plane * sizeof(int32_t));
        }
        memset(ze, 0xff, plane * sizeof(int32_t));
        const float *p0 = vol + (size_t)z * plane, *p1 = p0 + plane;
        for (int y = 0; y < ny - 1 && !rc; ++y) {
            for (int x = 0; x < nx - 1; ++x) {
                double v[8];
                const size_t o = (size_t)y * nx + x;
                v[0] = (double)p0[o] - level;          v[1] = (double)p0[o + 1] - level;
                v[2] = (double)p0[o + nx + 1] - level; v[3] = (double)p0[o + nx] - level;
                v[4] = (double)p1[o] - level;          v[5] = (double)p1[o + 1] - level;
                v[6] = (double)p1[o + nx + 1] - level; v[7] = (double)p1[o + nx] - level;
                int idx = 0;
                for (int c = 0; c < 8; ++c) if (v[c] > 0.0) idx |= 1 << c;
                if (case_hist) case_hist[idx]++;
                if (idx == 0 || idx == 255) continue;
                if (is_ambiguous(idx)) namb++;
                const int8_t *row = TRI[idx];
                for (int t = 0; row[t] >= 0; t += 3) {
                    int32_t tri[3];
                    for (int k = 0; k < 3; ++k) {
                        const int e = row[t + k];
                        const int ca = EDGE_CORNER[e][0], cb = EDGE_CORNER[e][1];
                        /* canonical owner = the lower corner of the edge */
                        const int lo = (CORNER_OFF[ca][0] + CORNER_OFF[ca][1] + CORNER_OFF[ca][2] <
                                        CORNER_OFF[cb][0] + CORNER_OFF[cb][1] + CORNER_OFF[cb][2]) ? ca : cb;
                        const int hi = (lo == ca) ? cb : ca;
                        const int dz = CORNER_OFF[lo][0], dy = CORNER_OFF[lo][1], dx = CORNER_OFF[lo][2];
                        const size_t oo = (size_t)(y + dy) * nx + (x + dx);
                        int32_t *slot;
                        int axis; /* 0 = z-edge, 1 = y-edge, 2 = x-edge */
                        if (CORNER_OFF[hi][2] != dx) { slot = &xe[dz][oo]; axis = 2; }
                        else if (CORNER_OFF[hi][1] != dy) { slot = &ye[dz][oo]; axis = 1; }
                        else { slot = &ze[oo]; axis = 0; }
                        if (*slot < 0) {
                            if (verts) {
                                if (nv >= cap_v) { rc = 2; break; }
                                const double wa = 1.0 / (FLT_EPSILON + fabs(v[lo]));
                                const double wb = 1.0 / (FLT_EPSILON + fabs(v[hi]));
                                const double frac = 1.0 * wb / (wa + wb);
                                double pz = (double)(z + dz), py = (double)(y + dy), px = (double)(x + dx);
                                if (axis == 0) pz += frac; else if (axis == 1) py += frac; else px += frac;
                                verts[3 * nv + 0] = (float)pz;
                                verts[3 * nv + 1] = (float)py;
                                verts[3 * nv + 2] = (float)px;
                            }
                            *slot = (int32_t)nv++;
                        }
                        tri[k] = *slot;
                    }
                    if (rc) break;
                    if (faces) {
                        if (nf >= cap_f) { rc = 2; break; }
                        /* gradient_direction='descent' => reversed winding */
                        faces[3 * nf + 0] = tri[2];
                        faces[3 * nf + 1] = tri[1];
                        faces[3 * nf + 2] = tri[0];
                    }
                    nf++;
                }
                if (rc) break;
            }
        }
    }
    for (int k = 0; k < 2; ++k) { free(xe[k]); free(ye[k]); }
    free(ze);
    *n_verts = nv; *n_faces = nf;
    if (n_ambiguous) *n_ambiguous = namb;
    return rc;
}

/* cube-case index volume (nz-1, ny-1, nx-1) uint8, for bit-exact classification tests */
int t3d_oracle_cube_cases(const float *vol, int nz, int ny, int nx, double level, uint8_t *out)
{
    const size_t plane = (size_t)ny * nx;
    for (int z = 0; z < nz - 1; ++z)
        for (int y = 0; y < ny - 1; ++y)
            for (int x = 0; x < nx - 1; ++x) {
                int idx = 0;
                for (int c = 0; c < 8; ++c) {
                    const float val = vol[(size_t)(z + CORNER_OFF[c][0]) * plane +
                                          (size_t)(y + CORNER_OFF[c][1]) * nx + (x + CORNER_OFF[c][2])];
                    if ((double)val - level > 0.0) idx |= 1 << c;
                }
                out[((size_t)z * (ny - 1) + y) * (nx - 1) + x] = (uint8_t)idx;
            }
    return 0;
}
