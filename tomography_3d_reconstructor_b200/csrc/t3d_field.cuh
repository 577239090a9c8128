// t3d_field.cuh -- padded-occupancy view and exact evaluation of the marched field
// (np.pad -> float64 -> scipy.ndimage.gaussian_filter(sigma=0.5) -> float32; surface_extractor.py:43-53).
#pragma once
#include "t3d_common.cuh"

// ------------------------------------------------------------------------------------------------
// padded-occupancy view + exact field evaluation
// ------------------------------------------------------------------------------------------------
struct OccView {
    const uint32_t* bits;  // packed occupancy: word (z, y, w) at bits[z*ps + y*rs + w], w < nw
    int Z, H, W, nw;
    int rs;                // row stride in words (= nw for a compact volume)
    long long ps;          // plane stride in words (= H*nw for a compact volume)
    int pad;               // 0 or 1
    int Zp, Hp, Wp;        // padded extents
    int gaussian;          // 1: field = gaussian(sigma 0.5) of padded occupancy; 0: field = occupancy
    double w0, w1, w2;     // scipy _gaussian_kernel1d(0.5, 0, 2): centre, +-1, +-2
};

// scipy 'reflect' (d c b a | a b c d | d c b a)
__device__ __forceinline__ int reflect_idx(int i, int n)
{
    if (i >= 0 && i < n) return i;
    if (n == 1) return 0;
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

// padded occupancy bit (coordinates already inside [0, Np))
__device__ __forceinline__ uint32_t pbit(const OccView& v, int zp, int yp, int xp)
{
    const int z = zp - v.pad, y = yp - v.pad, x = xp - v.pad;
    if (z < 0 || z >= v.Z || y < 0 || y >= v.H || x < 0 || x >= v.W) return 0u;
    return (v.bits[(z * v.ps + (long long)y * v.rs) + (x >> 5)] >> (x & 31)) & 1u;
}

// bits of padded row (zp, yp) at padded x = xs .. xs+4 (x reflected at the padded border)
__device__ __forceinline__ uint32_t get5(const OccView& v, int zp, int yp, int xs)
{
    const int z = zp - v.pad, y = yp - v.pad;
    if (z < 0 || z >= v.Z || y < 0 || y >= v.H) return 0u;
    const uint32_t* row = v.bits + (z * v.ps + (long long)y * v.rs);
    if (xs >= 0 && xs + 4 < v.Wp) {
        const int ox = xs - v.pad;
        const int w = ox >> 5, sh = ox & 31;
        const uint32_t lo = (w >= 0 && w < v.nw) ? row[w] : 0u;
        const uint32_t hi = (w + 1 >= 0 && w + 1 < v.nw) ? row[w + 1] : 0u;
        const unsigned long long win = (unsigned long long)lo | ((unsigned long long)hi << 32);
        return (uint32_t)(win >> sh) & 31u;
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int ox = reflect_idx(xs + k, v.Wp) - v.pad;
        if (ox >= 0 && ox < v.W) r |= ((row[ox >> 5] >> (ox & 31)) & 1u) << k;
    }
    return r;
}

// one symmetric 5-tap correlation in scipy's order: c*w0 + (a_-2 + a_+2)*w2 + (a_-1 + a_+1)*w1, no FMA
__device__ __forceinline__ double corr5(double m2, double m1, double c, double p1, double p2, double w0, double w1, double w2)
{
    double t = __dmul_rn(c, w0);
    t = __dadd_rn(t, __dmul_rn(__dadd_rn(m2, p2), w2));
    t = __dadd_rn(t, __dmul_rn(__dadd_rn(m1, p1), w1));
    return t;
}

// float32(field) at padded voxel (zp, yp, xp): separable float64 passes along z, then y, then x,
// each with reflect boundary, exactly like scipy.ndimage.gaussian_filter
__device__ __forceinline__ float field_value(const OccView& v, int zp, int yp, int xp)
{
    if (!v.gaussian) return pbit(v, zp, yp, xp) ? 1.0f : 0.0f;
    uint32_t b[5][5];
#pragma unroll
    for (int dz = 0; dz < 5; ++dz) {
        const int zr = reflect_idx(zp - 2 + dz, v.Zp);
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) b[dz][dy] = get5(v, zr, reflect_idx(yp - 2 + dy, v.Hp), xp - 2);
    }
    double Y[5];
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
        double A[5];
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
            const double c = (double)((b[2][dy] >> dx) & 1u);
            const double n1 = (double)(((b[1][dy] >> dx) & 1u) + ((b[3][dy] >> dx) & 1u));
            const double n2 = (double)(((b[0][dy] >> dx) & 1u) + ((b[4][dy] >> dx) & 1u));
            double t = __dmul_rn(c, v.w0);
            t = __dadd_rn(t, __dmul_rn(n2, v.w2));
            t = __dadd_rn(t, __dmul_rn(n1, v.w1));
            A[dy] = t;
        }
        Y[dx] = corr5(A[0], A[1], A[2], A[3], A[4], v.w0, v.w1, v.w2);
    }
    return __double2float_rn(corr5(Y[0], Y[1], Y[2], Y[3], Y[4], v.w0, v.w1, v.w2));
}

// padded occupancy word (zp, yp, wp) in padded x coordinates (no reflection; outside = 0)
__device__ __forceinline__ uint32_t pword(const OccView& v, int zp, int yp, int wp)
{
    const int z = zp - v.pad, y = yp - v.pad;
    if (z < 0 || z >= v.Z || y < 0 || y >= v.H || wp < 0) return 0u;
    const uint32_t* row = v.bits + (z * v.ps + (long long)y * v.rs);
    const uint32_t cur = (wp < v.nw) ? row[wp] : 0u;
    if (!v.pad) return cur;
    const uint32_t prev = (wp - 1 >= 0 && wp - 1 < v.nw) ? row[wp - 1] : 0u;
    return (cur << 1) | (prev >> 31);
}


// bits of padded row (zp, yp) at padded x = xs .. xs+5 (x reflected at the padded border)
__device__ __forceinline__ uint32_t get6(const OccView& v, int zp, int yp, int xs)
{
    const int z = zp - v.pad, y = yp - v.pad;
    if (z < 0 || z >= v.Z || y < 0 || y >= v.H) return 0u;
    const uint32_t* row = v.bits + (z * v.ps + (long long)y * v.rs);
    if (xs >= 0 && xs + 5 < v.Wp) {
        const int ox = xs - v.pad;
        const int w = ox >> 5, sh = ox & 31;
        const uint32_t lo = (w >= 0 && w < v.nw) ? row[w] : 0u;
        const uint32_t hi = (sh > 26 && w + 1 >= 0 && w + 1 < v.nw) ? row[w + 1] : 0u;
        const unsigned long long win = (unsigned long long)lo | ((unsigned long long)hi << 32);
        return (uint32_t)(win >> sh) & 63u;
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int ox = reflect_idx(xs + k, v.Wp) - v.pad;
        if (ox >= 0 && ox < v.W) r |= ((row[ox >> 5] >> (ox & 31)) & 1u) << k;
    }
    return r;
}

// z-pass lookup: the first (z) pass of the separable filter sees only 0/1 samples, so its result is one of 18
// values indexed by (centre c, n2 = a[-2]+a[+2], n1 = a[-1]+a[+1]):  c*w0 + n2*w2 + n1*w1 in scipy's order.
// SWAR form: the six x columns of one (plane, row) fetch are six bits; spread[] moves bit k to bit 5k, and the five planes
// of a z pass add up -- centre << 4, outer pair << 2, inner pair << 0 -- to six 5-bit indices (c << 4 | n2 << 2 | n1) at once.
struct ZLut {
    double z[32];           // indexed by c << 4 | n2 << 2 | n1 (n1, n2 in 0..2)
    uint32_t spread[64];    // 6 bits -> one bit per 5-bit field
};

__device__ __forceinline__ void fill_zlut(const OccView& v, ZLut& L)   // by the first 64 threads of the block; __syncthreads() after it
{
    const int i = threadIdx.x;
    if (i < 32) {
        const int c = i >> 4, n2 = (i >> 2) & 3, n1 = i & 3;
        double t = __dmul_rn((double)c, v.w0);
        t = __dadd_rn(t, __dmul_rn((double)n2, v.w2));
        t = __dadd_rn(t, __dmul_rn((double)n1, v.w1));
        L.z[i] = t;
    }
    if (i < 64) {
        uint32_t r = 0;
        for (int k = 0; k < 6; ++k) r |= ((uint32_t)(i >> k) & 1u) << (5 * k);
        L.spread[i] = r;
    }
}

// float32 field at both end points of the grid edge (z,y,x) -> +1 along AXIS (0 = z, 1 = y, 2 = x), sharing one
// neighbourhood fetch of NY rows x NZ planes x 6 x-bits.  Same arithmetic as field_value().
template <int AXIS>
__device__ __forceinline__ void edge_field_values(const OccView& v, const ZLut& L, int z, int y, int x, float& fa, float& fb)
{
    if (!v.gaussian) {
        fa = pbit(v, z, y, x) ? 1.0f : 0.0f;
        fb = pbit(v, z + (AXIS == 0), y + (AXIS == 1), x + (AXIS == 2)) ? 1.0f : 0.0f;
        return;
    }
    constexpr int NZ = 5 + (AXIS == 0), NY = 5 + (AXIS == 1), NE = 1 + (AXIS == 0);
    // idx[e][dy]: six 5-bit z-pass indices (one per x column) of row dy for the five planes e .. e+4
    uint32_t idx[NE][NY];
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
        for (int dy = 0; dy < NY; ++dy) idx[e][dy] = 0u;
    // plane j of a z pass contributes: outer pair (j = 0, 4) << 2, inner pair (j = 1, 3) << 0, centre (j = 2) << 4
    auto add_plane = [&](int dy, int dz, uint32_t six) {
        const uint32_t s = L.spread[six];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int j = dz - e;
            if (j < 0 || j > 4) continue;
            idx[e][dy] += s << (j == 2 ? 4 : (j == 0 || j == 4) ? 2 : 0);
        }
    };
    // un-padded coordinates of the first fetched voxel; interior = the whole fetch lies inside the occupancy volume
    const int oz0 = z - 2 - v.pad, oy0 = y - 2 - v.pad, ox0 = x - 2 - v.pad;
    if (oz0 >= 0 && oz0 + NZ <= v.Z && oy0 >= 0 && oy0 + NY <= v.H && ox0 >= 0 && ox0 + 6 <= v.W) {
        const int w = ox0 >> 5, sh = ox0 & 31;
        const bool two = sh > 26;  // the 6-bit window straddles two words
        const uint32_t* p = v.bits + (oz0 * v.ps + (long long)oy0 * v.rs) + w;
        const long long dzs = v.ps;
#pragma unroll
        for (int dy = 0; dy < NY; ++dy) {
#pragma unroll
            for (int dz = 0; dz < NZ; ++dz) {
                const uint32_t* q = p + dz * dzs + dy * v.rs;
                const uint32_t lo = q[0], hi = two ? q[1] : 0u;
                add_plane(dy, dz, __funnelshift_r(lo, hi, sh) & 63u);
            }
        }
    } else {
#pragma unroll
        for (int dy = 0; dy < NY; ++dy) {
            const int yr = reflect_idx(y - 2 + dy, v.Hp);
#pragma unroll
            for (int dz = 0; dz < NZ; ++dz) add_plane(dy, dz, get6(v, reflect_idx(z - 2 + dz, v.Zp), yr, x - 2));
        }
    }
    // z pass of column (dy, k) for end point e
    auto zpass = [&](int dy, int k, int e) -> double { return L.z[(idx[e][dy] >> (5 * k)) & 31u]; };
    if (AXIS == 2) {
        // end points differ by one in x: the z and y passes of the six x columns are shared
        double Y[6];
#pragma unroll
        for (int k = 0; k < 6; ++k)
            Y[k] = corr5(zpass(0, k, 0), zpass(1, k, 0), zpass(2, k, 0), zpass(3, k, 0), zpass(4, k, 0), v.w0, v.w1, v.w2);
        fa = __double2float_rn(corr5(Y[0], Y[1], Y[2], Y[3], Y[4], v.w0, v.w1, v.w2));
        fb = __double2float_rn(corr5(Y[1], Y[2], Y[3], Y[4], Y[5], v.w0, v.w1, v.w2));
    } else if (AXIS == 1) {
        // end points differ by one in y: the z pass of the six rows is shared
        double Ya[5], Yb[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double A[6];
#pragma unroll
            for (int dy = 0; dy < 6; ++dy) A[dy] = zpass(dy, k, 0);
            Ya[k] = corr5(A[0], A[1], A[2], A[3], A[4], v.w0, v.w1, v.w2);
            Yb[k] = corr5(A[1], A[2], A[3], A[4], A[5], v.w0, v.w1, v.w2);
        }
        fa = __double2float_rn(corr5(Ya[0], Ya[1], Ya[2], Ya[3], Ya[4], v.w0, v.w1, v.w2));
        fb = __double2float_rn(corr5(Yb[0], Yb[1], Yb[2], Yb[3], Yb[4], v.w0, v.w1, v.w2));
    } else {
        // end points differ by one in z: different taps in every column
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double Y[5];
#pragma unroll
            for (int k = 0; k < 5; ++k)
                Y[k] = corr5(zpass(0, k, e), zpass(1, k, e), zpass(2, k, e), zpass(3, k, e), zpass(4, k, e), v.w0, v.w1, v.w2);
            const float f = __double2float_rn(corr5(Y[0], Y[1], Y[2], Y[3], Y[4], v.w0, v.w1, v.w2));
            if (e == 0) fa = f; else fb = f;
        }
    }
}

static inline OccView t3d_make_view(const void* occ_bits, int Z, int H, int W, int pad, int gaussian, const double* w3)
{
    OccView v;
    v.bits = (const uint32_t*)occ_bits;
    v.Z = Z; v.H = H; v.W = W; v.nw = t3d_wpr(W);
    v.rs = v.nw; v.ps = (long long)H * v.nw;
    v.pad = pad ? 1 : 0;
    v.Zp = Z + 2 * v.pad; v.Hp = H + 2 * v.pad; v.Wp = W + 2 * v.pad;
    v.gaussian = gaussian ? 1 : 0;
    // scipy.ndimage._filters._gaussian_kernel1d(0.5, 0, 2)
    v.w0 = w3 ? w3[0] : 0x1.92b965ef5aaeep-1;
    v.w1 = w3 ? w3[1] : 0x1.b405b9842b206p-4;
    v.w2 = w3 ? w3[2] : 0x1.14aebe6a24088p-12;
    return v;
}

// the same (Z,H,W) occupancy stored inside a larger buffer: word (z,y,w) at origin[z*plane_stride + y*row_stride + w]
static inline OccView t3d_make_view_strided(const void* origin, int Z, int H, int W, int row_stride, long long plane_stride, int pad,
                                            int gaussian, const double* w3)
{
    OccView v = t3d_make_view(origin, Z, H, W, pad, gaussian, w3);
    v.rs = row_stride; v.ps = plane_stride;
    return v;
}

// t3d_mc.cu: vertices of a mesh emitted on a sign volume whose x coordinates are shifted by x_off against the padded
// grid of `view` (the padded-storage layout of the fused pipeline: x_off = 127)
int t3d_mc_vertices_view_dev(const OccView& view, int x_off, const void* vkeys_u64, const void* sizes_u64, uint32_t cap_verts,
                             int unpad_shift, int z_offset, const void* cum_f64, const void* adj_f64, int n_cum, double mm_per_pixel_y,
                             double mm_per_pixel_x, int scale_in_f64, int which_blocks, void* verts_f32, void* stream);
// t3d_mc.cu: counting / emission passes whose ambiguity tests evaluate the field on a strided view (same x_off convention)
int t3d_mc_words_view_dev(const OccView& view, int x_off, const void* sign_bits, int Zs, int Hs, int Ws, int z_begin, int z_end,
                          const void* ballots_u32, const void* chunkbase_u32, uint32_t cap_active, const void* sizes_u64,
                          void* aw_idx_u32, void* aw_cnt_u32, void* n_ambiguous_u64, void* stream);
int t3d_mc_emit_view_dev(const OccView& view, int x_off, const void* sign_bits, int Zs, int Hs, int Ws, int z_begin, int z_end,
                         const void* ballots_u32, const void* chunkbase_u32, const void* aw_idx_u32, const void* aw_base_u32,
                         uint32_t cap_active, const void* sizes_u64, uint32_t cap_verts, uint32_t cap_faces, void* vkeys_u64,
                         void* faces_i32, int parts, void* stream);
