// t3d_surface.cu -- SurfaceExtractor kernels (sm_100a).
//
// Reference semantics (file:line into the reference repository, and SURVEY.md section 8a):
//   field      : np.pad(1) -> float64 -> scipy gaussian_filter(sigma=0.5), surface_extractor.py:43-53
//   mc         : skimage.measure.marching_cubes(volume, level=0.5), surface_extractor.py:55
//   transform  : un-pad, _apply_variable_slice_depths, mm scaling, surface_extractor.py:57-65,82-113
//   canonical  : _ensure_manifold_mesh (np.unique rows + degenerate-face drop), :115-126
//   measures   : calculate_mesh_volume / calculate_surface_area, :128-148
//
// Design.  The marched field is a 5x5x5 Gaussian blur of a 0/1 occupancy.  Its centre tap is 0.48665 and
// each face tap 0.06586 (weights sum to 1), hence
//     voxel set   and >=1 (reflected) face neighbour set   => field >= 0.5525 > 0.5
//     voxel unset and >=1 (reflected) face neighbour unset => field <= 0.4475 < 0.5
// so the SIGN of (field - 0.5) equals the occupancy bit except at voxels whose six face neighbours all
// disagree with them; only those are evaluated exactly.  All topology (cube cases, vertex ownership,
// counts, ranks) is therefore pure bit arithmetic on a packed sign volume; float64 arithmetic in scipy's
// exact summation order is spent only on the two end points of each cut edge (vertex interpolation).
#include <cub/device/device_radix_sort.cuh>

#include "mc_tables.h"
#include "t3d_common.cuh"

// ------------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------------
__device__ __align__(16) const int8_t g_tri_table[256][T3D_MC_ROW] = {T3D_TRI_TABLE_ROWS};
static const int8_t h_tri_table[256][T3D_MC_ROW] = {T3D_TRI_TABLE_ROWS};

struct McLuts {
    uint8_t ntri[256];
    uint8_t amb[256];
};
__constant__ McLuts c_luts;
static bool g_luts_ready = false;

static int host_is_ambiguous(int idx)
{
    static const int FC[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 3, 7, 4}, {1, 2, 6, 5}};
    for (int f = 0; f < 6; ++f) {
        const int a = (idx >> FC[f][0]) & 1, b = (idx >> FC[f][1]) & 1, c = (idx >> FC[f][2]) & 1, d = (idx >> FC[f][3]) & 1;
        if (a == c && b == d && a != b) return 1;
    }
    static const int DG[4][2] = {{0, 6}, {1, 7}, {2, 4}, {3, 5}};
    for (int k = 0; k < 4; ++k) {
        const int m = (1 << DG[k][0]) | (1 << DG[k][1]);
        if (idx == m || idx == (255 ^ m)) return 1;
    }
    return 0;
}

static int ensure_luts()
{
    if (g_luts_ready) return 0;
    McLuts l;
    for (int i = 0; i < 256; ++i) {
        int n = 0;
        while (n < T3D_MC_ROW && h_tri_table[i][n] >= 0) n += 3;
        l.ntri[i] = (uint8_t)(n / 3);
        l.amb[i] = (uint8_t)host_is_ambiguous(i);
    }
    T3D_CUDA(cudaMemcpyToSymbol(c_luts, &l, sizeof(l)));
    g_luts_ready = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// padded-occupancy view + exact field evaluation
// ------------------------------------------------------------------------------------------------
struct OccView {
    const uint32_t* bits;  // (Z, H, nw) packed occupancy
    int Z, H, W, nw;
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
    return (v.bits[((int64_t)z * v.H + y) * v.nw + (x >> 5)] >> (x & 31)) & 1u;
}

// bits of padded row (zp, yp) at padded x = xs .. xs+4 (x reflected at the padded border)
__device__ __forceinline__ uint32_t get5(const OccView& v, int zp, int yp, int xs)
{
    const int z = zp - v.pad, y = yp - v.pad;
    if (z < 0 || z >= v.Z || y < 0 || y >= v.H) return 0u;
    const uint32_t* row = v.bits + ((int64_t)z * v.H + y) * v.nw;
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
__device__ float field_value(const OccView& v, int zp, int yp, int xp)
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
    const uint32_t* row = v.bits + ((int64_t)z * v.H + y) * v.nw;
    const uint32_t cur = (wp < v.nw) ? row[wp] : 0u;
    if (!v.pad) return cur;
    const uint32_t prev = (wp - 1 >= 0 && wp - 1 < v.nw) ? row[wp - 1] : 0u;
    return (cur << 1) | (prev >> 31);
}

// ------------------------------------------------------------------------------------------------
// sign volume: S(q) = float32(field(q)) > 0.5, packed in padded coordinates (Zp, Hp, nwp)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_field_sign(OccView v, uint32_t* __restrict__ sign, int nwp,
                                                    unsigned long long* __restrict__ n_exact)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)v.Zp * v.Hp * nwp;
    if (i >= total) return;
    const int wp = (int)(i % nwp);
    const int64_t r = i / nwp;
    const int yp = (int)(r % v.Hp), zp = (int)(r / v.Hp);
    const uint32_t vm = valid_mask(wp, v.Wp);
    const uint32_t c = pword(v, zp, yp, wp);
    const uint32_t zm = pword(v, reflect_idx(zp - 1, v.Zp), yp, wp), zq = pword(v, reflect_idx(zp + 1, v.Zp), yp, wp);
    const uint32_t ym = pword(v, zp, reflect_idx(yp - 1, v.Hp), wp), yq = pword(v, zp, reflect_idx(yp + 1, v.Hp), wp);
    const uint32_t lw = pword(v, zp, yp, wp - 1), rw = pword(v, zp, yp, wp + 1);
    uint32_t xm = (c << 1) | (lw >> 31);  // value at x-1
    uint32_t xq = (c >> 1) | (rw << 31);  // value at x+1
    if (wp == 0) xm = (xm & ~1u) | (c & 1u);  // reflect: x = -1 -> x = 0
    {
        const int last = v.Wp - 1;            // reflect: x = Wp -> x = Wp-1
        if ((last >> 5) == wp) {
            const uint32_t bit = 1u << (last & 31);
            xq = (xq & ~bit) | (c & bit);
        }
    }
    const uint32_t any1 = zm | zq | ym | yq | xm | xq;
    const uint32_t all1 = zm & zq & ym & yq & xm & xq;
    uint32_t s = c & any1;
    uint32_t need = ((c & ~any1) | (~c & all1)) & vm;
    if (need) {
        atomicAdd(n_exact, (unsigned long long)__popc(need));
        while (need) {
            const int b = __ffs(need) - 1;
            need &= need - 1;
            if (field_value(v, zp, yp, (wp << 5) + b) > 0.5f) s |= 1u << b; else s &= ~(1u << b);
        }
    }
    sign[i] = s & vm;
}

// ------------------------------------------------------------------------------------------------
// marching cubes pass 1: per voxel-row counts of owned cut edges (x, y, z) and triangles.
// One warp per row (zp, yp) of the sign volume; lane = word.
// ------------------------------------------------------------------------------------------------
struct RowWords {
    uint32_t s00, s01, s10, s11;  // rows (z,y) (z,y+1) (z+1,y) (z+1,y+1), word w
    uint32_t n00, n01, n10, n11;  // bit 0 of word w+1 of each row (0/1)
};

__device__ __forceinline__ uint32_t ldw(const uint32_t* row, int w, int nwp) { return (row && w < nwp) ? row[w] : 0u; }

__device__ __forceinline__ RowWords load_rows(const uint32_t* r00, const uint32_t* r01, const uint32_t* r10,
                                              const uint32_t* r11, int w, int nwp)
{
    RowWords q;
    q.s00 = ldw(r00, w, nwp); q.s01 = ldw(r01, w, nwp); q.s10 = ldw(r10, w, nwp); q.s11 = ldw(r11, w, nwp);
    q.n00 = ldw(r00, w + 1, nwp) & 1u; q.n01 = ldw(r01, w + 1, nwp) & 1u;
    q.n10 = ldw(r10, w + 1, nwp) & 1u; q.n11 = ldw(r11, w + 1, nwp) & 1u;
    return q;
}

__device__ __forceinline__ uint32_t shr1(uint32_t s, uint32_t nbit) { return (s >> 1) | (nbit << 31); }  // value at x+1

// 8-bit cube case of the cube whose origin is bit b of the current word
__device__ __forceinline__ int cube_case(const RowWords& q, int b)
{
    const uint32_t a00 = shr1(q.s00, q.n00), a01 = shr1(q.s01, q.n01), a10 = shr1(q.s10, q.n10), a11 = shr1(q.s11, q.n11);
    return (int)(((q.s00 >> b) & 1u) | (((a00 >> b) & 1u) << 1) | (((a01 >> b) & 1u) << 2) | (((q.s01 >> b) & 1u) << 3) |
                 (((q.s10 >> b) & 1u) << 4) | (((a10 >> b) & 1u) << 5) | (((a11 >> b) & 1u) << 6) | (((q.s11 >> b) & 1u) << 7));
}

__global__ void __launch_bounds__(256) k_mc_count(const uint32_t* __restrict__ sign, int Zp, int Hp, int Wp, int nwp,
                                                  uint32_t* __restrict__ rowcnt, int64_t n_rows,
                                                  unsigned long long* __restrict__ n_ambiguous)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int y = (int)(row % Hp), z = (int)(row / Hp);
    const uint32_t l = lane_id();
    const bool hy = (y + 1 < Hp), hz = (z + 1 < Zp);
    const uint32_t* r00 = sign + row * nwp;
    const uint32_t* r01 = hy ? r00 + nwp : nullptr;
    const uint32_t* r10 = hz ? r00 + (int64_t)Hp * nwp : nullptr;
    const uint32_t* r11 = (hy && hz) ? r10 + nwp : nullptr;
    uint32_t nX = 0, nY = 0, nZ = 0, nT = 0, nA = 0;
    for (int w = l; w < nwp; w += 32) {
        const RowWords q = load_rows(r00, r01, r10, r11, w, nwp);
        const uint32_t vm = valid_mask(w, Wp), em = valid_mask(w, Wp - 1);  // em: x+1 still inside
        nX += __popc((q.s00 ^ shr1(q.s00, q.n00)) & em);
        if (hy) nY += __popc((q.s00 ^ q.s01) & vm);
        if (hz) nZ += __popc((q.s00 ^ q.s10) & vm);
        if (hy && hz) {
            const uint32_t o = q.s00 | q.s01 | q.s10 | q.s11, a = q.s00 & q.s01 & q.s10 & q.s11;
            const uint32_t on = (q.n00 | q.n01 | q.n10 | q.n11), an = (q.n00 & q.n01 & q.n10 & q.n11);
            uint32_t act = ((o | shr1(o, on)) & ~(a & shr1(a, an))) & em;
            while (act) {
                const int b = __ffs(act) - 1;
                act &= act - 1;
                const int cs = cube_case(q, b);
                nT += c_luts.ntri[cs];
                nA += c_luts.amb[cs];
            }
        }
    }
    nX = warp_sum(nX); nY = warp_sum(nY); nZ = warp_sum(nZ); nT = warp_sum(nT); nA = warp_sum(nA);
    if (l == 0) {
        rowcnt[row] = nX;
        rowcnt[n_rows + row] = nY;
        rowcnt[2 * n_rows + row] = nZ;
        rowcnt[3 * n_rows + row] = nT;
        if (nA) atomicAdd(n_ambiguous, (unsigned long long)nA);
    }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of n_arrays independent uint32 arrays of length n (in place), totals -> totals[k]
// three kernels: block reduce, scan of block sums (one block per array), block scan + offset
// ------------------------------------------------------------------------------------------------
#define SC_THREADS 256
#define SC_ITEMS 8
#define SC_TILE (SC_THREADS * SC_ITEMS)

__global__ void __launch_bounds__(SC_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, int64_t n, int64_t stride,
                                                            unsigned long long* __restrict__ block_sums, int n_blocks)
{
    const uint32_t* a = in + (int64_t)blockIdx.y * stride;
    const int64_t base = (int64_t)blockIdx.x * SC_TILE;
    unsigned long long s = 0;
    for (int k = 0; k < SC_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * SC_THREADS + threadIdx.x;
        if (i < n) s += a[i];
    }
    s = warp_sum(s);
    __shared__ unsigned long long sh[SC_THREADS / 32];
    if (lane_id() == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < SC_THREADS / 32; ++k) t += sh[k];
        block_sums[(int64_t)blockIdx.y * n_blocks + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_scan_block_sums(unsigned long long* __restrict__ block_sums, int n_blocks,
                                                          unsigned long long* __restrict__ totals)
{
    unsigned long long* a = block_sums + (int64_t)blockIdx.x * n_blocks;
    __shared__ unsigned long long sh[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = (i < n_blocks) ? a[i] : 0ull;
        unsigned long long x = v;
        const uint32_t l = lane_id();
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, x, o);
            if (l >= (uint32_t)o) x += t;
        }
        if (l == 31) sh[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long y = sh[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, y, o);
                if (l >= (uint32_t)o) y += t;
            }
            sh[threadIdx.x] = y;
        }
        __syncthreads();
        const unsigned long long warp_off = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0ull;
        const unsigned long long carry = carry_s;
        if (i < n_blocks) a[i] = carry + warp_off + x - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

template <typename OutT>
__global__ void __launch_bounds__(SC_THREADS) k_scan_final(const uint32_t* __restrict__ in, OutT* __restrict__ out, int64_t n,
                                                           int64_t stride, const unsigned long long* __restrict__ block_sums,
                                                           int n_blocks)
{
    const uint32_t* a = in + (int64_t)blockIdx.y * stride;
    OutT* o = out + (int64_t)blockIdx.y * stride;
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;  // blocked arrangement
    uint32_t v[SC_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        v[k] = (base + k < n) ? a[base + k] : 0u;
        s += v[k];
    }
    const uint32_t incl = warp_incl_scan(s);
    __shared__ uint32_t sh[SC_THREADS / 32];
    if (lane_id() == 31) sh[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) woff += sh[k];
    unsigned long long run = block_sums[(int64_t)blockIdx.y * n_blocks + blockIdx.x] + woff + (incl - s);
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        if (base + k < n) o[base + k] = (OutT)run;
        run += v[k];
    }
}

extern "C" int64_t t3d_scan_workspace_bytes(int64_t n, int n_arrays)
{
    const int64_t nb = (n + SC_TILE - 1) / SC_TILE;
    return (nb * n_arrays + 16) * 8;
}

// in: n_arrays arrays of n uint32 (array k starts at in + k*n); out: same layout, uint32 (out_is_u64 = 0)
// or uint64 (1); may alias `in` only for uint32 output.  totals: n_arrays uint64 (device).
extern "C" int t3d_exclusive_scan_u32(const void* in, void* out, int64_t n, int n_arrays, int out_is_u64, void* totals_u64,
                                      void* workspace, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0 || n_arrays <= 0) {
        if (n_arrays > 0) T3D_CUDA(cudaMemsetAsync(totals_u64, 0, 8 * n_arrays, st));
        return 0;
    }
    const int64_t nb = (n + SC_TILE - 1) / SC_TILE;
    if (nb > 0x7fffffff) { t3d_set_error("t3d_exclusive_scan_u32: too many elements"); return 2; }
    unsigned long long* bs = (unsigned long long*)workspace;
    dim3 grid((unsigned)nb, n_arrays);
    k_scan_reduce<<<grid, SC_THREADS, 0, st>>>((const uint32_t*)in, n, n, bs, (int)nb);
    k_scan_block_sums<<<n_arrays, 1024, 0, st>>>(bs, (int)nb, (unsigned long long*)totals_u64);
    if (out_is_u64)
        k_scan_final<unsigned long long><<<grid, SC_THREADS, 0, st>>>((const uint32_t*)in, (unsigned long long*)out, n, n, bs, (int)nb);
    else
        k_scan_final<uint32_t><<<grid, SC_THREADS, 0, st>>>((const uint32_t*)in, (uint32_t*)out, n, n, bs, (int)nb);
    T3D_CHECK_LAUNCH("t3d_exclusive_scan_u32");
    t3d_count_launches(3);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// marching cubes pass 2: vertex + face emission.
// Vertex ids: [all x-edge vertices | all y-edge vertices | all z-edge vertices], each block in raster
// order of the owning voxel (edge owned by its lower corner) => id = block offset + row base + rank in row.
// Faces are emitted in the reference's order (cubes z-major, y, x fastest; table order inside a cube)
// with the winding reversed (gradient_direction='descent').
// ------------------------------------------------------------------------------------------------
struct EmitParams {
    OccView occ;
    const uint32_t* sign;
    int nwp;
    int64_t n_rows;
    const uint32_t* rowbase;  // 4 arrays of n_rows (exclusive scans of the counts): X, Y, Z, T
    uint32_t offY, offZ;      // offX = 0
    // vertex transform
    float shift;              // 1 if manifold else 0 (surface_extractor.py:57-60)
    const double* cum;        // cumulative adjusted depths, n_cum entries (nullptr / 0: no z map)
    const double* adj;        // adjusted depths, n_cum-1 entries
    int n_cum;
    double mm_y, mm_x;        // mm per pixel
    int scale_f64;            // 1: multiply in float64 then round (numpy float64 scalar operand), 0: float32 multiply
    float* verts;             // (V, 3) z, y, x
    int32_t* faces;           // (F, 3)
};

__device__ __forceinline__ void emit_vertex(const EmitParams& p, uint32_t id, int axis, int z, int y, int x)
{
    // end points of the edge: (z,y,x) and +1 along `axis` (0 = z, 1 = y, 2 = x)
    const double va = (double)field_value(p.occ, z, y, x) - 0.5;
    const double vb = (double)field_value(p.occ, z + (axis == 0), y + (axis == 1), x + (axis == 2)) - 0.5;
    const double wa = __ddiv_rn(1.0, __dadd_rn(1.1920928955078125e-07, fabs(va)));
    const double wb = __ddiv_rn(1.0, __dadd_rn(1.1920928955078125e-07, fabs(vb)));
    const double frac = __ddiv_rn(wb, __dadd_rn(wa, wb));
    double pz = (double)z, py = (double)y, px = (double)x;
    if (axis == 0) pz = __dadd_rn(pz, frac); else if (axis == 1) py = __dadd_rn(py, frac); else px = __dadd_rn(px, frac);
    float fz = __fsub_rn(__double2float_rn(pz), p.shift);
    float fy = __fsub_rn(__double2float_rn(py), p.shift);
    float fx = __fsub_rn(__double2float_rn(px), p.shift);
    if (p.n_cum > 0) {
        // surface_extractor.py:98-113, closed form verified bit-exact in SURVEY.md V8
        if (fz < 0.0f) fz = 0.0f;
        else if (fz >= (float)(p.n_cum - 1)) fz = __double2float_rn(p.cum[p.n_cum - 1]);
        else {
            const float fl = floorf(fz);
            const int lo = (int)fl;
            const float fr = __fsub_rn(fz, fl);
            const int ai = min(lo, p.n_cum - 2);
            fz = __double2float_rn(__dadd_rn(p.cum[lo], __dmul_rn((double)fr, p.adj[ai])));
        }
    }
    if (p.scale_f64) {
        fy = __double2float_rn(__dmul_rn((double)fy, p.mm_y));
        fx = __double2float_rn(__dmul_rn((double)fx, p.mm_x));
    } else {
        fy = __fmul_rn(fy, __double2float_rn(p.mm_y));
        fx = __fmul_rn(fx, __double2float_rn(p.mm_x));
    }
    float* o = p.verts + 3 * (int64_t)id;
    o[0] = fz; o[1] = fy; o[2] = fx;
}

__device__ __forceinline__ uint32_t lt_mask(int b) { return b >= 32 ? 0xffffffffu : ((1u << b) - 1u); }

#define EM_WARPS 8

__global__ void __launch_bounds__(EM_WARPS * 32) k_mc_emit(EmitParams p)
{
    __shared__ __align__(16) int8_t s_tri[256][T3D_MC_ROW];
    {
        const int4* src = reinterpret_cast<const int4*>(&g_tri_table[0][0]);
        int4* dst = reinterpret_cast<int4*>(&s_tri[0][0]);
        for (int i = threadIdx.x; i < 256; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * EM_WARPS + (threadIdx.x >> 5);
    if (row >= p.n_rows) return;
    const int Zp = p.occ.Zp, Hp = p.occ.Hp, Wp = p.occ.Wp, nwp = p.nwp;
    const int y = (int)(row % Hp), z = (int)(row / Hp);
    const uint32_t l = lane_id();
    const bool hy = (y + 1 < Hp), hz = (z + 1 < Zp);
    const uint32_t* r00 = p.sign + row * nwp;
    const uint32_t* r01 = hy ? r00 + nwp : nullptr;
    const uint32_t* r10 = hz ? r00 + (int64_t)Hp * nwp : nullptr;
    const uint32_t* r11 = (hy && hz) ? r10 + nwp : nullptr;
    const int64_t N = p.n_rows;
    // running bases (advance chunk by chunk)
    uint32_t bX00 = p.rowbase[row];
    uint32_t bX01 = hy ? p.rowbase[row + 1] : 0u;
    uint32_t bX10 = hz ? p.rowbase[row + Hp] : 0u;
    uint32_t bX11 = (hy && hz) ? p.rowbase[row + Hp + 1] : 0u;
    uint32_t bY0 = p.offY + p.rowbase[N + row];
    uint32_t bY1 = hz ? p.offY + p.rowbase[N + row + Hp] : 0u;
    uint32_t bZ0 = p.offZ + p.rowbase[2 * N + row];
    uint32_t bZ1 = hy ? p.offZ + p.rowbase[2 * N + row + 1] : 0u;
    uint32_t bT = p.rowbase[3 * N + row];

    for (int w0 = 0; w0 < nwp; w0 += 32) {
        const int w = w0 + l;
        const RowWords q = load_rows(r00, r01, r10, r11, w, nwp);
        const uint32_t vm = valid_mask(w, Wp), em = valid_mask(w, Wp - 1);
        const uint32_t X00 = (q.s00 ^ shr1(q.s00, q.n00)) & em;
        const uint32_t X01 = hy ? (q.s01 ^ shr1(q.s01, q.n01)) & em : 0u;
        const uint32_t X10 = hz ? (q.s10 ^ shr1(q.s10, q.n10)) & em : 0u;
        const uint32_t X11 = (hy && hz) ? (q.s11 ^ shr1(q.s11, q.n11)) & em : 0u;
        const uint32_t Y0 = hy ? (q.s00 ^ q.s01) & vm : 0u;
        const uint32_t Y1 = (hy && hz) ? (q.s10 ^ q.s11) & vm : 0u;
        const uint32_t Z0 = hz ? (q.s00 ^ q.s10) & vm : 0u;
        const uint32_t Z1 = (hy && hz) ? (q.s01 ^ q.s11) & vm : 0u;
        uint32_t act = 0;
        if (hy && hz) {
            const uint32_t o = q.s00 | q.s01 | q.s10 | q.s11, a = q.s00 & q.s01 & q.s10 & q.s11;
            const uint32_t on = (q.n00 | q.n01 | q.n10 | q.n11), an = (q.n00 & q.n01 & q.n10 & q.n11);
            act = ((o | shr1(o, on)) & ~(a & shr1(a, an))) & em;
        }
        uint32_t nt = 0;
        for (uint32_t m = act; m;) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            nt += c_luts.ntri[cube_case(q, b)];
        }
        // packed warp scans (two 16-bit fields per word; per-chunk sums <= 1024 and <= 5120 for triangles)
        const uint32_t c0 = __popc(X00) | (__popc(X01) << 16), c1 = __popc(X10) | (__popc(X11) << 16);
        const uint32_t c2 = __popc(Y0) | (__popc(Y1) << 16), c3 = __popc(Z0) | (__popc(Z1) << 16);
        const uint32_t i0 = warp_incl_scan(c0), i1 = warp_incl_scan(c1), i2 = warp_incl_scan(c2), i3 = warp_incl_scan(c3);
        const uint32_t it = warp_incl_scan(nt);
        const uint32_t e0 = i0 - c0, e1 = i1 - c1, e2 = i2 - c2, e3 = i3 - c3;
        const uint32_t pX00 = bX00 + (e0 & 0xffffu), pX01 = bX01 + (e0 >> 16);
        const uint32_t pX10 = bX10 + (e1 & 0xffffu), pX11 = bX11 + (e1 >> 16);
        const uint32_t pY0 = bY0 + (e2 & 0xffffu), pY1 = bY1 + (e2 >> 16);
        const uint32_t pZ0 = bZ0 + (e3 & 0xffffu), pZ1 = bZ1 + (e3 >> 16);
        uint32_t pT = bT + (it - nt);

        // ---- vertices owned by this row
        {
            uint32_t id = pX00;
            for (uint32_t m = X00; m;) { const int b = __ffs(m) - 1; m &= m - 1; emit_vertex(p, id++, 2, z, y, (w << 5) + b); }
            id = pY0;
            for (uint32_t m = Y0; m;) { const int b = __ffs(m) - 1; m &= m - 1; emit_vertex(p, id++, 1, z, y, (w << 5) + b); }
            id = pZ0;
            for (uint32_t m = Z0; m;) { const int b = __ffs(m) - 1; m &= m - 1; emit_vertex(p, id++, 0, z, y, (w << 5) + b); }
        }
        // ---- faces of the cubes of this row
        for (uint32_t m = act; m;) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int cs = cube_case(q, b);
            const uint32_t lb = lt_mask(b), lb1 = lt_mask(b + 1);
            const int8_t* rowt = s_tri[cs];
            for (int t = 0; t < T3D_MC_ROW && rowt[t] >= 0; t += 3) {
                uint32_t vid[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    uint32_t id;
                    switch (rowt[t + k]) {
                        case 0: id = pX00 + __popc(X00 & lb); break;
                        case 1: id = pY0 + __popc(Y0 & lb1); break;
                        case 2: id = pX01 + __popc(X01 & lb); break;
                        case 3: id = pY0 + __popc(Y0 & lb); break;
                        case 4: id = pX10 + __popc(X10 & lb); break;
                        case 5: id = pY1 + __popc(Y1 & lb1); break;
                        case 6: id = pX11 + __popc(X11 & lb); break;
                        case 7: id = pY1 + __popc(Y1 & lb); break;
                        case 8: id = pZ0 + __popc(Z0 & lb); break;
                        case 9: id = pZ0 + __popc(Z0 & lb1); break;
                        case 10: id = pZ1 + __popc(Z1 & lb1); break;
                        default: id = pZ1 + __popc(Z1 & lb); break;  // 11
                    }
                    vid[k] = id;
                }
                int32_t* f = p.faces + 3 * (int64_t)pT;
                f[0] = (int32_t)vid[2]; f[1] = (int32_t)vid[1]; f[2] = (int32_t)vid[0];
                ++pT;
            }
        }
        // ---- advance running bases by the chunk totals
        const uint32_t t0 = __shfl_sync(0xffffffffu, i0, 31), t1 = __shfl_sync(0xffffffffu, i1, 31);
        const uint32_t t2 = __shfl_sync(0xffffffffu, i2, 31), t3 = __shfl_sync(0xffffffffu, i3, 31);
        bX00 += t0 & 0xffffu; bX01 += t0 >> 16; bX10 += t1 & 0xffffu; bX11 += t1 >> 16;
        bY0 += t2 & 0xffffu; bY1 += t2 >> 16; bZ0 += t3 & 0xffffu; bZ1 += t3 >> 16;
        bT += __shfl_sync(0xffffffffu, it, 31);
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI: field sign / count / emit
// ------------------------------------------------------------------------------------------------
static OccView make_view(const void* occ_bits, int Z, int H, int W, int pad, int gaussian, const double* w3)
{
    OccView v;
    v.bits = (const uint32_t*)occ_bits;
    v.Z = Z; v.H = H; v.W = W; v.nw = t3d_wpr(W);
    v.pad = pad ? 1 : 0;
    v.Zp = Z + 2 * v.pad; v.Hp = H + 2 * v.pad; v.Wp = W + 2 * v.pad;
    v.gaussian = gaussian ? 1 : 0;
    // scipy.ndimage._filters._gaussian_kernel1d(0.5, 0, 2)
    v.w0 = w3 ? w3[0] : 0x1.92b965ef5aaeep-1;
    v.w1 = w3 ? w3[1] : 0x1.b405b9842b206p-4;
    v.w2 = w3 ? w3[2] : 0x1.14aebe6a24088p-12;
    return v;
}

// sign volume dims: (Z+2p, H+2p, words_per_row(W+2p)).  n_exact_u64 (device, zeroed here) receives the
// number of voxels that needed the exact float64 evaluation.
extern "C" int t3d_field_sign(const void* occ_bits, int Z, int H, int W, int pad, const double* weights3, void* sign_bits,
                              void* n_exact_u64, void* stream)
{
    if (Z <= 0 || H <= 0 || W <= 0) { t3d_set_error("t3d_field_sign: empty volume"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const OccView v = make_view(occ_bits, Z, H, W, pad, 1, weights3);
    const int nwp = t3d_wpr(v.Wp);
    const int64_t total = (int64_t)v.Zp * v.Hp * nwp;
    T3D_CUDA(cudaMemsetAsync(n_exact_u64, 0, 8, st));
    k_field_sign<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v, (uint32_t*)sign_bits, nwp, (unsigned long long*)n_exact_u64);
    T3D_CHECK_LAUNCH("t3d_field_sign");
    t3d_count_launches(1);
    return 0;
}

// rowcnt: 4 * Zs*Hs uint32 (x-edge, y-edge, z-edge vertex counts and triangle counts per voxel row)
extern "C" int t3d_mc_count(const void* sign_bits, int Zs, int Hs, int Ws, void* rowcnt_u32, void* n_ambiguous_u64,
                            void* stream)
{
    if (Zs <= 0 || Hs <= 0 || Ws <= 0) { t3d_set_error("t3d_mc_count: empty volume"); return 2; }
    if (ensure_luts()) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = (int64_t)Zs * Hs;
    T3D_CUDA(cudaMemsetAsync(n_ambiguous_u64, 0, 8, st));
    k_mc_count<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>((const uint32_t*)sign_bits, Zs, Hs, Ws, t3d_wpr(Ws),
                                                                    (uint32_t*)rowcnt_u32, rows,
                                                                    (unsigned long long*)n_ambiguous_u64);
    T3D_CHECK_LAUNCH("t3d_mc_count");
    t3d_count_launches(1);
    return 0;
}

// rowbase: the exclusive scans of rowcnt (4 arrays); n_x / n_y: totals of the x- and y-edge counts.
// occ_*: the occupancy the sign volume was derived from (pad/gaussian as in t3d_field_sign; gaussian = 0
// means the sign volume IS the occupancy and pad must be 0).
// cum/adj: device float64 arrays for the variable-slice-depth z map (n_cum = 0 disables it).
extern "C" int t3d_mc_emit(const void* sign_bits, const void* occ_bits, int Z, int H, int W, int pad, int gaussian,
                           const double* weights3, const void* rowbase_u32, uint32_t n_x, uint32_t n_y, int unpad_shift,
                           const void* cum_f64, const void* adj_f64, int n_cum, double mm_per_pixel_y, double mm_per_pixel_x,
                           int scale_in_f64, void* verts_f32, void* faces_i32, void* stream)
{
    if (Z <= 0 || H <= 0 || W <= 0) { t3d_set_error("t3d_mc_emit: empty volume"); return 2; }
    if (!gaussian && pad) { t3d_set_error("t3d_mc_emit: pad requires gaussian"); return 2; }
    if (ensure_luts()) return 1;
    EmitParams p;
    p.occ = make_view(occ_bits, Z, H, W, pad, gaussian, weights3);
    p.sign = (const uint32_t*)sign_bits;
    p.nwp = t3d_wpr(p.occ.Wp);
    p.n_rows = (int64_t)p.occ.Zp * p.occ.Hp;
    p.rowbase = (const uint32_t*)rowbase_u32;
    p.offY = n_x;
    p.offZ = n_x + n_y;
    p.shift = unpad_shift ? 1.0f : 0.0f;
    p.cum = (const double*)cum_f64;
    p.adj = (const double*)adj_f64;
    p.n_cum = n_cum;
    p.mm_y = mm_per_pixel_y;
    p.mm_x = mm_per_pixel_x;
    p.scale_f64 = scale_in_f64 ? 1 : 0;
    p.verts = (float*)verts_f32;
    p.faces = (int32_t*)faces_i32;
    k_mc_emit<<<(unsigned)((p.n_rows + EM_WARPS - 1) / EM_WARPS), EM_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
    T3D_CHECK_LAUNCH("t3d_mc_emit");
    t3d_count_launches(1);
    return 0;
}

// float32 field of the whole padded grid (test / debugging aid: lets the parity tests compare the field
// itself with scipy bit for bit)
__global__ void __launch_bounds__(256) k_field_dense(OccView v, float* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)v.Zp * v.Hp * v.Wp;
    if (i >= total) return;
    const int x = (int)(i % v.Wp);
    const int64_t r = i / v.Wp;
    out[i] = field_value(v, (int)(r / v.Hp), (int)(r % v.Hp), x);
}

extern "C" int t3d_field_dense(const void* occ_bits, int Z, int H, int W, int pad, int gaussian, const double* weights3,
                               void* out_f32, void* stream)
{
    if (Z <= 0 || H <= 0 || W <= 0) { t3d_set_error("t3d_field_dense: empty volume"); return 2; }
    const OccView v = make_view(occ_bits, Z, H, W, pad, gaussian, weights3);
    const int64_t total = (int64_t)v.Zp * v.Hp * v.Wp;
    k_field_dense<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(v, (float*)out_f32);
    T3D_CHECK_LAUNCH("t3d_field_dense");
    t3d_count_launches(1);
    return 0;
}

// cube-case volume (Zs-1, Hs-1, Ws-1) uint8 from a sign volume (parity tests: bit-exact cube cases)
__global__ void __launch_bounds__(256) k_cube_cases(const uint32_t* __restrict__ sign, int Zs, int Hs, int Ws, int nws,
                                                    uint8_t* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)(Zs - 1) * (Hs - 1) * (Ws - 1);
    if (i >= total) return;
    const int x = (int)(i % (Ws - 1));
    const int64_t r = i / (Ws - 1);
    const int y = (int)(r % (Hs - 1)), z = (int)(r / (Hs - 1));
    auto sb = [&](int zz, int yy, int xx) -> int {
        return (int)((sign[((int64_t)zz * Hs + yy) * nws + (xx >> 5)] >> (xx & 31)) & 1u);
    };
    out[i] = (uint8_t)(sb(z, y, x) | (sb(z, y, x + 1) << 1) | (sb(z, y + 1, x + 1) << 2) | (sb(z, y + 1, x) << 3) |
                       (sb(z + 1, y, x) << 4) | (sb(z + 1, y, x + 1) << 5) | (sb(z + 1, y + 1, x + 1) << 6) |
                       (sb(z + 1, y + 1, x) << 7));
}

extern "C" int t3d_cube_cases(const void* sign_bits, int Zs, int Hs, int Ws, void* out_u8, void* stream)
{
    if (Zs < 2 || Hs < 2 || Ws < 2) { t3d_set_error("t3d_cube_cases: volume smaller than 2x2x2"); return 2; }
    const int64_t total = (int64_t)(Zs - 1) * (Hs - 1) * (Ws - 1);
    k_cube_cases<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)sign_bits, Zs, Hs, Ws,
                                                                                  t3d_wpr(Ws), (uint8_t*)out_u8);
    T3D_CHECK_LAUNCH("t3d_cube_cases");
    t3d_count_launches(1);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// canonical mesh (_ensure_manifold_mesh): lexicographic sort of float32 rows (z, y, x), merge exact
// duplicates, remap faces, drop faces with a repeated index (order otherwise preserved).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_key(float f)
{
    uint32_t u = __float_as_uint(f);
    if (u == 0x80000000u) u = 0u;  // -0.0 == +0.0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256) k_make_keys(const float* __restrict__ verts, int64_t V, int col, const uint32_t* __restrict__ perm,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ iota)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const uint32_t src = perm ? perm[i] : (uint32_t)i;
    keys[i] = float_key(verts[3 * (int64_t)src + col]);
    if (iota) iota[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_unique_heads(const float* __restrict__ verts, const uint32_t* __restrict__ perm, int64_t V,
                                                      uint32_t* __restrict__ head)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    uint32_t h = 1;
    if (i > 0) {
        const float* a = verts + 3 * (int64_t)perm[i];
        const float* b = verts + 3 * (int64_t)perm[i - 1];
        h = (float_key(a[0]) != float_key(b[0]) || float_key(a[1]) != float_key(b[1]) || float_key(a[2]) != float_key(b[2])) ? 1u : 0u;
    }
    head[i] = h;
}

__global__ void __launch_bounds__(256) k_scatter_unique(const float* __restrict__ verts, const uint32_t* __restrict__ perm,
                                                        const uint32_t* __restrict__ head, const uint32_t* __restrict__ pos,
                                                        int64_t V, float* __restrict__ out_verts, uint32_t* __restrict__ newid)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    // pos = exclusive scan of head => unique index of element i is pos[i] + head[i] - 1
    const uint32_t u = pos[i] + head[i] - 1u;
    const uint32_t src = perm[i];
    newid[src] = u;
    if (head[i]) {
        const float* a = verts + 3 * (int64_t)src;
        float* o = out_verts + 3 * (int64_t)u;
        o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
    }
}

__global__ void __launch_bounds__(256) k_face_valid(const int32_t* __restrict__ faces, int64_t F, const uint32_t* __restrict__ newid,
                                                    uint32_t* __restrict__ valid)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F) return;
    const uint32_t a = newid[faces[3 * i]], b = newid[faces[3 * i + 1]], c = newid[faces[3 * i + 2]];
    valid[i] = (a != b && b != c && a != c) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_face_compact(const int32_t* __restrict__ faces, int64_t F, const uint32_t* __restrict__ newid,
                                                      const uint32_t* __restrict__ valid, const uint32_t* __restrict__ pos,
                                                      long long* __restrict__ out64, int32_t* __restrict__ out32)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F || !valid[i]) return;
    const uint32_t a = newid[faces[3 * i]], b = newid[faces[3 * i + 1]], c = newid[faces[3 * i + 2]];
    const int64_t o = 3 * (int64_t)pos[i];
    if (out64) { out64[o] = a; out64[o + 1] = b; out64[o + 2] = c; }
    if (out32) { out32[o] = (int32_t)a; out32[o + 1] = (int32_t)b; out32[o + 2] = (int32_t)c; }
}

static size_t sort_temp_bytes(int64_t V)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)V);
    return bytes;
}

static inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int64_t t3d_canonicalize_workspace_bytes(int64_t V, int64_t F)
{
    const int64_t n = V > F ? V : F;
    int64_t b = 0;
    b += 4 * align256(4 * V);                       // keys_a, keys_b, perm_a, perm_b
    b += 2 * align256(4 * n);                       // flags, positions (vertices then faces)
    b += align256(4 * V);                           // newid
    b += align256((int64_t)sort_temp_bytes(V > 0 ? V : 1));
    b += align256(t3d_scan_workspace_bytes(n, 1));
    b += 256;                                       // totals
    return b;
}

// verts_in (V,3) f32, faces_in (F,3) i32  ->  verts_out (<=V,3) f32 sorted+unique, faces_out (<=F,3) int64 and/or int32
// (either pointer may be null).  counts_u64[0] = V', counts_u64[1] = F' (device).
extern "C" int t3d_mesh_canonicalize(const void* verts_in, int64_t V, const void* faces_in, int64_t F, void* verts_out,
                                     void* faces_out_i64, void* faces_out_i32, void* counts_u64, void* workspace, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (V <= 0 || V > 0x7fffffff || F < 0) { t3d_set_error("t3d_mesh_canonicalize: bad sizes"); return 2; }
    const int64_t n = V > F ? V : F;
    char* ws = (char*)workspace;
    uint32_t* keys_a = (uint32_t*)ws; ws += align256(4 * V);
    uint32_t* keys_b = (uint32_t*)ws; ws += align256(4 * V);
    uint32_t* perm_a = (uint32_t*)ws; ws += align256(4 * V);
    uint32_t* perm_b = (uint32_t*)ws; ws += align256(4 * V);
    uint32_t* flags = (uint32_t*)ws; ws += align256(4 * n);
    uint32_t* pos = (uint32_t*)ws; ws += align256(4 * n);
    uint32_t* newid = (uint32_t*)ws; ws += align256(4 * V);
    size_t temp_bytes = sort_temp_bytes(V);
    void* temp = ws; ws += align256((int64_t)temp_bytes);
    void* scan_ws = ws; ws += align256(t3d_scan_workspace_bytes(n, 1));
    unsigned long long* totals = (unsigned long long*)ws;
    unsigned long long* counts = (unsigned long long*)counts_u64;
    const unsigned gv = (unsigned)((V + 255) / 256);
    const float* vin = (const float*)verts_in;

    // stable LSD over the three columns: x (least significant), y, z
    uint32_t *pin = nullptr, *pout = perm_a, *pspare = perm_b;
    for (int pass = 0; pass < 3; ++pass) {
        const int col = 2 - pass;
        // keys of the current order; first pass also creates the identity permutation (in pspare)
        k_make_keys<<<gv, 256, 0, st>>>(vin, V, col, pin, keys_a, pass == 0 ? pspare : nullptr);
        const uint32_t* vals_in = (pass == 0) ? pspare : pin;
        T3D_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const uint32_t*)keys_a, keys_b, vals_in, pout, (int)V, 0, 32, st));
        // rotate buffers: pout becomes the current permutation
        uint32_t* old_in = pin;
        pin = pout;
        pout = (pass == 0) ? pspare : old_in;
        if (pass == 0) pspare = nullptr;
    }
    const uint32_t* perm = pin;
    k_unique_heads<<<gv, 256, 0, st>>>(vin, perm, V, flags);
    if (t3d_exclusive_scan_u32(flags, pos, V, 1, 0, totals, scan_ws, stream)) return 1;
    T3D_CUDA(cudaMemcpyAsync(counts, totals, 8, cudaMemcpyDeviceToDevice, st));
    k_scatter_unique<<<gv, 256, 0, st>>>(vin, perm, flags, pos, V, (float*)verts_out, newid);
    if (F > 0) {
        const unsigned gf = (unsigned)((F + 255) / 256);
        k_face_valid<<<gf, 256, 0, st>>>((const int32_t*)faces_in, F, newid, flags);
        if (t3d_exclusive_scan_u32(flags, pos, F, 1, 0, totals, scan_ws, stream)) return 1;
        T3D_CUDA(cudaMemcpyAsync(counts + 1, totals, 8, cudaMemcpyDeviceToDevice, st));
        k_face_compact<<<gf, 256, 0, st>>>((const int32_t*)faces_in, F, newid, flags, pos, (long long*)faces_out_i64,
                                           (int32_t*)faces_out_i32);
    } else {
        T3D_CUDA(cudaMemsetAsync(counts + 1, 0, 8, st));
    }
    T3D_CHECK_LAUNCH("t3d_mesh_canonicalize");
    t3d_count_launches(F > 0 ? 7 : 5);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// mesh measures: signed volume sum_f v0.(v1 x v2)/6 and area sum_f 0.5|(v1-v0)x(v2-v0)|, float64 terms and
// float64 accumulation, fixed reduction order (warp-shuffle trees, then one block over the block partials)
// ------------------------------------------------------------------------------------------------
template <typename IdxT>
__global__ void __launch_bounds__(256) k_mesh_measure(const float* __restrict__ verts, const IdxT* __restrict__ faces, int64_t F,
                                                      double* __restrict__ partials)
{
    double vol = 0.0, area = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < F; i += (int64_t)gridDim.x * blockDim.x) {
        const float* a = verts + 3 * (int64_t)faces[3 * i];
        const float* b = verts + 3 * (int64_t)faces[3 * i + 1];
        const float* c = verts + 3 * (int64_t)faces[3 * i + 2];
        const double ax = a[0], ay = a[1], az = a[2], bx = b[0], by = b[1], bz = b[2], cx = c[0], cy = c[1], cz = c[2];
        const double nx = by * cz - bz * cy, ny = bz * cx - bx * cz, nz = bx * cy - by * cx;
        vol += (ax * nx + ay * ny + az * nz) / 6.0;
        const double ux = bx - ax, uy = by - ay, uz = bz - az, vx = cx - ax, vy = cy - ay, vz = cz - az;
        const double px = uy * vz - uz * vy, py = uz * vx - ux * vz, pz = ux * vy - uy * vx;
        area += 0.5 * sqrt(px * px + py * py + pz * pz);
    }
    vol = warp_sum(vol); area = warp_sum(area);
    __shared__ double sv[8], sa[8];
    if (lane_id() == 0) { sv[threadIdx.x >> 5] = vol; sa[threadIdx.x >> 5] = area; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0, a = 0;
        for (int k = 0; k < 8; ++k) { v += sv[k]; a += sa[k]; }
        partials[2 * blockIdx.x] = v; partials[2 * blockIdx.x + 1] = a;
    }
}

__global__ void __launch_bounds__(256) k_reduce_partials(const double* __restrict__ partials, int n, double* __restrict__ out)
{
    double v = 0, a = 0;
    for (int i = threadIdx.x; i < n; i += 256) { v += partials[2 * i]; a += partials[2 * i + 1]; }
    v = warp_sum(v); a = warp_sum(a);
    __shared__ double sv[8], sa[8];
    if (lane_id() == 0) { sv[threadIdx.x >> 5] = v; sa[threadIdx.x >> 5] = a; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tv = 0, ta = 0;
        for (int k = 0; k < 8; ++k) { tv += sv[k]; ta += sa[k]; }
        out[0] = tv; out[1] = ta;
    }
}

#define MM_BLOCKS (T3D_NUM_SMS * 4)

extern "C" int64_t t3d_mesh_measure_workspace_bytes(void) { return (int64_t)MM_BLOCKS * 2 * 8; }

// out_f64[0] = signed volume, out_f64[1] = area (device)
extern "C" int t3d_mesh_measure(const void* verts_f32, int64_t V, const void* faces, int64_t F, int faces_are_i64, void* out_f64,
                                void* workspace, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    (void)V;
    if (F <= 0) { T3D_CUDA(cudaMemsetAsync(out_f64, 0, 16, st)); return 0; }
    int blocks = (int)((F + 255) / 256);
    if (blocks > MM_BLOCKS) blocks = MM_BLOCKS;
    if (faces_are_i64)
        k_mesh_measure<long long><<<blocks, 256, 0, st>>>((const float*)verts_f32, (const long long*)faces, F, (double*)workspace);
    else
        k_mesh_measure<int32_t><<<blocks, 256, 0, st>>>((const float*)verts_f32, (const int32_t*)faces, F, (double*)workspace);
    k_reduce_partials<<<1, 256, 0, st>>>((const double*)workspace, blocks, (double*)out_f64);
    T3D_CHECK_LAUNCH("t3d_mesh_measure");
    t3d_count_launches(2);
    return 0;
}
