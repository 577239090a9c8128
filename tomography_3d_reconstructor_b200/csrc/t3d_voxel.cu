// t3d_voxel.cu -- VoxelProcessor / VolumeCalculator kernels (sm_100a).
//
// Reference semantics (file:line into the reference repository):
//   pack        : image_loader.py:108 (`img >= threshold`) + np.stack, voxel_processor.py:46
//   fill_holes  : scipy.ndimage.binary_fill_holes on slice 0 / Z-1, voxel_processor.py:60-70
//   gap_fill    : the z loop of _close_volume_ends, voxel_processor.py:72-75 (== 3-point z stencil)
//   morph       : skimage binary_opening / binary_closing, voxel_processor.py:87-91
//   stats       : np.sum per slice (volume_calculator.py:31-33) and np.where min/max (:62-79)
//   point cloud : np.where + subsample, voxel_processor.py:99-108
//
// All kernels are HBM/L2-bandwidth bound integer work on the bit-packed occupancy; no tensor cores.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "t3d_common.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

extern "C" void t3d_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* t3d_last_error(void) { return g_err; }
static long long g_launches = 0;
extern "C" void t3d_count_launches(int n) { g_launches += n; }
extern "C" int64_t t3d_launch_count(void) { return g_launches; }

struct ZeroRange { const char* p; size_t n; };
static thread_local ZeroRange g_zero[8];
static thread_local int g_nzero = 0;
void t3d_prezero_register(const void* p, size_t n) { if (g_nzero < 8) g_zero[g_nzero++] = {(const char*)p, n}; }
void t3d_prezero_clear(void) { g_nzero = 0; }
int t3d_zero_async(void* p, size_t n, cudaStream_t st)
{
    for (int k = 0; k < g_nzero; ++k)
        if ((const char*)p >= g_zero[k].p && (const char*)p + n <= g_zero[k].p + g_zero[k].n) return 0;   // zeroed up front
    T3D_CUDA(cudaMemsetAsync(p, 0, n, st));
    return 0;
}

int t3d_rows_per_thread(const char* env_name, int dflt)
{
    const char* e = getenv(env_name);
    if (!e) return dflt;
    const int v = atoi(e);
    return (v >= 1 && v <= 4096) ? v : dflt;
}
extern "C" int t3d_version(void) { return 100; }
extern "C" int64_t t3d_words_per_row(int W) { return t3d_wpr(W); }

// ------------------------------------------------------------------------------------------------
// pack: uint8 (Z,H,W) -> bits.  Fast path: W % 32 == 0 and 16-byte aligned input; each warp turns
// 1024 contiguous bytes into 32 contiguous words per iteration with fully coalesced 128-bit loads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_flat(const uint8_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                   int64_t n_words, uint32_t thr4)
{
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t l = lane_id();
    const int64_t n_chunks = n_words >> 5;  // chunks of 32 words = 1024 bytes
    for (int64_t c = warp; c < n_chunks; c += n_warps) {
        const uint8_t* p = src + (c << 10);
        const uint4 a = ld_stream_u4(p + 16 * l);
        const uint4 b = ld_stream_u4(p + 512 + 16 * l);
        const uint32_t ha = ge16(a, thr4), hb = ge16(b, thr4);
        const int s0 = (2 * l) & 31;
        const uint32_t a0 = __shfl_sync(0xffffffffu, ha, s0), a1 = __shfl_sync(0xffffffffu, ha, s0 + 1);
        const uint32_t b0 = __shfl_sync(0xffffffffu, hb, s0), b1 = __shfl_sync(0xffffffffu, hb, s0 + 1);
        dst[(c << 5) + l] = (l < 16) ? (a0 | (a1 << 16)) : (b0 | (b1 << 16));
    }
    // tail words (n_words % 32): one thread per word
    const int64_t tail0 = n_chunks << 5;
    const int64_t t = tail0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_words) {
        const uint4* p = reinterpret_cast<const uint4*>(src + (t << 5));
        dst[t] = ge16(p[0], thr4) | (ge16(p[1], thr4) << 16);
    }
}

// generic path: any W, any alignment; one thread per output word
__global__ void __launch_bounds__(256) k_pack_generic(const uint8_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                      int64_t n_rows, int W, int wpr, int thr)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * wpr) return;
    const int64_t row = i / wpr;
    const int w = (int)(i - row * wpr);
    const uint8_t* p = src + row * W + (w << 5);
    const int n = min(32, W - (w << 5));
    uint32_t v = 0;
    for (int k = 0; k < n; ++k) v |= (uint32_t)(p[k] >= thr) << k;
    dst[i] = v;
}

extern "C" int t3d_pack_masks(const void* masks_u8, int Z, int H, int W, int threshold, void* bits, void* stream)
{
    if (Z <= 0 || H <= 0 || W <= 0) { t3d_set_error("t3d_pack_masks: empty volume"); return 2; }
    if (threshold < 0) threshold = 0;
    if (threshold > 256) threshold = 256;
    cudaStream_t st = (cudaStream_t)stream;
    const int wpr = t3d_wpr(W);
    const int64_t rows = (int64_t)Z * H;
    if ((W & 127) == 0 && ((uintptr_t)masks_u8 & 15) == 0 && threshold >= 1 && threshold <= 255) {
        const int64_t n_words = rows * wpr;
        const uint32_t thr4 = 0x01010101u * (uint32_t)threshold;
        int64_t blocks = (n_words / 32 + 7) / 8;           // one warp-iteration per 32 words
        const int64_t cap = (int64_t)T3D_NUM_SMS * 16;     // persistent-ish: 16 CTAs of 8 warps per SM
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        k_pack_flat<<<(unsigned)blocks, 256, 0, st>>>((const uint8_t*)masks_u8, (uint32_t*)bits, n_words, thr4);
    } else {
        const int64_t n = rows * wpr;
        k_pack_generic<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const uint8_t*)masks_u8, (uint32_t*)bits, rows, W,
                                                                    wpr, threshold);
    }
    T3D_CHECK_LAUNCH("t3d_pack_masks");
    t3d_count_launches(1);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// unpack: bits -> uint8 0/1 (numpy bool) for the API-visible volumes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpack_flat(const uint32_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                     int64_t n_words)
{
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t l = lane_id();
    const int64_t n_chunks = (n_words + 31) >> 5;
    for (int64_t c = warp; c < n_chunks; c += n_warps) {
        const int64_t wi = (c << 5) + l;
        const uint32_t w = wi < n_words ? src[wi] : 0u;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            // lane writes 16 bytes at offset 16*l of this 512-byte half: bits of word (16*half + l/2), half-word l&1
            const uint32_t ww = __shfl_sync(0xffffffffu, w, 16 * half + (l >> 1));
            const uint32_t h = (ww >> (16 * (l & 1))) & 0xffffu;
            uint4 o;
            o.x = expand4(h & 15u); o.y = expand4((h >> 4) & 15u);
            o.z = expand4((h >> 8) & 15u); o.w = expand4((h >> 12) & 15u);
            const int64_t word_of_lane = (c << 5) + 16 * half + (l >> 1);
            if (word_of_lane < n_words) st_stream_u4(dst + (c << 10) + 512 * half + 16 * l, o);
        }
    }
}

__global__ void __launch_bounds__(256) k_unpack_generic(const uint32_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                        int64_t n_rows, int W, int wpr)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * wpr) return;
    const int64_t row = i / wpr;
    const int w = (int)(i - row * wpr);
    const uint32_t v = src[i];
    uint8_t* p = dst + row * W + (w << 5);
    const int n = min(32, W - (w << 5));
    for (int k = 0; k < n; ++k) p[k] = (v >> k) & 1u;
}

extern "C" int t3d_unpack_bits(const void* bits, int Z, int H, int W, void* out_u8, void* stream)
{
    if (Z <= 0 || H <= 0 || W <= 0) { t3d_set_error("t3d_unpack_bits: empty volume"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const int wpr = t3d_wpr(W);
    const int64_t rows = (int64_t)Z * H;
    const int64_t n = rows * wpr;
    if ((W & 127) == 0 && ((uintptr_t)out_u8 & 15) == 0) {
        int64_t blocks = ((n + 31) / 32 + 7) / 8;
        const int64_t cap = (int64_t)T3D_NUM_SMS * 16;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        k_unpack_flat<<<(unsigned)blocks, 256, 0, st>>>((const uint32_t*)bits, (uint8_t*)out_u8, n);
    } else {
        k_unpack_generic<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const uint32_t*)bits, (uint8_t*)out_u8, rows, W, wpr);
    }
    T3D_CHECK_LAUNCH("t3d_unpack_bits");
    t3d_count_launches(1);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// 2-D hole filling of one bit plane (in place): complement of the 4-connected flood fill of the
// background started from outside the image.  One CTA of 1024 threads per plane; every iteration
// closes reachability along whole rows (carry-lookahead over words) and along whole columns
// (segmented carry-lookahead over rows), so convex-ish objects converge in two iterations.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fill_up(uint32_t m, uint32_t s)  // flood s within m towards bit 31
{
    s |= m & (s << 1);  m &= m << 1;
    s |= m & (s << 2);  m &= m << 2;
    s |= m & (s << 4);  m &= m << 4;
    s |= m & (s << 8);  m &= m << 8;
    s |= m & (s << 16);
    return s;
}
__device__ __forceinline__ uint32_t fill_dn(uint32_t m, uint32_t s)  // towards bit 0
{
    s |= m & (s >> 1);  m &= m >> 1;
    s |= m & (s >> 2);  m &= m >> 2;
    s |= m & (s >> 4);  m &= m >> 4;
    s |= m & (s >> 8);  m &= m >> 8;
    s |= m & (s >> 16);
    return s;
}

// closure of `reach` along rows, both directions.  A warp owns 32 consecutive rows: it stages a 32-row x 16-word tile
// of the mask and of `reach` in shared memory with coalesced loads (row pitch 17 words: conflict-free when lane = row),
// every lane then runs the carry chain of ITS row over the 16 words, and the tile is written back coalesced.  The
// carry of each row crosses tiles in a register.
#define FH_TW 16
#define FH_PITCH 17

__device__ __forceinline__ void rows_closure_tile(const uint32_t* __restrict__ m, uint32_t* __restrict__ r, int nw, int nwv,
                                                  int y0, int H, uint32_t* tm, uint32_t* tr)
{
    const uint32_t l = lane_id();
    const int rsub = l >> 4, wsub = l & 15;
    const int n_tiles = (nwv + FH_TW - 1) / FH_TW;
    for (int dir = 0; dir < 2; ++dir) {
        uint32_t carry = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int c = dir == 0 ? t : n_tiles - 1 - t;
            const int w = c * FH_TW + wsub;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int rr = 2 * j + rsub, y = y0 + rr;
                const bool ok = (y < H) && (w < nwv);
                tm[rr * FH_PITCH + wsub] = ok ? m[(int64_t)y * nw + w] : 0u;
                tr[rr * FH_PITCH + wsub] = ok ? r[(int64_t)y * nw + w] : 0u;
            }
            __syncwarp();
            uint32_t* mr = tm + l * FH_PITCH;
            uint32_t* rr_ = tr + l * FH_PITCH;
            if (dir == 0) {
#pragma unroll
                for (int k = 0; k < FH_TW; ++k) {
                    const uint32_t mm = mr[k];
                    const uint32_t f = fill_up(mm, rr_[k] | (carry & mm & 1u));
                    carry = f >> 31;
                    rr_[k] = f;
                }
            } else {
#pragma unroll
                for (int k = FH_TW - 1; k >= 0; --k) {
                    const uint32_t mm = mr[k];
                    const uint32_t f = fill_dn(mm, rr_[k] | ((carry << 31) & mm));
                    carry = f & 1u;
                    rr_[k] = f;
                }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int rr = 2 * j + rsub, y = y0 + rr;
                if (y < H && w < nwv) r[(int64_t)y * nw + w] = tr[rr * FH_PITCH + wsub];
            }
            __syncwarp();
        }
    }
}

#define FH_THREADS 1024
#define FH_MAXSEG 64

__global__ void __launch_bounds__(FH_THREADS) k_fill_holes(uint32_t* bits_planes, int64_t plane_stride_words,
                                                           uint32_t* scratch, int H, int W, int nw)
{
    // plane handled by this CTA; scratch layout: [plane][0: mask m][1: reach r], each H*nw words
    uint32_t* bits = bits_planes + (int64_t)blockIdx.x * plane_stride_words;
    const int64_t pw = (int64_t)H * nw;
    uint32_t* m = scratch + (int64_t)blockIdx.x * 2 * pw;
    uint32_t* r = m + pw;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;

    // background mask + seeds (background pixels on the image border touch the "outside")
    for (int64_t i = tid; i < pw; i += FH_THREADS) {
        const int y = (int)(i / nw), w = (int)(i - (int64_t)y * nw);
        const uint32_t vm = valid_mask(w, W);
        const uint32_t bg = ~bits[i] & vm;
        uint32_t seed = 0;
        if (y == 0 || y == H - 1) seed = bg;
        if (w == 0) seed |= bg & 1u;
        if (w == ((W - 1) >> 5)) seed |= bg & (1u << ((W - 1) & 31));
        m[i] = bg;
        r[i] = seed;
    }
    __syncthreads();

    extern __shared__ uint32_t tile[];  // per warp: mask tile + reach tile, 32 rows x FH_PITCH words each
    // column segments: thread (seg, col) sweeps rows [seg*L, seg*L+L)
    __shared__ uint32_t sG[FH_MAXSEG][32];  // used in column tiles of 32 word-columns
    __shared__ uint32_t sP[FH_MAXSEG][32];
    const int nseg = min(FH_MAXSEG, min(FH_THREADS / 32, H));  // 32 segments
    const int L = (H + nseg - 1) / nseg;

    for (int iter = 0; iter < 4 * (H + W) + 8; ++iter) {
        // A round = row closure then column closure.  The row closure is idempotent, so the fixpoint is reached as
        // soon as a column closure changes nothing: only the column phase feeds `changed`.
        int changed = 0;
        // ---- rows
        for (int y0 = warp * 32; y0 < H; y0 += (FH_THREADS / 32) * 32)
            rows_closure_tile(m, r, nw, (W + 31) >> 5, y0, H, tile + warp * 2 * 32 * FH_PITCH,
                              tile + warp * 2 * 32 * FH_PITCH + 32 * FH_PITCH);
        __syncthreads();
        // ---- columns, 32 word-columns at a time: thread = (seg = warp, col = lane)
        for (int c0 = 0; c0 < nw; c0 += 32) {
            const int c = c0 + (tid & 31), seg = warp;
            const bool act = (c < nw) && (seg < nseg);
            const int y0 = seg * L, y1 = min(H, y0 + L);
            for (int dir = 0; dir < 2; ++dir) {
                // pass 1: segment summary with zero carry-in
                uint32_t G = 0, P = 0xffffffffu;
                if (act && y0 < y1) {
                    if (dir == 0) for (int y = y0; y < y1; ++y) { const uint32_t mm = m[(int64_t)y * nw + c]; G = r[(int64_t)y * nw + c] | (G & mm); P &= mm; }
                    else          for (int y = y1 - 1; y >= y0; --y) { const uint32_t mm = m[(int64_t)y * nw + c]; G = r[(int64_t)y * nw + c] | (G & mm); P &= mm; }
                } else { P = 0xffffffffu; G = 0; }
                if (seg < FH_MAXSEG) { sG[seg][tid & 31] = G; sP[seg][tid & 31] = (act && y0 < y1) ? P : 0xffffffffu; }
                __syncthreads();
                // pass 2: carry into this segment
                uint32_t cin = 0;
                if (act) {
                    if (dir == 0) for (int s = 0; s < seg; ++s) cin = sG[s][tid & 31] | (sP[s][tid & 31] & cin);
                    else          for (int s = nseg - 1; s > seg; --s) cin = sG[s][tid & 31] | (sP[s][tid & 31] & cin);
                }
                // pass 3: final sweep (rows fetched eight at a time ahead of the dependent chain)
                if (act && y0 < y1) {
                    uint32_t carry = cin;
                    const int n = y1 - y0;
                    for (int k0 = 0; k0 < n; k0 += 8) {
                        uint32_t mm[8], rr[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int kk = k0 + k;
                            const int y = dir == 0 ? y0 + kk : y1 - 1 - kk;
                            const bool ok = kk < n;
                            mm[k] = ok ? m[(int64_t)y * nw + c] : 0u;
                            rr[k] = ok ? r[(int64_t)y * nw + c] : 0u;
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int kk = k0 + k;
                            if (kk < n) {
                                const int y = dir == 0 ? y0 + kk : y1 - 1 - kk;
                                const uint32_t nv = rr[k] | (carry & mm[k]);
                                if (nv != rr[k]) { r[(int64_t)y * nw + c] = nv; changed = 1; }
                                carry = nv;
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    // holes = background never reached
    for (int64_t i = tid; i < pw; i += FH_THREADS) bits[i] |= m[i] & ~r[i];
}

extern "C" int64_t t3d_fill_holes_scratch_bytes(int n_planes, int H, int W)
{
    return (int64_t)n_planes * 2 * H * t3d_wpr(W) * 4;
}

extern "C" int t3d_fill_holes_2d(void* bits, int n_planes, int64_t plane_stride_words, int H, int W, void* scratch,
                                 void* stream)
{
    if (n_planes <= 0) return 0;
    if (H <= 0 || W <= 0) { t3d_set_error("t3d_fill_holes_2d: empty plane"); return 2; }
    const size_t smem = (size_t)(FH_THREADS / 32) * 2 * 32 * FH_PITCH * sizeof(uint32_t);  // 136 KB
    static bool attr_set = false;
    if (!attr_set) {
        T3D_CUDA(cudaFuncSetAttribute(k_fill_holes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    k_fill_holes<<<n_planes, FH_THREADS, smem, (cudaStream_t)stream>>>((uint32_t*)bits, plane_stride_words,
                                                                      (uint32_t*)scratch, H, W, t3d_wpr(W));
    T3D_CHECK_LAUNCH("t3d_fill_holes_2d");
    t3d_count_launches(1);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// z gap fill: out[z] = f[z] | (f[z-1] & f[z+1]) for 1 <= z <= Z-2, copy for z = 0, Z-1
// (voxel_processor.py:72-75; the np.any guards are redundant and the loop is not a recurrence,
// SURVEY.md V3).  `lo` / `hi` are optional neighbour planes for z-slab sharding: when given, local
// plane 0 / Z-1 is an interior plane of the global stack and uses them as f[-1] / f[Z].
// Optionally accumulates per-slice popcounts of the result.  128-bit accesses, two per thread in flight.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gap_fill(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                  const uint4* __restrict__ lo, const uint4* __restrict__ hi, int Z,
                                                  int64_t pw4, unsigned long long* __restrict__ counts)
{
    const int z = blockIdx.y;
    const uint4* c = in + (int64_t)z * pw4;
    const uint4* a = (z > 0) ? c - pw4 : lo;
    const uint4* b = (z < Z - 1) ? c + pw4 : hi;
    uint4* o = out + (int64_t)z * pw4;
    uint32_t cnt = 0;
    const bool fill = (a != nullptr) && (b != nullptr);
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < pw4; i += 2 * stride) {
        const int64_t j = i + stride;
        const bool two = j < pw4;
        uint4 v0 = c[i], v1 = two ? c[j] : make_uint4(0, 0, 0, 0);
        if (fill) {
            const uint4 a0 = a[i], b0 = b[i];
            v0 = or4(v0, and4(a0, b0));
            if (two) { const uint4 a1 = a[j], b1 = b[j]; v1 = or4(v1, and4(a1, b1)); }
        }
        o[i] = v0;
        if (two) o[j] = v1;
        cnt += popc4(v0) + popc4(v1);
    }
    if (counts) {
        cnt = warp_sum(cnt);
        __shared__ uint32_t s[8];
        if (lane_id() == 0) s[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int k = 0; k < 8; ++k) t += s[k];
            if (t) atomicAdd(counts + z, t);
        }
    }
}

extern "C" int t3d_gap_fill(const void* in_bits, void* out_bits, const void* lo_plane, const void* hi_plane, int Z, int H,
                            int W, void* slice_counts_u64, void* stream)
{
    if (Z <= 0 || H <= 0 || W <= 0) { t3d_set_error("t3d_gap_fill: empty volume"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t pw4 = (int64_t)H * (t3d_wpr(W) / 4);
    if (slice_counts_u64 && t3d_zero_async(slice_counts_u64, sizeof(unsigned long long) * Z, st)) return 1;
    int bx = (int)min((int64_t)32, (pw4 + 511) / 512);
    dim3 grid(bx, Z);
    k_gap_fill<<<grid, 256, 0, st>>>((const uint4*)in_bits, (uint4*)out_bits, (const uint4*)lo_plane, (const uint4*)hi_plane, Z,
                                     pw4, (unsigned long long*)slice_counts_u64);
    T3D_CHECK_LAUNCH("t3d_gap_fill");
    t3d_count_launches(1);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// 6-connected binary morphology.  Stage s of a call is an erosion (bit s of erode_mask set; out-of-volume = 1)
// or a dilation (out-of-volume = 0): exactly skimage's binary_erosion / binary_dilation with the default cross
// footprint.  opening then closing = stages E,D,D,E = erode_mask 0b1001.
//
// One launch per stage.  The packed volume is 1/8 byte per voxel (67 MB at 512x1024x1024: L2 resident on a B200),
// so a stage is bound by instruction issue and load latency, not HBM.  A thread owns one uint4 column (128 voxels)
// of one plane and marches down MY rows with the y neighbours in a register window: per 128 output voxels it issues
// 3 x 128-bit loads (next row, z-1, z+1) + 2 scalar loads (x neighbours, L1 hits), 8 funnel shifts and 12 LOP3.
// ------------------------------------------------------------------------------------------------
#define MY 8

template <bool ER, bool FIX>
__global__ void __launch_bounds__(256, 5) k_morph4(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int Z, int H,
                                                int W, int nw, int lanes_x, int pz_per_block, int my,
                                                unsigned long long* __restrict__ counts)
{
    constexpr uint32_t B = ER ? 0xffffffffu : 0u;
    const int lx = threadIdx.x % lanes_x, pz = threadIdx.x / lanes_x;
    const int nw4 = nw >> 2, nwv = (W + 31) >> 5;
    const int w4 = blockIdx.x * lanes_x + lx, z = blockIdx.z * pz_per_block + pz, y0 = blockIdx.y * my;
    extern __shared__ unsigned int s_cnt[];
    if (counts) {
        if ((int)threadIdx.x < pz_per_block) s_cnt[threadIdx.x] = 0;
        __syncthreads();
    }
    uint32_t cnt = 0;
    if (pz < pz_per_block && w4 < nw4 && z < Z) {
        const uint4 vm = valid_mask4(w4, W);
        // words beyond the volume (tail bits, padding words) read as the border value of this stage
        const bool fix = FIX && ER && (4 * w4 + 4 > (W >> 5));
        auto row4 = [&](int zz, int yy) -> uint4 {
            uint4 v = *reinterpret_cast<const uint4*>(in + ((int64_t)zz * H + yy) * nw + 4 * w4);
            if (fix) v = make_uint4(v.x | ~vm.x, v.y | ~vm.y, v.z | ~vm.z, v.w | ~vm.w);
            return v;
        };
        const bool has_l = (w4 > 0), has_r = (4 * w4 + 4 < nwv);
        const uint4 B4 = splat4(B);
        uint4 prev = (y0 > 0) ? row4(z, y0 - 1) : B4;
        uint4 cur = row4(z, y0);
        const int y1 = min(H, y0 + my);
#pragma unroll 4
        for (int y = y0; y < y1; ++y) {
            const uint4 next = (y + 1 < H) ? row4(z, y + 1) : B4;
            const uint4 zm = (z > 0) ? row4(z - 1, y) : B4;
            const uint4 zp = (z + 1 < Z) ? row4(z + 1, y) : B4;
            const uint32_t* rowp = in + ((int64_t)z * H + y) * nw + 4 * w4;
            const uint32_t l = has_l ? rowp[-1] : B;
            uint32_t r = has_r ? rowp[4] : B;
            if (FIX && ER && has_r && 4 * w4 + 5 >= nwv) r |= ~valid_mask(4 * w4 + 4, W);  // right neighbour is the partial last word
            const uint4 xm = shl1_4(cur, l), xp = shr1_4(cur, r);
            uint4 v;
            if (ER) v = and4(and4(and4(cur, xm), and4(xp, prev)), and4(and4(next, zm), zp));
            else v = or4(or4(or4(cur, xm), or4(xp, prev)), or4(or4(next, zm), zp));
            v = and4(v, vm);
            *reinterpret_cast<uint4*>(out + ((int64_t)z * H + y) * nw + 4 * w4) = v;
            cnt += popc4(v);
            prev = cur;
            cur = next;
        }
        if (counts && cnt) atomicAdd(&s_cnt[pz], cnt);
    }
    if (counts) {
        __syncthreads();
        const int zz = blockIdx.z * pz_per_block + threadIdx.x;
        if ((int)threadIdx.x < pz_per_block && zz < Z && s_cnt[threadIdx.x])
            atomicAdd(counts + zz, (unsigned long long)s_cnt[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------------
// The same stage with the input staged through shared memory.  A CTA owns a tile of MT_Z planes x MT_Y rows x 32 words
// (1024 voxels) and first copies the tile plus a one-cell halo into shared memory with coalesced 128-bit loads -- cells
// outside the volume (planes, rows, tail bits, padding words) get the border value of the stage there, so the stencil
// itself is branch-free.  Every input word is then read from L2 (6*34)/(4*32) = 1.6 times instead of 3 + the x-neighbour
// words.  EXPERIMENT (T3D_MORPH_TILE=1): despite the lower L2 traffic it is slower than k_morph4 (43 vs 33 us per stage at
// 512x1024x1024) -- the load / barrier / compute phases overlap worse than the register march's independent loads.
// ------------------------------------------------------------------------------------------------
#define MT_Z 4
#define MT_Y 32
#define MT_Q 8    // uint4 per tile row

template <bool ER>
__global__ void __launch_bounds__(256) k_morph_tile(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int Z, int H, int W,
                                                    int nw, unsigned long long* __restrict__ counts)
{
    constexpr uint32_t B = ER ? 0xffffffffu : 0u;
    __shared__ uint4 s_in[MT_Z + 2][MT_Y + 2][MT_Q];
    __shared__ uint32_t s_l[MT_Z + 2][MT_Y + 2], s_r[MT_Z + 2][MT_Y + 2];
    __shared__ unsigned int s_cnt[MT_Z];
    const int nw4 = nw >> 2, nwv = (W + 31) >> 5;
    const int q0 = blockIdx.x * MT_Q, y0 = blockIdx.y * MT_Y, z0 = blockIdx.z * MT_Z;
    const int tid = threadIdx.x;
    if (tid < MT_Z) s_cnt[tid] = 0;
    // ---- stage the tile + halo
    constexpr int ROWS = (MT_Z + 2) * (MT_Y + 2);
    for (int i = tid; i < ROWS * MT_Q; i += 256) {
        const int q = i % MT_Q, r = i / MT_Q, py = r % (MT_Y + 2), pz = r / (MT_Y + 2);
        const int z = z0 - 1 + pz, y = y0 - 1 + py, w4 = q0 + q;
        uint4 v = make_uint4(B, B, B, B);
        if (z >= 0 && z < Z && y >= 0 && y < H && w4 < nw4) {
            v = *reinterpret_cast<const uint4*>(in + ((int64_t)z * H + y) * nw + 4 * w4);
            if (ER) {   // tail bits and padding words read as 1
                const uint4 vm = valid_mask4(w4, W);
                v = make_uint4(v.x | ~vm.x, v.y | ~vm.y, v.z | ~vm.z, v.w | ~vm.w);
            }
        }
        s_in[pz][py][q] = v;
    }
    for (int r = tid; r < 2 * ROWS; r += 256) {
        const int side = r / ROWS, rr = r - side * ROWS, py = rr % (MT_Y + 2), pz = rr / (MT_Y + 2);
        const int z = z0 - 1 + pz, y = y0 - 1 + py;
        const int w = side ? 4 * (q0 + MT_Q) : 4 * q0 - 1;     // word right of the tile row / left of it
        uint32_t v = B;
        if (z >= 0 && z < Z && y >= 0 && y < H && w >= 0 && w < nwv) {
            v = in[((int64_t)z * H + y) * nw + w];
            if (ER) v |= ~valid_mask(w, W);
        }
        if (side) s_r[pz][py] = v; else s_l[pz][py] = v;
    }
    __syncthreads();
    // ---- stencil: thread = (row py, uint4 q), marching over the MT_Z planes with the z neighbours in registers
    const int q = tid % MT_Q, py = tid / MT_Q;     // 256 threads = 32 rows x 8 uint4
    const int y = y0 + py, w4 = q0 + q;
    const bool live = (y < H) && (w4 < nw4);
    const uint4 vm = valid_mask4(w4, W);
    uint4 zm = s_in[0][py + 1][q], cur = s_in[1][py + 1][q];
#pragma unroll
    for (int pz = 0; pz < MT_Z; ++pz) {
        const uint4 zp = s_in[pz + 2][py + 1][q];
        const uint4 prev = s_in[pz + 1][py][q], next = s_in[pz + 1][py + 2][q];
        const uint32_t l = q > 0 ? s_in[pz + 1][py + 1][q - 1].w : s_l[pz + 1][py + 1];
        const uint32_t r = q < MT_Q - 1 ? s_in[pz + 1][py + 1][q + 1].x : s_r[pz + 1][py + 1];
        const uint4 xm = shl1_4(cur, l), xp = shr1_4(cur, r);
        uint4 v;
        if (ER) v = and4(and4(and4(cur, xm), and4(xp, prev)), and4(and4(next, zm), zp));
        else v = or4(or4(or4(cur, xm), or4(xp, prev)), or4(or4(next, zm), zp));
        v = and4(v, vm);
        const int z = z0 + pz;
        uint32_t c = 0;
        if (live && z < Z) {
            *reinterpret_cast<uint4*>(out + ((int64_t)z * H + y) * nw + 4 * w4) = v;
            c = popc4(v);
        }
        if (counts) {
            c = warp_sum(c);
            if ((tid & 31) == 0 && c) atomicAdd(&s_cnt[pz], c);
        }
        zm = cur;
        cur = zp;
    }
    if (counts) {
        __syncthreads();
        if (tid < MT_Z && z0 + tid < Z && s_cnt[tid]) atomicAdd(counts + z0 + tid, (unsigned long long)s_cnt[tid]);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused multi-stage morphology (up to 4 stages, e.g. opening then closing = E,D,D,E) in ONE pass over the volume.
// A CTA owns a band of full-width rows and marches through a chunk of planes.  Level 0 = input, level k = output of
// stage k-1; levels 0..NST-1 live in shared memory as rings of 4 planes, the last level goes to global memory.  Level k
// lags level k-1 by two planes, so everything a step reads was written in an earlier step: one __syncthreads per
// plane.  Cells outside the volume (rows, planes, tail bits) hold the border value of the stage that will READ them
// (1 for an erosion, 0 for a dilation), which makes the inner loop branch-free.
// The halo is nst rows / planes on each side (recomputed by neighbouring CTAs); band and chunk sizes are chosen so
// that the grid is about one wave of the 148 SMs.
// ------------------------------------------------------------------------------------------------
struct FusedMorph {
    const uint32_t* in;
    uint32_t* out;
    int Z, H, W, nw, nw4;
    int nst;
    uint32_t erode_mask;
    int BY, BZ, RB;            // output rows per band, output planes per chunk, rows held in shared memory
    unsigned long long* counts;
};

__global__ void __launch_bounds__(1024, 1) k_morph_fused(FusedMorph p)
{
    extern __shared__ uint4 ring[];  // [level][slot 0..3][row 0..RB)[w4]
    __shared__ unsigned int s_cnt[2];
    const int R = p.nst, RB = p.RB, nw4 = p.nw4;
    const int yb0 = blockIdx.y * p.BY, yb1 = min(p.H, yb0 + p.BY);
    const int zc0 = blockIdx.z * p.BZ, zc1 = min(p.Z, zc0 + p.BZ);
    const int plane4 = RB * nw4;                  // uint4 per ring plane
    const int nwv = (p.W + 31) >> 5;
    const int T = (zc1 - zc0) + R + 2 * p.nst;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    auto border = [&](int stage) -> uint32_t { return ((p.erode_mask >> stage) & 1u) ? 0xffffffffu : 0u; };
    // items of one step: level 0 loads RB rows; level k (1..nst) computes rows [k, RB-k)
    int first[6];
    first[0] = 0;
    first[1] = RB * nw4;
    for (int k = 1; k <= p.nst; ++k) first[k + 1] = first[k] + (RB - 2 * k) * nw4;
    const int n_items = first[p.nst + 1];
    if (tid < 2) s_cnt[tid] = 0;
    __syncthreads();

    // input staging: every thread owns up to two uint4 of a plane; the global loads for plane zi+1 are issued at the
    // start of step t and land in shared memory at the start of step t+1, so their latency hides behind the compute
    const int n_in = first[1];
    const uint32_t B0 = border(0);
    auto load_in = [&](int zi, int j) -> uint4 {
        const int r = j / nw4, w4 = j - r * nw4;
        const int y = yb0 - R + r;
        uint4 v = splat4(B0);
        if (zi >= 0 && zi < p.Z && y >= 0 && y < p.H && zi < zc1 + R) {
            v = *reinterpret_cast<const uint4*>(p.in + ((int64_t)zi * p.H + y) * p.nw + 4 * w4);
            if (B0 && 4 * w4 + 4 > (p.W >> 5)) { const uint4 vm = valid_mask4(w4, p.W); v = make_uint4(v.x | ~vm.x, v.y | ~vm.y, v.z | ~vm.z, v.w | ~vm.w); }
        }
        return v;
    };
    uint4 pre0 = splat4(B0), pre1 = splat4(B0);
    if (tid < n_in) pre0 = load_in(zc0 - R, tid);
    if (tid + nthreads < n_in) pre1 = load_in(zc0 - R, tid + nthreads);

    for (int t = 0; t < T; ++t) {
        const int zi = zc0 - R + t;   // input plane of this step
        // ---- level 0: plane zi (prefetched) -> ring, then prefetch plane zi + 1
        {
            uint4* dst = ring + ((zi + 64) & 3) * plane4;
            if (tid < n_in) dst[tid] = pre0;
            if (tid + nthreads < n_in) dst[tid + nthreads] = pre1;
            if (tid < n_in) pre0 = load_in(zi + 1, tid);
            if (tid + nthreads < n_in) pre1 = load_in(zi + 1, tid + nthreads);
        }
        uint32_t pc = 0;
        for (int i = n_in + tid; i < n_items; i += nthreads) {
            int k = 1;
            while (i >= first[k + 1]) ++k;
            const int j = i - first[k];
            {
                // ---- level k = stage k-1 applied to level k-1, plane z = zi - 2k
                const int z = zi - 2 * k;
                const int halo = R - k;
                if (z < zc0 - halo || z >= zc1 + halo) continue;
                const int rr = j / nw4, w4 = j - rr * nw4;
                const int r = k + rr;
                const int y = yb0 - R + r;
                const bool er = (p.erode_mask >> (k - 1)) & 1u;
                const uint32_t B = er ? 0xffffffffu : 0u;
                const uint4* L = ring + (k - 1) * 4 * plane4;
                const uint4* pc0 = L + ((z + 64) & 3) * plane4 + r * nw4 + w4;
                const uint4 c = pc0[0], ym = pc0[-nw4], yp = pc0[nw4];
                const uint4 zm = L[((z - 1 + 64) & 3) * plane4 + r * nw4 + w4], zq = L[((z + 1 + 64) & 3) * plane4 + r * nw4 + w4];
                const uint32_t l = (w4 > 0) ? pc0[-1].w : B;
                const uint32_t rgt = (4 * w4 + 4 < nwv) ? pc0[1].x : B;
                const uint4 xm = shl1_4(c, l), xp = shr1_4(c, rgt);
                uint4 v = er ? and4(and4(and4(c, xm), and4(xp, ym)), and4(and4(yp, zm), zq))
                             : or4(or4(or4(c, xm), or4(xp, ym)), or4(or4(yp, zm), zq));
                const uint4 vm = valid_mask4(w4, p.W);
                const bool inside = (z >= 0 && z < p.Z && y >= 0 && y < p.H);
                if (k < p.nst) {
                    const uint32_t Bn = border(k);   // what the next stage must see outside the volume / beyond W
                    if (!inside) v = splat4(Bn);
                    else v = make_uint4((v.x & vm.x) | (Bn & ~vm.x), (v.y & vm.y) | (Bn & ~vm.y), (v.z & vm.z) | (Bn & ~vm.z), (v.w & vm.w) | (Bn & ~vm.w));
                    ring[(k * 4 + ((z + 64) & 3)) * plane4 + r * nw4 + w4] = v;
                } else if (inside && y >= yb0 && y < yb1) {
                    v = and4(v, vm);
                    *reinterpret_cast<uint4*>(p.out + ((int64_t)z * p.H + y) * p.nw + 4 * w4) = v;
                    pc += popc4(v);
                }
            }
        }
        if (p.counts) {
            pc = warp_sum(pc);
            if ((tid & 31) == 0 && pc) atomicAdd(&s_cnt[t & 1], pc);
        }
        __syncthreads();
        if (p.counts && tid == 0) {
            const int zf = zi - 2 * p.nst;
            const unsigned int c = s_cnt[t & 1];
            if (c && zf >= zc0 && zf < zc1) atomicAdd(p.counts + zf, (unsigned long long)c);
            s_cnt[t & 1] = 0;   // reused at step t+2, after the barrier of step t+1
        }
    }
}

// returns 1 if the fused kernel was launched, 0 if the shape does not fit (caller falls back to one launch per stage)
static int launch_morph_fused(const uint32_t* in, uint32_t* out, int Z, int H, int W, int n_stages, unsigned erode_mask,
                              unsigned long long* counts, cudaStream_t st, int* rc)
{
    *rc = 0;
    if (n_stages < 2 || n_stages > 4) return 0;
    // Correct but shared-memory-bandwidth bound (5 LDS.128 per 128 voxels and stage): 243 us vs 4 x 38 us for the
    // per-stage kernels at 512x1024x1024, so it is opt-in until the z neighbours are kept in registers.
    if (!getenv("T3D_FUSED_MORPH")) return 0;
    const int nw = t3d_wpr(W), nw4 = nw / 4;
    const int smem_budget = 200 * 1024;
    const int row_bytes = nw * 4;
    int RB = smem_budget / (n_stages * 4 * row_bytes);
    if (RB > 64) RB = 64;
    if (RB * nw4 > 2048) RB = 2048 / nw4;   // the input plane is staged by at most two uint4 per thread
    int BY = RB - 2 * n_stages;
    if (BY < 8) return 0;
    if (BY > H) { BY = H; RB = BY + 2 * n_stages; }
    const int n_bands = (H + BY - 1) / BY;
    // even out the bands, then choose the chunk count so that bands * chunks is about one wave
    BY = (H + n_bands - 1) / n_bands;
    RB = BY + 2 * n_stages;
    int n_chunks = T3D_NUM_SMS / n_bands;
    if (n_chunks < 1) n_chunks = 1;
    int BZ = (Z + n_chunks - 1) / n_chunks;
    if (BZ < 4 * n_stages) BZ = Z < 4 * n_stages ? Z : 4 * n_stages;   // keep the redundant halo work bounded
    n_chunks = (Z + BZ - 1) / BZ;
    FusedMorph p;
    p.in = in; p.out = out; p.Z = Z; p.H = H; p.W = W; p.nw = nw; p.nw4 = nw4; p.nst = n_stages; p.erode_mask = erode_mask;
    p.BY = BY; p.BZ = BZ; p.RB = RB; p.counts = counts;
    const size_t smem = (size_t)n_stages * 4 * RB * row_bytes;
    static size_t attr = 0;
    if (smem > attr) {
        if (cudaFuncSetAttribute(k_morph_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        attr = smem;
    }
    dim3 grid(1, n_bands, n_chunks);
    k_morph_fused<<<grid, 1024, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { t3d_set_error("t3d_morph (fused): launch failed: %s", cudaGetErrorString(e)); *rc = 1; }
    return 1;
}

extern "C" int64_t t3d_morph_scratch_bytes(int Z, int H, int W, int n_stages)
{
    return n_stages > 1 ? (int64_t)Z * H * t3d_wpr(W) * 4 * (n_stages > 2 ? 2 : 1) : 0;
}

// scratch: t3d_morph_scratch_bytes (intermediate stages ping-pong there); in/out must not alias.
extern "C" int t3d_morph(const void* in_bits, void* out_bits, int Z, int H, int W, int n_stages, unsigned erode_mask,
                         void* slice_counts_u64, void* scratch, void* stream)
{
    if (Z <= 0 || H <= 0 || W <= 0) { t3d_set_error("t3d_morph: empty volume"); return 2; }
    if (n_stages < 1 || n_stages > 32) { t3d_set_error("t3d_morph: n_stages must be 1..32"); return 2; }
    if (in_bits == out_bits) { t3d_set_error("t3d_morph: in-place is not supported"); return 2; }
    if (n_stages > 1 && !scratch) { t3d_set_error("t3d_morph: scratch required for more than one stage"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nw = t3d_wpr(W), nw4 = nw / 4;
    const int64_t vol_words = (int64_t)Z * H * nw;
    if (slice_counts_u64 && t3d_zero_async(slice_counts_u64, sizeof(unsigned long long) * Z, st)) return 1;
    {   // all stages in one pass when they fit in shared memory
        int rc = 0;
        if (launch_morph_fused((const uint32_t*)in_bits, (uint32_t*)out_bits, Z, H, W, n_stages, erode_mask,
                               (unsigned long long*)slice_counts_u64, st, &rc)) {
            if (rc) return rc;
            t3d_count_launches(1);
            return 0;
        }
    }
    static const bool use_tile = getenv("T3D_MORPH_TILE") != nullptr;   // measured slower than k_morph4 (43 vs 33 us per C1 stage)
    if (use_tile) {
        dim3 tgrid((nw4 + MT_Q - 1) / MT_Q, (H + MT_Y - 1) / MT_Y, (Z + MT_Z - 1) / MT_Z);
        uint32_t* tmp2[2] = {(uint32_t*)scratch, (uint32_t*)scratch + vol_words};
        const uint32_t* src2 = (const uint32_t*)in_bits;
        for (int s = 0; s < n_stages; ++s) {
            const bool last = (s == n_stages - 1);
            uint32_t* dst = last ? (uint32_t*)out_bits : tmp2[s & 1];
            unsigned long long* cnt = last ? (unsigned long long*)slice_counts_u64 : nullptr;
            if ((erode_mask >> s) & 1u) k_morph_tile<true><<<tgrid, 256, 0, st>>>(src2, dst, Z, H, W, nw, cnt);
            else k_morph_tile<false><<<tgrid, 256, 0, st>>>(src2, dst, Z, H, W, nw, cnt);
            src2 = dst;
        }
        T3D_CHECK_LAUNCH("t3d_morph");
        t3d_count_launches(n_stages);
        return 0;
    }
    const int lanes_x = nw4 < 256 ? nw4 : 256;
    const int pzb = 256 / lanes_x;
    static const int my = t3d_rows_per_thread("T3D_MORPH_ROWS", MY);
    dim3 grid((nw4 + lanes_x - 1) / lanes_x, (H + my - 1) / my, (Z + pzb - 1) / pzb);
    const size_t smem = sizeof(unsigned int) * pzb;
    uint32_t* tmp[2] = {(uint32_t*)scratch, (uint32_t*)scratch + vol_words};
    const uint32_t* src = (const uint32_t*)in_bits;
    for (int s = 0; s < n_stages; ++s) {
        const bool last = (s == n_stages - 1);
        uint32_t* dst = last ? (uint32_t*)out_bits : tmp[s & 1];
        unsigned long long* cnt = last ? (unsigned long long*)slice_counts_u64 : nullptr;
        const bool er = (erode_mask >> s) & 1u;
        if (er && (W & 127)) k_morph4<true, true><<<grid, 256, smem, st>>>(src, dst, Z, H, W, nw, lanes_x, pzb, my, cnt);
        else if (er) k_morph4<true, false><<<grid, 256, smem, st>>>(src, dst, Z, H, W, nw, lanes_x, pzb, my, cnt);
        else k_morph4<false, false><<<grid, 256, smem, st>>>(src, dst, Z, H, W, nw, lanes_x, pzb, my, cnt);
        src = dst;
    }
    T3D_CHECK_LAUNCH("t3d_morph");
    t3d_count_launches(n_stages);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// per-slice popcounts + bounding box of the set voxels (warp-shuffle tree reductions)
// bbox = {zmin, zmax, ymin, ymax, xmin, xmax}; empty volume leaves {INT_MAX, -1, ...}
// ------------------------------------------------------------------------------------------------
__global__ void k_stats_init(unsigned long long* counts, int Z, int* bbox)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (counts && i < Z) counts[i] = 0;
    if (bbox && i < 6) bbox[i] = (i & 1) ? -1 : 0x7fffffff;
}

#define ST_YSPLIT 4

__global__ void __launch_bounds__(256) k_stats(const uint32_t* __restrict__ bits, int H, int nw,
                                               unsigned long long* __restrict__ counts, int* __restrict__ bbox)
{
    const int z = blockIdx.z;
    const int w = blockIdx.x * 32 + threadIdx.x;
    const int rows_per = (H + ST_YSPLIT - 1) / ST_YSPLIT;
    const int ya = blockIdx.y * rows_per, yb = min(H, ya + rows_per);
    const uint32_t* p = bits + (int64_t)z * H * nw;
    unsigned long long cnt = 0;
    int ymin = 0x7fffffff, ymax = -1, xmin = 0x7fffffff, xmax = -1;
    if (w < nw) {
        for (int y = ya + threadIdx.y; y < yb; y += 8) {
            const uint32_t v = p[(int64_t)y * nw + w];
            if (v) {
                cnt += __popc(v);
                ymin = min(ymin, y); ymax = max(ymax, y);
                xmin = min(xmin, (w << 5) + __ffs(v) - 1);
                xmax = max(xmax, (w << 5) + 31 - __clz(v));
            }
        }
    }
    cnt = warp_sum(cnt);
    ymin = warp_min(ymin); xmin = warp_min(xmin); ymax = warp_max(ymax); xmax = warp_max(xmax);
    __shared__ unsigned long long sc[8];
    __shared__ int sb[8][4];
    const int wi = threadIdx.y;
    if (threadIdx.x == 0) { sc[wi] = cnt; sb[wi][0] = ymin; sb[wi][1] = ymax; sb[wi][2] = xmin; sb[wi][3] = xmax; }
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        for (int k = 1; k < 8; ++k) {
            cnt += sc[k]; ymin = min(ymin, sb[k][0]); ymax = max(ymax, sb[k][1]);
            xmin = min(xmin, sb[k][2]); xmax = max(xmax, sb[k][3]);
        }
        if (cnt) {
            if (counts) atomicAdd(counts + z, cnt);
            if (bbox) {
                atomicMin(bbox + 0, z); atomicMax(bbox + 1, z);
                atomicMin(bbox + 2, ymin); atomicMax(bbox + 3, ymax);
                atomicMin(bbox + 4, xmin); atomicMax(bbox + 5, xmax);
            }
        }
    }
}

extern "C" int t3d_volume_stats(const void* bits, int Z, int H, int W, void* slice_counts_u64, void* bbox_i32x6,
                                void* stream)
{
    if (Z <= 0 || H <= 0 || W <= 0) { t3d_set_error("t3d_volume_stats: empty volume"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nw = t3d_wpr(W);
    const int n_init = Z > 6 ? Z : 6;
    k_stats_init<<<(n_init + 255) / 256, 256, 0, st>>>((unsigned long long*)slice_counts_u64, Z, (int*)bbox_i32x6);
    dim3 grid((nw + 31) / 32, ST_YSPLIT, Z), block(32, 8);
    k_stats<<<grid, block, 0, st>>>((const uint32_t*)bits, H, nw, (unsigned long long*)slice_counts_u64, (int*)bbox_i32x6);
    T3D_CHECK_LAUNCH("t3d_volume_stats");
    t3d_count_launches(2);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// point cloud (voxel_processor.py:99-127): np.where order (C order), every `sub`-th set voxel,
// z -> cum[z] + depth[z]/2, y*mm_y, x*mm_x, float64 (N,3).
// Two kernels: per-row popcounts (then an exclusive scan by the caller, t3d_scan_u32) and emission.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_popc(const uint32_t* __restrict__ bits, int64_t n_rows, int nw,
                                                  uint32_t* __restrict__ row_counts)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    uint32_t c = 0;
    for (int w = lane_id(); w < nw; w += 32) c += __popc(bits[row * nw + w]);
    c = warp_sum(c);
    if (lane_id() == 0) row_counts[row] = c;
}

__global__ void __launch_bounds__(256) k_point_cloud(const uint32_t* __restrict__ bits, int64_t n_rows, int H, int nw,
                                                     const unsigned long long* __restrict__ row_base, int sub,
                                                     const double* __restrict__ zc_mm, double mm_y, double mm_x,
                                                     double* __restrict__ out)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int z = (int)(row / H), y = (int)(row - (int64_t)z * H);
    unsigned long long base = row_base[row];
    const uint32_t l = lane_id();
    for (int w0 = 0; w0 < nw; w0 += 32) {
        const int w = w0 + l;
        uint32_t v = (w < nw) ? bits[row * nw + w] : 0u;
        const uint32_t c = __popc(v);
        const uint32_t incl = warp_incl_scan(c);
        unsigned long long idx = base + incl - c;
        while (v) {
            const int b = __ffs(v) - 1;
            v &= v - 1;
            if (idx % (unsigned)sub == 0) {
                double* o = out + 3 * (idx / (unsigned)sub);
                o[0] = zc_mm[z];
                o[1] = (double)y * mm_y;
                o[2] = (double)((w << 5) + b) * mm_x;
            }
            ++idx;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

extern "C" int t3d_row_popcounts(const void* bits, int Z, int H, int W, void* row_counts_u32, void* stream)
{
    const int64_t rows = (int64_t)Z * H;
    if (rows <= 0) return 0;
    k_row_popc<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)bits, rows,
                                                                                      t3d_wpr(W), (uint32_t*)row_counts_u32);
    T3D_CHECK_LAUNCH("t3d_row_popcounts");
    t3d_count_launches(1);
    return 0;
}

extern "C" int t3d_point_cloud_emit(const void* bits, int Z, int H, int W, const void* row_base_u64, int subsample,
                                    const void* z_centre_mm_f64, double mm_per_pixel_y, double mm_per_pixel_x,
                                    void* out_f64, void* stream)
{
    const int64_t rows = (int64_t)Z * H;
    if (rows <= 0) return 0;
    if (subsample < 1) subsample = 1;
    k_point_cloud<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint32_t*)bits, rows, H, t3d_wpr(W), (const unsigned long long*)row_base_u64, subsample,
        (const double*)z_centre_mm_f64, mm_per_pixel_y, mm_per_pixel_x, (double*)out_f64);
    T3D_CHECK_LAUNCH("t3d_point_cloud_emit");
    t3d_count_launches(1);
    return 0;
}
