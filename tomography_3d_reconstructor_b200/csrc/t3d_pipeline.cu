// t3d_pipeline.cu -- the whole hot path as ONE enqueue: mask stack -> voxel grid -> smoothing -> canonical mesh + volumes.
//
// Order of work = the reference orchestrator's (tomography_3d_reconstruction.py:88-140, 194-229), every stage one of
// the kernels of this library.  Nothing here synchronises with the host: data-dependent sizes (active words, vertices,
// faces) stay in device memory, all buffers are capacity-sized by the caller and nothing is written beyond a
// capacity (the overflow flags in the result block tell the caller to retry with larger capacities).  The call can
// therefore be captured in a CUDA graph and replayed with a single launch; hole filling of the two end slices and the
// bounding-box reduction run on an internal side stream, forked and joined with events (also capturable).
#include "t3d.h"
#include "t3d_common.cuh"

// result block (uint64 slots); keep in sync with include/t3d.h and pipeline.py
enum {
    R_NACTIVE = 0, R_NX = 1, R_NY = 2, R_NZ = 3, R_NT = 4, R_VCANON = 5, R_FCANON = 6, R_UNVERIFIED = 7, R_OVERFLOW = 8,
    R_NAMBIGUOUS = 9, R_NEXACT = 10, R_VOLUME_F64 = 11, R_AREA_F64 = 12, R_BBOX_I32X6 = 13 /* 3 slots */, R_VRAW = 16, R_NEXC = 17,
    R_COUNTS = 32  /* Z raw per-slice counts, then Z smoothed per-slice counts */
};

#define EXC_CAP (1u << 20)  // capacity of the list of sign words that need the exact field evaluation

static inline int64_t al(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct Layout {
    int64_t bitsA, bitsB, bitsC, morph, fill, sign, exc, ballots, chunkbase, scan1, aw_idx, aw_cnt, aw_base, scan2, vkeys, verts_raw,
        faces_raw, canon, measure, total;
};

static Layout make_layout(int Z, int H, int W, int pad, uint32_t capNA, uint32_t capV, uint32_t capF, int n_stages)
{
    Layout L;
    const int64_t nw = t3d_words_per_row(W), vol = (int64_t)Z * H * nw * 4;
    const int Zp = Z + 2 * pad, Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int64_t nwp = t3d_words_per_row(Wp), n_chunks = t3d_mc_num_chunks(Zp, Hp, Wp);
    int64_t o = 0;
    L.bitsA = o; o += al(vol);
    L.bitsB = o; o += al(vol);
    L.bitsC = o; o += al(vol);
    L.morph = o; o += al(t3d_morph_scratch_bytes(Z, H, W, n_stages));
    L.fill = o; o += al(t3d_fill_holes_scratch_bytes(2, H, W));
    L.sign = o; o += al((int64_t)Zp * Hp * nwp * 4);
    L.exc = o; o += al((int64_t)EXC_CAP * 8);
    L.ballots = o; o += al(n_chunks * 4);
    L.chunkbase = o; o += al(n_chunks * 4);
    L.scan1 = o; o += al(t3d_scan_workspace_bytes(n_chunks, 1));
    L.aw_idx = o; o += al((int64_t)capNA * 4);
    L.aw_cnt = o; o += al((int64_t)capNA * 16);
    L.aw_base = o; o += al((int64_t)capNA * 16);
    L.scan2 = o; o += al(t3d_scan_workspace_bytes(capNA, 4));
    L.vkeys = o; o += al((int64_t)capV * 8);
    L.verts_raw = o; o += al((int64_t)capV * 12);
    L.faces_raw = o; o += al((int64_t)capF * 12);
    L.canon = o; o += al(t3d_canonicalize_fast_workspace_bytes(capV, capF));
    L.measure = o; o += al(t3d_mesh_measure_workspace_bytes());
    L.total = o;
    return L;
}

extern "C" int64_t t3d_reconstruct_workspace_bytes(int Z, int H, int W, int add_padding, int n_stages, uint32_t cap_active,
                                                   uint32_t cap_verts, uint32_t cap_faces)
{
    return make_layout(Z, H, W, add_padding ? 1 : 0, cap_active, cap_verts, cap_faces, n_stages).total;
}

extern "C" int64_t t3d_reconstruct_results_len(int Z) { return R_COUNTS + 2 * (int64_t)Z; }

__global__ void k_finalize_sizes(unsigned long long* r, unsigned long long capNA, unsigned long long capV, unsigned long long capF)
{
    unsigned long long of8 = (r[R_NEXC] > EXC_CAP) ? 8ull : 0ull;
    const unsigned long long v = r[R_NX] + r[R_NY] + r[R_NZ];
    r[R_VRAW] = v;
    unsigned long long of = 0;
    if (r[R_NACTIVE] > capNA) of |= 1;
    if (v > capV) of |= 2;
    if (r[R_NT] > capF) of |= 4;
    r[R_OVERFLOW] = of | of8;
}

struct SideStream {
    cudaStream_t s = nullptr;
    cudaEvent_t e[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
static SideStream g_side[64];

static int side_for_current_device(SideStream** out)
{
    int dev = 0;
    T3D_CUDA(cudaGetDevice(&dev));
    SideStream& s = g_side[dev & 63];
    if (!s.s) {
        T3D_CUDA(cudaStreamCreateWithFlags(&s.s, cudaStreamNonBlocking));
        for (int k = 0; k < 6; ++k) T3D_CUDA(cudaEventCreateWithFlags(&s.e[k], cudaEventDisableTiming));
    }
    *out = &s;
    return 0;
}

#define RUN(call) do { if (int rc__ = (call)) return rc__; } while (0)

extern "C" int t3d_reconstruct(const void* masks_u8, int Z, int H, int W, int threshold, int close_ends, int n_stages,
                               unsigned erode_mask, int add_padding, const double* weights3_host, const void* cum_f64,
                               const void* adj_f64, int n_cum, double mm_per_pixel_y, double mm_per_pixel_x, int scale_in_f64,
                               uint32_t cap_active, uint32_t cap_verts, uint32_t cap_faces, void* verts_out_f32, void* faces_out_i64,
                               void* results_u64, void* workspace, void* stream)
{
    if (Z <= 0 || H <= 0 || W <= 0) { t3d_set_error("t3d_reconstruct: empty volume"); return 2; }
    if (cap_active == 0 || cap_verts == 0 || cap_faces == 0) { t3d_set_error("t3d_reconstruct: zero capacity"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const int pad = add_padding ? 1 : 0;
    const Layout L = make_layout(Z, H, W, pad, cap_active, cap_verts, cap_faces, n_stages);
    char* ws = (char*)workspace;
    unsigned long long* R = (unsigned long long*)results_u64;
    const int64_t nw = t3d_words_per_row(W), plane_words = (int64_t)H * nw, plane_bytes = (int64_t)H * W;
    const int Zp = Z + 2 * pad, Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int64_t n_chunks = t3d_mc_num_chunks(Zp, Hp, Wp);
    uint32_t* bitsA = (uint32_t*)(ws + L.bitsA);
    uint32_t* bitsB = (uint32_t*)(ws + L.bitsB);
    uint32_t* bitsC = (uint32_t*)(ws + L.bitsC);
    const uint8_t* m = (const uint8_t*)masks_u8;
    SideStream* side;
    RUN(side_for_current_device(&side));
    T3D_CUDA(cudaMemsetAsync(R, 0, sizeof(unsigned long long) * R_COUNTS, st));

    // ---- create_voxel_data: pack, fill the holes of the end slices (side stream), z gap fill + per-slice counts
    if (close_ends && Z >= 3) {
        RUN(t3d_pack_masks(m, 1, H, W, threshold, bitsA, st));
        RUN(t3d_pack_masks(m + (int64_t)(Z - 1) * plane_bytes, 1, H, W, threshold, bitsA + (int64_t)(Z - 1) * plane_words, st));
        T3D_CUDA(cudaEventRecord(side->e[0], st));
        T3D_CUDA(cudaStreamWaitEvent(side->s, side->e[0], 0));
        RUN(t3d_fill_holes_2d(bitsA, 2, (int64_t)(Z - 1) * plane_words, H, W, ws + L.fill, side->s));
        T3D_CUDA(cudaEventRecord(side->e[1], side->s));
        RUN(t3d_pack_masks(m + plane_bytes, Z - 2, H, W, threshold, bitsA + plane_words, st));
        T3D_CUDA(cudaStreamWaitEvent(st, side->e[1], 0));
        RUN(t3d_gap_fill(bitsA, bitsB, nullptr, nullptr, Z, H, W, R + R_COUNTS, st));
    } else if (close_ends) {
        RUN(t3d_pack_masks(m, Z, H, W, threshold, bitsA, st));
        RUN(t3d_fill_holes_2d(bitsA, 1, 0, H, W, ws + L.fill, st));
        if (Z > 1) RUN(t3d_fill_holes_2d(bitsA + (int64_t)(Z - 1) * plane_words, 1, 0, H, W, ws + L.fill, st));
        RUN(t3d_gap_fill(bitsA, bitsB, nullptr, nullptr, Z, H, W, R + R_COUNTS, st));
    } else {
        RUN(t3d_pack_masks(m, Z, H, W, threshold, bitsB, st));
        RUN(t3d_volume_stats(bitsB, Z, H, W, R + R_COUNTS, nullptr, st));
    }
    // bounding box of the raw grid: side stream, joined at the end
    T3D_CUDA(cudaEventRecord(side->e[2], st));
    T3D_CUDA(cudaStreamWaitEvent(side->s, side->e[2], 0));
    RUN(t3d_volume_stats(bitsB, Z, H, W, nullptr, R + R_BBOX_I32X6, side->s));
    T3D_CUDA(cudaEventRecord(side->e[3], side->s));

    // ---- smooth_voxel_data
    const uint32_t* smoothed = bitsB;
    if (n_stages > 0) {
        RUN(t3d_morph(bitsB, bitsC, Z, H, W, n_stages, erode_mask, R + R_COUNTS + Z, ws + L.morph, st));
        smoothed = bitsC;
    } else {
        T3D_CUDA(cudaMemcpyAsync(R + R_COUNTS + Z, R + R_COUNTS, sizeof(unsigned long long) * Z, cudaMemcpyDeviceToDevice, st));
    }

    // ---- extract_manifold_surface: field sign, two-pass marching cubes, vertices
    RUN(t3d_field_sign_lean(smoothed, Z, H, W, pad, weights3_host, ws + L.sign, R + R_NEXACT, ws + L.exc, EXC_CAP, R + R_NEXC, st));
    RUN(t3d_mc_flags(ws + L.sign, Zp, Hp, Wp, 0, -1, ws + L.ballots, st));
    RUN(t3d_exclusive_scan_u32(ws + L.ballots, ws + L.chunkbase, n_chunks, 1, 0, 1, R + R_NACTIVE, ws + L.scan1, st));
    RUN(t3d_mc_words_dev(ws + L.sign, Zp, Hp, Wp, 0, -1, ws + L.ballots, ws + L.chunkbase, cap_active, R + R_NACTIVE, ws + L.aw_idx,
                         ws + L.aw_cnt, R + R_NAMBIGUOUS, st));
    RUN(t3d_exclusive_scan_u32_dev(ws + L.aw_cnt, ws + L.aw_base, cap_active, cap_active, 4, 0, 0, R + R_NACTIVE, R + R_NX,
                                   ws + L.scan2, st));
    k_finalize_sizes<<<1, 1, 0, st>>>(R, cap_active, cap_verts, cap_faces);
    RUN(t3d_mc_emit_dev(ws + L.sign, Zp, Hp, Wp, 0, -1, ws + L.ballots, ws + L.chunkbase, ws + L.aw_idx, ws + L.aw_base, cap_active,
                        R + R_NACTIVE, cap_verts, cap_faces, ws + L.vkeys, ws + L.faces_raw, st));
    RUN(t3d_mc_vertices_dev(smoothed, Z, H, W, pad, 1, weights3_host, ws + L.vkeys, R + R_NACTIVE, cap_verts, 1, 0, cum_f64, adj_f64,
                            n_cum, mm_per_pixel_y, mm_per_pixel_x, scale_in_f64, ws + L.verts_raw, st));
    // ---- mesh volume / area on the emitted mesh (side stream, concurrent with the canonical sort), canonical mesh
    T3D_CUDA(cudaStreamWaitEvent(st, side->e[3], 0));   // the bbox reduction is done with the side stream
    T3D_CUDA(cudaEventRecord(side->e[4], st));
    T3D_CUDA(cudaStreamWaitEvent(side->s, side->e[4], 0));
    RUN(t3d_mesh_measure_dev(ws + L.verts_raw, ws + L.faces_raw, cap_faces, R + R_NT, 0, R + R_VOLUME_F64, ws + L.measure, side->s));
    T3D_CUDA(cudaEventRecord(side->e[5], side->s));
    RUN(t3d_mesh_canonicalize_fast_dev(ws + L.verts_raw, cap_verts, R + R_VRAW, ws + L.faces_raw, cap_faces, R + R_NT, verts_out_f32,
                                       faces_out_i64, nullptr, R + R_VCANON, ws + L.canon, st));
    T3D_CUDA(cudaStreamWaitEvent(st, side->e[5], 0));
    T3D_CHECK_LAUNCH("t3d_reconstruct");
    t3d_count_launches(1);
    return 0;
}
