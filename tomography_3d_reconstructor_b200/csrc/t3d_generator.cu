// t3d_generator.cu -- half-ellipsoid end-cap slices on the device.
//
// The reference builds the Section_0 / Section_2 end caps by scaling the boundary Section_1 mask about the centre of its
// fitted ellipse, factor = sqrt(1 - (z/c)^2), one cv2.warpAffine (INTER_LINEAR, constant border 0) per slice
// (ellipsoid_slice_generator.py:61-77, 107-143).  This kernel is OpenCV's warpAffine arithmetic, bit for bit
// (imgproc/src/imgwarp.cpp: WarpAffineInvoker + remapBilinear on 8-bit data), for all slices of a cap in one launch:
//   * the 2x3 matrix is inverted by the host exactly as cv::warpAffine does (float64);
//   * source coordinates are fixed point with AB_BITS = 10: X = (rint((M1*y + M2) * 1024) + 16 + rint(M0 * x * 1024)) >> 5,
//     integer part X >> 5, 5-bit fraction X & 31 (same for Y);
//   * bilinear weights are 15-bit fixed point, (32 - ax)(32 - ay) * 32 etc. (they sum to 32768 exactly, so OpenCV's table
//     fix-up never applies), result = (sum + 16384) >> 15; samples outside the image read 0.
// tests/test_gpu_generator.py compares against cv2.warpAffine itself.
#include "t3d.h"
#include "t3d_common.cuh"

struct CapArgs {
    const uint8_t* base;   // (H, W) source mask
    uint8_t* out;          // (n, H, W)
    const double* m;       // per slice: inverted matrix {M0, M1, M2, M3, M4, M5}; M0 = 0 marks an empty slice
    int n, H, W;
};

__device__ __forceinline__ int cap_px(const uint8_t* __restrict__ b, int H, int W, int y, int x)
{
    return ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) ? (int)b[(int64_t)y * W + x] : 0;
}

__global__ void __launch_bounds__(256) k_endcap_slices(CapArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= a.W) return;
    const double* m = a.m + 6 * k;
    uint8_t* o = a.out + ((int64_t)k * a.H + y) * a.W + x;
    if (m[0] == 0.0 && m[4] == 0.0) { *o = 0; return; }         // factor <= 0 or z outside [0, c]: np.zeros_like
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
    const int X = (X0 + __double2int_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0))) >> 5;
    const int Y = (Y0 + __double2int_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0))) >> 5;
    // cv::remap stores the integer parts as saturated shorts
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int ax = X & 31, ay = Y & 31;
    const int s = (32 - ax) * (32 - ay) * 32 * cap_px(a.base, a.H, a.W, sy, sx) + ax * (32 - ay) * 32 * cap_px(a.base, a.H, a.W, sy, sx + 1) +
                  (32 - ax) * ay * 32 * cap_px(a.base, a.H, a.W, sy + 1, sx) + ax * ay * 32 * cap_px(a.base, a.H, a.W, sy + 1, sx + 1);
    *o = (uint8_t)((s + 16384) >> 15);
}

// out_u8 (n, H, W): slice k = warpAffine(base, M_k) with inv_matrices_f64 (device, 6 doubles per slice) the matrices already
// inverted as cv::warpAffine inverts them; an all-zero matrix yields an all-zero slice.
extern "C" int t3d_endcap_slices(const void* base_u8, int H, int W, const void* inv_matrices_f64, int n, void* out_u8, void* stream)
{
    if (H <= 0 || W <= 0 || n <= 0 || H > 32767 || W > 32767) { t3d_set_error("t3d_endcap_slices: bad sizes"); return 2; }
    CapArgs a;
    a.base = (const uint8_t*)base_u8; a.out = (uint8_t*)out_u8; a.m = (const double*)inv_matrices_f64; a.n = n; a.H = H; a.W = W;
    dim3 grid((W + 255) / 256, H, n);
    k_endcap_slices<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    T3D_CHECK_LAUNCH("t3d_endcap_slices");
    t3d_count_launches(1);
    return 0;
}
