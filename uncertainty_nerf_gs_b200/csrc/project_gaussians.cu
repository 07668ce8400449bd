// (f2) Producer side of the splat path: EWA projection of 3-D Gaussians and view-dependent colours.
//
// Reference call sites (file:line under /root/reference/nerfuncertainty):
//   models/activesplatfacto/activesplatfacto_model.py:221-234   project_gaussians(means, exp(scales), 1, quats / |quats|,
//                                                                viewmat[:3, :], fx, fy, cx, cy, H, W, 16)
//   models/activesplatfacto/activesplatfacto_model.py:242-249   rgbs = clamp(spherical_harmonics(n, viewdirs, coeffs) + 0.5, 0)
// Both functions live in gsplat 0.1.11 (README.md:30; not vendored: parity unpinned).  The arithmetic below restates
// gsplat's published kernels: near-plane clip on the camera-space z, cov3d = (R S)(R S)^T, EWA Jacobian with the
// 1.3 x tan(fov/2) clamp of x/z and y/z, +0.3 px^2 blur with its density compensation, conic = inverse of the blurred
// 2-D covariance, radius = ceil(3 sqrt(largest eigenvalue)), centre = (fx x/(z + 1e-6) + cx, ...), tile rectangle as in
// binning.cu; real spherical harmonics up to degree 3 in the 3DGS sign convention.  Outputs of culled Gaussians are
// zero, with gsplat's partial-write behaviour kept (cov3d is written before the determinant / tile-area culls, the
// conic before the tile-area cull).
//
// One thread per Gaussian: 40 B in, 60 B out -- a pure streaming kernel (HBM-bound, ~20 us per million Gaussians).
#include "ub_common.cuh"

namespace ub {

struct ProjectParams {
  const float* means3d;
  const float* scales;
  const float* quats;
  const float* viewmat;  // device, [3, 4] row-major (world -> camera)
  float glob_scale, fx, fy, cx, cy, clip_thresh;
  int height, width;
  long long num;
  float* xys;
  float* depths;
  int32_t* radii;
  float* conics;
  float* compensation;
  int32_t* num_tiles_hit;
  float* cov3d;
};

__global__ void __launch_bounds__(256) project_gaussians_kernel(const ProjectParams p) {
  __shared__ float vm[12];
  if (threadIdx.x < 12) vm[threadIdx.x] = p.viewmat[threadIdx.x];
  __syncthreads();
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.num) return;

  float xy0 = 0.f, xy1 = 0.f, depth = 0.f, comp = 0.f;
  float con0 = 0.f, con1 = 0.f, con2 = 0.f;
  float c3[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int radius_i = 0, tiles = 0;

  const float mx = p.means3d[3 * g + 0], my = p.means3d[3 * g + 1], mz = p.means3d[3 * g + 2];
  // camera-space position
  const float vx = vm[0] * mx + vm[1] * my + vm[2] * mz + vm[3];
  const float vy = vm[4] * mx + vm[5] * my + vm[6] * mz + vm[7];
  const float vz = vm[8] * mx + vm[9] * my + vm[10] * mz + vm[11];
  if (!(vz <= p.clip_thresh)) {
    // cov3d = (R S)(R S)^T, quaternion (w, x, y, z)
    const float qw0 = p.quats[4 * g + 0], qx0 = p.quats[4 * g + 1], qy0 = p.quats[4 * g + 2], qz0 = p.quats[4 * g + 3];
    const float qs = rsqrtf(qw0 * qw0 + qx0 * qx0 + qy0 * qy0 + qz0 * qz0);
    const float w = qw0 * qs, x = qx0 * qs, y = qy0 * qs, z = qz0 * qs;
    const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - w * z), 2.f * (x * z + w * y)},
                           {2.f * (x * y + w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - w * x)},
                           {2.f * (x * z - w * y), 2.f * (y * z + w * x), 1.f - 2.f * (x * x + y * y)}};
    const float s[3] = {p.glob_scale * p.scales[3 * g + 0], p.glob_scale * p.scales[3 * g + 1],
                        p.glob_scale * p.scales[3 * g + 2]};
    float M[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) M[r][c] = R[r][c] * s[c];
    float V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) V[r][c] = M[r][0] * M[c][0] + M[r][1] * M[c][1] + M[r][2] * M[c][2];
    c3[0] = V[0][0]; c3[1] = V[0][1]; c3[2] = V[0][2]; c3[3] = V[1][1]; c3[4] = V[1][2]; c3[5] = V[2][2];

    // EWA projection
    const float tan_fovx = 0.5f * (float)p.width / p.fx, tan_fovy = 0.5f * (float)p.height / p.fy;
    const float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
    const float tz = vz;
    const float tx = tz * fminf(lim_x, fmaxf(-lim_x, vx / tz));
    const float ty = tz * fminf(lim_y, fmaxf(-lim_y, vy / tz));
    const float rz = 1.f / tz, rz2 = rz * rz;
    const float J[2][3] = {{p.fx * rz, 0.f, -p.fx * tx * rz2}, {0.f, p.fy * rz, -p.fy * ty * rz2}};
    float T[2][3];  // J W, W = rotation part of the view matrix
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) T[r][c] = J[r][0] * vm[0 + c] + J[r][1] * vm[4 + c] + J[r][2] * vm[8 + c];
    float TV[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) TV[r][c] = T[r][0] * V[0][c] + T[r][1] * V[1][c] + T[r][2] * V[2][c];
    const float c00 = TV[0][0] * T[0][0] + TV[0][1] * T[0][1] + TV[0][2] * T[0][2];
    const float c01 = TV[0][0] * T[1][0] + TV[0][1] * T[1][1] + TV[0][2] * T[1][2];
    const float c11 = TV[1][0] * T[1][0] + TV[1][1] * T[1][1] + TV[1][2] * T[1][2];
    const float det_orig = c00 * c11 - c01 * c01;
    const float a = c00 + 0.3f, b = c01, c = c11 + 0.3f;
    const float det = a * c - b * b;
    const float comp_v = sqrtf(fmaxf(0.f, det_orig / det));
    if (det != 0.f) {
      const float inv_det = 1.f / det;
      con0 = c * inv_det;
      con1 = -b * inv_det;
      con2 = a * inv_det;
      const float mid = 0.5f * (a + c);
      const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
      const float radius = ceilf(3.f * sqrtf(fmaxf(mid + disc, mid - disc)));
      const float rw = 1.f / (vz + 1e-6f);
      const float px = vx * rw * p.fx + p.cx, py = vy * rw * p.fy + p.cy;
      const int tiles_x = (p.width + UB_TILE - 1) / UB_TILE, tiles_y = (p.height + UB_TILE - 1) / UB_TILE;
      const float tcx = px / (float)UB_TILE, tcy = py / (float)UB_TILE, tr = radius / (float)UB_TILE;
      const int x0 = min(max(0, (int)(tcx - tr)), tiles_x), x1 = min(max(0, (int)(tcx + tr + 1.f)), tiles_x);
      const int y0 = min(max(0, (int)(tcy - tr)), tiles_y), y1 = min(max(0, (int)(tcy + tr + 1.f)), tiles_y);
      const int area = (x1 - x0) * (y1 - y0);
      if (area > 0) {
        tiles = area;
        depth = vz;
        radius_i = (int)radius;
        xy0 = px;
        xy1 = py;
        comp = comp_v;
      }
    }
  }
  p.xys[2 * g + 0] = xy0;
  p.xys[2 * g + 1] = xy1;
  p.depths[g] = depth;
  p.radii[g] = radius_i;
  p.conics[3 * g + 0] = con0;
  p.conics[3 * g + 1] = con1;
  p.conics[3 * g + 2] = con2;
  if (p.compensation) p.compensation[g] = comp;
  if (p.num_tiles_hit) p.num_tiles_hit[g] = tiles;
  if (p.cov3d) {
#pragma unroll
    for (int k = 0; k < 6; ++k) p.cov3d[6 * g + k] = c3[k];
  }
}

// colours[g, c] = sum_k Y_k(viewdir_g) coeffs[g, k, c] for the first (degrees_to_use + 1)^2 bases
__global__ void __launch_bounds__(256)
spherical_harmonics_kernel(int num_bases, int degrees_to_use, const float* __restrict__ viewdirs,
                           const float* __restrict__ coeffs, long long num, float* __restrict__ out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= num) return;
  const float* cf = coeffs + (size_t)g * num_bases * 3;
  float col[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) col[c] = 0.28209479177387814f * cf[c];
  if (degrees_to_use >= 1) {
    float x = viewdirs[3 * g + 0], y = viewdirs[3 * g + 1], z = viewdirs[3 * g + 2];
    const float norm = sqrtf(x * x + y * y + z * z);
    x /= norm;
    y /= norm;
    z /= norm;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      col[c] += 0.4886025119029199f * (-y * cf[3 + c] + z * cf[6 + c] - x * cf[9 + c]);
    if (degrees_to_use >= 2) {
      const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        col[c] += 1.0925484305920792f * xy * cf[12 + c] + -1.0925484305920792f * yz * cf[15 + c] +
                  0.31539156525252005f * (2.f * zz - xx - yy) * cf[18 + c] +
                  -1.0925484305920792f * xz * cf[21 + c] + 0.5462742152960396f * (xx - yy) * cf[24 + c];
      if (degrees_to_use >= 3) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          col[c] += -0.5900435899266435f * y * (3.f * xx - yy) * cf[27 + c] + 2.890611442640554f * xy * z * cf[30 + c] +
                    -0.4570457994644658f * y * (4.f * zz - xx - yy) * cf[33 + c] +
                    0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy) * cf[36 + c] +
                    -0.4570457994644658f * x * (4.f * zz - xx - yy) * cf[39 + c] +
                    1.445305721320277f * z * (xx - yy) * cf[42 + c] +
                    -0.5900435899266435f * x * (xx - 3.f * yy) * cf[45 + c];
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) out[3 * g + c] = col[c];
}

}  // namespace ub

extern "C" {

int ub_project_gaussians(const float* means3d, const float* scales, float glob_scale, const float* quats,
                         const float* viewmat, float fx, float fy, float cx, float cy, int32_t img_height,
                         int32_t img_width, float clip_thresh, int64_t num_gaussians, float* out_xys,
                         float* out_depths, int32_t* out_radii, float* out_conics, float* out_compensation,
                         int32_t* out_num_tiles_hit, float* out_cov3d, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_gaussians >= 0 && img_height >= 1 && img_width >= 1, UB_ERR_BAD_ARG, "project_gaussians: bad sizes");
  UB_REQUIRE(fx > 0.f && fy > 0.f, UB_ERR_BAD_ARG, "project_gaussians: focal lengths must be positive");
  if (num_gaussians == 0) return UB_OK;
  UB_REQUIRE(means3d && scales && quats && viewmat, UB_ERR_BAD_ARG, "project_gaussians: input pointer is NULL");
  UB_REQUIRE(out_xys && out_depths && out_radii && out_conics, UB_ERR_BAD_ARG,
             "project_gaussians: xys / depths / radii / conics outputs must be non-NULL");
  ProjectParams p{};
  p.means3d = means3d;
  p.scales = scales;
  p.quats = quats;
  p.viewmat = viewmat;
  p.glob_scale = glob_scale;
  p.fx = fx;
  p.fy = fy;
  p.cx = cx;
  p.cy = cy;
  p.clip_thresh = clip_thresh;
  p.height = img_height;
  p.width = img_width;
  p.num = num_gaussians;
  p.xys = out_xys;
  p.depths = out_depths;
  p.radii = out_radii;
  p.conics = out_conics;
  p.compensation = out_compensation;
  p.num_tiles_hit = out_num_tiles_hit;
  p.cov3d = out_cov3d;
  const unsigned blocks = (unsigned)((num_gaussians + 255) / 256);
  project_gaussians_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(p);
  return check_launch("project_gaussians");
}

int ub_spherical_harmonics(int32_t degree, int32_t degrees_to_use, const float* viewdirs, const float* coeffs,
                           int64_t num_gaussians, float* out_colors, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(degree >= 0 && degree <= 3, UB_ERR_UNSUPPORTED, "spherical_harmonics: degree must be in [0, 3] (got %d)",
             degree);
  UB_REQUIRE(degrees_to_use >= 0 && degrees_to_use <= degree, UB_ERR_BAD_ARG,
             "spherical_harmonics: degrees_to_use must be in [0, degree]");
  UB_REQUIRE(num_gaussians >= 0, UB_ERR_BAD_ARG, "spherical_harmonics: bad size");
  if (num_gaussians == 0) return UB_OK;
  UB_REQUIRE(coeffs && out_colors && (viewdirs || degrees_to_use == 0), UB_ERR_BAD_ARG,
             "spherical_harmonics: pointer is NULL");
  const unsigned blocks = (unsigned)((num_gaussians + 255) / 256);
  spherical_harmonics_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      (degree + 1) * (degree + 1), degrees_to_use, viewdirs, coeffs, num_gaussians, out_colors);
  return check_launch("spherical_harmonics");
}

}  // extern "C"
