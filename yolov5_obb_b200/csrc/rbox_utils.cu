// Post-NMS geometry and CSL targets (HBM-bound elementwise kernels; one thread per box):
//   y5obb_rbox2poly_f32      /root/reference/utils/rboxs_utils.py:106-126  (torch branch of rbox2poly)
//   y5obb_poly2hbb_f32       /root/reference/utils/rboxs_utils.py:147-165  (torch branch of poly2hbb)
//   y5obb_scale_polys_f32    /root/reference/utils/general.py:636-650      (scale_polys, in place)
//   y5obb_gaussian_label     /root/reference/utils/rboxs_utils.py:9-26     (gaussian_label_cpu: 180-bin CSL row)
// Each arithmetic step is a separately rounded fp32 op (no FMA contraction), like the chain of ATen
// elementwise kernels it replaces; the CSL row is evaluated in fp64 like the numpy original.
#include <algorithm>

#include "common.cuh"

namespace y5obb {
namespace {

__global__ void k_rbox2poly(const float* __restrict__ r, float* __restrict__ p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float cx = r[5 * i], cy = r[5 * i + 1], w = r[5 * i + 2], h = r[5 * i + 3], th = r[5 * i + 4];
  const float c = cosf(th), s = sinf(th);
  const float w2 = __fdiv_rn(w, 2.0f), h2 = __fdiv_rn(h, 2.0f);
  const float v1x = __fmul_rn(w2, c), v1y = __fmul_rn(-w2, s);   // vector1 = (w/2 * Cos, -w/2 * Sin)
  const float v2x = __fmul_rn(-h2, s), v2y = __fmul_rn(-h2, c);  // vector2 = (-h/2 * Sin, -h/2 * Cos)
  float* o = p + 8 * i;
  o[0] = __fadd_rn(__fadd_rn(cx, v1x), v2x);  // point1 = center + vector1 + vector2
  o[1] = __fadd_rn(__fadd_rn(cy, v1y), v2y);
  o[2] = __fsub_rn(__fadd_rn(cx, v1x), v2x);  // point2 = center + vector1 - vector2
  o[3] = __fsub_rn(__fadd_rn(cy, v1y), v2y);
  o[4] = __fsub_rn(__fsub_rn(cx, v1x), v2x);  // point3 = center - vector1 - vector2
  o[5] = __fsub_rn(__fsub_rn(cy, v1y), v2y);
  o[6] = __fadd_rn(__fsub_rn(cx, v1x), v2x);  // point4 = center - vector1 + vector2
  o[7] = __fadd_rn(__fsub_rn(cy, v1y), v2y);
}

__global__ void k_poly2hbb(const float* __restrict__ p, float* __restrict__ hbb, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* q = p + 8 * i;
  const float xmax = fmaxf(fmaxf(q[0], q[2]), fmaxf(q[4], q[6])), xmin = fminf(fminf(q[0], q[2]), fminf(q[4], q[6]));
  const float ymax = fmaxf(fmaxf(q[1], q[3]), fmaxf(q[5], q[7])), ymin = fminf(fminf(q[1], q[3]), fminf(q[5], q[7]));
  hbb[4 * i] = __fdiv_rn(__fadd_rn(xmax, xmin), 2.0f);
  hbb[4 * i + 1] = __fdiv_rn(__fadd_rn(ymax, ymin), 2.0f);
  hbb[4 * i + 2] = __fsub_rn(xmax, xmin);
  hbb[4 * i + 3] = __fsub_rn(ymax, ymin);
}

__global__ void k_scale_polys(float* __restrict__ p, int64_t n, float padx, float pady, float gain) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 8) return;
  const float pad = (i & 1) ? pady : padx;
  p[i] = __fdiv_rn(__fsub_rn(p[i], pad), gain);
}

// out[k] = exp(-(x_j)^2 / (2 sigma^2)), j = (k + int(nc/2 - angle)) mod nc, x_j = j - nc/2
__global__ void k_gaussian_label(const double* __restrict__ angle, float* __restrict__ out, int64_t n, int nc, double sig) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * nc) return;
  const int64_t row = i / nc;
  const int k = (int)(i - row * nc);
  const int index = (int)((double)nc / 2.0 - angle[row]);  // int(): truncation toward zero
  int j = (k + index) % nc;
  if (j < 0) j += nc;
  const double x = (double)j - (double)nc / 2.0;
  out[i] = (float)exp(-(x * x) / (2.0 * sig * sig));
}

}  // namespace
}  // namespace y5obb

using namespace y5obb;

extern "C" {

int y5obb_rbox2poly_f32(const float* rboxes5, float* polys8, int64_t n, void* stream) {
  if (n < 0) return Y5OBB_EINVAL;
  if (n == 0) return Y5OBB_OK;
  if (!rboxes5 || !polys8) return Y5OBB_EINVAL;
  k_rbox2poly<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rboxes5, polys8, n);
  Y5_LAUNCH_CHECK();
  return Y5OBB_OK;
}

int y5obb_poly2hbb_f32(const float* polys8, float* hbb4, int64_t n, void* stream) {
  if (n < 0) return Y5OBB_EINVAL;
  if (n == 0) return Y5OBB_OK;
  if (!polys8 || !hbb4) return Y5OBB_EINVAL;
  k_poly2hbb<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(polys8, hbb4, n);
  Y5_LAUNCH_CHECK();
  return Y5OBB_OK;
}

int y5obb_scale_polys_f32(float* polys8, int64_t n, float pad_x, float pad_y, float gain, void* stream) {
  if (n < 0 || gain == 0.0f) return Y5OBB_EINVAL;
  if (n == 0) return Y5OBB_OK;
  if (!polys8) return Y5OBB_EINVAL;
  k_scale_polys<<<(unsigned)((n * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(polys8, n, pad_x, pad_y, gain);
  Y5_LAUNCH_CHECK();
  return Y5OBB_OK;
}

int y5obb_gaussian_label(const double* angle_deg, float* csl_out, int64_t n, int num_class, double sigma, void* stream) {
  if (n < 0 || num_class <= 0 || !(sigma > 0)) return Y5OBB_EINVAL;
  if (n == 0) return Y5OBB_OK;
  if (!angle_deg || !csl_out) return Y5OBB_EINVAL;
  k_gaussian_label<<<(unsigned)((n * num_class + 255) / 256), 256, 0, (cudaStream_t)stream>>>(angle_deg, csl_out, n,
                                                                                              num_class, sigma);
  Y5_LAUNCH_CHECK();
  return Y5OBB_OK;
}

}  // extern "C"
