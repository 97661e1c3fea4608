// Weight gradient of a convolution on the tensor cores (sm_100a): for every tap (kh, kw)
//   dW[tap][co][ci] = sum over (b, ho, wo) of dz[b, ho, wo, co] * x[b, s*ho + kh - ph, s*wo + kw - pw, ci]
// i.e. a GEMM with M = Cout, N = Cin and K = all output pixels.  Both operands are read straight from the NHWC bf16
// activations / gradients (channel slices allowed): a 4-D TMA box {64 channels, kwp, khp, 1} lands in shared memory
// as [64 pixels][128 B of channels], 128B-swizzled, which is the MN-major canonical layout of tcgen05.mma (the
// reduction axis K = pixel rows).  The tap is a shift of the x box's origin (zero padding = TMA out-of-bounds fill,
// stride 2 = tensor-map element strides on W and H, exactly as the forward kernel reads its input).  The pixel axis
// is split over CTAs (split-K); every CTA accumulates its slice in TMEM (fp32) and adds it into dW with fp32 atomics.
// Replaces the cuDNN wgrad call autograd makes for /root/reference/models/common.py:37-46 in train.py:333.
// Tensor-bound for wide layers (2 * pixels * Cout * Cin * taps flop), L2-bound for narrow ones.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace y5obb {
namespace {

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_SMEM = 200 * 1024;
constexpr int WG_BK = 64;                       // pixels per pipeline stage
constexpr uint32_t WG_CHUNK = WG_BK * 128;      // one 64-channel chunk of a stage: [64 pixels][128 B]

struct WgradK {
  CUtensorMap tmA;  // dz NHWC slice: (Cout, Wo, Ho, B), box {64, kwp, khp, 1}
  CUtensorMap tmB;  // x NHWC slice: (Cin, Wi, Hi, B), box {64, kwp*s, khp*s, 1}, element strides {1, s, s, 1}
  int B, Ho, Wo, Cout, Cin;
  int KH, KW, stride, pad_h, pad_w;
  int kwp, khp, BN;         // pixel tile kwp x khp = WG_BK; N tile (input channels, multiple of 16)
  int tiles_w, tiles_h;     // pixel tiles per image
  int ksteps;               // B * tiles_h * tiles_w
  int ksplit, co_blks, ci_blks;
  int stages, a_chunks, b_chunks;
  uint32_t idesc, tmem_cols;
  float* dW;                // element (tap, co, ci) at dW[tap * s_tap + row(co) * s_co + ci * s_ci]
  long long s_tap, s_co, s_ci;
  int co_group, co_group_pad;
};

// MN-major, 128B-swizzled operand: 64-channel chunks of [pixels][128 B]; 8-pixel groups 1024 B apart (SBO), chunks
// WG_CHUNK bytes apart (LBO).  (cute/atom/mma_traits_sm100.hpp: Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)).)
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(WG_CHUNK >> 4) << 16;   // LBO
  d |= (uint64_t)(1024u >> 4) << 32;      // SBO
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradK p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = 2u * WG_CHUNK;
  const uint32_t stage_bytes = a_bytes + (uint32_t)p.b_chunks * WG_CHUNK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item: (tap, co block, ci block, k split)
  int item = blockIdx.x;
  const int ks = item % p.ksplit;
  item /= p.ksplit;
  const int cib = item % p.ci_blks;
  item /= p.ci_blks;
  const int cob = item % p.co_blks;
  const int tap = item / p.co_blks;
  const int kh = tap / p.KW, kw = tap - kh * p.KW;
  const int k0 = (int)(((long long)p.ksteps * ks) / p.ksplit), k1 = (int)(((long long)p.ksteps * (ks + 1)) / p.ksplit);
  // 64-channel chunks of A this block really has (rows of D beyond Cout are never stored)
  const int a_chunks = min(2, (p.Cout - cob * 128 + 63) >> 6);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_base_smem, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = (uint32_t)(a_chunks + p.b_chunks) * WG_CHUNK;
      for (int k = k0; k < k1; ++k) {
        const int b = k / per_img;
        const int r = k - b * per_img;
        const int th = r / p.tiles_w;
        const int ho0 = th * p.khp, wo0 = (r - th * p.tiles_w) * p.kwp;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        ptx::mbar_expect_tx(&full_bar[s], tx);
        for (int c = 0; c < a_chunks; ++c)
          ptx::tma_load_4d(sa + (size_t)c * WG_CHUNK, &p.tmA, &full_bar[s], cob * 128 + c * 64, wo0, ho0, b);
        const int wi0 = wo0 * p.stride + kw - p.pad_w, hi0 = ho0 * p.stride + kh - p.pad_h;
        for (int c = 0; c < p.b_chunks; ++c)
          ptx::tma_load_4d(sa + a_bytes + (size_t)c * WG_CHUNK, &p.tmB, &full_bar[s], cib * p.BN + c * 64, wi0, hi0, b);
        if (++s == p.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint64_t desc_hi = make_mnmajor_desc(0u);
      const uint32_t ring = ptx::smem_u32(smem);
      uint32_t accumulate = 0u;
      for (int k = k0; k < k1; ++k) {
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t sa = ring + (uint32_t)s * stage_bytes;
        const uint64_t da = desc_hi | (uint64_t)((sa & 0x3FFFFu) >> 4);
        const uint64_t db = desc_hi | (uint64_t)(((sa + a_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
        for (int j = 0; j < WG_BK / 16; ++j) {  // 16 pixel rows = 2048 B per MMA
          ptx::umma_bf16(tmem_base, da + (uint64_t)(128 * j), db + (uint64_t)(128 * j), p.idesc, accumulate);
          accumulate = 1u;
        }
        ptx::umma_commit(&empty_bar[s]);
        if (++s == p.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      ptx::umma_commit(&done_bar);
    }
  } else {
    // epilogue: TMEM lane = output channel (row of dW), columns = input channels
    const int q = warp & 3;
    const int co = cob * 128 + q * 32 + lane;
    int co_row = co;
    bool co_ok = co < p.Cout;
    if (p.co_group_pad) {
      const int a = co / p.co_group_pad, c = co - a * p.co_group_pad;
      co_ok = co_ok && c < p.co_group;
      co_row = a * p.co_group + c;
    }
    if (k1 > k0) {
      ptx::mbar_wait(&done_bar, 0u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      float* row = p.dW + (long long)tap * p.s_tap + (long long)co_row * p.s_co + (long long)(cib * p.BN) * p.s_ci;
      const int ncols = min(p.BN, p.Cin - cib * p.BN);
      if (cob * 128 + q * 32 < p.Cout) {  // warp-uniform
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + (uint32_t)c0, r);
          ptx::tmem_ld_wait();
          if (co_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < ncols) atomicAdd(row + (long long)(c0 + j) * p.s_ci, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_tmapEncodeTiled wg_get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  });
  return fn;
}

struct WgradObj {
  WgradK k;
  int grid;
  size_t smem;
  double flops;
};

}  // namespace
}  // namespace y5obb

using namespace y5obb;

extern "C" {

int y5obb_wgrad_create(const y5obb_wgrad_desc* d, y5obb_wgrad_t** out) {
  if (!d || !out || !d->dz || !d->x || !d->dw) return Y5OBB_EINVAL;
  if (d->stride != 1 && d->stride != 2) return Y5OBB_EINVAL;
  if (d->B < 1 || d->Cout < 1 || d->Cin < 1 || d->KH < 1 || d->KW < 1 || d->Ho < 1 || d->Wo < 1) return Y5OBB_EINVAL;
  if ((d->dz_pix_stride & 7) || (d->x_pix_stride & 7) || d->dz_pix_stride < d->Cout || d->x_pix_stride < d->Cin)
    return Y5OBB_EINVAL;
  if ((reinterpret_cast<uintptr_t>(d->dz) | reinterpret_cast<uintptr_t>(d->x)) & 15) return Y5OBB_EINVAL;
  PFN_tmapEncodeTiled enc = wg_get_encode();
  if (!enc) return Y5OBB_ECUDA;
  WgradObj* o = new WgradObj();
  WgradK& k = o->k;
  memset(&k, 0, sizeof(k));
  k.B = d->B;
  k.Ho = d->Ho;
  k.Wo = d->Wo;
  k.Cout = d->Cout;
  k.Cin = d->Cin;
  k.KH = d->KH;
  k.KW = d->KW;
  k.stride = d->stride;
  k.pad_h = d->pad_h;
  k.pad_w = d->pad_w;
  k.dW = d->dw;
  if (d->dw_tap_stride || d->dw_co_stride || d->dw_ci_stride) {
    k.s_tap = d->dw_tap_stride;
    k.s_co = d->dw_co_stride;
    k.s_ci = d->dw_ci_stride;
  } else {
    k.s_tap = (long long)d->Cout * d->Cin;
    k.s_co = d->Cin;
    k.s_ci = 1;
  }
  if (d->co_group_pad < 0 || d->co_group < 0 || d->co_group > d->co_group_pad) {
    delete o;
    return Y5OBB_EINVAL;
  }
  k.co_group = d->co_group;
  k.co_group_pad = d->co_group_pad;
  // pixel tile kwp x khp = 64 output pixels: the shape that covers the map with the fewest tiles (ties: widest rows)
  long long best = -1;
  for (int kwp = 1; kwp <= WG_BK; kwp <<= 1) {
    const int khp = WG_BK / kwp;
    const long long cost = (long long)((d->Wo + kwp - 1) / kwp) * ((d->Ho + khp - 1) / khp);
    if (best < 0 || cost <= best) {
      best = cost;
      k.kwp = kwp;
      k.khp = khp;
    }
  }
  k.tiles_w = (d->Wo + k.kwp - 1) / k.kwp;
  k.tiles_h = (d->Ho + k.khp - 1) / k.khp;
  k.ksteps = d->B * k.tiles_w * k.tiles_h;
  k.ci_blks = (d->Cin + 255) / 256;
  k.BN = ((d->Cin + k.ci_blks - 1) / k.ci_blks + 15) / 16 * 16;
  k.b_chunks = (k.BN + 63) / 64;
  k.co_blks = (d->Cout + 127) / 128;
  const int items = d->KH * d->KW * k.co_blks * k.ci_blks;
  // split-K so that ~2 CTAs per SM exist, but never fewer than 16 pipeline steps per CTA: every split adds a
  // 128 x BN fp32 atomic epilogue
  k.ksplit = std::max(1, std::min(std::max(1, k.ksteps / 16), (2 * sm_count() + items - 1) / items));
  const uint32_t stage_bytes = (2u + (uint32_t)k.b_chunks) * WG_CHUNK;
  k.stages = (int)std::min<size_t>(WG_MAX_STAGES, WG_SMEM / stage_bytes);
  if (k.stages < 2) {
    delete o;
    return Y5OBB_EINVAL;
  }
  k.idesc = ptx::make_idesc_bf16(128, k.BN) | (1u << 15) | (1u << 16);  // A and B MN-major
  k.tmem_cols = 32;
  while ((int)k.tmem_cols < k.BN) k.tmem_cols <<= 1;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->dz_pix_stride * 2, (cuuint64_t)d->dz_pix_stride * d->Wo * 2,
                             (cuuint64_t)d->dz_pix_stride * d->Wo * d->Ho * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)k.kwp, (cuuint32_t)k.khp, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&k.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->dz), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      g_last_cuda_error = (int)r;
      delete o;
      return Y5OBB_ECUDA;
    }
  }
  {
    const cuuint32_t s = (cuuint32_t)d->stride;
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_pix_stride * 2, (cuuint64_t)d->x_pix_stride * d->Wi * 2,
                             (cuuint64_t)d->x_pix_stride * d->Wi * d->Hi * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)k.kwp * s, (cuuint32_t)k.khp * s, 1};
    cuuint32_t es[4] = {1, s, s, 1};
    CUresult r = enc(&k.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      g_last_cuda_error = (int)r;
      delete o;
      return Y5OBB_ECUDA;
    }
  }
  o->grid = items * k.ksplit;
  o->smem = std::max<size_t>((size_t)k.stages * stage_bytes + 1024, 116 * 1024);
  o->flops = 2.0 * d->B * d->Ho * d->Wo * (double)d->Cout * d->Cin * d->KH * d->KW;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM + 2048);
    if (e != cudaSuccess) {
      delete o;
      return cuda_fail(e);
    }
    attr_set = true;
  }
  *out = reinterpret_cast<y5obb_wgrad_t*>(o);
  return Y5OBB_OK;
}

int y5obb_wgrad_run(const y5obb_wgrad_t* w, void* stream) {
  if (!w) return Y5OBB_EINVAL;
  const WgradObj* o = reinterpret_cast<const WgradObj*>(w);
  wgrad_kernel<<<o->grid, WG_THREADS, o->smem, (cudaStream_t)stream>>>(o->k);
  Y5_LAUNCH_CHECK();
  return Y5OBB_OK;
}

void y5obb_wgrad_destroy(y5obb_wgrad_t* w) { delete reinterpret_cast<WgradObj*>(w); }

}  // extern "C"
