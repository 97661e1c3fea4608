// Weight gradient of a convolution on the tensor cores (sm_100a): for every tap (kh, kw)
//   dW[tap][co][ci] = sum over (b, ho, wo) of dz[b, co, ho, wo] * x[b, ci, s*ho + kh - p, s*wo + kw - p]
// i.e. a GEMM with M = Cout, N = Cin and K = all output pixels.  Both operands are read from NCHW bf16 copies, in
// which the pixel axis (K) is contiguous, so one 4-D TMA box {kw_px, kh_px, channels, 1} lands as the K-major,
// swizzled [channels][64 pixels] tile tcgen05.mma consumes; the tap is a shift of the x box's origin (zero padding =
// TMA out-of-bounds fill, stride 2 = tensor-map element strides).  The pixel axis is split over CTAs (split-K); every
// CTA accumulates its slice in TMEM (fp32) and adds it into dW with fp32 atomics.
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

struct WgradK {
  CUtensorMap tmA;  // dz NCHW: (Wo, Ho, Cout, B)
  CUtensorMap tmB;  // x NCHW: stride 1: (Wi, Hi, Cin, B); stride 2: phase-split (Wi/2, Hi/2, 4 phases, Cin, B)
  int B, Ho, Wo, Cout, Cin;
  int KH, KW, stride, pad_h, pad_w;
  int kwp, khp, BK, BN;     // pixel tile kwp x khp = BK; N tile
  int tiles_w, tiles_h;     // pixel tiles per image
  int ksteps;               // B * tiles_h * tiles_w
  int ksplit, co_blks, ci_blks;
  int stages;
  uint32_t a_bytes, b_bytes, b_stage_bytes, idesc, tmem_cols;
  float* dW;                // [KH*KW][Cout][Cin] fp32
};

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradK p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = p.a_bytes + p.b_stage_bytes;
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
      for (int k = k0; k < k1; ++k) {
        const int b = k / per_img;
        const int r = k - b * per_img;
        const int th = r / p.tiles_w;
        const int ho0 = th * p.khp, wo0 = (r - th * p.tiles_w) * p.kwp;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        ptx::mbar_expect_tx(&full_bar[s], p.a_bytes + p.b_bytes);
        ptx::tma_load_4d(sa, &p.tmA, &full_bar[s], wo0, ho0, cob * 128, b);
        if (p.stride == 1) {
          ptx::tma_load_4d(sa + p.a_bytes, &p.tmB, &full_bar[s], wo0 + kw - p.pad_w, ho0 + kh - p.pad_h, cib * p.BN, b);
        } else {  // input pixel 2*o + (k - pad): phase (k - pad) & 1 of the de-interleaved copy, offset floor((k - pad) / 2)
          const int oh = kh - p.pad_h, ow = kw - p.pad_w;
          const int ah = oh & 1, aw = ow & 1;
          ptx::tma_load_5d(sa + p.a_bytes, &p.tmB, &full_bar[s], wo0 + ((ow - aw) >> 1), ho0 + ((oh - ah) >> 1),
                           ah * 2 + aw, cib * p.BN, b);
        }
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
      const uint32_t row_bytes = (uint32_t)p.BK * 2u;
      const uint64_t desc_hi = ptx::make_kmajor_desc(0u, row_bytes);
      const uint32_t ring = ptx::smem_u32(smem);
      uint32_t accumulate = 0u;
      const int nk = p.BK >> 4;
      for (int k = k0; k < k1; ++k) {
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t sa = ring + (uint32_t)s * stage_bytes;
        const uint64_t da = desc_hi | (uint64_t)((sa & 0x3FFFFu) >> 4);
        const uint64_t db = desc_hi | (uint64_t)(((sa + p.a_bytes) & 0x3FFFFu) >> 4);
        for (int j = 0; j < nk; ++j) {
          ptx::umma_bf16(tmem_base, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), p.idesc, accumulate);
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
    if (k1 > k0) {
      ptx::mbar_wait(&done_bar, 0u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      float* row = p.dW + ((long long)tap * p.Cout + co) * p.Cin + cib * p.BN;
      const int ncols = min(p.BN, p.Cin - cib * p.BN);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr + (uint32_t)c0, r);
        ptx::tmem_ld_wait();
        if (co < p.Cout) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < ncols) atomicAdd(row + c0 + j, __uint_as_float(r[j]));
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
  if (!d || !out || !d->dz_nchw || !d->x_nchw || !d->dw) return Y5OBB_EINVAL;
  if (d->stride != 1 && d->stride != 2) return Y5OBB_EINVAL;
  if (d->B < 1 || d->Cout < 1 || d->Cin < 1 || d->KH < 1 || d->KW < 1) return Y5OBB_EINVAL;
  // NCHW row strides must be multiples of 16 bytes (the stride-2 copy of x is de-interleaved: rows of Wi/2)
  if ((d->Wo & 7) || ((d->stride == 1 ? d->Wi : d->Wi / 2) & 7) || (d->stride == 2 && ((d->Wi | d->Hi) & 1))) return Y5OBB_EINVAL;
  if ((reinterpret_cast<uintptr_t>(d->dz_nchw) | reinterpret_cast<uintptr_t>(d->x_nchw)) & 15) return Y5OBB_EINVAL;
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
  // pixel tile: kwp x khp = BK output pixels (K of one pipeline stage)
  int kwp = 8;
  while (kwp < 64 && kwp < d->Wo) kwp <<= 1;
  int bk = 64;
  int khp = bk / kwp;
  while (khp > 1 && khp / 2 >= d->Ho && bk > 16) {  // tiny maps: shrink K per stage rather than multiply zeros
    khp >>= 1;
    bk >>= 1;
  }
  k.kwp = kwp;
  k.khp = khp;
  k.BK = bk;
  k.tiles_w = (d->Wo + kwp - 1) / kwp;
  k.tiles_h = (d->Ho + khp - 1) / khp;
  k.ksteps = d->B * k.tiles_w * k.tiles_h;
  k.ci_blks = (d->Cin + 255) / 256;
  k.BN = ((d->Cin + k.ci_blks - 1) / k.ci_blks + 15) / 16 * 16;
  k.co_blks = (d->Cout + 127) / 128;
  const int items = d->KH * d->KW * k.co_blks * k.ci_blks;
  k.ksplit = std::max(1, std::min(k.ksteps, (2 * sm_count() + items - 1) / items));
  k.a_bytes = (uint32_t)128 * bk * 2;
  k.b_bytes = (uint32_t)k.BN * bk * 2;
  k.b_stage_bytes = (uint32_t)align_up(k.b_bytes, 1024);
  k.stages = (int)std::min<size_t>(WG_MAX_STAGES, WG_SMEM / (k.a_bytes + k.b_stage_bytes));
  if (k.stages < 2) {
    delete o;
    return Y5OBB_EINVAL;
  }
  k.idesc = ptx::make_idesc_bf16(128, k.BN);
  k.tmem_cols = 32;
  while ((int)k.tmem_cols < k.BN) k.tmem_cols <<= 1;
  const CUtensorMapSwizzle sw =
      bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->Cout, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->Wo * 2, (cuuint64_t)d->Wo * d->Ho * 2, (cuuint64_t)d->Wo * d->Ho * d->Cout * 2};
    cuuint32_t box[4] = {(cuuint32_t)kwp, (cuuint32_t)khp, 128, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&k.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->dz_nchw), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      g_last_cuda_error = (int)r;
      delete o;
      return Y5OBB_ECUDA;
    }
  }
  {
    CUresult r;
    if (d->stride == 1) {
      cuuint64_t dims[4] = {(cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->Cin, (cuuint64_t)d->B};
      cuuint64_t strides[3] = {(cuuint64_t)d->Wi * 2, (cuuint64_t)d->Wi * d->Hi * 2,
                               (cuuint64_t)d->Wi * d->Hi * d->Cin * 2};
      cuuint32_t box[4] = {(cuuint32_t)kwp, (cuuint32_t)khp, (cuuint32_t)k.BN, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      r = enc(&k.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x_nchw), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {  // x copy is de-interleaved: [B][Cin][2*2 phases][Hi/2][Wi/2]
      const cuuint64_t w2 = d->Wi / 2, h2 = d->Hi / 2;
      cuuint64_t dims[5] = {w2, h2, 4, (cuuint64_t)d->Cin, (cuuint64_t)d->B};
      cuuint64_t strides[4] = {w2 * 2, w2 * h2 * 2, w2 * h2 * 4 * 2, w2 * h2 * 4 * (cuuint64_t)d->Cin * 2};
      cuuint32_t box[5] = {(cuuint32_t)kwp, (cuuint32_t)khp, 1, (cuuint32_t)k.BN, 1};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      r = enc(&k.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->x_nchw), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
      g_last_cuda_error = (int)r;
      delete o;
      return Y5OBB_ECUDA;
    }
  }
  o->grid = items * k.ksplit;
  o->smem = std::max<size_t>((size_t)k.stages * (k.a_bytes + k.b_stage_bytes) + 1024, 116 * 1024);
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
