// Fused SGD-Nesterov + weight decay + EMA update over all tensors of the model in ONE launch:
//   train.py:148-162,336-342   optimizer groups (BatchNorm weights / other weights with decay / biases), optimizer.step()
//   torch.optim.SGD(nesterov=True, dampening=0):  g' = g + wd * p;  buf = g' (first step) | momentum * buf + g';
//                                                 p -= lr * (g' + momentum * buf)
//   utils/torch_utils.py:304-314                  ema = d * ema + (1 - d) * p  for every floating-point state_dict entry
// tests/test_sgd_ema_gpu.py: equal to torch.optim.SGD + ModelEMA over three steps; train_step.TrainStep uses it whenever every
// batch ends with an optimizer step (it replaces about 60 multi-tensor launches, 0.6 ms per yolov5m step).
// HBM-bound: 16 B read + 12 B written per parameter element.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace y5obb {
namespace {

struct SgdEntryK {
  float* p;
  const float* g;
  float* mom;
  float* ema;
  long long first;  // global index of this tensor's first element
  long long n;
  float weight_decay;
  int group;        // 0..2: index into the per-step lr table; -1: EMA only (buffers)
};

struct SgdPlan {
  SgdEntryK* table = nullptr;
  int n = 0;
  long long total = 0;
};

struct SgdStep {
  float lr[3];
  float momentum, ema_decay;
  int first_step;
};

__global__ void k_sgd_ema(const SgdEntryK* __restrict__ table, int n, long long total, SgdStep s) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const SgdEntryK e = table[lo];
    const long long j = i - e.first;
    float p = e.p[j];
    if (e.group >= 0) {
      float g = e.g[j];
      if (e.weight_decay != 0.f) g = g + e.weight_decay * p;          // torch: grad.add(param, alpha=weight_decay)
      float buf = s.first_step ? g : s.momentum * e.mom[j] + g;        // buf.mul_(momentum).add_(grad)
      e.mom[j] = buf;
      g = g + s.momentum * buf;                                         // nesterov: grad.add(buf, alpha=momentum)
      p = p - s.lr[e.group] * g;                                        // param.add_(grad, alpha=-lr)
      e.p[j] = p;
    }
    if (e.ema) e.ema[j] = s.ema_decay * e.ema[j] + (1.0f - s.ema_decay) * p;  // v *= d; v += (1 - d) * p
  }
}

}  // namespace
}  // namespace y5obb

using namespace y5obb;

extern "C" {

int y5obb_sgd_ema_plan_create(const y5obb_sgd_entry* entries, int n, y5obb_sgd_plan_t** out) {
  if (!entries || n <= 0 || !out) return Y5OBB_EINVAL;
  std::vector<SgdEntryK> h((size_t)n);
  long long total = 0;
  for (int i = 0; i < n; ++i) {
    const y5obb_sgd_entry& e = entries[i];
    if (!e.p || e.n <= 0 || e.group < -1 || e.group > 2) return Y5OBB_EINVAL;
    if (e.group >= 0 && (!e.g || !e.mom)) return Y5OBB_EINVAL;
    SgdEntryK& k = h[(size_t)i];
    k.p = e.p;
    k.g = e.g;
    k.mom = e.mom;
    k.ema = e.ema;
    k.first = total;
    k.n = e.n;
    k.weight_decay = e.weight_decay;
    k.group = e.group;
    total += e.n;
  }
  SgdPlan* p = new SgdPlan();
  p->n = n;
  p->total = total;
  cudaError_t err = cudaMalloc(&p->table, sizeof(SgdEntryK) * (size_t)n);
  if (err == cudaSuccess) err = cudaMemcpy(p->table, h.data(), sizeof(SgdEntryK) * (size_t)n, cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    if (p->table) cudaFree(p->table);
    delete p;
    return cuda_fail(err);
  }
  *out = reinterpret_cast<y5obb_sgd_plan_t*>(p);
  return Y5OBB_OK;
}

int y5obb_sgd_ema_plan_run(const y5obb_sgd_plan_t* plan, const float* lr3, float momentum, float ema_decay, int first_step,
                           void* stream) {
  if (!plan || !lr3) return Y5OBB_EINVAL;
  const SgdPlan* p = reinterpret_cast<const SgdPlan*>(plan);
  SgdStep s;
  s.lr[0] = lr3[0];
  s.lr[1] = lr3[1];
  s.lr[2] = lr3[2];
  s.momentum = momentum;
  s.ema_decay = ema_decay;
  s.first_step = first_step;
  const int grid = (int)std::min<long long>((p->total + 255) / 256, (long long)sm_count() * 32);
  k_sgd_ema<<<grid, 256, 0, (cudaStream_t)stream>>>(p->table, p->n, p->total, s);
  Y5_LAUNCH_CHECK();
  return Y5OBB_OK;
}

void y5obb_sgd_ema_plan_destroy(y5obb_sgd_plan_t* plan) {
  SgdPlan* p = reinterpret_cast<SgdPlan*>(plan);
  if (!p) return;
  if (p->table) cudaFree(p->table);
  delete p;
}

}  // extern "C"
