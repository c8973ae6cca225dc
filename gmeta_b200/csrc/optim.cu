// Flat-buffer parameter updates: the inner SGD step on per-task fast weights
// (reference G-Meta/meta.py:126,151), the meta-gradient sum over tasks, and the outer Adam step
// with the reference's NaN gate (meta.py:97,161-169).
#include <math.h>

#include "common.cuh"

namespace gmeta {
namespace {

__global__ void sgd_update_kernel(const float* __restrict__ w_in, long long w_in_stride,
                                  const float* __restrict__ grad, float lr, int n_tasks, int n_params,
                                  float* __restrict__ w_out) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  const long long total = (long long)n_tasks * n_params;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / n_params);
    const int p = (int)(i - (long long)t * n_params);
    // p - lr * g, rounded as torch does it (mul, then sub): meta.py:126
    w_out[i] = w_in[t * w_in_stride + p] - __fmul_rn(lr, grad[i]);
  }
}

__global__ void sum_over_tasks_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                      int n_tasks, int n_params, float* __restrict__ out) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_params; p += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < n_tasks; ++t) {
      s += a[(size_t)t * n_params + p];
      if (b) s += b[(size_t)t * n_params + p];
    }
    out[p] = s;
  }
}

__global__ void adam_update_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                   float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                   int n_params, float step_size, float one_minus_beta1, float beta2,
                                   float one_minus_beta2, float eps, float bias_c2_sqrt, float grad_scale,
                                   const float* __restrict__ loss_gate, int32_t* __restrict__ skipped) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  const bool skip = loss_gate != nullptr && isnan(*loss_gate);
  if (skipped && blockIdx.x == 0 && threadIdx.x == 0) *skipped = skip ? 1 : 0;
  if (skip) return;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_params; p += gridDim.x * blockDim.x) {
    const float g = grad[p] * grad_scale;
    // torch/optim/adam.py single-tensor path: lerp, addcmul, sqrt/div/add eps, addcdiv
    const float m = exp_avg[p] + one_minus_beta1 * (g - exp_avg[p]);
    const float v = exp_avg_sq[p] * beta2 + one_minus_beta2 * g * g;
    exp_avg[p] = m;
    exp_avg_sq[p] = v;
    const float denom = sqrtf(v) / bias_c2_sqrt + eps;
    param[p] = param[p] - step_size * (m / denom);
  }
}

// Scalar prologue of the graph-safe Adam step: NaN gate, step counter and bias corrections live on the device,
// so the launch parameters never change and the step can be replayed from a CUDA graph.
// state: [0] step count (int), [1] skipped flag of this step (int), [2] step_size, [3] sqrt(bias_c2) (float bits)
__global__ void adam_prepare_kernel(int32_t* __restrict__ state, double lr, double beta1, double beta2,
                                    const float* __restrict__ loss_sum, float loss_scale,
                                    const float* __restrict__ acc_sums, int n_acc, float* __restrict__ step_out) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  const float gate = loss_sum ? *loss_sum * loss_scale : 0.f;      // meta.py:161: losses_q[-1] / task_num
  const bool skip = isnan(gate);                                    // meta.py:163-164
  if (threadIdx.x == 0) {
    int step = state[0];
    if (!skip) {
      ++step;
      state[0] = step;
      // torch/optim/adam.py: bias_correction in Python floats (double)
      const double bias_c1 = 1.0 - pow(beta1, (double)step);
      const double bias_c2 = 1.0 - pow(beta2, (double)step);
      state[2] = __float_as_int((float)(lr / bias_c1));
      state[3] = __float_as_int((float)sqrt(bias_c2));
    }
    state[1] = skip ? 1 : 0;
    if (step_out) {
      step_out[n_acc] = gate;
      step_out[n_acc + 1] = skip ? 1.f : 0.f;
    }
  }
  if (step_out && acc_sums)
    for (int k = threadIdx.x; k < n_acc; k += blockDim.x) step_out[k] = acc_sums[k] * loss_scale;   // meta.py:171
}

__global__ void adam_apply_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                  float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, int n_params,
                                  float one_minus_beta1, float beta2, float one_minus_beta2, float eps, float grad_scale,
                                  const int32_t* __restrict__ state) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  if (state[1]) return;
  const float step_size = __int_as_float(state[2]), bias_c2_sqrt = __int_as_float(state[3]);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_params; p += gridDim.x * blockDim.x) {
    const float g = grad[p] * grad_scale;
    const float m = exp_avg[p] + one_minus_beta1 * (g - exp_avg[p]);
    const float v = exp_avg_sq[p] * beta2 + one_minus_beta2 * g * g;
    exp_avg[p] = m;
    exp_avg_sq[p] = v;
    const float denom = sqrtf(v) / bias_c2_sqrt + eps;
    param[p] = param[p] - step_size * (m / denom);
  }
}

}  // namespace
}  // namespace gmeta

using namespace gmeta;

extern "C" int gmeta_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t n_params,
                               double lr, double beta1, double beta2, double eps, int32_t* state, float grad_scale,
                               const float* loss_sum, float loss_scale, const float* acc_sums, int32_t n_acc,
                               float* step_out, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !state || n_params <= 0 || n_acc < 0) return GMETA_ERR_BAD_ARG;
  launch_pdl(adam_prepare_kernel, dim3(1), dim3(64), 0, (cudaStream_t)stream, state, lr, beta1, beta2, loss_sum, loss_scale, acc_sums,
                                                        n_acc, step_out);
  int rc = check_launch();
  if (rc != GMETA_OK) return rc;
  const int grid = ceil_div(n_params, 256) < 8 * kNumSMs ? ceil_div(n_params, 256) : 8 * kNumSMs;
  launch_pdl(adam_apply_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n_params,
                                                           (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                           (float)eps, grad_scale, state);
  return check_launch();
}

extern "C" int gmeta_sgd_update(const float* w_in, int64_t w_in_task_stride, const float* grad, float lr,
                                int32_t n_tasks, int32_t n_params, float* w_out, void* stream) {
  if (!w_in || !grad || !w_out || n_tasks <= 0 || n_params <= 0) return GMETA_ERR_BAD_ARG;
  const long long total = (long long)n_tasks * n_params;
  const int grid = (int)((total + 255) / 256 < 8 * kNumSMs ? (total + 255) / 256 : 8 * kNumSMs);
  launch_pdl(sgd_update_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, w_in, w_in_task_stride, grad, lr, n_tasks,
                                                           n_params, w_out);
  return check_launch();
}

extern "C" int gmeta_sum_over_tasks(const float* a, const float* b, int32_t n_tasks, int32_t n_params,
                                    float* out, void* stream) {
  if (!a || !out || n_tasks <= 0 || n_params <= 0) return GMETA_ERR_BAD_ARG;
  const int grid = ceil_div(n_params, 256) < 8 * kNumSMs ? ceil_div(n_params, 256) : 8 * kNumSMs;
  launch_pdl(sum_over_tasks_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a, b, n_tasks, n_params, out);
  return check_launch();
}

extern "C" int gmeta_adam_update(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                 int32_t n_params, double lr, double beta1, double beta2, double eps,
                                 int32_t step, float grad_scale, const float* loss_gate, int32_t* skipped,
                                 void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n_params <= 0 || step <= 0) return GMETA_ERR_BAD_ARG;
  // scalar prologue in double, exactly like the Python floats of torch/optim/adam.py
  const double bias_c1 = 1.0 - pow(beta1, (double)step);
  const double bias_c2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bias_c1);
  const int grid = ceil_div(n_params, 256) < 8 * kNumSMs ? ceil_div(n_params, 256) : 8 * kNumSMs;
  launch_pdl(adam_update_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n_params,
                                                            step_size, (float)(1.0 - beta1), (float)beta2,
                                                            (float)(1.0 - beta2), (float)eps,
                                                            (float)sqrt(bias_c2), grad_scale, loss_gate,
                                                            skipped);
  return check_launch();
}
