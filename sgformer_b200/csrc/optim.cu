// Fused multi-tensor Adam step: replaces torch.optim.Adam over the reference's two parameter groups
// (large/main.py:115-119: params1 = TransConv with trans_weight_decay, params2 = GraphConv + fc with gnn_weight_decay).
// One launch updates up to SGF_ADAM_MAX_TENSORS parameter tensors (pointer table passed by value); the step counts live on the
// device (one per tensor, advanced by the launch itself) so that a CUDA-graph replay of the training step advances the bias corrections.
// Semantics = torch.optim.Adam(amsgrad=False, maximize=False): g += wd*p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
#include "common.cuh"
#include "launch_count.h"
#include "../../include/sgformer_b200.h"

namespace sgf {
constexpr int kAdamChunk = 2048;      // elements per block

__global__ void adam_tick_kernel(const __grid_constant__ sgf_adam_args a) {
    if ((int)threadIdx.x < a.n_tensors) *a.step[threadIdx.x] += 1.f;
}

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ sgf_adam_args a) {
    int ti = 0;
    while (ti + 1 < a.n_tensors && (int)blockIdx.x >= a.chunk0[ti + 1]) ++ti;
    const int64_t base = (int64_t)(blockIdx.x - a.chunk0[ti]) * kAdamChunk;
    const int64_t n = a.numel[ti];
    float* __restrict__ p = a.param[ti];
    const float* __restrict__ g = a.grad[ti];
    float* __restrict__ m = a.exp_avg[ti];
    float* __restrict__ v = a.exp_avg_sq[ti];
    const float b1 = a.beta1[ti], b2 = a.beta2[ti], lr = a.lr[ti], eps = a.eps[ti], wd = a.weight_decay[ti];
    const float t = *a.step[ti];
    const float bc1 = 1.f - powf(b1, t);
    const float bc2s = sqrtf(1.f - powf(b2, t));
    const float step_size = lr / bc1;
    for (int64_t i = base + threadIdx.x; i < base + kAdamChunk && i < n; i += 256) {
        const float pv = p[i];
        const float gv = g[i] + wd * pv;
        const float mv = b1 * m[i] + (1.f - b1) * gv;
        const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
        m[i] = mv;
        v[i] = vv;
        p[i] = pv - step_size * mv / (sqrtf(vv) / bc2s + eps);
    }
}
}  // namespace sgf

using namespace sgf;

extern "C" int sgf_adam_step(sgf_adam_args* a, void* stream) {
    if (!a || a->n_tensors <= 0 || a->n_tensors > SGF_ADAM_MAX_TENSORS) return SGF_ERR_ARG;
    int total = 0;
    for (int i = 0; i < a->n_tensors; ++i) {
        if (!a->param[i] || !a->grad[i] || !a->exp_avg[i] || !a->exp_avg_sq[i] || !a->step[i] || a->numel[i] < 0) return SGF_ERR_ARG;
        a->chunk0[i] = total;
        total += (int)((a->numel[i] + kAdamChunk - 1) / kAdamChunk);
    }
    a->chunk0[a->n_tensors] = total;
    // every tensor carries its own step count (torch.optim.Adam semantics: a parameter without a gradient is skipped and keeps
    // its count): advance the counts of this launch's tensors, then update with t = the new count
    adam_tick_kernel<<<1, SGF_ADAM_MAX_TENSORS, 0, (cudaStream_t)stream>>>(*a);
    SGF_LAUNCH_CHECK(); count_launch();
    if (total == 0) return SGF_OK;
    adam_kernel<<<total, 256, 0, (cudaStream_t)stream>>>(*a);
    SGF_LAUNCH_CHECK(); count_launch();
    return SGF_OK;
}
