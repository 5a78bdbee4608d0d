// Clip-level score aggregation + loss of the training loops, pool_method "lse"
// (src/tasks/run_video_retrieval.py:404-422, src/tasks/run_video_qa.py:484-501):
//     logits = stack(per-clip logits).permute(1, 0, 2)                        (B', n_clips, C)
//     out    = logsumexp(logits.view(B', -1), -1, keepdim) - logsumexp(logits, dim=1)   (B', C)
//     loss   = gather(out, -1, labels).mean()
// and its backward, in ONE launch: the reference spends ~45 ATen launches (two logsumexp, gather, mean and their autograd
// nodes) on a (n_clips, B', C) fp32 tensor of a few hundred values; inside the step's CUDA graph those are ~45 dependent
// nodes of 2-4 us each on the critical path between the last forward kernel and the first backward kernel.
//     d loss / d z[k, b, c] = (1 / B') * ( softmax over all (k', c') of example b  -  [c == y_b] * softmax over k' of z[:, b, y_b] )
#include "common.cuh"
#include "host_util.h"

namespace cb {

__global__ void __launch_bounds__(256) clip_lse_loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                            float* __restrict__ loss, float* __restrict__ dlogits, int n_clips,
                                                            int nseq, int ncls, float inv_n, float grad_scale) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (b < nseq) {
    const int64_t y64 = labels[b];
    const int y = static_cast<int>(y64 < 0 ? 0 : (y64 >= ncls ? ncls - 1 : y64));   // torch.gather would raise; stay in bounds
    const int64_t clip_pitch = static_cast<int64_t>(nseq) * ncls;
    const float* z = logits + static_cast<int64_t>(b) * ncls;
    float m_all = -INFINITY, m_y = -INFINITY;
    for (int k = 0; k < n_clips; ++k) {
      for (int c = 0; c < ncls; ++c) m_all = fmaxf(m_all, z[k * clip_pitch + c]);
      m_y = fmaxf(m_y, z[k * clip_pitch + y]);
    }
    float s_all = 0.f, s_y = 0.f;
    for (int k = 0; k < n_clips; ++k) {
      for (int c = 0; c < ncls; ++c) s_all += expf(z[k * clip_pitch + c] - m_all);
      s_y += expf(z[k * clip_pitch + y] - m_y);
    }
    const float lse_all = m_all + logf(s_all), lse_y = m_y + logf(s_y);
    l = lse_all - lse_y;
    if (dlogits != nullptr) {
      float* d = dlogits + static_cast<int64_t>(b) * ncls;
      const float gs = inv_n * grad_scale;
      for (int k = 0; k < n_clips; ++k)
        for (int c = 0; c < ncls; ++c) {
          const float v = z[k * clip_pitch + c];
          float g = expf(v - lse_all);
          if (c == y) g -= expf(v - lse_y);
          d[k * clip_pitch + c] = g * gs;
        }
    }
  }
  // block sum -> one atomic per block
  l = warp_sum(l);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss, v * inv_n);
  }
}

}  // namespace cb

using namespace cb;

extern "C" int cb_clip_lse_loss(const float* logits, const int64_t* labels, float* loss, float* dlogits, int n_clips, int nseq,
                                int ncls, float grad_scale, void* stream) {
  CB_REQUIRE(logits && labels && loss, "cb_clip_lse_loss: null pointer");
  CB_REQUIRE(n_clips > 0 && nseq > 0 && ncls > 0, "cb_clip_lse_loss: empty problem (n_clips=%d nseq=%d ncls=%d)", n_clips, nseq, ncls);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (e != cudaSuccess) {
    set_error("cb_clip_lse_loss: memset failed: %s", cudaGetErrorString(e));
    return CB_ERR_CUDA;
  }
  launch_k(clip_lse_loss_kernel, ceil_div(nseq, 256), 256, 0, st, logits, labels, loss, dlogits, n_clips, nseq, ncls, 1.0f / nseq,
           grad_scale);
  return check_launch("cb_clip_lse_loss");
}
