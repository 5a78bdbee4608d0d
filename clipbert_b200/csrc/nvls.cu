// Hand-written all-reduce of a flat fp32 gradient buffer through the NVSwitch (NVLink SHARP, "NVLS") - the B200-native form of
// the data-parallel exchange that Horovod + NCCL perform for the reference (src/tasks/run_video_retrieval.py:299-301,432).
//
// Every rank allocates the buffer as symmetric memory and maps it at a MULTICAST address (torch.distributed._symmetric_memory:
// rendezvous(...).multicast_ptr). Rank r owns the r-th 1/world slice of the element range:
//     multimem.ld_reduce.add.f32  pulls the slice through the switch, which sums the copies of ALL ranks in flight,
//     (x scale: 1/world for the average),
//     multimem.st                 writes the result back through the switch into EVERY rank's copy.
// Per rank that is n/world elements pulled and n/world pushed - the reduce-scatter + all-gather traffic of a ring, without the
// ring: no per-hop latency, no staging buffers, and only as many CTAs as the caller allows (NCCL's all-reduce kernel pins its
// CTAs for the whole exchange, which with this library's persistent one-CTA-per-SM GEMMs costs a second wave).
// The caller brackets the launch with cross-rank barriers (all gradients written before; all slices stored after).
#include "common.cuh"
#include "host_util.h"

namespace cb {

__global__ void __launch_bounds__(256) nvls_allreduce_f32_kernel(float* __restrict__ mc, int64_t v4_begin, int64_t v4_end, float scale) {
  pdl_wait();
  pdl_trigger();
  for (int64_t i = v4_begin + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < v4_end;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float* p = mc + i * 4;
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
}

}  // namespace cb

using namespace cb;

extern "C" int cb_nvls_allreduce_f32(void* multicast_ptr, int64_t n, int rank, int world, float scale, int max_ctas, void* stream) {
  CB_REQUIRE(multicast_ptr != nullptr, "cb_nvls_allreduce_f32: null multicast pointer (no NVLS mapping)");
  CB_REQUIRE((reinterpret_cast<uintptr_t>(multicast_ptr) & 15) == 0 && n > 0 && n % 4 == 0,
             "cb_nvls_allreduce_f32: pointer must be 16-byte aligned and n a positive multiple of 4 (got %lld)", static_cast<long long>(n));
  CB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "cb_nvls_allreduce_f32: bad rank %d / world %d", rank, world);
  const int64_t v4 = n / 4;
  const int64_t per = (v4 + world - 1) / world;
  const int64_t begin = rank * per < v4 ? rank * per : v4;
  const int64_t end = begin + per < v4 ? begin + per : v4;
  if (end <= begin) return CB_OK;                      // nothing in this rank's slice (tiny buffers)
  int grid = ceil_div(end - begin, 256);
  const int cap = max_ctas > 0 ? max_ctas : 32;
  if (grid > cap) grid = cap;
  launch_k(nvls_allreduce_f32_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), static_cast<float*>(multicast_ptr), begin, end, scale);
  return check_launch("cb_nvls_allreduce_f32");
}
