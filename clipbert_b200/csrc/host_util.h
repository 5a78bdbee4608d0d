// Host-side helpers shared by the C-ABI translation units: error reporting, launch counting,
// TMA tensor-map encoding (driver entry point fetched at run time, no link-time libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/clipbert_b200.h"

namespace cb {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return CB_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return CB_OK;
}

// 2D bf16 row-major tensor [rows, inner] (row pitch ld elements), 128B-swizzled boxes.
// Cached by value; returns nullptr (and sets the error) on failure.
const CUtensorMap* get_tmap_2d(const void* base, uint64_t inner, uint64_t rows, uint64_t ld,
                               uint32_t box_inner, uint32_t box_rows);

#define CB_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      cb::set_error(__VA_ARGS__);    \
      return CB_ERR_INVALID;         \
    }                                \
  } while (0)

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace cb
