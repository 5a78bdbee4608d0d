// tcgen05 tensor-core contraction for every dense op on the ClipBERT path (see cb_gemm in
// include/clipbert_b200.h). One warp-specialised kernel, two operand modes:
//   MODE 0 (TN)    : A [rows, K] and B [N, K] both K-major; optional 9-tap row-shifted K loop
//                    (3x3 conv over a zero-bordered NHWC activation) ; fused epilogue.
//   MODE 1 (WGRAD) : dW = dY^T X with both operands read MN-major from the activation layout,
//                    split over the pixel/token dimension, fp32 red.global accumulation.
// Pipeline: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
// (TMEM -> registers -> global). smem ring of STAGES x (A 128x64 | B BNx64) bf16 tiles, 128B swizzle.
#include "common.cuh"
#include "host_util.h"

namespace cb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;

struct GemmEpi {
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int64_t res_ld;
  const __nv_bfloat16* aux;
  int64_t aux_ld;
  int aux_mode;
  int act;
  void* out;
  int64_t out_ld;
  int out_fp32;
  __nv_bfloat16* out2;
  int64_t out2_ld;
  int rowmap, H, W;
  uint32_t drop_thresh;
  float drop_inv_keep;
  uint64_t seed;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 64) ? 4 : (BN == 128 ? 3 : 4);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

template <int BN, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                int M, int N, int K, int ntaps, int tap_w, int tap_sign, int iters_per_split,
                GemmEpi epi) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;

  // ---- per-CTA iteration space ----
  int tap = 0, it_begin = 0, it_end = 0;
  if (MODE == 0 || MODE == 2) {
    const int kc = (K + BK - 1) / BK;
    it_begin = 0;
    it_end = ntaps * kc;
  } else {
    tap = blockIdx.z % ntaps;
    const int split = blockIdx.z / ntaps;
    const int total = (K + BK - 1) / BK;
    it_begin = split * iters_per_split;
    it_end = min(total, it_begin + iters_per_split);
  }
  const int n_iters = it_end - it_begin;  // may be <= 0 for a trailing empty split

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (n_iters > 0) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        const int kc_per_tap = (K + BK - 1) / BK;
        for (int i = 0; i < n_iters; ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          const int it = it_begin + i;
          if (MODE == 0) {
            const int t = it / kc_per_tap;
            const int kc = it - t * kc_per_tap;
            int shift = 0;
            if (ntaps == 9) shift = tap_sign * ((t / 3 - 1) * tap_w + (t % 3 - 1));
            tma_load_2d(sa, &tmA, &full_bar[s], kc * BK, m0 + shift);
            tma_load_2d(sb, &tmB, &full_bar[s], t * K + kc * BK, n0);
          } else if (MODE == 2) {
            const int t = it / kc_per_tap;
            const int kc = it - t * kc_per_tap;
            int shift = 0;
            if (ntaps == 9) shift = tap_sign * ((t / 3 - 1) * tap_w + (t % 3 - 1));
            tma_load_2d(sa, &tmA, &full_bar[s], kc * BK, m0 + shift);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (BK * 128), &tmB, &full_bar[s], t * N + n0 + j * 64, kc * BK);
          } else {
            int shift = 0;
            if (ntaps == 9) shift = tap_sign * ((tap / 3 - 1) * tap_w + (tap % 3 - 1));
            const int p = it * BK;
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sa + j * (BK * 128), &tmA, &full_bar[s], m0 + j * 64, p);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (BK * 128), &tmB, &full_bar[s], n0 + j * 64, p + shift);
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, MODE == 1, MODE != 0);
        for (int i = 0; i < n_iters; ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t ad, bd;
            if (MODE == 0) {
              ad = umma_smem_desc(a_addr + k * 32, 16, 1024);
              bd = umma_smem_desc(b_addr + k * 32, 16, 1024);
            } else if (MODE == 2) {
              ad = umma_smem_desc(a_addr + k * 32, 16, 1024);
              bd = umma_smem_desc(b_addr + k * 2048, BK * 128, 1024);
            } else {
              ad = umma_smem_desc(a_addr + k * 2048, BK * 128, 1024);
              bd = umma_smem_desc(b_addr + k * 2048, BK * 128, 1024);
            }
            umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs have read it
        }
        umma_commit(accum_bar);  // accumulator complete
      }
    } else {
      // ===================== epilogue warps =====================
      const int q = warp & 3;  // TMEM lane quarter this warp may access
      const int m = m0 + q * 32 + lane;
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

      if (MODE == 1) {
        const bool row_ok = m < M;
        const float rs = (row_ok && epi.scale) ? epi.scale[m] : 1.0f;
        float* orow = reinterpret_cast<float*>(epi.out) + static_cast<int64_t>(m) * epi.out_ld +
                      static_cast<int64_t>(tap) * N;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(trow + c, v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int n = n0 + c + j;
              if (n + 4 <= N) {
                red_add_f32x4(orow + n, __uint_as_float(v[j]) * rs, __uint_as_float(v[j + 1]) * rs,
                              __uint_as_float(v[j + 2]) * rs, __uint_as_float(v[j + 3]) * rs);
              }
            }
          }
        }
      } else {
        bool row_ok = m < M;
        int64_t orow = m;
        if (epi.rowmap == CB_ROWMAP_PAD) {
          const int hw = epi.H * epi.W;
          const int img = m / hw;
          const int r = m - img * hw;
          const int y = r / epi.W, x = r - y * epi.W;
          orow = (static_cast<int64_t>(img) * (epi.H + 2) + y + 1) * (epi.W + 2) + x + 1;
        } else if (epi.rowmap == CB_ROWMAP_UNPAD) {
          const int wp = epi.W + 2, hp = epi.H + 2;
          const int img = m / (hp * wp);
          const int r = m - img * (hp * wp);
          const int y = r / wp, x = r - y * wp;
          row_ok = row_ok && y >= 1 && y <= epi.H && x >= 1 && x <= epi.W;
          orow = (static_cast<int64_t>(img) * epi.H + (y - 1)) * epi.W + (x - 1);
        }
        const __nv_bfloat16* res_row = epi.residual ? epi.residual + static_cast<int64_t>(m) * epi.res_ld : nullptr;
        const __nv_bfloat16* aux_row = epi.aux ? epi.aux + static_cast<int64_t>(m) * epi.aux_ld : nullptr;
        __nv_bfloat16* out2_row = epi.out2 ? epi.out2 + orow * epi.out2_ld : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(trow + c, v);
          tmem_ld_wait();
          if (!row_ok) continue;
#pragma unroll
          for (int g = 0; g < 32; g += 8) {
            const int n = n0 + c + g;
            if (n + 8 > N) continue;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g + j]);
            if (epi.scale) {
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(epi.scale + n));
              const float4 s1 = __ldg(reinterpret_cast<const float4*>(epi.scale + n + 4));
              f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
              f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
            }
            if (epi.shift) {
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(epi.shift + n));
              const float4 s1 = __ldg(reinterpret_cast<const float4*>(epi.shift + n + 4));
              f[0] += s0.x; f[1] += s0.y; f[2] += s0.z; f[3] += s0.w;
              f[4] += s1.x; f[5] += s1.y; f[6] += s1.z; f[7] += s1.w;
            }
            if (epi.drop_thresh) {
              const uint64_t base = static_cast<uint64_t>(orow) * static_cast<uint64_t>(N) + n;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                f[j] *= dropout_mult(epi.seed, base + j, epi.drop_thresh, epi.drop_inv_keep);
            }
            if (res_row) {
              const uint4 r = *reinterpret_cast<const uint4*>(res_row + n);
              const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), cc = unpack_bf16x2(r.z),
                           d = unpack_bf16x2(r.w);
              f[0] += a.x; f[1] += a.y; f[2] += b.x; f[3] += b.y;
              f[4] += cc.x; f[5] += cc.y; f[6] += d.x; f[7] += d.y;
            }
            if (out2_row) {
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
              o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
              *reinterpret_cast<uint4*>(out2_row + n) = o;
            }
            if (epi.act == CB_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.0f);
            } else if (epi.act == CB_ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
            } else if (epi.act == CB_ACT_TANH) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = tanhf(f[j]);
            }
            if (aux_row) {
              const uint4 r = *reinterpret_cast<const uint4*>(aux_row + n);
              float a[8];
              float2 t;
              t = unpack_bf16x2(r.x); a[0] = t.x; a[1] = t.y;
              t = unpack_bf16x2(r.y); a[2] = t.x; a[3] = t.y;
              t = unpack_bf16x2(r.z); a[4] = t.x; a[5] = t.y;
              t = unpack_bf16x2(r.w); a[6] = t.x; a[7] = t.y;
              if (epi.aux_mode == CB_AUX_RELU_MASK) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = a[j] > 0.0f ? f[j] : 0.0f;
              } else if (epi.aux_mode == CB_AUX_GELU_GRAD) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] *= gelu_erf_grad(a[j]);
              } else if (epi.aux_mode == CB_AUX_TANH_GRAD) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] *= (1.0f - a[j] * a[j]);
              }
            }
            if (epi.out_fp32) {
              float* o = reinterpret_cast<float*>(epi.out) + orow * epi.out_ld + n;
              *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
            } else {
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
              o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(epi.out) + orow * epi.out_ld + n) = o;
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int MODE>
static int launch_gemm(const cb_gemm_desc& d, const GemmEpi& epi, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  auto kern = gemm_kernel<BN, MODE>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return CB_ERR_CUDA;
    }
    attr_set = true;
  }
  const CUtensorMap *ta, *tb;
  dim3 grid;
  int iters_per_split = 0;
  if (MODE == 0) {
    ta = get_tmap_2d(d.a, d.k, d.a_rows, d.a_ld, BK, BM);
    tb = get_tmap_2d(d.b, static_cast<uint64_t>(d.k) * d.ntaps, d.b_rows, d.b_ld, BK, BN);
    grid = dim3(ceil_div(d.n, BN), ceil_div(d.m, BM), 1);
  } else if (MODE == 2) {
    ta = get_tmap_2d(d.a, d.k, d.a_rows, d.a_ld, BK, BM);
    tb = get_tmap_2d(d.b, static_cast<uint64_t>(d.n) * d.ntaps, d.b_rows, d.b_ld, 64, BK);
    grid = dim3(ceil_div(d.n, BN), ceil_div(d.m, BM), 1);
  } else {
    ta = get_tmap_2d(d.a, d.m, d.a_rows, d.a_ld, 64, BK);
    tb = get_tmap_2d(d.b, d.n, d.b_rows, d.b_ld, 64, BK);
    const int total = ceil_div(d.k, BK);
    int splits = d.split_k < 1 ? 1 : d.split_k;
    if (splits > total) splits = total;
    iters_per_split = ceil_div(total, splits);
    splits = ceil_div(total, iters_per_split);
    grid = dim3(ceil_div(d.n, BN), ceil_div(d.m, BM), splits * d.ntaps);
  }
  if (!ta || !tb) return CB_ERR_CUDA;
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(*ta, *tb, d.m, d.n, d.k, d.ntaps, d.tap_w,
                                                         d.tap_sign, iters_per_split, epi);
  return check_launch("cb_gemm");
}

}  // namespace cb

extern "C" int cb_gemm(const cb_gemm_desc* dp, void* stream_v) {
  using namespace cb;
  CB_REQUIRE(dp != nullptr, "cb_gemm: null descriptor");
  const cb_gemm_desc& d = *dp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  CB_REQUIRE(d.a && d.b && d.out, "cb_gemm: null operand pointer");
  CB_REQUIRE(d.m > 0 && d.n > 0 && d.k > 0, "cb_gemm: empty problem m=%d n=%d k=%d", d.m, d.n, d.k);
  CB_REQUIRE(d.ntaps == 1 || d.ntaps == 9, "cb_gemm: ntaps must be 1 or 9 (got %d)", d.ntaps);
  CB_REQUIRE(d.mode == CB_GEMM_TN || d.mode == CB_GEMM_WGRAD || d.mode == CB_GEMM_NN, "cb_gemm: bad mode %d", d.mode);
  CB_REQUIRE(d.dropout_p >= 0.0f && d.dropout_p < 1.0f, "cb_gemm: dropout_p out of range");

  GemmEpi epi;
  epi.scale = d.scale;
  epi.shift = d.shift;
  epi.residual = static_cast<const __nv_bfloat16*>(d.residual);
  epi.res_ld = d.res_ld;
  epi.aux = static_cast<const __nv_bfloat16*>(d.aux);
  epi.aux_ld = d.aux_ld;
  epi.aux_mode = d.aux ? d.aux_mode : CB_AUX_NONE;
  epi.act = d.act;
  epi.out = d.out;
  epi.out_ld = d.out_ld;
  epi.out_fp32 = d.out_fp32;
  epi.out2 = static_cast<__nv_bfloat16*>(d.out2);
  epi.out2_ld = d.out2_ld;
  epi.rowmap = d.rowmap;
  epi.H = d.map_h;
  epi.W = d.map_w;
  epi.seed = d.dropout_seed;
  if (d.dropout_p > 0.0f) {
    double t = static_cast<double>(d.dropout_p) * 4294967296.0;
    epi.drop_thresh = t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
    if (epi.drop_thresh == 0) epi.drop_thresh = 1;
    epi.drop_inv_keep = 1.0f / (1.0f - d.dropout_p);
  } else {
    epi.drop_thresh = 0;
    epi.drop_inv_keep = 1.0f;
  }

  if (d.mode == CB_GEMM_TN || d.mode == CB_GEMM_NN) {
    const bool nn = d.mode == CB_GEMM_NN;
    CB_REQUIRE(d.n % 8 == 0, "cb_gemm(TN): n must be a multiple of 8 (got %d)", d.n);
    CB_REQUIRE(d.k % 8 == 0, "cb_gemm(TN): k must be a multiple of 8 (got %d)", d.k);
    CB_REQUIRE(d.out_ld % 8 == 0, "cb_gemm(TN): out_ld must be a multiple of 8");
    CB_REQUIRE(!d.residual || d.res_ld % 8 == 0, "cb_gemm(TN): res_ld must be a multiple of 8");
    CB_REQUIRE(!d.aux || d.aux_ld % 8 == 0, "cb_gemm(TN): aux_ld must be a multiple of 8");
    CB_REQUIRE(!d.out2 || d.out2_ld % 8 == 0, "cb_gemm(TN): out2_ld must be a multiple of 8");
    CB_REQUIRE(d.rowmap == CB_ROWMAP_NONE || (d.map_h > 0 && d.map_w > 0), "cb_gemm: rowmap needs map_h/map_w");
    CB_REQUIRE(d.ntaps == 1 || d.tap_w > 2, "cb_gemm: 9-tap mode needs tap_w = W + 2");
    int bn = d.block_n;
    if (bn == 0) {
      // fill the 148 SMs: prefer the widest tile that still yields >= ~1 wave of CTAs
      const int64_t mt = ceil_div(d.m, BM);
      if (d.n >= 256 && mt * ceil_div(d.n, 256) >= 148) bn = 256;
      else if (d.n >= 128 && mt * ceil_div(d.n, 128) >= 120) bn = 128;
      else bn = (d.n >= 128 && mt * ceil_div(d.n, 64) > 2 * 296) ? 128 : 64;
    }
    switch (bn) {
      case 64: return nn ? launch_gemm<64, 2>(d, epi, stream) : launch_gemm<64, 0>(d, epi, stream);
      case 128: return nn ? launch_gemm<128, 2>(d, epi, stream) : launch_gemm<128, 0>(d, epi, stream);
      case 256: return nn ? launch_gemm<256, 2>(d, epi, stream) : launch_gemm<256, 0>(d, epi, stream);
      default: CB_REQUIRE(false, "cb_gemm: block_n must be 0, 64, 128 or 256 (got %d)", bn);
    }
  } else {
    CB_REQUIRE(d.out_fp32 == 1, "cb_gemm(WGRAD): output must be fp32");
    CB_REQUIRE(d.m % 8 == 0 && d.n % 8 == 0, "cb_gemm(WGRAD): m, n must be multiples of 8 (got %d, %d)", d.m, d.n);
    CB_REQUIRE(d.out_ld % 4 == 0, "cb_gemm(WGRAD): out_ld must be a multiple of 4");
    CB_REQUIRE((reinterpret_cast<uintptr_t>(d.out) & 15) == 0, "cb_gemm(WGRAD): out must be 16-byte aligned");
    int bn = d.block_n;
    if (bn == 0) bn = (d.n >= 128) ? 128 : 64;
    switch (bn) {
      case 64: return launch_gemm<64, 1>(d, epi, stream);
      case 128: return launch_gemm<128, 1>(d, epi, stream);
      default: CB_REQUIRE(false, "cb_gemm(WGRAD): block_n must be 0, 64 or 128 (got %d)", bn);
    }
  }
  return CB_ERR_INVALID;
}
