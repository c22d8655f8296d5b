// Warm-up scoring kernels (HBM-bound): elementwise fp32 accumulation of gradients, per-block
// signed-sum accumulation, block-score reduction with the four reference strategies, and the
// activation (channel) scoring pair.
//
// Reference call sites replaced (paths relative to the reference root):
//   deepspeed/fine_tune.py:716-768   D2H copy + CPU `+=` of every q/k/v(/MLP) gradient
//   deepspeed/smt/smt_helper.py:55-78, 233-251   reshape to [R/b, b, C/b, b] and reduce dims (1,3)
//   deepspeed/fine_tune.py:649-678   activation hook (|x|, allreduce, D2H, `+=`)
//   deepspeed/smt/smt_helper.py:168-183   channel scores
#include "common.cuh"

namespace smt {
namespace {

constexpr int kThreads = 256;

// ---- acc += grad ---------------------------------------------------------------------------

#ifndef SMT_SCORE_UNROLL
#define SMT_SCORE_UNROLL 2     // vec8 groups in flight per thread (measured best of {1, 2, 4}: profiles/r01_kernels.md)
#endif
#ifndef SMT_SCORE_CTAS
#define SMT_SCORE_CTAS 4       // grid cap in CTAs per SM for the streaming kernels (2 x 4 measured best: 6 125 GB/s)
#endif

template <int DT>
__device__ __forceinline__ void load_grad8(const void* __restrict__ grad, int64_t i, float (&g)[8]) {
  if (DT == SMT_F32) {
    const float4 a = ld_stream_f4(reinterpret_cast<const float4*>(grad) + 2 * i);
    const float4 b = ld_stream_f4(reinterpret_cast<const float4*>(grad) + 2 * i + 1);
    g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w;
    g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
  } else {
    unpack8<DT>(ld_stream_u4(reinterpret_cast<const uint4*>(grad) + i), g);
  }
}

template <int DT>
__global__ void __launch_bounds__(kThreads) score_accumulate_vec_kernel(
    float* __restrict__ acc, const void* __restrict__ grad, int64_t n_vec8) {
  // one "vec8" = 8 consecutive elements: 32 B of fp32 accumulator, 16 B (16-bit) or 32 B (fp32) of grad
  constexpr int U = SMT_SCORE_UNROLL;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n_vec8; i += U * stride) {
    float g[U][8];
    float4 a0[U], a1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {                       // all loads first: U x 48 B in flight per thread
      load_grad8<DT>(grad, i + u * stride, g[u]);
      const float4* ap = reinterpret_cast<const float4*>(acc) + 2 * (i + u * stride);
      a0[u] = ap[0];
      a1[u] = ap[1];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4* ap = reinterpret_cast<float4*>(acc) + 2 * (i + u * stride);
      a0[u].x += g[u][0]; a0[u].y += g[u][1]; a0[u].z += g[u][2]; a0[u].w += g[u][3];
      a1[u].x += g[u][4]; a1[u].y += g[u][5]; a1[u].z += g[u][6]; a1[u].w += g[u][7];
      ap[0] = a0[u];
      ap[1] = a1[u];
    }
  }
  for (; i < n_vec8; i += stride) {                     // remainder
    float g[8];
    load_grad8<DT>(grad, i, g);
    float4* ap = reinterpret_cast<float4*>(acc) + 2 * i;
    float4 a0 = ap[0], a1 = ap[1];
    a0.x += g[0]; a0.y += g[1]; a0.z += g[2]; a0.w += g[3];
    a1.x += g[4]; a1.y += g[5]; a1.z += g[6]; a1.w += g[7];
    ap[0] = a0;
    ap[1] = a1;
  }
}

template <int DT>
__global__ void __launch_bounds__(kThreads) score_accumulate_scalar_kernel(
    float* __restrict__ acc, const void* __restrict__ grad, int64_t begin, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    acc[i] += load_as_float<DT>(grad, i);
}

// ---- per-block reductions --------------------------------------------------------------------
// One CTA per b x b block.  A row of the block is contiguous (b*sizeof(T) bytes), so a group of
// b/VEC threads reads one row with 128-bit loads and the CTA walks down the rows.

enum { kSigned = 0, kAbs = 1, kSquare = 2 };

template <int MODE>
__device__ __forceinline__ float term(float x) {
  if (MODE == kSigned) return x;
  if (MODE == kAbs) return fabsf(x);
  return x * x;
}

template <int B, int DT, int MODE>
__device__ __forceinline__ float block_partial(const void* __restrict__ src, int64_t ld, int brow,
                                               int bcol) {
  constexpr int VEC = (DT == SMT_F32) ? 4 : 8;          // elements per 128-bit load
  constexpr int TPR = B / VEC;                          // threads per row
  constexpr int RPP = kThreads / TPR;                   // rows per pass
  constexpr int PASSES = B / RPP;
  constexpr int UNROLL = PASSES >= 4 ? 4 : PASSES;
  static_assert(kThreads % TPR == 0 && B % RPP == 0 && PASSES % UNROLL == 0, "tiling");
  const int tcol = threadIdx.x % TPR, trow = threadIdx.x / TPR;
  const char* base = reinterpret_cast<const char*>(src) +
                     ((int64_t)(brow * B + trow) * ld + (int64_t)bcol * B + tcol * VEC) *
                         (DT == SMT_F32 ? 4 : 2);
  const int64_t pass_stride = (int64_t)RPP * ld * (DT == SMT_F32 ? 4 : 2);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
  for (int p = 0; p < PASSES; p += UNROLL) {
    uint4 u[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) u[j] = ld_stream_u4(base + (int64_t)(p + j) * pass_stride);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      if (DT == SMT_F32) {
        s0 += term<MODE>(__uint_as_float(u[j].x)) + term<MODE>(__uint_as_float(u[j].y));
        s1 += term<MODE>(__uint_as_float(u[j].z)) + term<MODE>(__uint_as_float(u[j].w));
      } else {
        float f[8];
        unpack8<DT>(u[j], f);
        s0 += (term<MODE>(f[0]) + term<MODE>(f[1])) + (term<MODE>(f[2]) + term<MODE>(f[3]));
        s1 += (term<MODE>(f[4]) + term<MODE>(f[5])) + (term<MODE>(f[6]) + term<MODE>(f[7]));
      }
    }
  }
  return s0 + s1;
}

template <int B, int MODE>
__global__ void __launch_bounds__(kThreads) block_score_reduce_kernel(
    const float* __restrict__ acc, int64_t ld, int strategy, float* __restrict__ scores) {
  __shared__ float scratch[32];
  const int bcol = blockIdx.x, brow = blockIdx.y;
  float s = block_partial<B, SMT_F32, MODE>(acc, ld, brow, bcol);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    const float cnt = (float)(B * B);
    float r;
    if (strategy == SMT_MEAN_ABS) r = fabsf(s / cnt);
    else if (strategy == SMT_ABS_MEAN) r = s / cnt;
    else if (strategy == SMT_L1) r = s;
    else r = sqrtf(s);
    scores[(int64_t)brow * gridDim.x + bcol] = r;
  }
}

template <int B, int DT>
__global__ void __launch_bounds__(kThreads) block_sum_accumulate_kernel(
    float* __restrict__ sums, const void* __restrict__ grad, int64_t ld) {
  __shared__ float scratch[32];
  const int bcol = blockIdx.x, brow = blockIdx.y;
  float s = block_partial<B, DT, kSigned>(grad, ld, brow, bcol);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) sums[(int64_t)brow * gridDim.x + bcol] += s;
}

__global__ void block_sum_finalize_kernel(const float* __restrict__ sums, float* __restrict__ scores,
                                          int64_t n, float cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) scores[i] = fabsf(sums[i] / cnt);
}

// ---- activation scoring ------------------------------------------------------------------------

template <int DT>
__global__ void __launch_bounds__(kThreads) act_score_accumulate_kernel(
    float* __restrict__ acc, const void* __restrict__ x, int batch, int64_t sc_vec8) {
  // sc_vec8 = S*C/8 vectors per batch entry
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sc_vec8; i += stride) {
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < batch; ++b) {
      float g[8];
      if (DT == SMT_F32) {
        const float4* p = reinterpret_cast<const float4*>(x) + 2 * ((int64_t)b * sc_vec8 + i);
        const float4 a = ld_stream_f4(p), c = ld_stream_f4(p + 1);
        g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w;
        g[4] = c.x; g[5] = c.y; g[6] = c.z; g[7] = c.w;
      } else {
        const uint4 u = ld_stream_u4(reinterpret_cast<const uint4*>(x) + (int64_t)b * sc_vec8 + i);
        unpack8<DT>(u, g);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += fabsf(g[j]);
    }
    float4* ap = reinterpret_cast<float4*>(acc) + 2 * i;
    float4 a0 = ap[0], a1 = ap[1];
    a0.x += s[0]; a0.y += s[1]; a0.z += s[2]; a0.w += s[3];
    a1.x += s[4]; a1.y += s[5]; a1.z += s[6]; a1.w += s[7];
    ap[0] = a0;
    ap[1] = a1;
  }
}

// out[c] over rows of acc[S, C]; CTA = 32 column-quads x 8 row lanes.
__global__ void __launch_bounds__(kThreads) channel_score_reduce_kernel(
    const float* __restrict__ acc, int seq, int channels, int strategy, float* __restrict__ out) {
  __shared__ float4 part[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c4 = blockIdx.x * 32 + tx;  // column quad
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < channels) {
    for (int r = ty; r < seq; r += 8) {
      const float4 v = *reinterpret_cast<const float4*>(acc + (int64_t)r * channels + c4 * 4);
      if (strategy == SMT_L2) {
        s.x += v.x * v.x; s.y += v.y * v.y; s.z += v.z * v.z; s.w += v.w * v.w;
      } else if (strategy == SMT_ABS_MEAN) {
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      } else {
        s.x += fabsf(v.x); s.y += fabsf(v.y); s.z += fabsf(v.z); s.w += fabsf(v.w);
      }
    }
  }
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c4 * 4 < channels) {
    float4 t = part[0][tx];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      t.x += part[j][tx].x; t.y += part[j][tx].y; t.z += part[j][tx].z; t.w += part[j][tx].w;
    }
    const float cnt = (float)seq;
    if (strategy == SMT_MEAN_ABS) { t.x /= cnt; t.y /= cnt; t.z /= cnt; t.w /= cnt; }
    else if (strategy == SMT_ABS_MEAN) {
      t.x = fabsf(t.x / cnt); t.y = fabsf(t.y / cnt); t.z = fabsf(t.z / cnt); t.w = fabsf(t.w / cnt);
    } else if (strategy == SMT_L2) {
      t.x = sqrtf(t.x); t.y = sqrtf(t.y); t.z = sqrtf(t.z); t.w = sqrtf(t.w);
    }
    *reinterpret_cast<float4*>(out + c4 * 4) = t;
  }
}

inline int streaming_grid(int64_t work_items) {
  // enough CTAs for ~8 resident per SM, capped by the work
  int64_t want = (work_items + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)sm_count() * SMT_SCORE_CTAS;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace smt

using namespace smt;

extern "C" SMT_API int smt_score_accumulate(float* acc, const void* grad, int grad_dtype, int64_t n,
                                    void* stream) {
  SMT_CHECK_ARG(acc && grad, "smt_score_accumulate: null pointer");
  SMT_CHECK_ARG(n >= 0, "smt_score_accumulate: n < 0");
  SMT_CHECK_ARG(grad_dtype >= SMT_F32 && grad_dtype <= SMT_F16, "smt_score_accumulate: bad dtype %d",
                grad_dtype);
  if (n == 0) return SMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n_vec = (aligned16(acc) && aligned16(grad)) ? n / 8 : 0;
  if (n_vec > 0) {
    int grid = streaming_grid(n_vec);
    if (grad_dtype == SMT_F32) score_accumulate_vec_kernel<SMT_F32><<<grid, kThreads, 0, st>>>(acc, grad, n_vec);
    else if (grad_dtype == SMT_BF16) score_accumulate_vec_kernel<SMT_BF16><<<grid, kThreads, 0, st>>>(acc, grad, n_vec);
    else score_accumulate_vec_kernel<SMT_F16><<<grid, kThreads, 0, st>>>(acc, grad, n_vec);
    SMT_CHECK_LAUNCH();
  }
  if (n_vec * 8 < n) {
    int grid = streaming_grid(n - n_vec * 8);
    if (grad_dtype == SMT_F32) score_accumulate_scalar_kernel<SMT_F32><<<grid, kThreads, 0, st>>>(acc, grad, n_vec * 8, n);
    else if (grad_dtype == SMT_BF16) score_accumulate_scalar_kernel<SMT_BF16><<<grid, kThreads, 0, st>>>(acc, grad, n_vec * 8, n);
    else score_accumulate_scalar_kernel<SMT_F16><<<grid, kThreads, 0, st>>>(acc, grad, n_vec * 8, n);
    SMT_CHECK_LAUNCH();
  }
  return SMT_OK;
}

namespace {
int check_blocked(const char* who, const void* p, int rows, int cols, int64_t ld, int block, int elem_bytes) {
  SMT_CHECK_ARG(p != nullptr, "%s: null pointer", who);
  SMT_CHECK_ARG(block_ok(block), "%s: block size %d not in {64,128,256}", who, block);
  SMT_CHECK_ARG(rows > 0 && cols > 0 && rows % block == 0 && cols % block == 0,
                "%s: matrix %dx%d is not a multiple of block %d", who, rows, cols, block);
  SMT_CHECK_ARG(ld >= cols, "%s: ld %lld < cols %d", who, (long long)ld, cols);
  SMT_CHECK_ARG(aligned16(p) && (ld * elem_bytes) % 16 == 0, "%s: matrix must be 16-byte aligned (ptr and row pitch)", who);
  SMT_CHECK_ARG(rows / block <= 65535, "%s: too many block rows", who);
  return SMT_OK;
}
}  // namespace

extern "C" SMT_API int smt_block_score_reduce(const float* acc, int rows, int cols, int64_t ld, int block,
                                      int strategy, float* scores, void* stream) {
  if (int rc = check_blocked("smt_block_score_reduce", acc, rows, cols, ld, block, 4)) return rc;
  SMT_CHECK_ARG(scores != nullptr, "smt_block_score_reduce: null scores");
  SMT_CHECK_ARG(strategy >= SMT_MEAN_ABS && strategy <= SMT_L2, "smt_block_score_reduce: unknown strategy %d", strategy);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cols / block, rows / block);
  const int mode = strategy == SMT_MEAN_ABS ? kSigned : (strategy == SMT_L2 ? kSquare : kAbs);
#define SMT_LAUNCH_REDUCE(B)                                                                       \
  do {                                                                                             \
    if (mode == kSigned) block_score_reduce_kernel<B, kSigned><<<grid, kThreads, 0, st>>>(acc, ld, strategy, scores); \
    else if (mode == kAbs) block_score_reduce_kernel<B, kAbs><<<grid, kThreads, 0, st>>>(acc, ld, strategy, scores);  \
    else block_score_reduce_kernel<B, kSquare><<<grid, kThreads, 0, st>>>(acc, ld, strategy, scores);                 \
  } while (0)
  if (block == 256) SMT_LAUNCH_REDUCE(256);
  else if (block == 128) SMT_LAUNCH_REDUCE(128);
  else SMT_LAUNCH_REDUCE(64);
#undef SMT_LAUNCH_REDUCE
  SMT_CHECK_LAUNCH();
  return SMT_OK;
}

extern "C" SMT_API int smt_block_sum_accumulate(float* block_sums, const void* grad, int grad_dtype, int rows,
                                        int cols, int64_t ld, int block, void* stream) {
  SMT_CHECK_ARG(grad_dtype >= SMT_F32 && grad_dtype <= SMT_F16, "smt_block_sum_accumulate: bad dtype %d", grad_dtype);
  if (int rc = check_blocked("smt_block_sum_accumulate", grad, rows, cols, ld, block, dtype_bytes(grad_dtype))) return rc;
  SMT_CHECK_ARG(block_sums != nullptr, "smt_block_sum_accumulate: null block_sums");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cols / block, rows / block);
#define SMT_LAUNCH_BS(B)                                                                                      \
  do {                                                                                                        \
    if (grad_dtype == SMT_F32) block_sum_accumulate_kernel<B, SMT_F32><<<grid, kThreads, 0, st>>>(block_sums, grad, ld);   \
    else if (grad_dtype == SMT_BF16) block_sum_accumulate_kernel<B, SMT_BF16><<<grid, kThreads, 0, st>>>(block_sums, grad, ld); \
    else block_sum_accumulate_kernel<B, SMT_F16><<<grid, kThreads, 0, st>>>(block_sums, grad, ld);            \
  } while (0)
  if (block == 256) SMT_LAUNCH_BS(256);
  else if (block == 128) SMT_LAUNCH_BS(128);
  else SMT_LAUNCH_BS(64);
#undef SMT_LAUNCH_BS
  SMT_CHECK_LAUNCH();
  return SMT_OK;
}

extern "C" SMT_API int smt_block_sum_finalize(const float* block_sums, float* scores, int64_t n, int block,
                                      void* stream) {
  SMT_CHECK_ARG(block_sums && scores, "smt_block_sum_finalize: null pointer");
  SMT_CHECK_ARG(block_ok(block), "smt_block_sum_finalize: bad block %d", block);
  if (n <= 0) return SMT_OK;
  int grid = (int)((n + 255) / 256);
  block_sum_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(block_sums, scores, n, (float)(block * block));
  SMT_CHECK_LAUNCH();
  return SMT_OK;
}

extern "C" SMT_API int smt_act_score_accumulate(float* acc, const void* x, int x_dtype, int batch, int seq,
                                        int channels, void* stream) {
  SMT_CHECK_ARG(acc && x, "smt_act_score_accumulate: null pointer");
  SMT_CHECK_ARG(x_dtype >= SMT_F32 && x_dtype <= SMT_F16, "smt_act_score_accumulate: bad dtype %d", x_dtype);
  SMT_CHECK_ARG(batch > 0 && seq > 0 && channels > 0, "smt_act_score_accumulate: empty input");
  SMT_CHECK_ARG(channels % 8 == 0 && aligned16(acc) && aligned16(x), "smt_act_score_accumulate: channels must be a multiple of 8 and pointers 16-byte aligned");
  const int64_t sc_vec8 = (int64_t)seq * channels / 8;
  int grid = streaming_grid(sc_vec8);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == SMT_F32) act_score_accumulate_kernel<SMT_F32><<<grid, kThreads, 0, st>>>(acc, x, batch, sc_vec8);
  else if (x_dtype == SMT_BF16) act_score_accumulate_kernel<SMT_BF16><<<grid, kThreads, 0, st>>>(acc, x, batch, sc_vec8);
  else act_score_accumulate_kernel<SMT_F16><<<grid, kThreads, 0, st>>>(acc, x, batch, sc_vec8);
  SMT_CHECK_LAUNCH();
  return SMT_OK;
}

extern "C" SMT_API int smt_channel_score_reduce(const float* acc, int seq, int channels, int strategy,
                                        float* out, void* stream) {
  SMT_CHECK_ARG(acc && out, "smt_channel_score_reduce: null pointer");
  SMT_CHECK_ARG(seq > 0 && channels > 0 && channels % 4 == 0, "smt_channel_score_reduce: channels must be a positive multiple of 4");
  SMT_CHECK_ARG(aligned16(acc) && aligned16(out), "smt_channel_score_reduce: pointers must be 16-byte aligned");
  SMT_CHECK_ARG(strategy >= SMT_MEAN_ABS && strategy <= SMT_L2, "smt_channel_score_reduce: unknown strategy %d", strategy);
  int grid = (channels / 4 + 31) / 32;
  channel_score_reduce_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(acc, seq, channels, strategy, out);
  SMT_CHECK_LAUNCH();
  return SMT_OK;
}
