// Epilogue helpers shared by the block-gradient kernels: accumulator sub-tile -> global memory through a per-warp
// shared-memory transpose (every store instruction writes whole rows), and the L2-coherent loads of the split-K fix-up.
#pragma once

#include "common.cuh"

namespace smt {
namespace {

constexpr int kStageRow = 36;                 // floats per row of the epilogue transpose buffer (32 + 4 pad)

// Stores 4 values (optionally added to what is there) and returns the sum of squares of the values as STORED, i.e.
// after the rounding to the output type (the clip norm is defined on the gradient buffer's contents).
template <int ODT, bool ACC>
__device__ __forceinline__ float store4(void* out_base, int64_t off, float4 v) {
  if (ODT == SMT_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_base) + off);
    if (ACC) {
      const float4 old = *o;
      v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
    }
    *o = v;
    return (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  } else {
    uint2* o = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out_base) + off);
    if (ACC) {
      const uint2 old = *o;
      if (ODT == SMT_BF16) {
        v.x += __uint_as_float(old.x << 16); v.y += __uint_as_float(old.x & 0xffff0000u);
        v.z += __uint_as_float(old.y << 16); v.w += __uint_as_float(old.y & 0xffff0000u);
      } else {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&old.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&old.y));
        v.x += a.x; v.y += a.y; v.z += b.x; v.w += b.y;
      }
    }
    uint2 u;
    float4 w;
    if (ODT == SMT_BF16) {
      u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
      w = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                      __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    } else {
      u.x = pack_f16x2(v.x, v.y); u.y = pack_f16x2(v.z, v.w);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      w = make_float4(a.x, a.y, b.x, b.y);
    }
    *o = u;
    return (w.x * w.x + w.y * w.y) + (w.z * w.z + w.w * w.w);
  }
}

// One 32x32 fp32 accumulator sub-tile (lane = row, r[] = 32 columns) -> global, transposed through a per-warp
// shared-memory buffer so that each store instruction covers 4 rows x 128 B (fp32) / 64 B (16-bit).
template <int ODT, bool ACC>
__device__ __forceinline__ float store_subtile(float* stage, const uint32_t (&r)[32], int lane, void* out_base,
                                               int64_t off00, int ld) {
  float4* mine = reinterpret_cast<float4*>(stage + lane * kStageRow);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    mine[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                          __uint_as_float(r[4 * q + 3]));
  __syncwarp();
  const int sub = lane >> 3, cv = (lane & 7) * 4;
  float sq = 0.f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + sub;
    const float4 v = *reinterpret_cast<const float4*>(stage + row * kStageRow + cv);
    sq += store4<ODT, ACC>(out_base, off00 + (int64_t)row * ld + cv, v);
  }
  __syncwarp();
  return sq;
}

// Dispatch on (output type, accumulate) for one sub-tile; returns the thread's sum of squares of the stored values.
__device__ __forceinline__ float store_subtile_any(int out_dtype, bool acc, float* stage, const uint32_t (&r)[32],
                                                   int lane, void* out_base, int64_t off, int ld) {
  if (out_dtype == SMT_F32)
    return acc ? store_subtile<SMT_F32, true>(stage, r, lane, out_base, off, ld)
               : store_subtile<SMT_F32, false>(stage, r, lane, out_base, off, ld);
  if (out_dtype == SMT_BF16)
    return acc ? store_subtile<SMT_BF16, true>(stage, r, lane, out_base, off, ld)
               : store_subtile<SMT_BF16, false>(stage, r, lane, out_base, off, ld);
  return acc ? store_subtile<SMT_F16, true>(stage, r, lane, out_base, off, ld)
             : store_subtile<SMT_F16, false>(stage, r, lane, out_base, off, ld);
}

__device__ __forceinline__ float4 ld_cg_f4(const float* p) {   // L2-coherent load (data written by other CTAs of this grid)
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace
}  // namespace smt
