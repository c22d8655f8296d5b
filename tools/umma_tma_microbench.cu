// Microbenchmarks behind the block-gradient GEMM's design choices (sm_100a).  Standalone:
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_tma_microbench tools/umma_tma_microbench.cu -lcuda
//   /tmp/umma_tma_microbench
//
// (1) UMMA issue rate: one thread issues back-to-back tcgen05.mma (kind::f16, M = 128, both operands MN-major from static
//     shared memory, no TMA in the loop) for N in {64, 128, 256}, into 1 or 2 accumulators -> clocks per instruction.
// (2) TMA fill rate of one SM and of the whole chip: boxes of {64 features x R tokens} (R = 16 ... 256, 128-byte swizzle)
//     kept `inflight` deep from an L2-resident source -> bytes per clock per SM, and clocks per box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);             \
      exit(1);                                                                                    \
    }                                                                                             \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {     // non-blocking poll
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
// mode 0: blocking try_wait (default suspend limit); 1: test_wait polling; 2: try_wait with a 20 ns suspend hint
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int mode = 0) {
  const long long t0 = clock64();
  for (;;) {
    const bool ok = mode == 0 ? mbar_try_wait(bar, parity) : mode == 1 ? mbar_test_wait(bar, parity)
                                                                       : mbar_try_wait_hint(bar, parity, 20u);
    if (ok) return;
    if (clock64() - t0 > 2000000000ll) { printf("microbench: mbarrier timeout\n"); __trap(); }
  }
}
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {   // bf16 x bf16 -> fp32, A and B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- (1) UMMA issue rate ----------------------------------------------------------------------------
template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int iters, long long* out_clk) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;   // finite bf16 values
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N);
    // A: 2 chunks of {64 rows x 64 tokens} at 8 KiB stride; B: N/64 chunks behind them
    const uint32_t a0 = base, b0 = base + 16384;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = make_desc_mn_sw128(a0 + k * 2048, 8192, 1024);
        const uint64_t bd = make_desc_mn_sw128(b0 + k * 2048, 8192, 1024);
        const uint32_t d = tmem + (uint32_t)(((it * 4 + k) % NACC) * N);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    mbar_wait(smem_u32(&done_bar), 0);
    out_clk[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ---- (2) TMA fill rate ------------------------------------------------------------------------------
constexpr int kMaxInflight = 24;

__global__ void __launch_bounds__(32, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap map, int rows, int inflight,
                                                         int n_boxes, int n_col_chunks, int n_row_boxes,
                                                         long long* out_clk) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[kMaxInflight];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t box_bytes = (uint32_t)rows * 128u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < inflight; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // every SM walks the (L2-resident) source from a different offset
    int cc = (int)(blockIdx.x % (unsigned)n_col_chunks), rb = (int)((blockIdx.x * 37u) % (unsigned)n_row_boxes);
    auto issue = [&](int slot) {
      if (++cc == n_col_chunks) { cc = 0; if (++rb == n_row_boxes) rb = 0; }      // no divisions in the timed loop
      mbar_expect_tx(smem_u32(&bar[slot]), box_bytes);
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(base + slot * box_bytes), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar[slot])),
                     "r"(cc * 64), "r"(rb * rows) : "memory");
    };
    const long long t0 = clock64();
    for (int i = 0; i < inflight; ++i) issue(i);
    int slot = 0;
    uint32_t par = 0;
    for (int i = 0; i < n_boxes; ++i) {
      mbar_wait(smem_u32(&bar[slot]), par);
      if (i + inflight < n_boxes) issue(slot);
      if (++slot == inflight) { slot = 0; par ^= 1u; }
    }
    out_clk[blockIdx.x] = clock64() - t0;
  }
}

// Same, structured like the GEMM's producer: `group` boxes of R rows share one mbarrier ("stage"), `stages` stages in
// flight, one thread issues and waits.  wait_mode as in mbar_wait.
__global__ void __launch_bounds__(32, 1) tma_stage_kernel(const __grid_constant__ CUtensorMap map, int rows, int group,
                                                          int stages, int n_stages_total, int n_col_chunks,
                                                          int n_row_boxes, int wait_mode, long long* out_clk) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[kMaxInflight];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t box_bytes = (uint32_t)rows * 128u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int cc = (int)(blockIdx.x % (unsigned)n_col_chunks), rb = (int)((blockIdx.x * 37u) % (unsigned)n_row_boxes);
    auto issue = [&](int slot) {
      mbar_expect_tx(smem_u32(&bar[slot]), box_bytes * group);
      for (int g = 0; g < group; ++g) {
        if (++cc == n_col_chunks) { cc = 0; if (++rb == n_row_boxes) rb = 0; }    // no divisions in the timed loop
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(base + (slot * group + g) * box_bytes), "l"(reinterpret_cast<uint64_t>(&map)),
                       "r"(smem_u32(&bar[slot])), "r"(cc * 64), "r"(rb * rows) : "memory");
      }
    };
    const long long t0 = clock64();
    for (int i = 0; i < stages; ++i) issue(i);
    int slot = 0;
    uint32_t par = 0;
    for (int i = 0; i < n_stages_total; ++i) {
      mbar_wait(smem_u32(&bar[slot]), par, wait_mode);
      if (i + stages < n_stages_total) issue(slot);
      if (++slot == stages) { slot = 0; par ^= 1u; }
    }
    out_clk[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int N, int NACC>
void run_umma(const char* label, long long* d_clk) {
  auto k = umma_rate_kernel<N, NACC>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  const int iters = 2000;
  for (int grid : {1, 148}) {
    k<<<grid, 128, 100 * 1024>>>(iters, d_clk);
    CK(cudaDeviceSynchronize());
    k<<<grid, 128, 100 * 1024>>>(iters, d_clk);
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, d_clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = (double)mx / (iters * 4.0);
    printf("| %s | %d | %.1f | %.0f %% |\n", label, grid, per, 100.0 * (N / 2.0) / per);
  }
}

int main() {
  long long* d_clk;
  CK(cudaMalloc(&d_clk, sizeof(long long) * 148));
  printf("## (1) tcgen05.mma kind::f16, M = 128, K = 16, MN-major operands from shared memory, one issuing thread\n\n");
  printf("| N / accumulators | CTAs (1 per SM) | clk per UMMA | of the nominal N/2 clk |\n|---|---:|---:|---:|\n");
  run_umma<64, 1>("N = 64, 1 accumulator", d_clk);
  run_umma<64, 2>("N = 64, 2 accumulators", d_clk);
  run_umma<128, 1>("N = 128, 1 accumulator", d_clk);
  run_umma<128, 2>("N = 128, 2 accumulators", d_clk);
  run_umma<256, 1>("N = 256, 1 accumulator", d_clk);
  run_umma<256, 2>("N = 256, 2 accumulators", d_clk);

  // source: [8192 tokens, 4096 features] bf16 = 64 MiB (L2-resident after the first pass)
  const int64_t T = 8192, F = 4096;
  void* src;
  CK(cudaMalloc(&src, T * F * 2));
  CK(cudaMemset(src, 0, T * F * 2));
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
  CK(cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("\n## (2) TMA boxes {64 features x R tokens} (R x 128 B, SWIZZLE_128B) from an L2-resident 64 MiB source, one issuing thread per SM\n\n");
  printf("| R (rows per box) | boxes in flight | bytes in flight | CTAs | clk per box | B/clk per SM |\n|---:|---:|---:|---:|---:|---:|\n");
  for (int rows : {16, 32, 64, 128, 256}) {
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)F, (cuuint64_t)T};
    const cuuint64_t gstride[1] = {(cuuint64_t)F * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int inflight : {1, 2, 4, 8, 12, 24}) {
      if ((int64_t)inflight * rows * 128 > 192 * 1024) continue;
      for (int grid : {1, 148}) {
        const int n_boxes = 4096;
        for (int rep = 0; rep < 2; ++rep) {
          tma_rate_kernel<<<grid, 32, 200 * 1024>>>(map, rows, inflight, n_boxes, (int)(F / 64), (int)(T / rows), d_clk);
          CK(cudaDeviceSynchronize());
        }
        long long h[148];
        CK(cudaMemcpy(h, d_clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("| %d | %d | %d KiB | %d | %.0f | %.1f |\n", rows, inflight, inflight * rows * 128 / 1024, grid,
               (double)mx / n_boxes, (double)n_boxes * rows * 128 / (double)mx);
      }
    }
  }
  CK(cudaFuncSetAttribute(tma_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("\n## (3) stages of `group` boxes on one mbarrier, 148 CTAs, by wait flavour (0 = try_wait, 1 = test_wait polling, 2 = try_wait + 20 ns hint)\n\n");
  printf("| R | boxes per stage | stages in flight | KiB per stage | wait | clk per stage | clk per box | B/clk per SM |\n|---:|---:|---:|---:|---:|---:|---:|---:|\n");
  for (int rows : {64, 128, 256}) {
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)F, (cuuint64_t)T};
    const cuuint64_t gstride[1] = {(cuuint64_t)F * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    for (int group : {1, 2, 4, 8}) {
      for (int stages : {2, 3, 4, 6, 8}) {
        const int64_t bytes = (int64_t)group * stages * rows * 128;
        if (bytes > 192 * 1024 || bytes < 64 * 1024) continue;
        for (int mode : {0, 1, 2}) {
          const int n_total = 2048;
          for (int rep = 0; rep < 2; ++rep) {
            tma_stage_kernel<<<148, 32, 200 * 1024>>>(map, rows, group, stages, n_total, (int)(F / 64), (int)(T / rows), mode, d_clk);
            CK(cudaDeviceSynchronize());
          }
          long long h[148];
          CK(cudaMemcpy(h, d_clk, sizeof(long long) * 148, cudaMemcpyDeviceToHost));
          long long mx = 0;
          for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("| %d | %d | %d | %d | %d | %.0f | %.0f | %.1f |\n", rows, group, stages, group * rows * 128 / 1024, mode,
                 (double)mx / n_total, (double)mx / n_total / group, (double)n_total * group * rows * 128 / (double)mx);
        }
      }
    }
  }
  return 0;
}
