// Micro-benchmark: how fast does ONE thread feed tcgen05.mma (M = 128, K = 16, bf16) of various N / operand majors,
// and how long do they take to retire?  Prints cycles per MMA as seen by the issuing thread (issue) and until the commit
// arrives (retire).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mtvaf_b200/csrc
//   tools/micro/umma_rate.cu -o gpurun_out/umma_rate -lcuda     (operands are zeros in shared memory; values do not matter)
#ifndef UNIFORM_TMEM
#define UNIFORM_TMEM 1
#endif
#include <cstdio>
#include <cuda.h>
#include "common.cuh"
#include "ptx.cuh"
using namespace mtvaf;
using namespace mtvaf::ptx;

struct Cfg { int N, a_mn, b_mn, n_acc, group, reps, unroll; };

__global__ void __launch_bounds__(128, 1) umma_rate_kernel(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncwarp();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // REDUX writes a UNIFORM register: the MMA's TMEM operand is then known to be warp-uniform and the compiler drops the
  // ELECT / R2UR.BROADCAST / BRA.U.ANY loop it otherwise wraps around every tcgen05.mma issued from a divergent branch
  const uint32_t tmem_base = UNIFORM_TMEM ? __reduce_or_sync(0xffffffffu, tmem_slot) : tmem_slot;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t idesc = make_idesc_bf16(128, c.N, c.a_mn, c.b_mn);
    const uint64_t dA = c.a_mn ? make_smem_desc_sw128(smem_u32(smem), 8192, 1024) : make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t dB = c.b_mn ? make_smem_desc_sw128(smem_u32(smem) + 65536, 8192, 1024)
                               : make_smem_desc_sw128(smem_u32(smem) + 65536, 16, 1024);
    const int stepA = c.a_mn ? 128 : 2, stepB = c.b_mn ? 128 : 2;
    long long t_issue = 0, t_retire = 0;
    uint32_t ph = 0;
    for (int r = 0; r < c.reps; ++r) {
      const long long t0 = clock64();
      if (c.n_acc == 1) {
        if (c.unroll) {
#pragma unroll 8
          for (int g = 0; g < c.group; ++g)
            umma_f16_ss(tmem_base, dA + (g & 3) * stepA, dB + (g & 3) * stepB, idesc, g > 0 ? 1u : 0u);
        } else {
#pragma unroll 1
          for (int g = 0; g < c.group; ++g)
            umma_f16_ss(tmem_base, dA + (g & 3) * stepA, dB + (g & 3) * stepB, idesc, g > 0 ? 1u : 0u);
        }
      } else {
        int acc = 0;
#pragma unroll 1
        for (int g = 0; g < c.group; ++g) {
          umma_f16_ss(tmem_base + acc * 128, dA + (g & 3) * stepA, dB + (g & 3) * stepB, idesc, g >= c.n_acc ? 1u : 0u);
          acc = acc + 1 == c.n_acc ? 0 : acc + 1;
        }
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
      const long long t2 = clock64();
      t_issue += t1 - t0;
      t_retire += t2 - t0;
    }
    out[0] = t_issue; out[1] = t_retire;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const Cfg cfgs[] = {
      {64, 0, 0, 1, 8, 50, 0},  {64, 0, 0, 1, 32, 50, 0}, {64, 0, 0, 1, 32, 50, 1}, {64, 0, 0, 3, 24, 50, 0}, {64, 1, 1, 1, 8, 50, 0},
      {64, 1, 1, 1, 32, 50, 0}, {64, 1, 1, 1, 32, 50, 1}, {64, 0, 1, 1, 32, 50, 1}, {64, 1, 0, 1, 32, 50, 1}, {16, 1, 1, 1, 32, 50, 1},
      {32, 1, 1, 1, 32, 50, 1}, {144, 0, 0, 1, 4, 50, 1}, {144, 0, 0, 1, 32, 50, 1}, {128, 0, 0, 1, 32, 50, 1}, {256, 0, 0, 1, 32, 50, 1},
      {256, 0, 0, 1, 32, 50, 0}, {64, 1, 1, 3, 24, 50, 0}, {8, 0, 0, 1, 32, 50, 1},
  };
  printf("%5s %4s %4s %5s %6s %3s | %12s %12s\n", "N", "A_mn", "B_mn", "n_acc", "group", "unr", "issue/MMA", "retire/MMA");
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      umma_rate_kernel<<<1, 128, 200 * 1024>>>(c, d);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double n = (double)c.reps * c.group;
    printf("%5d %4d %4d %5d %6d %3d | %12.1f %12.1f\n", c.N, c.a_mn, c.b_mn, c.n_acc, c.group, c.unroll, h[0] / n, h[1] / n);
  }
  return 0;
}
