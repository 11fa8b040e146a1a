// Tensor-pipe rate of the MMA shapes the marching kernels issue (measurement tool, not part of the library):
// one CTA per SM, `nw` warps each issue `iters` back-to-back tcgen05.mma 128 x N x 16 (bf16, fp32 accumulate) on
// operands resident in shared memory, commit, and wait; cycles per MMA = elapsed SM clocks / MMAs issued by the CTA.
//   N         : 32 ... 256 (the marching kernels use 96 = three stacked 32-channel taps, 48 for Cout = 16)
//   layout    : K-major 64-byte rows (fprop / dgrad operands) or MN-major (wgrad operands)
//   nw        : 1 or 3 issuing warps (three = one per dz slab copy, the training mode)
//   spread    : consecutive MMAs of a warp go to the SAME accumulator columns (0) or rotate over two sets (1)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fetal-mri-segmentation_b200/csrc tools/mma_rate.cu -o tools/mma_rate.bin
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
typedef __nv_bfloat16 bf16;
#include "tc_ptx.cuh"

using namespace tcp;

struct Args {
  int N, mn_major, nw, spread, iters;
  int group, gap;  // second experiment: after every `group` MMAs the issuing warp spends `gap` clocks elsewhere
  long long* cycles;  // per CTA
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: A 128 rows (K-major: 144-row slab like the marching kernels) and B up to 256 rows, 64-byte rows
  const uint32_t a_base = smem0, b_base = smem0 + 32768u, bar0 = smem0 + 32768u + 32768u, tmem_slot = bar0 + 64u;
  for (uint32_t i = threadIdx.x; i < 65536u / 16u; i += blockDim.x)
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(smem0 + i * 16u), "r"(0u));
  if (warp == 0) {
    if (lane == 0) {
      for (int w = 0; w < 4; ++w) mbar_init(bar0 + 8u * w, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (warp_u < a.nw) {
    const uint32_t idesc = make_idesc(128, a.N, a.mn_major, a.mn_major);
    uint32_t a_lo, b_lo, a_hi, b_hi, kstep;
    if (!a.mn_major) {
      // K-major, 64-byte rows, SWIZZLE_64B: 8-row groups 512 B apart, k-step = 32 B = 2 descriptor units
      a_hi = b_hi = desc_hi(8u * 64u, layout_code(64));
      a_lo = desc_lo(a_base, 16u);
      b_lo = desc_lo(b_base, 16u);
      kstep = 2u;
    } else {
      // MN-major, 64-byte rows (32 channels): channel blocks one 512-byte atom apart (LBO = SBO), k-step = 16 voxels
      a_hi = b_hi = desc_hi(512u, layout_code(64));
      a_lo = desc_lo(a_base, 512u);
      b_lo = desc_lo(b_base, 8192u);
      kstep = (2u * 512u) >> 4;
    }
    const uint32_t acc0 = tmem_base + (uint32_t)warp_u * 128u;  // each warp its own accumulator columns (N <= 128) ...
    const uint32_t accN = a.N > 128 ? tmem_base : acc0;          // ... one warp only beyond
    t0 = clock64();
    if (a.group <= 1) {
      // rate table: nothing but the MMAs in the loop (even a counter and a compare per MMA make ONE issuing warp the
      // bottleneck: 75 cycles per MMA instead of 56 at N = 96)
      for (int i = 0; i < a.iters; ++i) {
        const uint32_t d = accN + ((a.spread && (i & 1) && a.N <= 64) ? 64u : 0u);
        umma_bf16_lh_elect(d, a_lo + (uint32_t)(i & 3) * kstep, a_hi, b_lo + (uint32_t)(i & 3) * kstep, b_hi, idesc, 1u);
      }
    } else {
      const bool gaps = a.gap > 0;
      int in_group = 0;  // (a counter, not i % group: an integer division per MMA costs 180 cycles of issue path)
      for (int i = 0; i < a.iters; ++i) {
        umma_bf16_lh_elect(accN, a_lo + (uint32_t)(i & 3) * kstep, a_hi, b_lo + (uint32_t)(i & 3) * kstep, b_hi, idesc, 1u);
        if (gaps && ++in_group == a.group) {
          in_group = 0;
          const long long tg = clock64();
          while (clock64() - tg < a.gap) {
          }
          __syncwarp();
        }
      }
    }
    umma_commit_elect(bar0 + 8u * (uint32_t)warp_u);
    mbar_wait(bar0 + 8u * (uint32_t)warp_u, 0);
    t1 = clock64();
    if (lane == 0) atomicMax((unsigned long long*)&a.cycles[blockIdx.x], (unsigned long long)(t1 - t0));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* cyc;
  cudaMalloc(&cyc, sizeof(long long) * sms);
  const size_t smem = 65536 + 1024 + 256;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 4096;
  printf("layout    N  warps  spread  cycles/MMA  ideal  (ideal = N/2 cycles: 4096 dense bf16 FMA = 8192 flop per clock and SM)\n");
  for (int mn = 0; mn < 2; ++mn)
    for (int nw = 1; nw <= 3; nw += 2)
      for (int N : {32, 48, 64, 96, 128, 192, 256}) {
        if (nw == 3 && N > 128) continue;
        if (mn && N > 128) continue;  // MN-major B tile laid out for <= 4 blocks of 32 channels
        for (int spread = 0; spread < 2; ++spread) {
          if (spread && N > 64) continue;
          cudaMemset(cyc, 0, sizeof(long long) * sms);
          Args a{N, mn, nw, spread, iters, 1, 0, cyc};
          mma_rate_kernel<<<sms, 128, smem>>>(a);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("%s N=%d nw=%d: %s\n", mn ? "MN-major" : "K-major ", N, nw, cudaGetErrorString(e));
            return 1;
          }
          long long h[256];
          cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("%s %4d  %5d  %6d  %10.1f  %5.0f\n", mn ? "MN-major" : "K-major ", N, nw, spread,
                 (double)mx / ((double)iters * nw), N / 2.0);
        }
      }
  // does the pipe keep working while the issuing warps are busy elsewhere? Rounds of `group` MMAs per warp (N = 96,
  // K-major) followed by `gap` clocks of other work in that warp: a deep MMA queue gives max(group * 56 * warps, gap + issue)
  // per round, a shallow one the sum.
  printf("\nrounds: N = 96, K-major; cycles per round of (group MMAs per warp, then gap clocks in the issuing warp)\n");
  printf("warps  group   gap  cycles/round  pipe work/round\n");
  for (int nw = 1; nw <= 3; nw += 2)
    for (int group : {6, 18})
      for (int gap : {0, 200, 400, 700, 1000, 1500}) {
        cudaMemset(cyc, 0, sizeof(long long) * sms);
        const int it = 4320;  // divisible by 6 and 18
        Args a{96, 0, nw, 0, it, group, gap, cyc};
        mma_rate_kernel<<<sms, 128, smem>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("rounds nw=%d group=%d gap=%d: %s\n", nw, group, gap, cudaGetErrorString(e));
          return 1;
        }
        long long h[256];
        cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%5d  %5d  %4d  %12.1f  %15d\n", nw, group, gap, (double)mx / (it / group), group * 56 * nw);
      }
  cudaFree(cyc);
  return 0;
}
