// conv_wgrad_march2.cu — second-generation plane-marching weight gradient of Conv3D 3x3x3 on tcgen05:
//
//     dW[kx,ky,kz][ci][co] = sum_u  X[u][ci] * dY[u - (kx-1, ky-1, kz-1)][co]
//
// Same march as conv_wgrad_march.cu (a CTA owns a 16 (y) x 8 (z) column and walks along x; both operands MN-major;
// the three ky taps are the M blocks of ONE A operand, one swizzle atom apart; the kx taps are N blocks = dY tiles of
// consecutive output planes in a ring), rebuilt around what the round-2 ncu capture showed
// (profiles/r2_ncu_key_metrics_train.json): the first generation issues three 128x96x16 MMAs per 16-voxel k-step,
// one per kz, each reading its own z-shifted copy of the X slab (4 KB) and the SAME three dY tiles (3 KB) - 21 KB of
// shared-memory operand reads per 144 tensor-pipe cycles = 146 B/clk against a 128 B/clk port, plus the TMA writes
// of three slab copies per plane. Here the kz shift moves to the OTHER operand:
//   * ONE X slab per input plane (19 y rows x 8 z, no z shift) is the A operand of every MMA of the plane;
//   * every output plane's dY arrives as THREE z-shifted tiles (z origin z0 + 1 - kz; TMA zero-fills outside the
//     volume = the 'same' padding) stored back to back, so the N blocks (kx, kz) = (plane j, copy kz) sit one tile
//     apart: 9 blocks x Cout columns. 288 > 256, so a k-step is TWO MMAs on two issuing warps with their own
//     accumulator columns: planes j, j+1 (N = 6 blocks = 192) and plane j+2 (N = 3 blocks = 96).
//   Operand reads per k-step: 2 x 4 KB (A) + 9 KB (B) = 17 KB per 144 cycles = 118 B/clk, and one slab TMA per plane
//   instead of three.
// The M waste is unchanged (a 3-tap stencil fills 96 of 128 rows with 32-channel blocks).
// TF autodiff gradient of Conv3D in create_convolution_block (fetal_net/model/unet3d/unet.py:102).
#include <algorithm>

#include "tc_ptx.cuh"

using namespace tcp;

namespace {

constexpr int kThreadsW2 = 256;  // warp 0 X producer, warp 1 dY producer, warps 2-3 MMA issue, warps 4-7 epilogue
constexpr int kMmaW0 = 2, kEpiW0 = 4;
constexpr int kBY = 16, kBZ = 8;
constexpr int kCC = 32;                 // channels per operand block (64 B rows, SWIZZLE_64B); 16 -> SWIZZLE_32B
constexpr uint32_t kXSlot = 10240;      // 19 (23 for 16-channel blocks) x 8 rows, rounded up to 1 KB
constexpr int kMirror = 2;              // ring slots R, R+1 repeat slots 0, 1

struct alignas(64) WgMarch2Params {
  CUtensorMap tmX;   // box (kcx, 8, 19 | 23, 1, 1)
  CUtensorMap tmDY;  // box (kcy, 8, 16, 1, 1)
  int N, X, Y, Z;
  int ny, nz;
  PlaneSplit split;      // how the (column, x) plane-tiles are dealt to the CTAs of a pair (common.cuh)
  int n_ci, n_co;
  int ctas_per_pair;
  int S;               // X slab slots
  int R;               // dY ring slots (planes), + kMirror mirror slots
  int kcx, kcy;        // channels per M block / N block: 32 or 16
  int Ct, cofs;
  float* dw;
  float* db;
};

__global__ void __launch_bounds__(kThreadsW2, 1) conv3d_wgrad_march2_kernel(const __grid_constant__ WgMarch2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t tile = 128u * (uint32_t)p.kcy * 2u;   // one z-shifted dY tile (8 KB at 32 channels)
  const uint32_t slot_b = 3u * tile;                   // the three copies of one output plane
  const uint32_t x_base = smem0;
  const uint32_t dy_base = x_base + (uint32_t)p.S * kXSlot;
  const uint32_t bar0 = dy_base + (uint32_t)(p.R + kMirror) * slot_b;
  auto xfull_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto xempty_bar = [&](uint32_t s) { return bar0 + 8u * ((uint32_t)p.S + s); };
  auto dyfull_bar = [&](uint32_t s) { return bar0 + 8u * ((uint32_t)(2 * p.S) + s); };
  auto dyempty_bar = [&](uint32_t s) { return bar0 + 8u * ((uint32_t)(2 * p.S + p.R) + s); };
  const uint32_t zero_bar = bar0 + 8u * (uint32_t)(2 * p.S + 2 * p.R);
  const uint32_t done_bar = zero_bar + 8u;
  const uint32_t tmem_slot = done_bar + 8u;
  constexpr uint32_t tmem_cols = 512;  // 9 blocks x <= 32 columns = 288, power of two

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmX);
    prefetch_tmap(&p.tmDY);
  }
  const int pair = blockIdx.x / p.ctas_per_pair;
  const int rank = blockIdx.x % p.ctas_per_pair;
  const int cic = pair % p.n_ci, coc = pair / p.n_ci;
  const bool do_bias = p.db != nullptr && cic == 0;
  if (warp == kMmaW0) {
    if (lane == 0) {
      for (int s = 0; s < p.S; ++s) {
        mbar_init(xfull_bar(s), 1);
        mbar_init(xempty_bar(s), 2);                  // both MMA warps
      }
      for (int s = 0; s < p.R; ++s) {
        mbar_init(dyfull_bar(s), 1);
        mbar_init(dyempty_bar(s), do_bias ? 6 : 2);   // both MMA warps (+ one arrival per bias-summing warp)
      }
      mbar_init(zero_bar, 128);
      mbar_init(done_bar, 2);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t R = (uint32_t)p.R;

  // k-th segment / work unit of this CTA (see PlaneSplit in common.cuh)
  auto next_seg = [&](int& it, int& n, int& iy, int& iz, int& xa, int& xb) -> bool {
    int col;
    if (!plane_split_next(p.split, rank, p.ctas_per_pair, it, col, xa, xb)) return false;
    iz = col % p.nz;
    const int r = col / p.nz;
    iy = r % p.ny;
    n = r / p.ny;
    ++it;
    return true;
  };

  if (warp_u == 0) {
    // ===== X producer: one slab per input plane =====
    pdl_wait();
    pdl_launch_dependents();
    const uint32_t x_bytes = (uint32_t)(kBY + 128 / p.kcx - 1) * kBZ * (uint32_t)p.kcx * 2u;
    uint32_t sidx = 0, sph = 0;
    for (int it = 0, n, iy, iz, xa, xb; next_seg(it, n, iy, iz, xa, xb);) {
      const int x_first = max(xa - 1, 0), x_last = min(xb, p.X - 1);
      for (int xi = x_first; xi <= x_last; ++xi) {
        mbar_wait(xempty_bar(sidx), sph ^ 1u);
        mbar_expect_tx_elect(xfull_bar(sidx), x_bytes);
        tma_load_5d_elect(x_base + sidx * kXSlot, &p.tmX, xfull_bar(sidx), cic * p.kcx, iz * kBZ, iy * kBY - 1, xi, n);
        if (++sidx == (uint32_t)p.S) {
          sidx = 0;
          sph ^= 1u;
        }
      }
    }
  } else if (warp_u == 1) {
    // ===== dY producer: three z-shifted tiles per output plane (copy kz at z origin z0 + 1 - kz) =====
    pdl_wait();
    const uint32_t dy_bytes = 3u * tile;
    uint32_t dcount = 0;
    for (int it = 0, n, iy, iz, xa, xb; next_seg(it, n, iy, iz, xa, xb);) {
      for (int xo = xa; xo < xb; ++xo, ++dcount) {
        const uint32_t slot = dcount % R;
        mbar_wait(dyempty_bar(slot), ((dcount / R) & 1u) ^ 1u);
        const bool mirror = slot < (uint32_t)kMirror;
        mbar_expect_tx_elect(dyfull_bar(slot), mirror ? 2u * dy_bytes : dy_bytes);
#pragma unroll
        for (int kz = 0; kz < 3; ++kz) {
          tma_load_5d_elect(dy_base + slot * slot_b + (uint32_t)kz * tile, &p.tmDY, dyfull_bar(slot), coc * p.kcy,
                            iz * kBZ + 1 - kz, iy * kBY, xo, n);
          if (mirror)
            tma_load_5d_elect(dy_base + (slot + R) * slot_b + (uint32_t)kz * tile, &p.tmDY, dyfull_bar(slot), coc * p.kcy,
                              iz * kBZ + 1 - kz, iy * kBY, xo, n);
        }
      }
    }
  } else if (warp_u < kEpiW0) {
    // ===== MMA warps: warp 0 issues N blocks of output planes j, j+1 (6 blocks), warp 1 those of plane j+2 =====
    const int w = warp_u - kMmaW0;
    const uint32_t ncy = (uint32_t)p.kcy;
    const uint32_t idesc3 = make_idesc(128, 3 * (int)ncy, 1, 1), idesc6 = make_idesc(128, 6 * (int)ncy, 1, 1);
    const uint32_t b_row = ncy * 2u, b_sbo = 8u * b_row;
    const uint32_t hi32 = desc_hi(b_sbo, layout_code((int)b_row));
    const uint32_t kstep = (2u * b_sbo) >> 4;  // 16 voxels per MMA, in 16-byte units
    const uint32_t a_row = (uint32_t)p.kcx * 2u, a_sbo = 8u * a_row;
    const uint32_t a_hi32 = desc_hi(a_sbo, layout_code((int)a_row));
    const uint32_t a_kstep = (2u * a_sbo) >> 4;
    mbar_wait(zero_bar, 0);
    tc_fence_after();
    uint32_t sidx = 0, sph = 0, dcount = 0, dwaited = 0;
    for (int it = 0, n, iy, iz, xa, xb; next_seg(it, n, iy, iz, xa, xb);) {
      const int x_first = max(xa - 1, 0), x_last = min(xb, p.X - 1);
      for (int xi = x_first; xi <= x_last; ++xi) {
        const int lo = max(xa, xi - 1), hi = min(xb - 1, xi + 1);  // output planes paired with input plane xi
        const uint32_t j_lo = (uint32_t)(lo - (xi - 1));           // block group j <-> xo = xi - 1 + j <-> kx = 2 - j
        const uint32_t nblk = (uint32_t)(hi - lo + 1);
        const uint32_t seq_lo = dcount + (uint32_t)(lo - xa);
        const uint32_t seq_need = dcount + (uint32_t)(hi - xa);
        for (; dwaited <= seq_need; ++dwaited) mbar_wait(dyfull_bar(dwaited % R), (dwaited / R) & 1u);
        mbar_wait(xfull_bar(sidx), sph);
        tc_fence_after();
        const uint32_t rs = seq_lo % R;  // the <= 3 consecutive planes never wrap: slots R, R+1 mirror slots 0, 1
        // this warp's share: planes [g0, g0 + gn) of the nblk
        const uint32_t g0 = w == 0 ? 0u : 2u;
        const uint32_t gn = w == 0 ? min(nblk, 2u) : (nblk == 3u ? 1u : 0u);
        if (gn > 0) {
          uint32_t a_lo = desc_lo(x_base + sidx * kXSlot, a_sbo);                 // M blocks (ky) one atom apart
          uint32_t b_lo = desc_lo(dy_base + (rs + g0) * slot_b, tile);            // N blocks (plane, kz) one tile apart
          const uint32_t d = tmem_base + (j_lo + g0) * 3u * ncy;
          const uint32_t id = gn == 2 ? idesc6 : idesc3;
#pragma unroll
          for (int ks = 0; ks < (kBY * kBZ) / 16; ++ks) {
            umma_bf16_lh_elect(d, a_lo, a_hi32, b_lo, hi32, id, 1u);
            a_lo += a_kstep;
            b_lo += kstep;
          }
        }
        umma_commit_elect(xempty_bar(sidx));
        if (++sidx == (uint32_t)p.S) {
          sidx = 0;
          sph ^= 1u;
        }
        // dY planes this warp is done with: output plane xi-1 always, and the rest at the end of the item
        if (xi - 1 >= xa) umma_commit_elect(dyempty_bar((dcount + (uint32_t)(xi - 1 - xa)) % R));
        if (xi == x_last && xi <= xb - 1) umma_commit_elect(dyempty_bar((dcount + (uint32_t)(xi - xa)) % R));
      }
      dcount += (uint32_t)(xb - xa);
    }
    umma_commit_elect(done_bar);
  } else {
    // ===== epilogue warps: zero the accumulators, sum the bias gradient on the way, flush once at the end =====
    const int q = warp & 3;
    const int row = q * 32 + lane;  // (ky, ci) = (row / kcx, row % kcx); ky >= 3: discarded blocks
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ncol16 = 9 * p.kcy / 16;
    for (int c16 = 0; c16 < ncol16; ++c16) tmem_st16_zero(lane_base + (uint32_t)c16 * 16u);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(zero_bar);
    if (do_bias) {
      // column sums of the UNSHIFTED dY copy (kz = 1) of every plane as it passes through shared memory
      const int et = (warp - kEpiW0) * 32 + lane;
      const int CH = p.kcy / 8, RP = 128 / CH;
      const int c = et % CH, r0 = et / CH;
      const uint32_t rowb = (uint32_t)p.kcy * 2u;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      uint32_t dcount = 0;
      for (int it = 0, n, iy, iz, xa, xb; next_seg(it, n, iy, iz, xa, xb);) {
        for (int xo = xa; xo < xb; ++xo, ++dcount) {
          const uint32_t slot = dcount % R;
          mbar_wait(dyfull_bar(slot), (dcount / R) & 1u);
          const uint32_t t0 = dy_base + slot * slot_b + tile;
          for (int k = 0; k < CH; ++k) {
            const int r = r0 + k * RP;
            const int sw = CH == 4 ? ((r >> 1) & 3) : ((r >> 2) & 1);
            uint32_t v0, v1, v2, v3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                         : "r"(t0 + (uint32_t)r * rowb + (uint32_t)((c ^ sw) * 16)));
            const uint32_t vv[4] = {v0, v1, v2, v3};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv[j]));
              acc[2 * j] += f.x;
              acc[2 * j + 1] += f.y;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(dyempty_bar(slot));
        }
      }
      for (int ofs = CH; ofs < 32; ofs <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], ofs);
      }
      if (lane < CH) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(p.db + coc * p.kcy + c * 8 + j, acc[j]);
      }
    }
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int ky = row / p.kcx, ci = row % p.kcx;
    for (int c16 = 0; c16 < ncol16; ++c16) {
      uint32_t r[16];
      tmem_ld16(lane_base + (uint32_t)(c16 * 16), r);
      tmem_ld_wait();
      if (ky < 3) {
        const int col = c16 * 16;
        const int blk = col / p.kcy;           // (plane group j, copy kz)
        const int kx = 2 - blk / 3, kz = blk % 3;
        const int tap = (kx * 3 + ky) * 3 + kz;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = coc * p.kcy + col % p.kcy + j;
          atomicAdd(p.dw + ((int64_t)co * 27 + tap) * p.Ct + p.cofs + cic * p.kcx + ci, __uint_as_float(r[j]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaW0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_w2() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int make_map2(CUtensorMap* tm, const bf16* base, int N, int X, int Y, int Z, int C, int by, int cbox) {
  PFN_encodeTiled enc = get_encode_w2();
  FM_CHECK(enc != nullptr, FM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)Z * C * 2, (cuuint64_t)Y * Z * C * 2,
                           (cuuint64_t)X * Y * Z * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)cbox, (cuuint32_t)kBZ, (cuuint32_t)by, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, cbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FM_CHECK(r == CUDA_SUCCESS, FM_ECUDA, "cuTensorMapEncodeTiled(wgrad march2) failed: %d", (int)r);
  return FM_OK;
}

const int kMaxDynSmemW2 = 227 * 1024;

}  // namespace

int k_conv3d_wgrad_march2(fm_ctx* ctx, const bf16* x, const bf16* dy, float* dw_packed, int N, int X, int Y, int Z,
                          int Cin, int Cin_total, int cin_ofs, int Cout, float* db) {
  FM_CHECK(conv_wgrad_march_supported(X, Y, Z, Cin, Cout, 3), FM_EINVAL,
           "conv3d wgrad march2: unsupported shape %dx%dx%d Cin=%d Cout=%d", X, Y, Z, Cin, Cout);
  WgMarch2Params p;
  memset(&p, 0, sizeof(p));
  p.N = N;
  p.X = X;
  p.Y = Y;
  p.Z = Z;
  p.ny = Y / kBY;
  p.nz = Z / kBZ;
  p.kcx = (Cin % kCC == 0) ? kCC : 16;
  p.n_ci = Cin / p.kcx;
  p.kcy = (Cout % kCC == 0) ? kCC : 16;
  p.n_co = Cout / p.kcy;
  p.Ct = Cin_total;
  p.cofs = cin_ofs;
  p.dw = dw_packed;
  p.db = db;
  const int pairs = p.n_ci * p.n_co;
  const int cols = N * p.ny * p.nz;
  // CTAs per (ci, co) pair: at least 4 planes each; work units dealt as PlaneSplit describes (common.cuh)
  p.ctas_per_pair = std::max(1, std::min(ctx->num_sms / pairs, cols * X / 4));
  plane_split_setup(&p.split, cols, X, p.ctas_per_pair, 1.0);
  if (p.split.mode == 1) p.ctas_per_pair = std::min(p.ctas_per_pair, cols * p.split.nxc);
  FM_TRY(make_map2(&p.tmX, x, N, X, Y, Z, Cin, kBY + 128 / p.kcx - 1, p.kcx));
  FM_TRY(make_map2(&p.tmDY, dy, N, X, Y, Z, Cout, kBY, p.kcy));
  // shared memory: S slab slots of 10 KB + (R + 2) plane slots of three tiles (24 KB at 32 channels, 12 KB at 16)
  const size_t slot_b = (size_t)3 * 128 * p.kcy * 2;
  p.S = 4;
  p.R = p.kcy == 32 ? 5 : 8;
  const size_t smem = (size_t)p.S * kXSlot + (size_t)(p.R + kMirror) * slot_b + 1024 + 512;
  FM_CHECK(smem <= (size_t)kMaxDynSmemW2, FM_EINVAL, "conv3d wgrad march2: %zu B of shared memory", smem);
  static bool attr_set = false;
  if (!attr_set) {
    FM_CUDA(cudaFuncSetAttribute(conv3d_wgrad_march2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kMaxDynSmemW2));
    attr_set = true;
  }
  const double vox = (double)N * X * Y * Z;
  ProfScope prof(ctx, "conv3d_wgrad_march", 2.0 * 27 * Cin * Cout * vox, vox * (Cin + Cout) * 2.0);
  FM_CUDA(launch_pdl(conv3d_wgrad_march2_kernel, dim3(pairs * p.ctas_per_pair), dim3(kThreadsW2), smem, ctx->stream, p));
  FM_LAUNCH_OK(ctx);
  return FM_OK;
}
