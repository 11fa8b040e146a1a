// conv_march_shared.cu — TRAINING variant of conv_march.cu: the three MMA-issue warps accumulate into SHARED TMEM
// blocks (16-deep ring instead of 4, one TMEM read per plane; the epilogue hands blocks back cleared and every
// MMA accumulates). The fp32 summation order then follows the interleaving of the three issue streams, so
// results can differ in the last bits from run to run - fine for the training passes (cuDNN's default algorithms
// behave the same), while inference keeps the bit-reproducible kernel of conv_march.cu.
// "plane-marching" implicit-GEMM Conv3D 3x3x3 for layers whose whole filter bank fits
// in shared memory (the full-resolution layers that hold most of the U-Net's FLOPs).
//
// The per-tap kernel (conv_tc.cu) re-reads every activation tile 27x from L2 and, for Cout <= 64, is
// bound by the A-operand read from shared memory. This kernel removes both limits:
//
//   * a CTA owns a column of the volume: 16 (y) x 8 (z) output voxels = the 128 rows of one UMMA tile,
//     and marches along x. Per input plane it loads, once, three z-shifted y-haloed slabs
//     (18 x 8 rows x KC channels, one TMA box each; out-of-bounds rows zero-filled = 'same' padding).
//     A dy shift is an 8-row (= one swizzle atom) offset of the UMMA descriptor start address, a dz
//     shift selects the slab copy, so all 9 (dy,dz) taps read the same resident data.
//   * the three dx taps are stacked along N: one MMA with B = [W(dx=+1); W(dx=0); W(dx=-1)] (3*Cout
//     rows) accumulates plane xi into the TMEM accumulators of output planes xi-1, xi, xi+1 at once
//     (a ring of accumulator blocks), so each A byte read from shared memory feeds 3x the MMA work
//     and every input plane is consumed exactly once.
//   * weights are loaded to shared memory once per CTA and stay resident for its whole lifetime.
//
// fprop: epilogue = bias + ReLU.  dgrad: same kernel on dY with flipped/transposed weights, epilogue =
// ReLU mask of the producing block. Keras call site: Conv3D in create_convolution_block
// (fetal_net/model/unet3d/unet.py:102).
#include <algorithm>

#include "tc_ptx.cuh"

using namespace tcp;

namespace {

constexpr int kThreadsM = 256;  // warp 0 TMA, warps 1-3 MMA issue (one per dz slab copy), warps 4-7 epilogue
constexpr int kBY = 16, kBZ = 8, kSlabRows = (kBY + 2) * kBZ;  // 144 rows per slab
constexpr uint32_t kSlot = kSlabRows * 128;                    // 18432 B (1024-aligned)
constexpr int kMaxRing = 16;

struct alignas(64) MarchParams {
  CUtensorMap tmA[2];  // activations, box (KC, 8, 18, 1, 1)
  CUtensorMap tmW[2];  // march-packed weights, 2-D (KC, rows), box (KC, 3*Cn)
  int nsrc;
  int nchunks[2];
  int KC[2];
  uint32_t wofs[2];  // byte offset of the source's resident weights inside the W region
  int N, X, Y, Z;
  int ny, nz, nxc, xchunk, items;
  int Cn;  // output channels (MMA N of one accumulator block)
  int R;   // accumulator ring blocks
  int stages;
  int out_C, out_cofs, relu;
  uint32_t w_bytes;   // bytes the weight TMA loads deliver (mbarrier expect_tx)
  uint32_t w_region;  // shared-memory bytes reserved for them (1024-aligned per source)
  uint32_t slot;      // bytes per slab slot (1024-aligned)
  int debug;          // FETAL_B200_DEBUG ablation bits: 1 skip slab TMA, 2 skip MMAs, 4 skip epilogue body
  const float* bias;
  bf16* out;
  const bf16* mask;
};

__global__ void __launch_bounds__(kThreadsM, 1) conv3d_march_shared_kernel(const __grid_constant__ MarchParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t w_base = smem0;
  const uint32_t a_base = smem0 + p.w_region;
  const uint32_t bar0 = a_base + (uint32_t)p.stages * p.slot;
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (uint32_t)(2 * p.stages + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (uint32_t)(2 * p.stages + kMaxRing + b); };
  const uint32_t wfull_bar = bar0 + 8u * (uint32_t)(2 * p.stages + 2 * kMaxRing);
  const uint32_t tmem_slot = wfull_bar + 8u;

  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(p.R * p.Cn)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) {
      prefetch_tmap(&p.tmA[s]);
      prefetch_tmap(&p.tmW[s]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int b = 0; b < p.R; ++b) {
        mbar_init(tfull_bar(b), 3);     // one tcgen05.commit per MMA warp
        mbar_init(tempty_bar(b), 128);
      }
      mbar_init(wfull_bar, 1);
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

  auto decode = [&](int item, int& n, int& iy, int& iz, int& xa, int& xb) {
    const int jx = item % p.nxc;
    int t = item / p.nxc;
    iz = t % p.nz;
    t /= p.nz;
    iy = t % p.ny;
    n = t / p.ny;
    xa = jx * p.xchunk;
    xb = min(p.X, xa + p.xchunk);
  };

  const int dbg = p.debug;
  // make the values the producer / MMA warps compute on provably warp-uniform
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);

  if (warp_u == 0) {
    // ===== producer warp (all lanes converged, one elected lane issues): resident weights once, then one
    //       slab per (plane, dz, source, chunk) =====
    mbar_expect_tx_elect(wfull_bar, p.w_bytes);
    for (int s = 0; s < p.nsrc; ++s) {
      const uint32_t tile = 3u * (uint32_t)p.Cn * (uint32_t)p.KC[s] * 2u;
      for (int t = 0; t < p.nchunks[s] * 9; ++t)
        tma_load_2d_elect(w_base + p.wofs[s] + (uint32_t)t * tile, &p.tmW[s], wfull_bar, 0, t * 3 * p.Cn);
    }
    // the weights do not depend on the previous kernel in the stream; the activation slabs do
    pdl_wait();
    pdl_launch_dependents();
    // every dz slab copy has its own ring of S3 = stages/3 slots, consumed strictly in order by "its" MMA
    // warp (a consumer that skipped slots of a shared ring could not tell mbarrier phases apart)
    const uint32_t S3 = (uint32_t)p.stages / 3u;
    uint32_t sidx[3] = {0u, 0u, 0u}, sph[3] = {0u, 0u, 0u};
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n, iy, iz, xa, xb;
      decode(item, n, iy, iz, xa, xb);
      const int x_first = max(xa - 1, 0), x_last = min(xb, p.X - 1);
      const int yc = iy * kBY - 1, zc = iz * kBZ - 1;
      for (int xi = x_first; xi <= x_last; ++xi) {
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
          for (int s = 0; s < p.nsrc; ++s) {
            const uint32_t bytes = (uint32_t)kSlabRows * (uint32_t)p.KC[s] * 2u;
            for (int ch = 0; ch < p.nchunks[s]; ++ch) {
              const uint32_t stage = (uint32_t)dz * S3 + sidx[dz];
              mbar_wait(empty_bar(stage), sph[dz] ^ 1u);
              if (dbg & 1) {
                mbar_expect_tx_elect(full_bar(stage), 0);
              } else {
                mbar_expect_tx_elect(full_bar(stage), bytes);
                tma_load_5d_elect(a_base + stage * p.slot, &p.tmA[s], full_bar(stage), ch * p.KC[s], zc + dz, yc, xi, n);
              }
              if (++sidx[dz] == S3) {
                sidx[dz] = 0;
                sph[dz] ^= 1u;
              }
            }
          }
      }
    }
  } else if (warp_u <= 3) {
    // ===== MMA warps 1..3: warp w owns the slab copy dz = w - 1 of every input plane. A single lane cannot
    // issue one 48-cycle MMA every 48 cycles (each issue is ~15 dependent uniform-datapath instructions),
    // so the issue stream is split three ways. All MMAs accumulate (the epilogue hands accumulator blocks
    // back zeroed), which makes the result independent of the interleaving of the three issue streams. =====
    const int dz = warp_u - 1;
    const uint32_t Cn = (uint32_t)p.Cn;
    const uint32_t ring_mask = (uint32_t)p.R - 1u;  // R is a power of two
    const uint32_t ring_shift = 31u - (uint32_t)__clz(p.R);
    const uint32_t idesc1 = make_idesc(128, (int)Cn, 0, 0);
    const uint32_t idesc2 = make_idesc(128, 2 * (int)Cn, 0, 0);
    const uint32_t idesc3 = make_idesc(128, 3 * (int)Cn, 0, 0);
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    // this warp's private ring of S3 slots
    const uint32_t S3 = (uint32_t)p.stages / 3u, slot0 = (uint32_t)dz * S3;
    uint32_t sidx = 0, ph = 0, ocount = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n, iy, iz, xa, xb;
      decode(item, n, iy, iz, xa, xb);
      const int x_first = max(xa - 1, 0), x_last = min(xb, p.X - 1);
      for (int xi = x_first; xi <= x_last; ++xi) {
        const int lo = max(xa, xi - 1), hi = min(xb - 1, xi + 1);  // output planes fed by plane xi
        const uint32_t j_lo = (uint32_t)(lo - (xi - 1));
        const uint32_t nblk = (uint32_t)(hi - lo + 1);
        const uint32_t seq_lo = ocount + (uint32_t)(lo - xa);
        const uint32_t rb_lo = seq_lo & ring_mask;
        // blocks first touched by this plane must have been drained (and zeroed) by the epilogue
        for (uint32_t j = 0; j < nblk; ++j) {
          if (xi == x_first || lo + (int)j == xi + 1) {
            const uint32_t seq = seq_lo + j;
            mbar_wait(tempty_bar(seq & ring_mask), (seq >> ring_shift) & 1u);
          }
        }
        tc_fence_after();
        // the <= 3 consecutive ring blocks, split only where the ring wraps
        const uint32_t nA = min(nblk, (uint32_t)p.R - rb_lo), nB = nblk - nA;
        const uint32_t colA = tmem_base + rb_lo * Cn, colB = tmem_base;
        const uint32_t idA = nA == 1 ? idesc1 : (nA == 2 ? idesc2 : idesc3);
        const uint32_t idB = nB == 1 ? idesc1 : idesc2;
        for (int s = 0; s < p.nsrc; ++s) {
          const uint32_t row_bytes = (uint32_t)p.KC[s] * 2u;
          const uint32_t sbo = 8u * row_bytes;
          const uint32_t hi32 = desc_hi(sbo, layout_code((int)row_bytes));
          const uint32_t btile16 = (3u * Cn * row_bytes) >> 4;  // B tile stride, 16-byte units
          const uint32_t blk16 = (Cn * row_bytes) >> 4;         // one N block of B rows
          const uint32_t dy16 = sbo >> 4;                       // one y row = 8 slab rows
          const int nk = p.KC[s] >> 4;
          const uint32_t b_lo0 = desc_lo(w_base + p.wofs[s], 16u) + (uint32_t)(dz * 3) * btile16 + j_lo * blk16;
          for (int ch = 0; ch < p.nchunks[s]; ++ch) {
            const uint32_t stage = slot0 + sidx;
            mbar_wait(full_bar(stage), ph);
            tc_fence_after();
            uint32_t a_lo = desc_lo(a_base + stage * p.slot, 16u);
            uint32_t b_lo = b_lo0 + (uint32_t)(ch * 9) * btile16;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              for (int k = 0; k < nk; ++k) {
                const uint32_t ak = a_lo + 2u * (uint32_t)k, bk = b_lo + 2u * (uint32_t)k;
                umma_bf16_lh_elect(colA, ak, hi32, bk, hi32, idA, 1u);
                if (nB) umma_bf16_lh_elect(colB, ak, hi32, bk + nA * blk16, hi32, idB, 1u);
              }
              a_lo += dy16;
              b_lo += btile16;
            }
            umma_commit_elect(empty_bar(stage));
            if (++sidx == S3) {
              sidx = 0;
              ph ^= 1u;
            }
          }
        }
        // output planes completed by this input plane (each MMA warp contributes one arrival)
        if (xi - 1 >= xa) umma_commit_elect(tfull_bar((ocount + (uint32_t)(xi - 1 - xa)) & ring_mask));
        if (xi == x_last && xi <= xb - 1) umma_commit_elect(tfull_bar((ocount + (uint32_t)(xi - xa)) & ring_mask));
      }
      ocount += (uint32_t)(xb - xa);
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int yl = row >> 3, zl = row & 7;
    uint32_t ocount = 0;
    // hand every accumulator block to the MMA warps zeroed (TMEM is not initialised by the allocation)
    for (int b = 0; b < p.R; ++b) {
      for (int c16 = 0; c16 < p.Cn / 16; ++c16)
        tmem_st16_zero(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * p.Cn + c16 * 16));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(tempty_bar(b));
    }
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n, iy, iz, xa, xb;
      decode(item, n, iy, iz, xa, xb);
      const int y = iy * kBY + yl, z = iz * kBZ + zl;
      for (int xo = xa; xo < xb; ++xo) {
        const uint32_t seq = ocount + (uint32_t)(xo - xa);
        const uint32_t rb = seq & ((uint32_t)p.R - 1u);
        const int64_t v = (((int64_t)n * p.X + xo) * p.Y + y) * p.Z + z;
        const int64_t off = v * p.out_C + p.out_cofs;
        const int ncg = (dbg & 4) ? 0 : p.Cn / 16;
        // dgrad: fetch this row's ReLU mask BEFORE waiting for the accumulator, so its latency hides behind the MMAs
        uint4 mk[8];
        if (p.mask != nullptr) {
          const uint4* mp = reinterpret_cast<const uint4*>(p.mask + off);
#pragma unroll
          for (int h = 0; h < 8; ++h)
            if (h < 2 * ncg) mk[h] = __ldg(mp + h);
        }
        mbar_wait(tfull_bar(rb), (seq >> (31u - (uint32_t)__clz(p.R))) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + rb * (uint32_t)p.Cn;
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          if (c16 >= ncg) break;
          uint32_t r[16];
          tmem_ld16(taddr + (uint32_t)c16 * 16u, r);
          tmem_ld_wait();
          tmem_st16_zero(taddr + (uint32_t)c16 * 16u);
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(r[j]);
          if (p.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + c16 * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b4 = __ldg(bp + j);
              f[4 * j] += b4.x;
              f[4 * j + 1] += b4.y;
              f[4 * j + 2] += b4.z;
              f[4 * j + 3] += b4.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (p.mask != nullptr) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint4 mv = mk[2 * c16 + h];
              const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mv);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 mf = __bfloat1622float2(mh[j]);
                if (!(mf.x > 0.f)) f[8 * h + 2 * j] = 0.f;
                if (!(mf.y > 0.f)) f[8 * h + 2 * j + 1] = 0.f;
              }
            }
          }
          uint4 o[2];
          __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
          for (int j = 0; j < 8; ++j) oh[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          uint4* op = reinterpret_cast<uint4*>(p.out + off + c16 * 16);
          op[0] = o[0];
          op[1] = o[1];
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(tempty_bar(rb));
      }
      ocount += (uint32_t)(xb - xa);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_m() {
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
CUtensorMapSwizzle swz(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
bool chan_ok_m(int C) { return C == 16 || C == 32 || C == 64 || (C > 64 && C % 64 == 0); }
const int kMaxDynSmemM = 227 * 1024;

// resident filter bank: 27 taps * Cin * Cout * 2 B per source, each source region 1024-aligned
uint32_t march_w_region(int C1, int C2, int Cn) {
  const uint32_t a = ((uint32_t)27 * C1 * Cn * 2u + 1023u) & ~1023u;
  const uint32_t b = ((uint32_t)27 * C2 * Cn * 2u + 1023u) & ~1023u;
  return a + b;
}

}  // namespace

int k_conv3d_march_shared(fm_ctx* ctx, const bf16* x1, const bf16* x2, const bf16* wm1, const bf16* wm2,
                   const float* bias, bf16* y, const bf16* mask, int N, int X, int Y, int Z, int C1, int C2,
                   int Cout, int relu, int out_C, int out_cofs) {
  if (!fm_march_v1())
    return k_conv3d_march2(ctx, x1, x2, wm1, wm2, bias, y, mask, N, X, Y, Z, C1, C2, Cout, relu, out_C, out_cofs, 3);
  FM_CHECK(conv_march_supported(X, Y, Z, C1, C2, Cout, 3), FM_EINVAL,
           "conv3d march: unsupported shape %dx%dx%d C1=%d C2=%d Cout=%d", X, Y, Z, C1, C2, Cout);
  PFN_encodeTiled enc = get_encode_m();
  FM_CHECK(enc != nullptr, FM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  MarchParams p;
  memset(&p, 0, sizeof(p));
  p.nsrc = C2 > 0 ? 2 : 1;
  p.N = N;
  p.X = X;
  p.Y = Y;
  p.Z = Z;
  p.ny = Y / kBY;
  p.nz = Z / kBZ;
  p.Cn = Cout;
  p.R = std::min(kMaxRing, 512 / Cout);
  p.out_C = out_C;
  p.out_cofs = out_cofs;
  p.relu = relu;
  p.bias = bias;
  p.out = y;
  p.mask = mask;
  // x-chunking: trade wave quantisation against the 2 halo planes each chunk re-loads
  {
    const int cols = N * p.ny * p.nz;
    double best = -1.0;
    int best_nxc = 1;
    for (int nxc = 1; nxc <= 16; nxc *= 2) {
      const int xc = ceil_div(X, nxc);
      if (xc < 4 && nxc > 1) break;
      const int items = cols * ceil_div(X, xc);
      const int waves = ceil_div(items, ctx->num_sms);
      const double eff = (double)items / ((double)waves * ctx->num_sms) * (double)xc / (double)(xc + 1.4);
      if (eff > best) {
        best = eff;
        best_nxc = nxc;
      }
    }
    p.xchunk = ceil_div(X, best_nxc);
    p.nxc = ceil_div(X, p.xchunk);
    p.items = cols * p.nxc;
  }
  const int Cs[2] = {C1, C2};
  const bf16* xs[2] = {x1, x2};
  const bf16* wms[2] = {wm1, wm2};
  uint32_t wofs = 0, wbytes = 0;
  for (int s = 0; s < p.nsrc; ++s) {
    const int KC = conv_march_kc(C1, C2, Cout, Cs[s]);
    p.KC[s] = KC;
    p.nchunks[s] = Cs[s] / KC;
    p.wofs[s] = wofs;
    wofs += ((uint32_t)27 * Cs[s] * Cout * 2u + 1023u) & ~1023u;
    wbytes += (uint32_t)27 * Cs[s] * Cout * 2u;
    {
      cuuint64_t dims[5] = {(cuuint64_t)Cs[s], (cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)N};
      cuuint64_t strides[4] = {(cuuint64_t)Cs[s] * 2, (cuuint64_t)Z * Cs[s] * 2, (cuuint64_t)Y * Z * Cs[s] * 2,
                               (cuuint64_t)X * Y * Z * Cs[s] * 2};
      cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)kBZ, (cuuint32_t)(kBY + 2), 1, 1};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&p.tmA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)xs[s], dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz(KC * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      FM_CHECK(r == CUDA_SUCCESS, FM_ECUDA, "cuTensorMapEncodeTiled(march act) failed: %d", (int)r);
    }
    {
      const cuuint64_t rows = (cuuint64_t)p.nchunks[s] * 27 * Cout;
      cuuint64_t dims[2] = {(cuuint64_t)KC, rows};
      cuuint64_t strides[1] = {(cuuint64_t)KC * 2};
      cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)(3 * Cout)};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&p.tmW[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wms[s], dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz(KC * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      FM_CHECK(r == CUDA_SUCCESS, FM_ECUDA, "cuTensorMapEncodeTiled(march weights) failed: %d", (int)r);
    }
  }
  {
    const char* e = getenv("FETAL_B200_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }
  p.w_bytes = wbytes;
  p.w_region = wofs;
  const uint32_t w_region = wofs;
  // slot = one slab at the widest K chunk in use (18 KB at KC = 64, 9 KB when every source runs at KC <= 32)
  p.slot = (uint32_t)kSlabRows * (uint32_t)std::max(p.KC[0], p.nsrc > 1 ? p.KC[1] : 0) * 2u;
  p.slot = (p.slot + 1023u) & ~1023u;
  int stages = (kMaxDynSmemM - 2048 - (int)w_region) / (int)p.slot;
  stages = std::min(stages, 9) / 3 * 3;  // three private rings (one per dz slab copy / MMA warp)
  FM_CHECK(stages >= 3, FM_EINVAL, "conv3d march: filter bank leaves no room for the slab rings");
  p.stages = stages;
  const size_t smem = (size_t)w_region + (size_t)stages * p.slot + 1024 + 512;
  FM_CHECK(smem <= (size_t)kMaxDynSmemM, FM_EINVAL, "conv3d march: %zu B of shared memory needed", smem);
  static bool attr_set = false;
  if (!attr_set) {
    FM_CUDA(cudaFuncSetAttribute(conv3d_march_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmemM));
    attr_set = true;
  }
  const int grid = std::min(p.items, ctx->num_sms);
  const double vox = (double)N * X * Y * Z;
  ProfScope prof(ctx, mask != nullptr || bias == nullptr ? "conv3d_march_dgrad" : "conv3d_march_fprop",
                 2.0 * 27 * (C1 + C2) * Cout * vox, vox * (C1 + C2 + Cout) * 2.0);
  FM_CUDA(launch_pdl(conv3d_march_shared_kernel, dim3(grid), dim3(kThreadsM), smem, ctx->stream, p));
  FM_LAUNCH_OK(ctx);
  return FM_OK;
}
