// sampler.cu — training patches cut on the device.
//
// The reference's training generator (fetal_net/generator.py:222-348) keeps every case in host memory, draws a random
// patch corner per sample (np.random.randint, generator.py:266-269), slices the data patch, the truth slice(s) at
// `truth_index` and - for the 2.5D models - the previous-truth slice(s) at `prev_truth_index` with
// get_patch_from_3d_data (edge-replicated where a slice sticks out, utils/patches.py:57-91), concatenates the previous
// truth to the data on the last axis (generator.py:305-306) and stacks a batch (convert_data, generator.py:380-401).
// At a few milliseconds per training step that host loop (and the H2D copy of every batch) is the bottleneck, so here
// the cases live in HBM (float32, the dtype Keras feeds) and ONE kernel cuts the whole batch straight into the
// model's input / target buffers. The random decisions stay on the host (fetal_net/device_sampler.py draws case order,
// corners and the skip-blank rejections with the reference's own np.random call sequence), so the device only needs
// (case, corner) per sample: index mapping bit-exact by construction, values bit-exact float32 copies.
//
// Optional cheap augmentations on the way (NOT the reference's nilearn / imgaug pipeline, generator.py:271-295, which
// stays on the host through the overlay): axis flips, an intensity scale, additive Gaussian noise from a counter-based
// hash - all per sample, parameters drawn by the host.
#include <vector>

#include "common.cuh"

struct fm_volset {
  fm_ctx* ctx = nullptr;
  struct Case {
    float *data = nullptr, *truth = nullptr;
    int32_t dims[3] = {0, 0, 0};
  };
  std::vector<Case> cases;
  // device table of the cases (pointers + dims), rebuilt when a case changes
  struct DevCase {
    const float *data, *truth;
    int32_t dims[3];
    int32_t pad;
  };
  DevCase* table = nullptr;
  bool table_dirty = true;
  int32_t* d_args = nullptr;  // per-sample (case, corner[3], flip bits) + float (scale, sigma) staging on the device
  size_t d_args_cap = 0;
};

namespace {

constexpr int kThreadsS = 256;

struct SampleGeom {
  int32_t patch[3];          // P0, P1, P2 of the data patch
  int32_t truth_index, truth_size, prev_index, prev_size;  // prev_size == 0: no previous-truth channels
  int32_t x_pitch;           // P2 + prev_size (last-axis extent of the network input)
};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// standard normal from two counter-based uniforms (Box-Muller)
__device__ __forceinline__ float gauss(uint32_t seed, uint32_t idx) {
  const uint32_t a = hash32(idx * 2u + seed), b = hash32(idx * 2u + 1u + seed * 0x9e3779b9u);
  const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777217.0f), u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2);
}

// args per sample: int32 [8] = {case, cx, cy, cz, flip bits (1: x, 2: y, 4: z), noise seed, bits of scale, bits of sigma}
// x: [B][P0][P1][P2 + prev_size] (data patch | previous-truth slices), y: [B][P0][P1][truth_size]
// Batches of up to kInlineSamples samples carry their arguments in the kernel parameters (no staging buffer, no
// synchronisation with the previous step); larger ones read them from `args`.
constexpr int kInlineSamples = 64;
struct SampleArgs {
  int32_t v[kInlineSamples * 8];
};
__global__ void __launch_bounds__(kThreadsS) sample_patches_kernel(const fm_volset::DevCase* __restrict__ cases,
                                                                   const int32_t* __restrict__ args,
                                                                   const __grid_constant__ SampleArgs inl, SampleGeom gm,
                                                                   float* __restrict__ x, float* __restrict__ y) {
  const int b = blockIdx.y;
  const int32_t* a = args != nullptr ? args + b * 8 : inl.v + b * 8;
  const fm_volset::DevCase cs = cases[a[0]];
  const int cx = a[1], cy = a[2], cz = a[3], flip = a[4];
  const uint32_t seed = (uint32_t)a[5];
  const float scale = __int_as_float(a[6]), sigma = __int_as_float(a[7]);
  const int P0 = gm.patch[0], P1 = gm.patch[1], P2 = gm.patch[2];
  const int rowx = gm.x_pitch, rowy = gm.truth_size;
  const int64_t nx = (int64_t)P0 * P1 * rowx, ny = (int64_t)P0 * P1 * rowy;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nx + ny; g += (int64_t)gridDim.x * blockDim.x) {
    const bool is_y = g >= nx;
    const int64_t q = is_y ? g - nx : g;
    const int row = is_y ? rowy : rowx;
    const int k = (int)(q % row);
    const int64_t ij = q / row;
    int j = (int)(ij % P1), i = (int)(ij / P1);
    // flips act on the assembled sample (data, previous truth and target alike), like flipping the arrays on the host
    if (flip & 1) i = P0 - 1 - i;
    if (flip & 2) j = P1 - 1 - j;
    const float* src;
    int zz;
    bool is_data = false;
    if (is_y) {
      src = cs.truth;
      zz = cz + gm.truth_index + ((flip & 4) ? rowy - 1 - k : k);
    } else if (k < P2) {
      src = cs.data;
      zz = cz + ((flip & 4) ? P2 - 1 - k : k);
      is_data = true;
    } else {
      src = cs.truth;
      const int kk = k - P2;
      zz = cz + gm.prev_index + ((flip & 4) ? gm.prev_size - 1 - kk : kk);
    }
    // get_patch_from_3d_data completes a patch that sticks out with the nearest edge sample (np.pad(mode='edge')):
    // clamping the coordinates yields the same values
    const int xx = min(max(cx + i, 0), cs.dims[0] - 1);
    const int yy = min(max(cy + j, 0), cs.dims[1] - 1);
    zz = min(max(zz, 0), cs.dims[2] - 1);
    float v = __ldg(src + ((int64_t)xx * cs.dims[1] + yy) * cs.dims[2] + zz);
    if (is_data) {
      v *= scale;
      if (sigma > 0.f) v += sigma * gauss(seed, (uint32_t)q);
    }
    (is_y ? y + (int64_t)b * ny : x + (int64_t)b * nx)[q] = v;
  }
}

int build_table(fm_volset* s) {
  if (!s->table_dirty) return FM_OK;
  std::vector<fm_volset::DevCase> host(s->cases.size());
  for (size_t i = 0; i < s->cases.size(); ++i) {
    host[i].data = s->cases[i].data;
    host[i].truth = s->cases[i].truth;
    for (int a = 0; a < 3; ++a) host[i].dims[a] = s->cases[i].dims[a];
    host[i].pad = 0;
  }
  if (!s->table) FM_CUDA(cudaMalloc((void**)&s->table, host.size() * sizeof(fm_volset::DevCase)));
  FM_CUDA(cudaMemcpy(s->table, host.data(), host.size() * sizeof(fm_volset::DevCase), cudaMemcpyHostToDevice));
  s->table_dirty = false;
  return FM_OK;
}

}  // namespace

extern "C" int fm_volset_create(fm_ctx* ctx, int n_cases, fm_volset** out) {
  FM_CHECK(ctx && out && n_cases > 0, FM_EINVAL, "fm_volset_create: bad argument");
  fm_volset* s = new fm_volset();
  s->ctx = ctx;
  s->cases.resize((size_t)n_cases);
  *out = s;
  return FM_OK;
}

extern "C" int fm_volset_destroy(fm_volset* s) {
  if (!s) return FM_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  for (auto& c : s->cases) {
    if (c.data) cudaFree(c.data);
    if (c.truth) cudaFree(c.truth);
  }
  if (s->table) cudaFree(s->table);
  if (s->d_args) cudaFree(s->d_args);
  delete s;
  return FM_OK;
}

extern "C" int fm_volset_set_case(fm_volset* s, int index, const float* data, const float* truth, const int32_t dims[3]) {
  FM_CHECK(s && data && truth && dims && index >= 0 && index < (int)s->cases.size(), FM_EINVAL,
           "fm_volset_set_case: bad argument");
  FM_CHECK(dims[0] > 0 && dims[1] > 0 && dims[2] > 0, FM_EINVAL, "fm_volset_set_case: empty volume");
  FM_CUDA(cudaSetDevice(s->ctx->device));
  FM_CUDA(cudaStreamSynchronize(s->ctx->stream));
  fm_volset::Case& c = s->cases[(size_t)index];
  if (c.data) cudaFree(c.data);
  if (c.truth) cudaFree(c.truth);
  c.data = c.truth = nullptr;
  const size_t n = (size_t)dims[0] * dims[1] * dims[2];
  FM_CUDA(cudaMalloc((void**)&c.data, n * sizeof(float)));
  FM_CUDA(cudaMalloc((void**)&c.truth, n * sizeof(float)));
  FM_CUDA(cudaMemcpy(c.data, data, n * sizeof(float), cudaMemcpyHostToDevice));
  FM_CUDA(cudaMemcpy(c.truth, truth, n * sizeof(float), cudaMemcpyHostToDevice));
  for (int a = 0; a < 3; ++a) c.dims[a] = dims[a];
  s->table_dirty = true;
  return FM_OK;
}

// Cuts `batch` samples into x_dev [B][P0][P1][P2 + prev_size] and y_dev [B][P0][P1][truth_size] (device pointers).
int sampler_gather_device(fm_volset* s, const int32_t* cases, const int32_t* corners, const fm_sample_aug* aug, int batch,
                          const int32_t patch[3], int truth_index, int truth_size, int prev_truth_index,
                          int prev_truth_size, float* x_dev, float* y_dev) {
  FM_CHECK(s && cases && corners && patch && batch > 0 && x_dev && y_dev, FM_EINVAL, "sampler: bad argument");
  FM_CHECK(truth_size > 0 && prev_truth_size >= 0 && patch[0] > 0 && patch[1] > 0 && patch[2] > 0, FM_EINVAL,
           "sampler: bad patch / truth extents");
  fm_ctx* ctx = s->ctx;
  FM_CUDA(cudaSetDevice(ctx->device));
  for (int b = 0; b < batch; ++b) {
    FM_CHECK(cases[b] >= 0 && cases[b] < (int)s->cases.size() && s->cases[(size_t)cases[b]].data != nullptr, FM_EINVAL,
             "sampler: case %d of sample %d is not loaded", cases[b], b);
  }
  FM_TRY(build_table(s));
  // per-sample arguments: inside the kernel parameters for ordinary batches (the call never waits for the previous
  // step, which is still in its backward pass when the next batch is drawn); through a pinned staging area otherwise
  SampleArgs inl;
  const bool inline_args = batch <= kInlineSamples;
  void* pin = nullptr;
  if (!inline_args) {
    FM_CUDA(cudaStreamSynchronize(ctx->stream));  // the previous step's copy out of the staging area has completed
    FM_TRY(fm_ctx_pinned(ctx, (size_t)batch * 32, &pin));
  }
  int32_t* h = inline_args ? inl.v : (int32_t*)pin;
  for (int b = 0; b < batch; ++b) {
    h[b * 8] = cases[b];
    h[b * 8 + 1] = corners[b * 3];
    h[b * 8 + 2] = corners[b * 3 + 1];
    h[b * 8 + 3] = corners[b * 3 + 2];
    const float scale = aug ? aug[b].intensity_scale : 1.0f, sigma = aug ? aug[b].noise_sigma : 0.0f;
    h[b * 8 + 4] = aug ? (int32_t)(aug[b].flip & 7u) : 0;
    h[b * 8 + 5] = aug ? (int32_t)aug[b].noise_seed : 0;
    memcpy(&h[b * 8 + 6], &scale, 4);
    memcpy(&h[b * 8 + 7], &sigma, 4);
  }
  if (!inline_args) {
    if (s->d_args_cap < (size_t)batch * 8) {
      if (s->d_args) cudaFree(s->d_args);
      s->d_args = nullptr;
      FM_CUDA(cudaMalloc((void**)&s->d_args, (size_t)batch * 32));
      s->d_args_cap = (size_t)batch * 8;
    }
    FM_CUDA(cudaMemcpyAsync(s->d_args, h, (size_t)batch * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  SampleGeom gm;
  for (int a = 0; a < 3; ++a) gm.patch[a] = patch[a];
  gm.truth_index = truth_index;
  gm.truth_size = truth_size;
  gm.prev_index = prev_truth_index;
  gm.prev_size = prev_truth_size;
  gm.x_pitch = patch[2] + prev_truth_size;
  const int64_t per = (int64_t)patch[0] * patch[1] * (gm.x_pitch + truth_size);
  const dim3 grid((unsigned)std::min<int64_t>(ceil_div64(per, kThreadsS), 4096), (unsigned)batch);
  ProfScope prof(ctx, "sample_patches", 0.0, (double)batch * per * 8.0);
  sample_patches_kernel<<<grid, kThreadsS, 0, ctx->stream>>>(s->table, inline_args ? nullptr : s->d_args, inl, gm, x_dev,
                                                             y_dev);
  FM_LAUNCH_OK(ctx);
  return FM_OK;
}

extern "C" int fm_volset_gather(fm_volset* s, const int32_t* cases, const int32_t* corners, const fm_sample_aug* aug,
                                int batch, const int32_t patch[3], int truth_index, int truth_size, int prev_truth_index,
                                int prev_truth_size, float* x_out, float* y_out) {
  FM_CHECK(s && x_out && y_out && patch && batch > 0, FM_EINVAL, "fm_volset_gather: bad argument");
  FM_CUDA(cudaSetDevice(s->ctx->device));
  const size_t nx = (size_t)batch * patch[0] * patch[1] * (patch[2] + prev_truth_size);
  const size_t ny = (size_t)batch * patch[0] * patch[1] * truth_size;
  float *dx = nullptr, *dy = nullptr;
  FM_CUDA(cudaMalloc((void**)&dx, nx * sizeof(float)));
  cudaError_t e = cudaMalloc((void**)&dy, ny * sizeof(float));
  int rc = e == cudaSuccess ? FM_OK : FM_ENOMEM;
  if (rc == FM_OK)
    rc = sampler_gather_device(s, cases, corners, aug, batch, patch, truth_index, truth_size, prev_truth_index,
                               prev_truth_size, dx, dy);
  if (rc == FM_OK) {
    e = cudaMemcpyAsync(x_out, dx, nx * sizeof(float), cudaMemcpyDeviceToHost, s->ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(y_out, dy, ny * sizeof(float), cudaMemcpyDeviceToHost, s->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->ctx->stream);
    if (e != cudaSuccess) {
      fm_set_error("fm_volset_gather: %s", cudaGetErrorString(e));
      rc = FM_ECUDA;
    }
  }
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(dx);
  if (dy) cudaFree(dy);
  return rc;
}
