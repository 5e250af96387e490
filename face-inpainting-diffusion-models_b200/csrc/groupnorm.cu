// K2: GroupNorm(32) [* (1+scale) + shift] [SiLU] [2x avg-pool | 2x nearest-up] on NHWC tensors.
//
// Replaces nn.GroupNorm + nn.SiLU (nn.py:46-48,151-152,173-174; unet.py:149-150; nn.py:251), the
// scale/shift modulation (nn.py:203-206) and the h_upd / x_upd resampling of up/down ResBlocks
// (nn.py:190-195).  HBM-bound: 2 B read (stats pass, normally L2-resident for the second pass) +
// 2 B read + 2 B write per bf16 element.
//
// Thread mapping (both passes): a thread owns one VEC-wide channel vector and walks pixels, so a
// warp reads consecutive 16-byte vectors of one pixel (fully coalesced) and keeps its per-channel
// affine coefficients in registers.  Statistics are accumulated in fp32 per thread, reduced per
// group in shared memory, and added to the [batch][groups][2] workspace in double precision.
#include "common.cuh"

namespace fidm {

struct GnParams {
  fidm_gn_args a;
  int cv;          // channel vectors per pixel
  int ppi;         // pixels processed per block iteration
  int pix_per_blk; // pixels (of the iteration space) per block
};

template <typename T, int VEC>
__global__ void gn_stats_kernel(const GnParams p) {
  const fidm_gn_args& a = p.a;
  __shared__ float red[64][2];
  const int n = blockIdx.y;
  const int hw = a.height * a.width;
  const int v = threadIdx.x % p.cv;
  const int pl = threadIdx.x / p.cv;
  const int cpg = a.channels / a.groups;
  for (int i = threadIdx.x; i < a.groups; i += blockDim.x) red[i][0] = red[i][1] = 0.0f;
  __syncthreads();
  const int p0 = blockIdx.x * p.pix_per_blk;
  const int p1 = min(hw, p0 + p.pix_per_blk);
  const T* base = reinterpret_cast<const T*>(a.x) + (long long)n * hw * a.ld_x + v * VEC;
  float s = 0.0f, ss = 0.0f;
  for (int px = p0 + pl; px < p1; px += p.ppi) {
    float f[VEC];
    load_vec<T, VEC>(base + (long long)px * a.ld_x, f);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      s += f[i];
      ss = fmaf(f[i], f[i], ss);
    }
  }
  const int g = (v * VEC) / cpg;
  atomicAdd(&red[g][0], s);
  atomicAdd(&red[g][1], ss);
  __syncthreads();
  for (int i = threadIdx.x; i < a.groups * 2; i += blockDim.x)
    atomicAdd(&a.stats[((long long)n * a.groups) * 2 + i], (double)red[i >> 1][i & 1]);
}

template <bool FAST>
__device__ __forceinline__ float silu_f(float v) {
  if (FAST) return __fdividef(v, 1.0f + __expf(-v));
  return v / (1.0f + expf(-v));
}

template <typename T, int VEC, int RESAMPLE>
__global__ void gn_apply_kernel(const GnParams p) {
  const fidm_gn_args& a = p.a;
  constexpr bool FAST = (sizeof(T) == 2);
  const int n = blockIdx.y;
  const int H = a.height, W = a.width, hw = H * W;
  const int v = threadIdx.x % p.cv;
  const int pl = threadIdx.x / p.cv;
  const int cpg = a.channels / a.groups;
  const int c0 = v * VEC;
  const int g = c0 / cpg;

  // per-thread affine coefficients  y = x * A + B
  float A[VEC], B[VEC];
  if (a.skip_norm) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) { A[i] = 1.0f; B[i] = 0.0f; }
  } else {
    const double cnt = (double)cpg * hw;
    const double s = a.stats[((long long)n * a.groups + g) * 2 + 0];
    const double ss = a.stats[((long long)n * a.groups + g) * 2 + 1];
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const float meanf = (float)mean;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float ga = a.gamma ? a.gamma[c0 + i] : 1.0f;
      const float be = a.beta ? a.beta[c0 + i] : 0.0f;
      float Ai = rstd * ga;
      float Bi = be - meanf * Ai;
      if (a.scale_shift) {
        const float sc = 1.0f + a.scale_shift[(long long)n * a.ld_ss + c0 + i];
        const float sh = a.scale_shift[(long long)n * a.ld_ss + a.channels + c0 + i];
        Ai *= sc;
        Bi = Bi * sc + sh;
      }
      A[i] = Ai;
      B[i] = Bi;
    }
  }
  const T* xin = reinterpret_cast<const T*>(a.x) + (long long)n * hw * a.ld_x + c0;
  T* yo = reinterpret_cast<T*>(a.y);
  T* yr = reinterpret_cast<T*>(a.y_raw);

  if (RESAMPLE == FIDM_RESAMPLE_NONE) {
    const int p0 = blockIdx.x * p.pix_per_blk, p1 = min(hw, p0 + p.pix_per_blk);
    for (int px = p0 + pl; px < p1; px += p.ppi) {
      float f[VEC];
      load_vec<T, VEC>(xin + (long long)px * a.ld_x, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float y = fmaf(f[i], A[i], B[i]);
        f[i] = a.silu ? silu_f<FAST>(y) : y;
      }
      store_vec<T, VEC>(yo + ((long long)n * hw + px) * a.ld_y + c0, f);
    }
  } else if (RESAMPLE == FIDM_RESAMPLE_DOWN) {
    const int Ho = H / 2, Wo = W / 2, ohw = Ho * Wo;
    const int p0 = blockIdx.x * p.pix_per_blk, p1 = min(ohw, p0 + p.pix_per_blk);
    for (int px = p0 + pl; px < p1; px += p.ppi) {
      const int ho = px / Wo, wo = px % Wo;
      float acc[VEC], raw[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = raw[i] = 0.0f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int ip = (2 * ho + (d >> 1)) * W + 2 * wo + (d & 1);
        float f[VEC];
        load_vec<T, VEC>(xin + (long long)ip * a.ld_x, f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          raw[i] += f[i];
          float y = fmaf(f[i], A[i], B[i]);
          acc[i] += a.silu ? silu_f<FAST>(y) : y;
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        acc[i] *= 0.25f;
        raw[i] *= 0.25f;
      }
      store_vec<T, VEC>(yo + ((long long)n * ohw + px) * a.ld_y + c0, acc);
      if (yr) store_vec<T, VEC>(yr + ((long long)n * ohw + px) * a.ld_raw + c0, raw);
    }
  } else {  // nearest 2x up: one input pixel feeds four outputs
    const int Wo = 2 * W, ohw = 4 * hw;
    const int p0 = blockIdx.x * p.pix_per_blk, p1 = min(hw, p0 + p.pix_per_blk);
    for (int px = p0 + pl; px < p1; px += p.ppi) {
      const int h = px / W, w = px % W;
      float f[VEC], y[VEC];
      load_vec<T, VEC>(xin + (long long)px * a.ld_x, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float t = fmaf(f[i], A[i], B[i]);
        y[i] = a.silu ? silu_f<FAST>(t) : t;
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const long long op = (long long)n * ohw + (long long)(2 * h + (d >> 1)) * Wo + 2 * w + (d & 1);
        store_vec<T, VEC>(yo + op * a.ld_y + c0, y);
        if (yr) store_vec<T, VEC>(yr + op * a.ld_raw + c0, f);
      }
    }
  }
}

template <typename T, int VEC>
static int launch_gn(const fidm_gn_args& a, cudaStream_t st) {
  GnParams p;
  p.a = a;
  p.cv = a.channels / VEC;
  FIDM_REQUIRE(p.cv <= 1024, FIDM_E_SHAPE, "groupnorm: %d channels not supported", a.channels);
  p.ppi = p.cv >= 256 ? 1 : 256 / p.cv;
  const int threads = p.cv * p.ppi;
  const int hw = a.height * a.width;
  // iteration space of the apply pass (output pixels for DOWN, input pixels otherwise)
  const int it_hw = (a.resample == FIDM_RESAMPLE_DOWN) ? hw / 4 : hw;
  auto plan = [&](int npix) {
    int chunks = (num_sms() * 8 + a.batch - 1) / a.batch;
    const int max_chunks = (npix + p.ppi - 1) / p.ppi;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    p.pix_per_blk = (npix + chunks - 1) / chunks;
    return (npix + p.pix_per_blk - 1) / p.pix_per_blk;
  };
  int chunks;
  if (!a.skip_norm) {
    FIDM_CUDA(cudaMemsetAsync(a.stats, 0, sizeof(double) * 2 * a.batch * a.groups, st));
    chunks = plan(hw);
    gn_stats_kernel<T, VEC><<<dim3(chunks, a.batch), threads, 0, st>>>(p);
    FIDM_CHECK_LAUNCH("groupnorm stats");
  }
  chunks = plan(it_hw);
  dim3 grid(chunks, a.batch);
  if (a.resample == FIDM_RESAMPLE_NONE)
    gn_apply_kernel<T, VEC, FIDM_RESAMPLE_NONE><<<grid, threads, 0, st>>>(p);
  else if (a.resample == FIDM_RESAMPLE_DOWN)
    gn_apply_kernel<T, VEC, FIDM_RESAMPLE_DOWN><<<grid, threads, 0, st>>>(p);
  else
    gn_apply_kernel<T, VEC, FIDM_RESAMPLE_UP><<<grid, threads, 0, st>>>(p);
  FIDM_CHECK_LAUNCH("groupnorm apply");
  return 0;
}

template <typename T>
static int dispatch_vec(const fidm_gn_args& a, cudaStream_t st) {
  const int cpg = a.channels / a.groups;
  constexpr int MAXV = 16 / sizeof(T);  // 16-byte vectors
  auto aligned = [&](int v) {
    const size_t bytes = sizeof(T) * v;
    bool ok = (cpg % v == 0) && (a.ld_x % v == 0) && (a.ld_y % v == 0) && ((uintptr_t)a.x % bytes == 0) &&
              ((uintptr_t)a.y % bytes == 0);
    if (a.y_raw) ok = ok && (a.ld_raw % v == 0) && ((uintptr_t)a.y_raw % bytes == 0);
    return ok;
  };
  if (MAXV >= 8 && aligned(8)) return launch_gn<T, 8>(a, st);
  if (aligned(4)) return launch_gn<T, 4>(a, st);
  if (aligned(2)) return launch_gn<T, 2>(a, st);
  return launch_gn<T, 1>(a, st);
}

}  // namespace fidm

extern "C" int fidm_groupnorm_silu_nhwc(const fidm_gn_args* a, fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(a && a->x && a->y && (a->stats || a->skip_norm), FIDM_E_BADARG, "groupnorm: null x/y/stats");
  FIDM_REQUIRE(a->batch > 0 && a->height > 0 && a->width > 0 && a->channels > 0, FIDM_E_BADARG, "groupnorm: empty shape");
  FIDM_REQUIRE(a->groups > 0 && a->groups <= 64 && a->channels % a->groups == 0, FIDM_E_SHAPE,
               "groupnorm: channels %d not divisible into %d groups", a->channels, a->groups);
  FIDM_REQUIRE(a->ld_x >= a->channels && a->ld_y >= a->channels, FIDM_E_BADARG, "groupnorm: ld < channels");
  FIDM_REQUIRE(a->resample >= 0 && a->resample <= 2, FIDM_E_BADARG, "groupnorm: bad resample %d", a->resample);
  if (a->resample == FIDM_RESAMPLE_DOWN)
    FIDM_REQUIRE(a->height % 2 == 0 && a->width % 2 == 0, FIDM_E_SHAPE, "groupnorm: odd size for 2x pooling");
  if (a->scale_shift) FIDM_REQUIRE(a->ld_ss >= 2 * a->channels, FIDM_E_BADARG, "groupnorm: ld_ss < 2*channels");
  if (a->dtype == FIDM_BF16) return dispatch_vec<__nv_bfloat16>(*a, (cudaStream_t)stream);
  if (a->dtype == FIDM_F32) return dispatch_vec<float>(*a, (cudaStream_t)stream);
  FIDM_REQUIRE(false, FIDM_E_BADARG, "groupnorm: bad dtype %d", a->dtype);
}
