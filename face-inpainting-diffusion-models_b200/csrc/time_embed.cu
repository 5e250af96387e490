// K5: timestep path.  timestep_embedding (nn.py:51-61) and the small-batch Linear layers of
// time_embed (unet.py:44-48) / ResBlock.emb_layers (nn.py:167-170), which depend only on t.
// Bandwidth-bound GEMV: one warp per output feature streams its weight row once for all samples.
#include "common.cuh"

namespace fidm {

__global__ void timestep_embedding_kernel(const float* __restrict__ t, const float* __restrict__ freqs,
                                          float* __restrict__ out, int batch, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * dim) return;
  const int b = i / dim, j = i % dim;
  float v = 0.0f;                                    // odd dim: trailing zero column (nn.py:59-60)
  if (j < half) v = cosf(__fmul_rn(t[b], freqs[j]));  // cos first, then sin (nn.py:58)
  else if (j < 2 * half) v = sinf(__fmul_rn(t[b], freqs[j - half]));
  out[i] = v;
}

constexpr int LIN_BCHUNK = 8;

template <typename WT>
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, const WT* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y,
                                                            int batch, int K, int O, int silu_in) {
  extern __shared__ float xs[];  // [LIN_BCHUNK][K]
  const int b0 = blockIdx.y * LIN_BCHUNK;
  const int nb = min(LIN_BCHUNK, batch - b0);
  for (int i = threadIdx.x; i < LIN_BCHUNK * K; i += blockDim.x) {
    const int b = i / K, k = i % K;
    float v = 0.0f;
    if (b < nb) {
      v = x[(long long)(b0 + b) * K + k];
      if (silu_in) v = v / (1.0f + expf(-v));
    }
    xs[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int o = blockIdx.x * warps + warp; o < O; o += gridDim.x * warps) {
    float acc[LIN_BCHUNK];
#pragma unroll
    for (int b = 0; b < LIN_BCHUNK; ++b) acc[b] = 0.0f;
    const WT* wr = w + (long long)o * K;
    for (int k = lane; k < K; k += 32) {
      const float wv = to_f32<WT>(wr[k]);
#pragma unroll
      for (int b = 0; b < LIN_BCHUNK; ++b) acc[b] = fmaf(wv, xs[b * K + k], acc[b]);
    }
#pragma unroll
    for (int b = 0; b < LIN_BCHUNK; ++b) acc[b] = warp_sum(acc[b]);
    if (lane == 0) {
      const float bv = bias ? bias[o] : 0.0f;
      for (int b = 0; b < nb; ++b) y[(long long)(b0 + b) * O + o] = acc[b] + bv;
    }
  }
}

}  // namespace fidm

extern "C" int fidm_timestep_embedding(const float* t, const float* freqs, float* out, int32_t batch, int32_t dim,
                                       fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(t && freqs && out && batch > 0 && dim > 1, FIDM_E_BADARG, "timestep_embedding: bad args");
  const int n = batch * dim;
  timestep_embedding_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, freqs, out, batch, dim);
  FIDM_CHECK_LAUNCH("timestep_embedding");
  return 0;
}

extern "C" int fidm_linear_small(const float* x, const void* w, int32_t w_dtype, const float* bias, float* y,
                                 int32_t batch, int32_t in_features, int32_t out_features, int32_t silu_input,
                                 fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(x && w && y && batch > 0 && in_features > 0 && out_features > 0, FIDM_E_BADARG, "linear_small: bad args");
  const size_t smem = (size_t)LIN_BCHUNK * in_features * sizeof(float);
  FIDM_REQUIRE(smem <= 200 * 1024, FIDM_E_SHAPE, "linear_small: in_features %d too large", in_features);
  const int warps = 8;
  int gx = (out_features + warps - 1) / warps;
  const int cap = num_sms() * 8;
  if (gx > cap) gx = cap;
  dim3 grid(gx, (batch + LIN_BCHUNK - 1) / LIN_BCHUNK);
  if (w_dtype == FIDM_BF16) {
    if (smem > 48 * 1024) FIDM_CUDA(cudaFuncSetAttribute(linear_small_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linear_small_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>(x, (const __nv_bfloat16*)w, bias, y, batch, in_features, out_features, silu_input);
  } else {
    if (smem > 48 * 1024) FIDM_CUDA(cudaFuncSetAttribute(linear_small_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linear_small_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>(x, (const float*)w, bias, y, batch, in_features, out_features, silu_input);
  }
  FIDM_CHECK_LAUNCH("linear_small");
  return 0;
}
