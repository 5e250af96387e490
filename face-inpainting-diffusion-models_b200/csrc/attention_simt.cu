// FFMA flash-style QKV attention (fp32 math, fp32|bf16 storage) over the NHWC qkv buffer.
//
// Verification-mode twin of the tcgen05 attention kernel and the path for head dims it does not
// take (e.g. the 128-wide heads of the T64 middle block).  Semantics of nn.py:222-235:
//   w = softmax_fp32((q*s)^T (k*s)),  a = w v^T,  s = head_dim^-1/4,
// head h owns channels [h*d,(h+1)*d) of each of the Q | K | V thirds (nn.py:226-234).
#include "common.cuh"

namespace fidm {

struct AttnSimtParams { fidm_attn_args a; };

constexpr int AQ = 64, AK = 64;  // queries / keys per tile

template <typename T, int D>
__global__ void __launch_bounds__(256) attn_simt_kernel(const AttnSimtParams p) {
  const fidm_attn_args& a = p.a;
  extern __shared__ float sm[];
  float* Qs = sm;                       // [AQ][D+1]
  float* Ks = Qs + AQ * (D + 1);        // [AK][D+1]
  float* Vs = Ks + AK * (D + 1);        // [AK][D]
  float* Ps = Vs + AK * D;              // [AQ][AK+1]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
  const int C = a.heads * D;
  const float s = 1.0f / sqrtf(sqrtf((float)D));
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.tokens * a.ld_qkv;

  for (int i = tid; i < AQ * (D / 4); i += 256) {
    const int r = i / (D / 4), d0 = (i % (D / 4)) * 4;
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    if (q0 + r < a.tokens) load_vec<T, 4>(base + (long long)(q0 + r) * a.ld_qkv + h * D + d0, f);
#pragma unroll
    for (int k = 0; k < 4; ++k) Qs[r * (D + 1) + d0 + k] = f[k] * s;
  }

  constexpr int NO = D / 16;
  float m[4], l[4], o[4][NO];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.0f;
#pragma unroll
    for (int j = 0; j < NO; ++j) o[i][j] = 0.0f;
  }

  for (int k0 = 0; k0 < a.tokens; k0 += AK) {
    __syncthreads();
    for (int i = tid; i < AK * (D / 4); i += 256) {
      const int r = i / (D / 4), d0 = (i % (D / 4)) * 4;
      float fk[4] = {0.f, 0.f, 0.f, 0.f}, fv[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + r < a.tokens) {
        const T* row = base + (long long)(k0 + r) * a.ld_qkv + h * D + d0;
        load_vec<T, 4>(row + C, fk);
        load_vec<T, 4>(row + 2 * C, fv);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        Ks[r * (D + 1) + d0 + k] = fk[k] * s;
        Vs[r * D + d0 + k] = fv[k];
      }
    }
    __syncthreads();
    float sc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[i][j] = 0.0f;
    for (int d = 0; d < D; ++d) {
      float qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = Qs[(ty * 4 + i) * (D + 1) + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = Ks[(tx + 16 * j) * (D + 1) + d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sc[i][j] = fmaf(qv[i], kv[j], sc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k0 + tx + 16 * j >= a.tokens) sc[i][j] = -INFINITY;
        mx = fmaxf(mx, sc[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mn = fmaxf(m[i], mx);
      const float corr = expf(m[i] - mn);
      float rs = 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pv = expf(sc[i][j] - mn);
        rs += pv;
        Ps[(ty * 4 + i) * (AK + 1) + tx + 16 * j] = pv;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l[i] = l[i] * corr + rs;
      m[i] = mn;
#pragma unroll
      for (int j = 0; j < NO; ++j) o[i][j] *= corr;
    }
    __syncthreads();
    for (int k = 0; k < AK; ++k) {
      float pv[4], vv[NO];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = Ps[(ty * 4 + i) * (AK + 1) + k];
#pragma unroll
      for (int j = 0; j < NO; ++j) vv[j] = Vs[k * D + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NO; ++j) o[i][j] = fmaf(pv[i], vv[j], o[i][j]);
    }
  }
  T* out = reinterpret_cast<T*>(a.out) + (long long)b * a.tokens * a.ld_out + h * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= a.tokens) continue;
    const float inv = 1.0f / l[i];
#pragma unroll
    for (int j = 0; j < NO; ++j) out[(long long)q * a.ld_out + tx + 16 * j] = from_f32<T>(o[i][j] * inv);
  }
}

template <typename T, int D>
static int launch_attn_simt(const fidm_attn_args& a, cudaStream_t st) {
  AttnSimtParams p;
  p.a = a;
  const size_t smem = sizeof(float) * (AQ * (D + 1) + AK * (D + 1) + AK * D + AQ * (AK + 1));
  FIDM_CUDA(cudaFuncSetAttribute(attn_simt_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a.tokens + AQ - 1) / AQ, a.heads, a.batch);
  attn_simt_kernel<T, D><<<grid, 256, smem, st>>>(p);
  FIDM_CHECK_LAUNCH("attention_simt");
  return 0;
}

// Any other head dim (the reference takes heads = C / num_head_channels or num_heads, nn.py:245-249, so d can be 48, 96,
// 192, 256 ... 1024): one warp per query row, online softmax over 32-key tiles, lane = key for the scores and
// lane = channel (stride 32) for the output.  A correctness path for unusual constructor arguments, not a tuned kernel.
constexpr int kGenericMaxD = 1024;
template <typename T>
__global__ void __launch_bounds__(128) attn_generic_kernel(const AttnSimtParams p) {
  const fidm_attn_args& a = p.a;
  extern __shared__ float sm[];
  const int D = a.head_dim, C = a.heads * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q_s = sm + warp * D;                         // this warp's query row, pre-scaled
  const int h = blockIdx.y, b = blockIdx.z;
  const float s = 1.0f / sqrtf(sqrtf((float)D));
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.tokens * a.ld_qkv + h * D;
  constexpr int kMaxO = kGenericMaxD / 32;
  for (int q = blockIdx.x * 4 + warp; q < a.tokens; q += gridDim.x * 4) {
    __syncwarp();
    for (int d = lane; d < D; d += 32) q_s[d] = to_f32<T>(base[(long long)q * a.ld_qkv + d]) * s;
    __syncwarp();
    float m = -INFINITY, l = 0.0f, o[kMaxO];
#pragma unroll
    for (int i = 0; i < kMaxO; ++i) o[i] = 0.0f;
    for (int k0 = 0; k0 < a.tokens; k0 += 32) {
      const int key = k0 + lane;
      float sc = -INFINITY;
      if (key < a.tokens) {
        const T* kr = base + (long long)key * a.ld_qkv + C;
        float acc = 0.0f;
        for (int d = 0; d < D; ++d) acc = fmaf(q_s[d], to_f32<T>(kr[d]) * s, acc);
        sc = acc;
      }
      const float mn = fmaxf(m, warp_max(sc));
      const float corr = expf(m - mn);
      const float pv = expf(sc - mn);                 // 0 for keys past the end
      l = l * corr + warp_sum(pv);
      m = mn;
#pragma unroll
      for (int i = 0; i < kMaxO; ++i) o[i] *= corr;
      const int nk = min(32, a.tokens - k0);
      for (int kk = 0; kk < nk; ++kk) {
        const float pk = __shfl_sync(0xffffffffu, pv, kk);
        const T* vr = base + (long long)(k0 + kk) * a.ld_qkv + 2 * C;
#pragma unroll
        for (int i = 0; i < kMaxO; ++i)
          if (i * 32 + lane < D) o[i] = fmaf(pk, to_f32<T>(vr[i * 32 + lane]), o[i]);
      }
    }
    T* out = reinterpret_cast<T*>(a.out) + ((long long)b * a.tokens + q) * a.ld_out + h * D;
    const float inv = 1.0f / l;
#pragma unroll
    for (int i = 0; i < kMaxO; ++i)
      if (i * 32 + lane < D) out[i * 32 + lane] = from_f32<T>(o[i] * inv);
  }
}

template <typename T>
static int dispatch_attn_simt(const fidm_attn_args& a, cudaStream_t st) {
  switch (a.head_dim) {
    case 16: return launch_attn_simt<T, 16>(a, st);
    case 32: return launch_attn_simt<T, 32>(a, st);
    case 64: return launch_attn_simt<T, 64>(a, st);
    case 128: return launch_attn_simt<T, 128>(a, st);
    default: break;
  }
  FIDM_REQUIRE(a.head_dim > 0 && a.head_dim <= kGenericMaxD, FIDM_E_SHAPE, "attention_simt: head_dim %d not in [1, %d]",
               a.head_dim, kGenericMaxD);
  AttnSimtParams p;
  p.a = a;
  const int qblocks = (a.tokens + 3) / 4;
  dim3 grid(qblocks < 1024 ? qblocks : 1024, a.heads, a.batch);
  attn_generic_kernel<T><<<grid, 128, 4 * a.head_dim * sizeof(float), st>>>(p);
  FIDM_CHECK_LAUNCH("attention_generic");
  return 0;
}

}  // namespace fidm

extern "C" int fidm_attention_qkv_nhwc_simt(const fidm_attn_args* a, fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(a && a->qkv && a->out, FIDM_E_BADARG, "attention_simt: null qkv/out");
  FIDM_REQUIRE(a->batch > 0 && a->tokens > 0 && a->heads > 0, FIDM_E_BADARG, "attention_simt: empty shape");
  FIDM_REQUIRE(a->ld_qkv >= 3 * a->heads * a->head_dim && a->ld_out >= a->heads * a->head_dim, FIDM_E_BADARG,
               "attention_simt: ld too small");
  FIDM_REQUIRE(a->ld_qkv % 4 == 0, FIDM_E_ALIGN, "attention_simt: ld_qkv must be a multiple of 4");
  if (a->dtype == FIDM_BF16) return dispatch_attn_simt<__nv_bfloat16>(*a, (cudaStream_t)stream);
  if (a->dtype == FIDM_F32) return dispatch_attn_simt<float>(*a, (cudaStream_t)stream);
  FIDM_REQUIRE(false, FIDM_E_BADARG, "attention_simt: bad dtype %d", a->dtype);
}
