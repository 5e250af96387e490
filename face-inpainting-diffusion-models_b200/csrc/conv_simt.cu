// Generic FFMA implicit-GEMM convolution (NHWC, KRSC weights), fp32 or bf16 storage, fp32 math.
//
// This is the "fp32 verification mode" of the UNet (rel-L2 1e-5 bar of the north star) and the
// kernel for the few shapes the tcgen05 kernel does not take (stride 2, channel counts that are
// not multiples of 64).  Same fused epilogue as the tensor-core kernel.  Reference call sites:
// every conv_nd of nn.py:102,126,153,176,182,184,252,254 and unet.py:55,151.
#include "common.cuh"

namespace fidm {

struct ConvSimtParams {
  fidm_conv_args a;
  int Ho, Wo, pad;
  long long M;  // batch * Ho * Wo
};

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvSimtParams p) {
  const fidm_conv_args& a = p.a;
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;

  // loader role: one pixel / one cout row, 4 consecutive k
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const long long lm = m0 + lrow;
  const bool lm_ok = lm < p.M;
  int ln = 0, lho = 0, lwo = 0;
  if (lm_ok) {
    ln = (int)(lm / ((long long)p.Ho * p.Wo));
    const int r = (int)(lm % ((long long)p.Ho * p.Wo));
    lho = r / p.Wo;
    lwo = r % p.Wo;
  }
  const int lco = n0 + lrow;
  const bool lco_ok = lco < a.cout;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  auto mainloop = [&](const T* __restrict__ X, int ldx, int cin, const T* __restrict__ Wt, int ks, int stride,
                      int pad, int H, int W) {
    const int taps = ks * ks;
    for (int tap = 0; tap < taps; ++tap) {
      const int r = tap / ks, s = tap % ks;
      const int hi = lho * stride + r - pad, wi = lwo * stride + s - pad;
      const bool in_ok = lm_ok && hi >= 0 && hi < H && wi >= 0 && wi < W;
      const T* xrow = X + (((long long)ln * H + hi) * W + wi) * ldx;
      const T* wrow = Wt + ((long long)lco * taps + tap) * cin;
      for (int c0 = 0; c0 < cin; c0 += TK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (in_ok) load_vec<T, 4>(xrow + c0 + lk, av);
        if (lco_ok) load_vec<T, 4>(wrow + c0 + lk, bv);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          As[lk + i][lrow] = av[i];
          Bs[lk + i][lrow] = bv[i];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
          const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
          const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
      }
    }
  };

  mainloop(reinterpret_cast<const T*>(a.x), a.ld_x, a.cin, reinterpret_cast<const T*>(a.w), a.ksize, a.stride,
           p.pad, a.height, a.width);
  if (a.x2)  // fused 1x1 second source at output resolution (ResBlock skip_connection)
    mainloop(reinterpret_cast<const T*>(a.x2), a.ld_x2, a.cin2, reinterpret_cast<const T*>(a.w2), 1, 1, 0, p.Ho, p.Wo);

  // ---- epilogue
  const int ohw = p.Ho * p.Wo;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const int n = (int)(m / ohw);
    const int px = (int)(m % ohw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= a.cout_valid) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[co];
      if (a.row_add) v += a.row_add[(long long)n * a.ld_row_add + co];
      if (a.residual) v += to_f32<T>(reinterpret_cast<const T*>(a.residual)[m * a.ld_res + co]);
      if (a.y_nchw_f32)
        reinterpret_cast<float*>(a.y)[((long long)n * a.cout_valid + co) * ohw + px] = v;
      else
        reinterpret_cast<T*>(a.y)[m * a.ld_y + co] = from_f32<T>(v);
    }
  }
}

}  // namespace fidm

extern "C" int fidm_conv2d_nhwc_simt(const fidm_conv_args* a, fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(a && a->x && a->w && a->y, FIDM_E_BADARG, "conv_simt: null x/w/y");
  FIDM_REQUIRE(a->ksize == 1 || a->ksize == 3, FIDM_E_SHAPE, "conv_simt: ksize %d", a->ksize);
  FIDM_REQUIRE(a->stride == 1 || a->stride == 2, FIDM_E_SHAPE, "conv_simt: stride %d", a->stride);
  FIDM_REQUIRE(a->cin % 16 == 0 && a->ld_x % 4 == 0, FIDM_E_SHAPE, "conv_simt: cin %d must be a multiple of 16", a->cin);
  if (a->x2) FIDM_REQUIRE(a->w2 && a->cin2 % 16 == 0 && a->ld_x2 % 4 == 0, FIDM_E_SHAPE, "conv_simt: bad second source");
  FIDM_REQUIRE(a->cout_valid > 0 && a->cout_valid <= a->cout, FIDM_E_BADARG, "conv_simt: cout_valid %d", a->cout_valid);
  ConvSimtParams p;
  p.a = *a;
  p.pad = a->ksize / 2;
  p.Ho = (a->height + 2 * p.pad - a->ksize) / a->stride + 1;
  p.Wo = (a->width + 2 * p.pad - a->ksize) / a->stride + 1;
  p.M = (long long)a->batch * p.Ho * p.Wo;
  dim3 grid((unsigned)((p.M + TM - 1) / TM), (a->cout + TN - 1) / TN);
  if (a->dtype == FIDM_BF16)
    conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else if (a->dtype == FIDM_F32)
    conv_simt_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else
    FIDM_REQUIRE(false, FIDM_E_BADARG, "conv_simt: bad dtype %d", a->dtype);
  FIDM_CHECK_LAUNCH("conv_simt");
  return 0;
}
