// Layout converters between the reference's fp32 NCHW tensors and the NHWC (bf16|fp32) activations
// used inside the UNet, plus the one-time OIHW -> KRSC weight repack.
#include "common.cuh"

namespace fidm {

struct PackParams { fidm_pack_args a; };

// One thread per pixel: reads are coalesced across the warp (consecutive pixels of one channel
// plane), writes are one contiguous c_pad-wide run per thread.
template <typename T>
__global__ void __launch_bounds__(256) pack_kernel(const PackParams p) {
  const fidm_pack_args& a = p.a;
  const long long total = (long long)a.batch * a.hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / a.hw);
    const int px = (int)(i % a.hw);
    T* dst = reinterpret_cast<T*>(a.dst) + i * a.ld_dst;
    int c = 0;
    for (int s = 0; s < a.n_src; ++s) {
      const float* src = a.src[s] + (long long)b * a.src_channels[s] * a.hw + px;
      for (int r = 0; r < a.src_repeat[s]; ++r)
        for (int k = 0; k < a.src_channels[s]; ++k) dst[c++] = from_f32<T>(src[(long long)k * a.hw]);
    }
    for (; c < a.c_pad; ++c) dst[c] = from_f32<T>(0.0f);
  }
}

// Fast path for the network input (<= 16 channels written, e.g. the 9-channel inpainting stem padded
// to 64: channels >= c_pad are zero-filled once at allocation and never touched again).
template <typename T>
__global__ void __launch_bounds__(256) pack16_kernel(const PackParams p) {
  const fidm_pack_args& a = p.a;
  const long long total = (long long)a.batch * a.hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / a.hw);
    const int px = (int)(i % a.hw);
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.0f;
    int c = 0;
    for (int s = 0; s < a.n_src; ++s) {
      const float* src = a.src[s] + (long long)b * a.src_channels[s] * a.hw + px;
      for (int k = 0; k < a.src_channels[s]; ++k) {
        const float f = src[(long long)k * a.hw];
        for (int r = 0; r < a.src_repeat[s]; ++r) {
          const int cc = c + r * a.src_channels[s] + k;
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q == cc) v[q] = f;
        }
      }
      c += a.src_channels[s] * a.src_repeat[s];
    }
    T* dst = reinterpret_cast<T*>(a.dst) + i * a.ld_dst;
    float lo[8], hi[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { lo[q] = v[q]; hi[q] = v[8 + q]; }
    if (sizeof(T) == 2) {
      store_vec<T, 8>(dst, lo);
      if (a.c_pad > 8) store_vec<T, 8>(dst + 8, hi);
    } else {
      for (int q = 0; q < a.c_pad; ++q) dst[q] = from_f32<T>(v[q]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) unpack_kernel(const T* __restrict__ src, int ld, float* __restrict__ dst,
                                                      int batch, int hw, int channels) {
  const long long total = (long long)batch * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const int px = (int)(i % hw);
    const T* s = src + i * ld;
    for (int c = 0; c < channels; ++c) dst[((long long)b * channels + c) * hw + px] = to_f32<T>(s[c]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) repack_kernel(const float* __restrict__ w, T* __restrict__ dst, int cout,
                                                      int cin, int ks, int cout_pad, int cin_pad) {
  const long long total = (long long)cout_pad * ks * ks * cin_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    const int tap = (int)((i / cin_pad) % (ks * ks));
    const int co = (int)(i / ((long long)cin_pad * ks * ks));
    float v = 0.0f;
    if (co < cout && ci < cin) v = w[((long long)co * cin + ci) * ks * ks + tap];
    dst[i] = from_f32<T>(v);
  }
}

static unsigned grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace fidm

extern "C" int fidm_pack_nchw_to_nhwc(const fidm_pack_args* a, fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(a && a->dst, FIDM_E_BADARG, "pack: null args");
  FIDM_REQUIRE(a->n_src >= 1 && a->n_src <= 4, FIDM_E_BADARG, "pack: n_src %d", a->n_src);
  int c = 0;
  for (int s = 0; s < a->n_src; ++s) {
    FIDM_REQUIRE(a->src[s] && a->src_channels[s] > 0 && a->src_repeat[s] > 0, FIDM_E_BADARG, "pack: bad source %d", s);
    c += a->src_channels[s] * a->src_repeat[s];
  }
  FIDM_REQUIRE(c <= a->c_pad && a->c_pad <= a->ld_dst, FIDM_E_SHAPE, "pack: %d channels > c_pad %d (ld %d)", c,
               a->c_pad, a->ld_dst);
  PackParams p;
  p.a = *a;
  const unsigned g = grid_for((long long)a->batch * a->hw);
  if (a->c_pad <= 16 && (a->c_pad % 8 == 0) && a->ld_dst % 8 == 0 && (uintptr_t)a->dst % 16 == 0 &&
      a->dst_dtype != FIDM_F32) {
    if (a->dst_dtype == FIDM_BF16)
      pack16_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(p);
    else
      pack16_kernel<__half><<<g, 256, 0, (cudaStream_t)stream>>>(p);
    FIDM_CHECK_LAUNCH("pack_nchw_to_nhwc");
    return 0;
  }
  if (a->dst_dtype == FIDM_BF16)
    pack_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(p);
  else if (a->dst_dtype == FIDM_F16)
    pack_kernel<__half><<<g, 256, 0, (cudaStream_t)stream>>>(p);
  else
    pack_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(p);
  FIDM_CHECK_LAUNCH("pack_nchw_to_nhwc");
  return 0;
}

extern "C" int fidm_unpack_nhwc_to_nchw(const void* src, int32_t src_dtype, int32_t ld_src, float* dst, int32_t batch,
                                        int32_t hw, int32_t channels, fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(src && dst && channels <= ld_src, FIDM_E_BADARG, "unpack: bad args");
  const unsigned g = grid_for((long long)batch * hw);
  if (src_dtype == FIDM_BF16)
    unpack_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, ld_src, dst, batch, hw, channels);
  else if (src_dtype == FIDM_F16)
    unpack_kernel<__half><<<g, 256, 0, (cudaStream_t)stream>>>((const __half*)src, ld_src, dst, batch, hw, channels);
  else
    unpack_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)src, ld_src, dst, batch, hw, channels);
  FIDM_CHECK_LAUNCH("unpack_nhwc_to_nchw");
  return 0;
}

extern "C" int fidm_repack_weight_oihw_to_krsc(const float* w, void* dst, int32_t dst_dtype, int32_t cout, int32_t cin,
                                               int32_t ksize, int32_t cout_pad, int32_t cin_pad, fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(w && dst && cout_pad >= cout && cin_pad >= cin && (ksize == 1 || ksize == 3), FIDM_E_BADARG,
               "repack: bad args");
  const unsigned g = grid_for((long long)cout_pad * ksize * ksize * cin_pad);
  if (dst_dtype == FIDM_BF16)
    repack_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)dst, cout, cin, ksize, cout_pad, cin_pad);
  else if (dst_dtype == FIDM_F16)
    repack_kernel<__half><<<g, 256, 0, (cudaStream_t)stream>>>(w, (__half*)dst, cout, cin, ksize, cout_pad, cin_pad);
  else
    repack_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(w, (float*)dst, cout, cin, ksize, cout_pad, cin_pad);
  FIDM_CHECK_LAUNCH("repack_weight");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// I/O adapters: the dataset / PNG conventions of the reference as GPU kernels.
// ------------------------------------------------------------------------------------------------
namespace fidm {

// data/dataset.py:38-42,130-142: image = (u8/255 - 0.5)/0.5, mask = (gray/255 < 0.5) ? 1 : 0 (1 = inpaint),
// masked_image = image * (1 - mask), keep = 1 - mask.  uint8 NHWC in, fp32 NCHW out; one rounding per torch op.
__global__ void __launch_bounds__(256) prepare_inputs_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ msk,
                                                              float* __restrict__ image, float* __restrict__ masked,
                                                              float* __restrict__ mask, float* __restrict__ keep, int batch,
                                                              int hw) {
  const long long total = (long long)batch * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw), px = (int)(i % hw);
    const float m = (__fdiv_rn((float)msk[i], 255.0f) < 0.5f) ? 1.0f : 0.0f;
    const float k = __fsub_rn(1.0f, m);
    mask[i] = m;
    keep[i] = k;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)img[i * 3 + c], 255.0f), 0.5f), 0.5f);
      const long long o = ((long long)b * 3 + c) * hw + px;
      image[o] = v;
      masked[o] = __fmul_rn(v, k);
    }
  }
}

// test_inp_ddim_100.py:693-696 then :33-41: out = toU8(sample * mask + gt * (1 - mask)), NCHW fp32 -> NHWC uint8.
__global__ void __launch_bounds__(256) blend_to_u8_kernel(const float* __restrict__ sample, const float* __restrict__ gt,
                                                           const float* __restrict__ mask, uint8_t* __restrict__ out,
                                                           int batch, int hw) {
  const long long total = (long long)batch * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw), px = (int)(i % hw);
    const float m = mask ? mask[i] : 1.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long o = ((long long)b * 3 + c) * hw + px;
      float v = sample[o];
      if (mask) v = __fadd_rn(__fmul_rn(v, m), __fmul_rn(gt[o], __fsub_rn(1.0f, m)));
      v = __fmul_rn(__fadd_rn(v, 1.0f), 127.5f);
      v = fminf(fmaxf(v, 0.0f), 255.0f);
      out[i * 3 + c] = (uint8_t)v;                     // truncation, as Tensor.to(torch.uint8)
    }
  }
}

}  // namespace fidm

extern "C" int fidm_prepare_inputs_u8(const uint8_t* image_u8_nhwc, const uint8_t* mask_u8, float* image, float* masked_image,
                                      float* mask, float* keep_mask, int32_t batch, int32_t hw, fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(image_u8_nhwc && mask_u8 && image && masked_image && mask && keep_mask && batch > 0 && hw > 0, FIDM_E_BADARG,
               "prepare_inputs_u8: bad args");
  prepare_inputs_kernel<<<grid_for((long long)batch * hw), 256, 0, (cudaStream_t)stream>>>(image_u8_nhwc, mask_u8, image,
                                                                                          masked_image, mask, keep_mask, batch, hw);
  FIDM_CHECK_LAUNCH("prepare_inputs_u8");
  return 0;
}

extern "C" int fidm_blend_to_u8(const float* sample, const float* gt, const float* mask, uint8_t* out_u8_nhwc, int32_t batch,
                                int32_t hw, fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(sample && out_u8_nhwc && batch > 0 && hw > 0 && ((mask == nullptr) == (gt == nullptr)), FIDM_E_BADARG,
               "blend_to_u8: bad args");
  blend_to_u8_kernel<<<grid_for((long long)batch * hw), 256, 0, (cudaStream_t)stream>>>(sample, gt, mask, out_u8_nhwc, batch, hw);
  FIDM_CHECK_LAUNCH("blend_to_u8");
  return 0;
}
