// K4: fused reverse-process step  [update at t] -> [known-region injection at t_inject].
//
// One launch replaces the per-step elementwise chain of the reference
// (gaussian_diffusion.py:114-157 inject, :213-298 p_mean_variance tail, :357-388 p_sample,
// :447-485 ddim_sample).  HBM-bound: DDIM eta=0 moves 64 B per pixel (read x, eps, gt, n_inj:
// 4x12 B, mask 4 B; write x 12 B).
//
// Rounding contract: every reference ATen op is one IEEE-754 fp32 rounding, so every product and
// sum below is an explicit __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn -- the compiler may not contract
// them into FMAs.  With identical model output and noise the DDIM path is bit-identical to the
// reference; the DDPM path differs only by the ulp of expf.
#include "common.cuh"

namespace fidm {

struct StepParams {
  fidm_step_args a;
};

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }

template <int VEC>
__global__ void __launch_bounds__(256) sampler_step_kernel(const StepParams p) {
  const fidm_step_args& a = p.a;
  const int hw_v = a.hw / VEC;
  const long long total = (long long)a.batch * a.channels * hw_v;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int pv = (int)(idx % hw_v);
    const int bc = (int)(idx / hw_v);
    const int c = bc % a.channels;
    const int b = bc / a.channels;
    const long long off = ((long long)b * a.channels + c) * a.hw + (long long)pv * VEC;

    int t_upd = a.t_update, t_inj = a.t_inject;
    if (a.t_dev) {
      t_upd = (int)a.t_dev[b];
      t_inj = (a.mode == FIDM_STEP_INJECT_ONLY) ? t_upd : t_upd - 1;
    }

    // Every load of this element is issued up front (one memory round trip instead of two dependent ones), and only
    // what the mode needs is read: the variance channels are skipped when neither the DDPM update nor a caller wants
    // the log-variance (DDIM: 64 B per pixel instead of 76).
    const bool upd = a.mode != FIDM_STEP_INJECT_ONLY, inj = a.mode != FIDM_STEP_UPDATE_ONLY;
    const bool need_var = upd && a.var_type != FIDM_VAR_FIXED && (a.sampler == FIDM_SAMPLER_DDPM || a.logvar_out != nullptr);
    const bool need_z = upd && (a.z != nullptr);
    const int oc = (a.var_type == FIDM_VAR_FIXED) ? a.channels : 2 * a.channels;
    const long long ooff = ((long long)b * oc + c) * a.hw + (long long)pv * VEC;
    const int mc = (a.mask_channels == 1) ? 0 : c;
    const long long moff = ((long long)b * a.mask_channels + mc) * a.hw + (long long)pv * VEC;
    float xs[VEC], mo[VEC], vv[VEC], zz[VEC], g[VEC], n[VEC], m[VEC];
    load_vec<float, VEC>(a.x + off, xs);
    if (upd) load_vec<float, VEC>(a.model_out + ooff, mo);
    if (need_var) load_vec<float, VEC>(a.model_out + ooff + (long long)a.channels * a.hw, vv);
    if (need_z) load_vec<float, VEC>(a.z + off, zz);
    if (inj) {
      load_vec<float, VEC>(a.gt + off, g);
      load_vec<float, VEC>(a.inject_noise + off, n);
      load_vec<float, VEC>(a.keep_mask + moff, m);
    }
    float smp[VEC], x0v[VEC];

    if (upd) {
      const float* cf = a.coef + (long long)t_upd * FIDM_COEF_COLS;
      float mv[VEC], lv[VEC];
      const float c1 = cf[FIDM_C_RECIP], c2 = cf[FIDM_C_RECIPM1];
      const float k1 = cf[FIDM_C_POST1], k2 = cf[FIDM_C_POST2];
      const float nz = cf[FIDM_C_NONZERO];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float x = xs[i];
        // ---- log-variance (gaussian_diffusion.py:241-265)
        float logvar = 0.0f;
        if (!need_var && a.var_type != FIDM_VAR_FIXED) {
          // log-variance not needed (DDIM update, caller does not ask for it)
        } else if (a.var_type == FIDM_VAR_LEARNED_RANGE) {
          const float frac = __fmul_rn(__fadd_rn(vv[i], 1.0f), 0.5f);                 // (v + 1) / 2
          logvar = __fadd_rn(__fmul_rn(frac, cf[FIDM_C_MAX_LOG]),
                             __fmul_rn(__fsub_rn(1.0f, frac), cf[FIDM_C_MIN_LOG]));
        } else if (a.var_type == FIDM_VAR_LEARNED) {
          logvar = vv[i];
        } else {
          logvar = cf[FIDM_C_FIXED_LOGVAR];
        }
        // ---- pred_xstart and posterior mean (:267-286)
        float x0, mean;
        if (a.sampler == FIDM_SAMPLER_DDIM_SCRIPT) {
          // test_inp_ddim_100.py:539-542: pred_x0 = (img - sqrt(1-ab_t) * eps) / sqrt(ab_t)
          x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(c2, mo[i])), c1);
          if (a.clip_denoised) x0 = clamp1(x0);
          mean = x0;
        } else if (a.mean_type == FIDM_MEAN_EPSILON) {
          x0 = __fsub_rn(__fmul_rn(c1, x), __fmul_rn(c2, mo[i]));
          if (a.clip_denoised) x0 = clamp1(x0);
          mean = __fadd_rn(__fmul_rn(k1, x0), __fmul_rn(k2, x));
        } else if (a.mean_type == FIDM_MEAN_START_X) {
          x0 = a.clip_denoised ? clamp1(mo[i]) : mo[i];
          mean = __fadd_rn(__fmul_rn(k1, x0), __fmul_rn(k2, x));
        } else {  // PREVIOUS_X (:274-278)
          x0 = __fsub_rn(__fmul_rn(cf[FIDM_C_XPREV_A], mo[i]), __fmul_rn(cf[FIDM_C_XPREV_B], x));
          if (a.clip_denoised) x0 = clamp1(x0);
          mean = mo[i];
        }
        x0v[i] = x0;
        mv[i] = mean;
        lv[i] = logvar;
        const float z = need_z ? zz[i] : 0.0f;
        if (a.sampler == FIDM_SAMPLER_DDIM_SCRIPT) {
          // :551-557  img = sqrt(ab_prev) * x0 + sqrt(1 - ab_prev - sigma^2) * eps_raw + sigma * noise
          const float mp = __fadd_rn(__fmul_rn(cf[FIDM_C_DDIM_SQRT_ABP], x0), __fmul_rn(cf[FIDM_C_DDIM_DIR], mo[i]));
          smp[i] = need_z ? __fadd_rn(mp, __fmul_rn(cf[FIDM_C_DDIM_SIGMA], z)) : mp;
        } else if (a.sampler == FIDM_SAMPLER_DDIM) {
          // eps re-derived from the (clamped) x0 (:470, :316-319), then :479-484
          const float e = __fdiv_rn(__fsub_rn(__fmul_rn(c1, x), x0), c2);
          const float mp = __fadd_rn(__fmul_rn(x0, cf[FIDM_C_DDIM_SQRT_ABP]), __fmul_rn(cf[FIDM_C_DDIM_DIR], e));
          smp[i] = need_z ? __fadd_rn(mp, __fmul_rn(__fmul_rn(nz, cf[FIDM_C_DDIM_SIGMA]), z)) : mp;
        } else {
          // :382-387  mean + nonzero * exp(0.5 * logvar) * z
          const float sd = expf(__fmul_rn(0.5f, logvar));
          smp[i] = __fadd_rn(mean, __fmul_rn(__fmul_rn(nz, sd), z));
        }
      }
      if (a.sample) store_vec<float, VEC>(a.sample + off, smp);
      if (a.pred_xstart) store_vec<float, VEC>(a.pred_xstart + off, x0v);
      if (a.mean_out) store_vec<float, VEC>(a.mean_out + off, mv);
      if (a.logvar_out) store_vec<float, VEC>(a.logvar_out + off, lv);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) smp[i] = xs[i];
    }

    if (a.mode != FIDM_STEP_UPDATE_ONLY) {
      // apply_inpainting_injection (:139-155): x = m*wg + (1-m)*x, wg = a*gt + s*n
      const float* cf = a.coef + (long long)t_inj * FIDM_COEF_COLS;
      const float ca = a.cumulative ? cf[FIDM_C_SQRT_AB] : cf[FIDM_C_SQRT_AB_F32];
      const float cs = a.cumulative ? cf[FIDM_C_SQRT_1MAB] : cf[FIDM_C_SQRT_1MAB_F32];
      float xn[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float wg = __fadd_rn(__fmul_rn(ca, g[i]), __fmul_rn(cs, n[i]));
        xn[i] = __fadd_rn(__fmul_rn(m[i], wg), __fmul_rn(__fsub_rn(1.0f, m[i]), smp[i]));
      }
      if (a.x_next) store_vec<float, VEC>(a.x_next + off, xn);
#pragma unroll
      for (int i = 0; i < VEC; ++i) smp[i] = xn[i];
    } else if (a.x_next) {
      store_vec<float, VEC>(a.x_next + off, smp);
    }
    // Step-boundary fusion: the next evaluation's stem input (channel c of every pixel of the NHWC network input) and
    // its timestep are written here -- the same roundings as the pack kernel (layout.cu), so the fused loop is
    // bit-identical to pack + evaluate.
    if (a.stem_out) {
      const long long px = (long long)b * a.hw + (long long)pv * VEC;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const long long e = (px + i) * a.stem_ld + c;
        if (a.stem_dtype == FIDM_F16) reinterpret_cast<__half*>(a.stem_out)[e] = from_f32<__half>(smp[i]);
        else if (a.stem_dtype == FIDM_BF16) reinterpret_cast<__nv_bfloat16*>(a.stem_out)[e] = from_f32<__nv_bfloat16>(smp[i]);
        else reinterpret_cast<float*>(a.stem_out)[e] = smp[i];
      }
    }
    if (a.t_out && idx < a.batch) a.t_out[idx] = a.t_out_value;
  }
}

}  // namespace fidm

extern "C" int fidm_sampler_step(const fidm_step_args* a, fidm_stream_t stream) {
  using namespace fidm;
  FIDM_REQUIRE(a != nullptr, FIDM_E_BADARG, "sampler_step: null args");
  FIDM_REQUIRE(a->batch > 0 && a->channels > 0 && a->hw > 0, FIDM_E_BADARG, "sampler_step: empty shape");
  FIDM_REQUIRE(a->mode >= 0 && a->mode <= 2, FIDM_E_BADARG, "sampler_step: bad mode %d", a->mode);
  FIDM_REQUIRE(a->coef && a->x, FIDM_E_BADARG, "sampler_step: coef/x are required");
  FIDM_REQUIRE(a->mask_channels == 1 || a->mask_channels == a->channels, FIDM_E_BADARG,
               "sampler_step: mask_channels must be 1 or channels");
  if (a->mode != FIDM_STEP_INJECT_ONLY) {
    FIDM_REQUIRE(a->model_out != nullptr, FIDM_E_BADARG, "sampler_step: model_out required for an update");
    FIDM_REQUIRE(a->t_dev || (a->t_update >= 0 && a->t_update < a->num_timesteps), FIDM_E_BADARG,
                 "sampler_step: t_update %d out of range [0,%d)", a->t_update, a->num_timesteps);
  }
  if (a->mode != FIDM_STEP_UPDATE_ONLY) {
    FIDM_REQUIRE(a->gt && a->keep_mask && a->inject_noise, FIDM_E_BADARG,
                 "sampler_step: gt/keep_mask/inject_noise required for an injection");
    FIDM_REQUIRE(a->t_dev || (a->t_inject >= 0 && a->t_inject < a->num_timesteps), FIDM_E_BADARG,
                 "sampler_step: t_inject %d out of range [0,%d)", a->t_inject, a->num_timesteps);
  }
  if (a->stem_out)
    FIDM_REQUIRE(a->stem_ld >= a->channels && (a->stem_dtype == FIDM_F32 || a->stem_dtype == FIDM_BF16 || a->stem_dtype == FIDM_F16),
                 FIDM_E_BADARG, "sampler_step: stem_out needs stem_ld >= channels and a dtype of F32 | BF16 | F16");
  StepParams p;
  p.a = *a;
  const bool vec4 = (a->hw % 4 == 0) && (((uintptr_t)a->x | (uintptr_t)a->model_out | (uintptr_t)a->z |
                                          (uintptr_t)a->gt | (uintptr_t)a->keep_mask | (uintptr_t)a->inject_noise |
                                          (uintptr_t)a->sample | (uintptr_t)a->pred_xstart | (uintptr_t)a->x_next |
                                          (uintptr_t)a->mean_out | (uintptr_t)a->logvar_out) % 16 == 0);
  const long long total = (long long)a->batch * a->channels * (a->hw / (vec4 ? 4 : 1));
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (vec4)
    sampler_step_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  else
    sampler_step_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  FIDM_CHECK_LAUNCH("sampler_step");
  return 0;
}
