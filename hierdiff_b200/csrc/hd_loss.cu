// hd_loss.cu - forward value of the diffusion loss / negative log-likelihood: the glue of
// DiffusionQM9.forward -> nll -> compute_loss (train_module/diffusion_qm9.py:701-751, :675-699, :530-673) around the
// network calls, which go through the same fused EGNN kernels as the sampler.  No gradients: this is the
// validation_step / test_step value (:779-785) and the training objective's forward value.
//
//   loss_prepare_k   x - CoG (models/utils.py:43-57), normalize (:165-172), [x | h] concatenation           one CTA / molecule
//   noise_mix_k      z_t = alpha_t * xh + sigma_t * eps on the real nodes (:573)                               one CTA / molecule
//   loss_terms_k     per-molecule reductions and the scalar algebra of compute_loss (:579-668) + nll (:694-697)
#include <math.h>

#include "hd_common.cuh"

namespace hd {
namespace loss {

constexpr int NT = 128;

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // s_red may still be read from the previous reduction
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(NT) loss_prepare_k(const float* __restrict__ x, const float* __restrict__ h,
                                                     const int32_t* __restrict__ sizes, int N, int F, float norm_x,
                                                     float norm_h, float bias_h, int center, float* __restrict__ xh,
                                                     int32_t* __restrict__ flags) {
  __shared__ float s_red[NT / 32];
  const int b = blockIdx.x, n = sizes[b], D = 3 + F, tid = threadIdx.x;
  const float* xb = x + (int64_t)b * N * 3;
  float m[3] = {0.f, 0.f, 0.f}, masked = 0.f;
  for (int i = tid; i < N; i += NT)
    for (int c = 0; c < 3; ++c) {
      const float v = xb[i * 3 + c];
      if (i < n) m[c] += v; else masked += fabsf(v);
    }
  float mean[3];
  for (int c = 0; c < 3; ++c) mean[c] = center ? block_sum(m[c], s_red) / (float)n : 0.f;
  masked = block_sum(masked, s_red);
  if (tid == 0 && masked > 1e-5f && flags) atomicOr(flags, HD_FLAG_MASK);   // models/utils.py:47-50
  for (int idx = tid; idx < N * D; idx += NT) {
    const int i = idx / D, c = idx % D;
    float v = 0.f;
    if (i < n) v = c < 3 ? (xb[i * 3 + c] - mean[c]) / norm_x : (h[((int64_t)b * N + i) * F + (c - 3)] - bias_h) / norm_h;
    xh[(int64_t)b * N * D + idx] = v;
  }
}

__device__ __forceinline__ float sigmoidf_(float g) { return 1.0f / (1.0f + expf(-g)); }

__global__ void __launch_bounds__(NT) noise_mix_k(const float* __restrict__ xh, const float* __restrict__ eps,
                                                  const float* __restrict__ gamma, const int32_t* __restrict__ sizes, int N,
                                                  int D, float* __restrict__ z, int32_t* __restrict__ flags) {
  __shared__ float s_red[NT / 32];
  const int b = blockIdx.x, n = sizes[b], tid = threadIdx.x;
  const float g = gamma[b];
  const float alpha = sqrtf(sigmoidf_(-g)), sigma = sqrtf(sigmoidf_(g));   // diffusion_qm9.py:148-154
  float m[3] = {0.f, 0.f, 0.f}, big = 0.f;
  for (int idx = tid; idx < N * D; idx += NT) {
    const int i = idx / D, c = idx % D;
    const int64_t o = (int64_t)b * N * D + idx;
    const float v = i < n ? alpha * xh[o] + sigma * eps[o] : 0.f;
    z[o] = v;
    if (c < 3) {
      m[c] += v;
      big = fmaxf(big, fabsf(v));
    }
  }
  // assert_mean_zero_with_mask(z_t[:, :, :3]) (models/utils.py:65-70): |sum| / (largest + 1e-10) < 1e-2
  float tot = 0.f;
  for (int c = 0; c < 3; ++c) tot = fmaxf(tot, fabsf(block_sum(m[c], s_red)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) big = fmaxf(big, __shfl_xor_sync(0xffffffffu, big, o));
  __syncthreads();
  if ((tid & 31) == 0) s_red[tid >> 5] = big;
  __syncthreads();
  if (tid == 0) {
    float mx = 0.f;
    for (int w = 0; w < NT / 32; ++w) mx = fmaxf(mx, s_red[w]);
    if (flags && tot / (mx + 1e-10f) >= 1e-2f) atomicOr(flags, HD_FLAG_COG);
  }
}

// log( Phi((c + 0.5)/s) - Phi((c - 0.5)/s) + 1e-10 ), Phi(v) = 0.5 (1 + erf(v / sqrt 2))   (diffusion_qm9.py:500-504)
__device__ __forceinline__ float log_bin_mass(float c, float s) {
  const float a = 0.5f * (1.0f + erff(((c + 0.5f) / s) / 1.4142135623730951f));
  const float b = 0.5f * (1.0f + erff(((c - 0.5f) / s) / 1.4142135623730951f));
  return logf(a - b + 1e-10f);
}

struct TermArgs {
  const float *xh, *z_t, *eps_t, *net_t, *z_0, *eps_0, *net_0, *t_int, *gamma_s, *gamma_t, *gamma_0, *gamma_T;
  const int32_t* sizes;
  float *nll, *loss, *error, *terms;
  int N, F;
  hd_loss_config cfg;
};

__global__ void __launch_bounds__(NT) loss_terms_k(const TermArgs a) {
  __shared__ float s_red[NT / 32];
  const int b = blockIdx.x, n = a.sizes[b], N = a.N, F = a.F, D = 3 + F, tid = threadIdx.x;
  const hd_loss_config c = a.cfg;
  // which network call carries the L0 term: the separate t = 0 call (t0_always) or the same one (:626-652)
  const float* z0 = c.t0_always ? a.z_0 : a.z_t;
  const float* e0 = c.t0_always ? a.eps_0 : a.eps_t;
  const float* n0 = c.t0_always ? a.net_0 : a.net_t;
  const float g0 = c.t0_always ? a.gamma_0[b] : a.gamma_t[b];
  const float sigma0_int = sqrtf(sigmoidf_(g0)) * c.norm_int;
  const int64_t base = (int64_t)b * N * D;
  float err = 0.f, mu_x = 0.f, mu_h = 0.f, sx = 0.f, sh = 0.f, li = 0.f;
  for (int idx = tid; idx < N * D; idx += NT) {
    const int i = idx / D, ch = idx % D;
    const int64_t o = base + idx;
    const float d = a.eps_t[o] - a.net_t[o];
    err += d * d;                                        // compute_error (:250-258)
    const float v = a.xh[o];
    if (i < n) { if (ch < 3) mu_x += v * v; else mu_h += v * v; }
    const float d0 = e0[o] - n0[o];
    if (ch < 3) sx += d0 * d0;
    // continuous features: the reference slices the network output with `[:, :, :n_dims + int_nf : ...]` (:473), i.e.
    // start None / stop 3 + int_nf / step 3 + int_nf + cont_nf = channel 0 only, broadcast against the cont_nf noise
    // channels; reproduced as written
    if (ch >= 3 + c.int_nf && ch < 3 + c.int_nf + c.cont_nf) {
      const float dh = e0[o] - n0[base + (int64_t)i * D];
      sh += dh * dh;
    }
    if (i < n && ch >= 3 && ch < 3 + c.int_nf) {
      const float h_int = rintf(v * c.norm_int + c.bias_int);          // :490 torch.round (half to even)
      const float est = z0[o] * c.norm_int + c.bias_int;
      li += log_bin_mass(h_int - est, sigma0_int);
    }
  }
  err = block_sum(err, s_red);
  mu_x = block_sum(mu_x, s_red);
  mu_h = block_sum(mu_h, s_red);
  sx = block_sum(sx, s_red);
  sh = block_sum(sh, s_red);
  li = block_sum(li, s_red);
  if (tid != 0) return;
  const float LOG_2PI_HALF = 0.9189385332046727f;
  const float fn = (float)n, d = (fn - 1.0f) * 3.0f;
  // kl_prior (:206-234): q = N(alpha_T xh, sigma_T), p = N(0, 1)
  const float gT = a.gamma_T[b];
  const float aT2 = sigmoidf_(-gT), sT = sqrtf(sigmoidf_(gT));
  const float kl_h = fn * F * (logf(1.0f / sT) + 0.5f * sT * sT - 0.5f) + 0.5f * aT2 * mu_h;
  const float kl_x = d * logf(1.0f / sT) + 0.5f * (d * sT * sT + aT2 * mu_x) - 0.5f * d;
  const float kl = kl_x + kl_h;
  // constants of -log p(x, h | z0) (:260-289)
  const float gz = a.gamma_0[b];
  float nlc = -(d * (-0.5f * gz - LOG_2PI_HALF)) - (fn * F * (-0.5f * gz - LOG_2PI_HALF));
  float error = err;
  if (c.l2_training) {
    error = err / (float)(D * N);
    nlc = 0.f;
  }
  const float snr_w = c.l2_training ? 1.0f : expf(-(a.gamma_s[b] - a.gamma_t[b])) - 1.0f;
  const float loss_gt0 = 0.5f * snr_w * error;
  const float L0 = -(-0.5f * sx - 0.5f * sh + li);
  float est, loss;
  if (c.t0_always) {
    est = (float)c.T * loss_gt0;
    loss = kl + est + nlc + L0;
  } else {
    const float lt = a.t_int[b] == 0.f ? L0 : loss_gt0;
    est = c.l2_training ? lt : (float)(c.T + 1) * lt;
    loss = kl + est + nlc;
  }
  const float delta_log_px = c.l2_training ? 0.f : -d * logf(c.norm_x);   // :165-167, :681-683
  a.loss[b] = loss;
  a.nll[b] = loss - delta_log_px;
  a.error[b] = error;
  if (a.terms) {
    a.terms[4 * b] = kl;
    a.terms[4 * b + 1] = est;
    a.terms[4 * b + 2] = nlc;
    a.terms[4 * b + 3] = L0;
  }
}

}  // namespace loss
}  // namespace hd

using namespace hd;

extern "C" {

HD_API int32_t hd_loss_prepare(const float* x, const float* h, const int32_t* sizes, int32_t B, int32_t N, int32_t F,
                               float norm_x, float norm_h, float bias_h, int32_t center, float* xh, int32_t* flags,
                               hd_stream_t stream) {
  if (!x || !h || !sizes || !xh || B < 1 || N < 1 || F < 1) {
    set_error("hd_loss_prepare: bad argument");
    return HD_E_INVALID;
  }
  loss::loss_prepare_k<<<B, loss::NT, 0, static_cast<cudaStream_t>(stream)>>>(x, h, sizes, N, F, norm_x, norm_h, bias_h,
                                                                           center, xh, flags);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

HD_API int32_t hd_loss_noise_mix(const float* xh, const float* eps, const float* gamma, const int32_t* sizes, int32_t B,
                                 int32_t N, int32_t F, float* z, int32_t* flags, hd_stream_t stream) {
  if (!xh || !eps || !gamma || !sizes || !z || B < 1 || N < 1 || F < 1) {
    set_error("hd_loss_noise_mix: bad argument");
    return HD_E_INVALID;
  }
  loss::noise_mix_k<<<B, loss::NT, 0, static_cast<cudaStream_t>(stream)>>>(xh, eps, gamma, sizes, N, 3 + F, z, flags);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

HD_API int32_t hd_loss_terms(const hd_loss_config* cfg, const float* xh, const float* z_t, const float* eps_t,
                             const float* net_t, const float* z_0, const float* eps_0, const float* net_0,
                             const float* t_int, const float* gamma_s, const float* gamma_t, const float* gamma_0,
                             const float* gamma_T, const int32_t* sizes, int32_t B, int32_t N, int32_t F, float* nll,
                             float* loss, float* error, float* terms, hd_stream_t stream) {
  if (!cfg || !xh || !z_t || !eps_t || !net_t || !t_int || !gamma_s || !gamma_t || !gamma_0 || !gamma_T || !sizes || !nll ||
      !loss || !error || B < 1 || N < 1 || F < 1 || (cfg->t0_always && (!z_0 || !eps_0 || !net_0)) ||
      cfg->int_nf < 0 || cfg->cont_nf < 0 || cfg->int_nf + cfg->cont_nf > F) {
    set_error("hd_loss_terms: bad argument");
    return HD_E_INVALID;
  }
  loss::TermArgs a{xh, z_t, eps_t, net_t, z_0, eps_0, net_0, t_int, gamma_s, gamma_t, gamma_0, gamma_T, sizes,
                   nll, loss, error, terms, N, F, *cfg};
  loss::loss_terms_k<<<B, loss::NT, 0, static_cast<cudaStream_t>(stream)>>>(a);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

}  // extern "C"
