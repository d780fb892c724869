/*
 * hd_oracle.c - see hd_oracle.h.  TEST INFRASTRUCTURE ONLY (parity: pinned).
 * Plain C, fp32 storage, as-written operation order of the reference.
 */
#include "hd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* elementary ops                                                            */
/* ------------------------------------------------------------------------ */
static inline float silu_f(float v) { return v / (1.0f + expf(-v)); }
static inline float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }
/* F.softplus(beta=1, threshold=20) */
static inline float softplus_f(float v) { return v > 20.0f ? v : log1pf(expf(v)); }
/* F.logsigmoid */
static inline float logsigmoid_f(float v) { return fminf(v, 0.0f) - log1pf(expf(-fabsf(v))); }

/* y[o] = b[o] + sum_k W[o][k]*in[k], one rounding per output (nn.Linear) */
static void linear(float* y, const float* in, const float* W, const float* b, int n_out, int n_in) {
  for (int o = 0; o < n_out; ++o) {
    const float* w = W + (size_t)o * n_in;
    double acc = 0.0;
#pragma omp simd reduction(+ : acc)
    for (int k = 0; k < n_in; ++k) acc += (double)w[k] * (double)in[k];
    y[o] = (float)(acc + (b ? (double)b[o] : 0.0));
  }
}

/* ------------------------------------------------------------------------ */
/* weight views into the flat buffer                                         */
/* ------------------------------------------------------------------------ */
typedef struct {
  const float *e1w, *e1b, *e2w, *e2b, *n1w, *n1b, *n2w, *n2b, *aw, *ab;
} gcl_w;
typedef struct {
  const float *c1w, *c1b, *c2w, *c2b, *c3w;
} equiv_w;

static const float* take(const float** p, int64_t n) {
  const float* r = *p;
  *p += n;
  return r;
}

int64_t hdo_weight_count(const hdo_config* c) {
  int64_t H = c->hidden_nf, Fi = c->in_node_nf;
  int64_t gcl = H * (2 * H + 2) + H + H * H + H + H * (2 * H) + H + H * H + H + (c->attention ? H + 1 : 0);
  int64_t eq = H * (2 * H + 2) + H + H * H + H + H;
  return H * Fi + H + Fi * H + Fi + (int64_t)c->n_layers * (c->inv_sublayers * gcl + eq);
}

/* ------------------------------------------------------------------------ */
/* egnn_new.py:260-266 coord2diff for one ordered pair                        */
/* ------------------------------------------------------------------------ */
static inline float radial_cd(const float* xi, const float* xj, float norm_constant, float cd[3]) {
  float d0 = xi[0] - xj[0], d1 = xi[1] - xj[1], d2 = xi[2] - xj[2];
  float r = d0 * d0 + d1 * d1 + d2 * d2;
  if (cd) {
    float nrm = sqrtf(r + 1e-8f) + norm_constant;
    cd[0] = d0 / nrm;
    cd[1] = d1 / nrm;
    cd[2] = d2 / nrm;
  }
  return r;
}

/* ------------------------------------------------------------------------ */
/* egnn_new.py:35-70  one GCL over the whole padded batch                     */
/* h [BN,H] in/out, x [BN,3] block-input coords, x0 [BN,3] egnn-input coords  */
/* ------------------------------------------------------------------------ */
static void gcl_forward(const hdo_config* c, const gcl_w* w, float* h, const float* x, const float* x0,
                        const int32_t* sizes, int B, int N) {
  const int H = c->hidden_nf, K1 = 2 * H + 2;
  float* hn = (float*)calloc((size_t)B * N * H, sizeof(float));
#pragma omp parallel
  {
    float* in = (float*)malloc(sizeof(float) * (size_t)(K1 + 6 * H));
    float *m1 = in + K1, *m = m1 + H, *agg = m + H, *in2 = agg + H; /* in2: 2H */
    float* o1 = in2 + 2 * H;                                        /* H */
#pragma omp for collapse(2) schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < N; ++i) {
        const int n = sizes[b];
        if (i >= n) continue; /* h * node_mask == 0, hn is zero-initialised */
        const size_t ri = (size_t)b * N + i;
        for (int k = 0; k < H; ++k) agg[k] = 0.0f;
        for (int j = 0; j < n; ++j) {
          if (j == i) continue; /* edge_mask = 1 - eye: contributes exactly 0 */
          const size_t rj = (size_t)b * N + j;
          memcpy(in, h + ri * H, sizeof(float) * H);     /* source = h[row] */
          memcpy(in + H, h + rj * H, sizeof(float) * H); /* target = h[col] */
          in[2 * H] = radial_cd(x + ri * 3, x + rj * 3, 0.f, NULL);
          in[2 * H + 1] = radial_cd(x0 + ri * 3, x0 + rj * 3, 0.f, NULL);
          linear(m1, in, w->e1w, w->e1b, H, K1);
          for (int k = 0; k < H; ++k) m1[k] = silu_f(m1[k]);
          linear(m, m1, w->e2w, w->e2b, H, H);
          for (int k = 0; k < H; ++k) m[k] = silu_f(m[k]);
          float att = 1.0f;
          if (c->attention) {
            float a;
            linear(&a, m, w->aw, w->ab, 1, H);
            att = sigmoid_f(a);
          }
          for (int k = 0; k < H; ++k) agg[k] += m[k] * att; /* scatter_add_, j ascending */
        }
        for (int k = 0; k < H; ++k) agg[k] = agg[k] / c->normalization_factor;
        memcpy(in2, h + ri * H, sizeof(float) * H);
        memcpy(in2 + H, agg, sizeof(float) * H);
        linear(o1, in2, w->n1w, w->n1b, H, 2 * H);
        for (int k = 0; k < H; ++k) o1[k] = silu_f(o1[k]);
        linear(m, o1, w->n2w, w->n2b, H, H);
        for (int k = 0; k < H; ++k) hn[ri * H + k] = h[ri * H + k] + m[k];
      }
    free(in);
  }
  memcpy(h, hn, sizeof(float) * (size_t)B * N * H);
  free(hn);
}

/* ------------------------------------------------------------------------ */
/* egnn_new.py:91-110  EquivariantUpdate; x updated in place                  */
/* ------------------------------------------------------------------------ */
static void equiv_forward(const hdo_config* c, const equiv_w* w, const float* h, float* x, const float* x0,
                          const int32_t* sizes, int B, int N) {
  const int H = c->hidden_nf, K1 = 2 * H + 2;
  const float range = (float)((double)c->coords_range / (double)c->n_layers);
  float* xn = (float*)calloc((size_t)B * N * 3, sizeof(float));
#pragma omp parallel
  {
    float* in = (float*)malloc(sizeof(float) * (size_t)(K1 + 2 * H));
    float *m1 = in + K1, *m = m1 + H;
#pragma omp for collapse(2) schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < N; ++i) {
        const int n = sizes[b];
        if (i >= n) continue;
        const size_t ri = (size_t)b * N + i;
        float ax = 0.f, ay = 0.f, az = 0.f;
        for (int j = 0; j < n; ++j) {
          if (j == i) continue;
          const size_t rj = (size_t)b * N + j;
          float cd[3];
          memcpy(in, h + ri * H, sizeof(float) * H);
          memcpy(in + H, h + rj * H, sizeof(float) * H);
          in[2 * H] = radial_cd(x + ri * 3, x + rj * 3, c->norm_constant, cd);
          in[2 * H + 1] = radial_cd(x0 + ri * 3, x0 + rj * 3, 0.f, NULL);
          linear(m1, in, w->c1w, w->c1b, H, K1);
          for (int k = 0; k < H; ++k) m1[k] = silu_f(m1[k]);
          linear(m, m1, w->c2w, w->c2b, H, H);
          for (int k = 0; k < H; ++k) m[k] = silu_f(m[k]);
          float phi;
          linear(&phi, m, w->c3w, NULL, 1, H);
          if (c->tanh) {
            float th = tanhf(phi);
            ax += cd[0] * th * range;
            ay += cd[1] * th * range;
            az += cd[2] * th * range;
          } else {
            ax += cd[0] * phi;
            ay += cd[1] * phi;
            az += cd[2] * phi;
          }
        }
        xn[ri * 3 + 0] = x[ri * 3 + 0] + ax / c->normalization_factor;
        xn[ri * 3 + 1] = x[ri * 3 + 1] + ay / c->normalization_factor;
        xn[ri * 3 + 2] = x[ri * 3 + 2] + az / c->normalization_factor;
      }
    free(in);
  }
  memcpy(x, xn, sizeof(float) * (size_t)B * N * 3);
  free(xn);
}

/* ------------------------------------------------------------------------ */
/* en_dynamics.py:49-122 + egnn_new.py:192-205                                */
/* ------------------------------------------------------------------------ */
int hdo_dynamics_forward(const hdo_config* c, const float* wbuf, const float* z, const float* t,
                         const int32_t* sizes, int32_t B, int32_t N, float* eps, hdo_trace* tr) {
  return hdo_dynamics_forward_ctx(c, wbuf, z, t, NULL, 0, sizes, B, N, eps, tr);
}

/* with `context` [B*N, C] appended after the time channel, unmasked (en_dynamics.py:76-79); the EGNN then has
 * in_node_nf = F + 1 + C channels and the context / time channels are sliced off its output (:99-105) */
int hdo_dynamics_forward_ctx(const hdo_config* c, const float* wbuf, const float* z, const float* t,
                             const float* context, int32_t C, const int32_t* sizes, int32_t B, int32_t N,
                             float* eps, hdo_trace* tr) {
  const int H = c->hidden_nf, Fi = c->in_node_nf, F = Fi - 1 - C, D = 3 + F;
  const size_t BN = (size_t)B * N;
  const float* p = wbuf;
  const float* emb_w = take(&p, (int64_t)H * Fi);
  const float* emb_b = take(&p, H);
  const float* out_w = take(&p, (int64_t)Fi * H);
  const float* out_b = take(&p, Fi);

  float* x = (float*)calloc(BN * 3, sizeof(float));
  float* x0 = (float*)calloc(BN * 3, sizeof(float));
  float* hin = (float*)calloc(BN * Fi, sizeof(float));
  float* h = (float*)calloc(BN * H, sizeof(float));
  /* xh * node_mask ; h = cat[h, t]  (time is NOT masked, en_dynamics.py:66-74) */
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N; ++i) {
      size_t r = (size_t)b * N + i;
      float mk = i < sizes[b] ? 1.f : 0.f;
      for (int d = 0; d < 3; ++d) x[r * 3 + d] = z[r * D + d] * mk;
      for (int f = 0; f < F; ++f) hin[r * Fi + f] = z[r * D + 3 + f] * mk;
      hin[r * Fi + F] = t[b];
      for (int k = 0; k < C; ++k) hin[r * Fi + F + 1 + k] = context[r * C + k];
    }
  memcpy(x0, x, sizeof(float) * BN * 3);
  for (size_t r = 0; r < BN; ++r) linear(h + r * H, hin + r * Fi, emb_w, emb_b, H, Fi);
  if (tr && tr->h_embed) memcpy(tr->h_embed, h, sizeof(float) * BN * H);

  for (int l = 0; l < c->n_layers; ++l) {
    float* xb = (float*)malloc(sizeof(float) * BN * 3); /* coords at block entry */
    memcpy(xb, x, sizeof(float) * BN * 3);
    for (int s = 0; s < c->inv_sublayers; ++s) {
      gcl_w g;
      g.e1w = take(&p, (int64_t)H * (2 * H + 2));
      g.e1b = take(&p, H);
      g.e2w = take(&p, (int64_t)H * H);
      g.e2b = take(&p, H);
      g.n1w = take(&p, (int64_t)H * 2 * H);
      g.n1b = take(&p, H);
      g.n2w = take(&p, (int64_t)H * H);
      g.n2b = take(&p, H);
      g.aw = g.ab = NULL;
      if (c->attention) {
        g.aw = take(&p, H);
        g.ab = take(&p, 1);
      }
      gcl_forward(c, &g, h, xb, x0, sizes, B, N);
      if (tr && l == 0 && s == 0 && tr->h_gcl0) memcpy(tr->h_gcl0, h, sizeof(float) * BN * H);
      if (tr && l == 0 && s == 1 && tr->h_gcl1) memcpy(tr->h_gcl1, h, sizeof(float) * BN * H);
    }
    equiv_w e;
    e.c1w = take(&p, (int64_t)H * (2 * H + 2));
    e.c1b = take(&p, H);
    e.c2w = take(&p, (int64_t)H * H);
    e.c2b = take(&p, H);
    e.c3w = take(&p, H);
    equiv_forward(c, &e, h, x, x0, sizes, B, N);
    free(xb);
    if (tr && l == 0 && tr->x_block0) memcpy(tr->x_block0, x, sizeof(float) * BN * 3);
  }

  /* embedding_out, * node_mask ; velocity ; drop time channel */
  float* ho = (float*)calloc(BN * Fi, sizeof(float));
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < sizes[b] && i < N; ++i) {
      size_t r = (size_t)b * N + i;
      linear(ho + r * Fi, h + r * H, out_w, out_b, Fi, H);
    }
  if (tr && tr->h_final) memcpy(tr->h_final, ho, sizeof(float) * BN * Fi);
  if (tr && tr->x_final) memcpy(tr->x_final, x, sizeof(float) * BN * 3);

  int nan = 0;
  for (size_t r = 0; r < BN; ++r) {
    int b = (int)(r / N), i = (int)(r % N);
    float mk = i < sizes[b] ? 1.f : 0.f;
    for (int d = 0; d < 3; ++d) {
      float v = (x[r * 3 + d] - x0[r * 3 + d]) * mk;
      if (isnan(v)) nan = 1;
      eps[r * D + d] = v;
    }
    for (int f = 0; f < F; ++f) eps[r * D + 3 + f] = ho[r * Fi + f];
  }
  if (nan)
    for (size_t r = 0; r < BN; ++r)
      for (int d = 0; d < 3; ++d) eps[r * D + d] = 0.f;
  /* remove_mean_with_mask on vel */
  for (int b = 0; b < B; ++b)
    for (int d = 0; d < 3; ++d) {
      float s = 0.f;
      for (int i = 0; i < N; ++i) s += eps[((size_t)b * N + i) * D + d];
      float mean = s / (float)sizes[b];
      for (int i = 0; i < sizes[b] && i < N; ++i) eps[((size_t)b * N + i) * D + d] -= mean;
    }
  free(x);
  free(x0);
  free(hin);
  free(h);
  free(ho);
  return nan;
}

/* ------------------------------------------------------------------------ */
/* schedule algebra                                                           */
/* ------------------------------------------------------------------------ */
void hdo_step_scalars(float gs, float gt, float out[3]) {
  float sigma2_ts = -expm1f(softplus_f(gs) - softplus_f(gt));
  float log_a2_ts = logsigmoid_f(-gt) - logsigmoid_f(-gs);
  float alpha_ts = expf(0.5f * log_a2_ts);
  float sigma_ts = sqrtf(sigma2_ts);
  float sigma_s = sqrtf(sigmoid_f(gs)), sigma_t = sqrtf(sigmoid_f(gt));
  out[0] = alpha_ts;
  out[1] = sigma2_ts / alpha_ts / sigma_t;
  out[2] = sigma_ts * sigma_s / sigma_t;
}

void hdo_final_scalars(float g0, float out[3]) {
  out[0] = sqrtf(sigmoid_f(-g0));      /* alpha_0 */
  out[1] = sqrtf(sigmoid_f(g0));       /* sigma_0 */
  out[2] = expf(-(-0.5f * g0));        /* SNR(-0.5 gamma_0) */
}

static void remove_mean_x(float* v, const int32_t* sizes, int B, int N, int D) {
  for (int b = 0; b < B; ++b)
    for (int d = 0; d < 3; ++d) {
      float s = 0.f;
      for (int i = 0; i < N; ++i) s += v[((size_t)b * N + i) * D + d];
      float mean = s / (float)sizes[b];
      for (int i = 0; i < sizes[b] && i < N; ++i) v[((size_t)b * N + i) * D + d] -= mean;
    }
}

/* masked, CoG-free noise of diffusion_qm9.py:445-456 into nz [B,N,3+F] */
static void combined_noise(const float* rx, const float* rh, const int32_t* sizes, int B, int N, int F, float* nz) {
  const int D = 3 + F;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N; ++i) {
      size_t r = (size_t)b * N + i;
      float mk = i < sizes[b] ? 1.f : 0.f;
      for (int d = 0; d < 3; ++d) nz[r * D + d] = rx[r * 3 + d] * mk;
      for (int f = 0; f < F; ++f) nz[r * D + 3 + f] = rh[r * F + f] * mk;
    }
  remove_mean_x(nz, sizes, B, N, D);
}

void hdo_reverse_step(const float* zt, const float* eps_in, const float* rx, const float* rh, const int32_t* sizes,
                      int32_t B, int32_t N, int32_t F, const float* sc, float* zs) {
  const int D = 3 + F;
  const size_t n = (size_t)B * N * D;
  float* eps = (float*)malloc(sizeof(float) * n);
  float* nz = (float*)malloc(sizeof(float) * n);
  memcpy(eps, eps_in, sizeof(float) * n);
  remove_mean_x(eps, sizes, B, N, D); /* :330 */
  combined_noise(rx, rh, sizes, B, N, F, nz);
  for (size_t k = 0; k < n; ++k) {
    const float* s = sc + 3 * (k / ((size_t)N * D)); /* [B,1,1] scalars: one row per molecule */
    float mu = zt[k] / s[0] - s[1] * eps[k];         /* :331 */
    zs[k] = mu + s[2] * nz[k];                       /* :337, :442 */
  }
  remove_mean_x(zs, sizes, B, N, D); /* :340-344 */
  free(eps);
  free(nz);
}

void hdo_final_decode(const float* z0, const float* eps0, const float* rx, const float* rh, const int32_t* sizes,
                      int32_t B, int32_t N, int32_t F, const float* sc_all, float norm_x, float norm_h, float bias_h,
                      float* x, float* h) {
  const int D = 3 + F;
  const size_t n = (size_t)B * N * D;
  float* nz = (float*)malloc(sizeof(float) * n);
  combined_noise(rx, rh, sizes, B, N, F, nz);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N; ++i) {
      size_t r = (size_t)b * N + i;
      float mk = i < sizes[b] ? 1.f : 0.f;
      const float* sc = sc_all + 3 * b;
      for (int d = 0; d < 3; ++d) {
        float mu = 1.0f / sc[0] * (z0[r * D + d] - sc[1] * eps0[r * D + d]); /* :244 */
        x[r * 3 + d] = (mu + sc[2] * nz[r * D + d]) * norm_x;                /* :305-309 */
      }
      for (int f = 0; f < F; ++f) h[r * F + f] = (z0[r * D + 3 + f] * norm_h + bias_h) * mk;
    }
  free(nz);
}

/* ------------------------------------------------------------------------ */
/* noise_model.py:163-200                                                     */
/* ------------------------------------------------------------------------ */
static float gamma_tilde(const float* p, float t) {
  const float l1w = softplus_f(p[2]), l1b = p[3];
  const float *l2w = p + 4, *l2b = p + 4 + 1024, *l3w = p + 4 + 2048;
  const float l3b = p[4 + 3072];
  float l1t = l1w * t + l1b;
  double acc = 0.0;
  for (int k = 0; k < 1024; ++k) {
    float hdn = sigmoid_f(softplus_f(l2w[k]) * l1t + l2b[k]);
    acc += (double)softplus_f(l3w[k]) * (double)hdn;
  }
  return l1t + (float)(acc + (double)l3b);
}

float hdo_gamma(const float* p, float t) {
  float g0 = gamma_tilde(p, 0.f), g1 = gamma_tilde(p, 1.f), gt = gamma_tilde(p, t);
  float nrm = (gt - g0) / (g1 - g0);
  return p[0] + (p[1] - p[0]) * nrm;
}
