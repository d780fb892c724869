// hd_fp32.cu - CUDA-core fp32 engine (HD_ENGINE_FP32): every EGNN sub-layer with plain FFMA
// arithmetic in the reference's summation structure (per edge j ascending).  It is the
// numerically closest engine to the reference and the on-device cross-check of the
// tensor-core engines; the tensor-core kernels in hd_tc.cu are the performance path.
#include "hd_common.cuh"

namespace hd {

// ---------------------------------------------------------------------------------------
// Y[r, o] = epilogue( bias[o] + sum_k [X1 | X2][r, k] * WT[k, o] )        (nn.Linear)
//   mode 0: store; mode 1: SiLU; mode 2: Y = (resid + v) * node_mask  (GCL residual, egnn_new.py:58-61,68-69)
// ---------------------------------------------------------------------------------------
constexpr int LIN_ROWS = 8;
__global__ void __launch_bounds__(256) linear_k(const float* __restrict__ X1, int ld1, int K1,
                                                const float* __restrict__ X2, int ld2, int K2,
                                                const float* __restrict__ WT, int n_out,
                                                const float* __restrict__ bias, float* __restrict__ Y, int ldy,
                                                int rows, int mode, const float* __restrict__ resid,
                                                const int32_t* __restrict__ sizes, int N) {
  extern __shared__ float xs[];  // [LIN_ROWS][K1+K2]
  const int K = K1 + K2;
  const int r0 = blockIdx.x * LIN_ROWS;
  for (int idx = threadIdx.x; idx < LIN_ROWS * K; idx += blockDim.x) {
    int r = idx / K, k = idx % K, row = r0 + r;
    float v = 0.f;
    if (row < rows) v = k < K1 ? X1[(int64_t)row * ld1 + k] : X2[(int64_t)row * ld2 + (k - K1)];
    xs[idx] = v;
  }
  __syncthreads();
  const int o = blockIdx.y * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  float acc[LIN_ROWS];
  const float b = bias ? bias[o] : 0.f;
#pragma unroll
  for (int r = 0; r < LIN_ROWS; ++r) acc[r] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float w = WT[(int64_t)k * n_out + o];
#pragma unroll
    for (int r = 0; r < LIN_ROWS; ++r) acc[r] = fmaf(xs[r * K + k], w, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < LIN_ROWS; ++r) {
    const int row = r0 + r;
    if (row >= rows) break;
    float v = acc[r] + b;
    if (mode == 1) v = silu_acc(v);
    if (mode == 2) {
      const int bm = row / N, i = row % N;
      v = i < sizes[bm] ? resid[(int64_t)row * ldy + o] + v : 0.f;
    }
    Y[(int64_t)row * ldy + o] = v;
  }
}

static int linear(const FwdCtx& c, const float* X1, int ld1, int K1, const float* X2, int ld2, int K2,
                  const float* WT, int n_out, const float* bias, float* Y, int ldy, int mode,
                  const float* resid) {
  const int rows = c.B * c.N;
  dim3 grid((rows + LIN_ROWS - 1) / LIN_ROWS, (n_out + 255) / 256);
  size_t smem = sizeof(float) * LIN_ROWS * (K1 + K2);
  linear_k<<<grid, 256, smem, c.stream>>>(X1, ld1, K1, X2, ld2, K2, WT, n_out, bias, Y, ldy, rows, mode, resid,
                                          c.sizes, c.N);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

// ---------------------------------------------------------------------------------------
// fused edge kernel, one CTA per receiver node (b,i); thread = channel.
//   m1_ij = SiLU(A_i + B_j + r_ij*wr + d0_ij*wd)            (layer 1 of edge_mlp / coord_mlp, split per node)
//   m_ij  = SiLU(W2 m1_ij + b2)
//   GCL  : agg_i = sum_j m_ij * sigmoid(wa.m_ij + ba) * mask_ij / norm               (egnn_new.py:35-62)
//   EQUIV: x_i  += sum_j cd_ij * tanh(u.m_ij) * range * mask_ij / norm               (egnn_new.py:91-110)
// ---------------------------------------------------------------------------------------
constexpr int JB = 8;
template <bool GCL>
__global__ void __launch_bounds__(256) edge_fp32_k(const float* __restrict__ AB, const float* __restrict__ x,
                                                   const float* __restrict__ x0, const float* __restrict__ wr,
                                                   const float* __restrict__ wd, const float* __restrict__ W2T,
                                                   const float* __restrict__ b2, const float* __restrict__ wa,
                                                   const float* __restrict__ ba, const int32_t* __restrict__ sizes,
                                                   int N, int attention, int use_tanh, float range,
                                                   float norm_constant, float inv_norm_div, float* __restrict__ out) {
  __shared__ float sm1[JB][H];
  __shared__ float red[8][JB];
  __shared__ float s_att[JB];
  const int b = blockIdx.x / N, i = blockIdx.x % N, t = threadIdx.x;
  const int n = sizes[b];
  const int64_t ri = (int64_t)b * N + i;
  if (i >= n) {  // padded receiver: h*mask / x*mask == 0
    if (GCL) out[ri * H + t] = 0.f;
    else if (t < 3) out[ri * 3 + t] = 0.f;
    return;
  }
  const float a = AB[ri * 2 * H + t];
  const float w_r = wr[t], w_d = wd[t], bias2 = b2[t], w_a = wa[t];
  const float xi0 = x[ri * 3], xi1 = x[ri * 3 + 1], xi2 = x[ri * 3 + 2];
  const float oi0 = x0[ri * 3], oi1 = x0[ri * 3 + 1], oi2 = x0[ri * 3 + 2];
  float agg = 0.f;            // GCL: channel t ; EQUIV: t<3 -> coordinate t
  const int warp = t >> 5, lane = t & 31;
  for (int j0 = 0; j0 < n; j0 += JB) {
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) {
      const int j = j0 + jj;
      float v = 0.f;
      if (j < n) {
        const int64_t rj = (int64_t)b * N + j;
        const float d0 = xi0 - x[rj * 3], d1 = xi1 - x[rj * 3 + 1], d2 = xi2 - x[rj * 3 + 2];
        const float e0 = oi0 - x0[rj * 3], e1 = oi1 - x0[rj * 3 + 1], e2 = oi2 - x0[rj * 3 + 2];
        const float r = d0 * d0 + d1 * d1 + d2 * d2, rr = e0 * e0 + e1 * e1 + e2 * e2;
        v = silu_acc(a + AB[rj * 2 * H + H + t] + r * w_r + rr * w_d);
      }
      sm1[jj][t] = v;
    }
    __syncthreads();
    float acc[JB];
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) acc[jj] = 0.f;
    for (int k = 0; k < H; k += 4) {
      const float w0 = W2T[(k + 0) * H + t], w1 = W2T[(k + 1) * H + t], w2 = W2T[(k + 2) * H + t],
                  w3 = W2T[(k + 3) * H + t];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) {
        const float4 m = *reinterpret_cast<const float4*>(&sm1[jj][k]);
        acc[jj] = fmaf(m.x, w0, acc[jj]);
        acc[jj] = fmaf(m.y, w1, acc[jj]);
        acc[jj] = fmaf(m.z, w2, acc[jj]);
        acc[jj] = fmaf(m.w, w3, acc[jj]);
      }
    }
    float m[JB];
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) {
      m[jj] = silu_acc(acc[jj] + bias2);
      float p = warp_sum(m[jj] * w_a);   // att_mlp / coord_mlp.4 dot over channels
      if (lane == 0) red[warp][jj] = p;
    }
    __syncthreads();
    if (t < JB) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][t];
      s_att[t] = s;
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) {
      const int j = j0 + jj;
      if (j >= n || j == i) continue;  // edge_mask
      if (GCL) {
        const float att = attention ? sigmoid_acc(s_att[jj] + ba[0]) : 1.0f;
        agg += m[jj] * att;
      } else if (t < 3) {
        const int64_t rj = (int64_t)b * N + j;
        const float d0 = xi0 - x[rj * 3], d1 = xi1 - x[rj * 3 + 1], d2 = xi2 - x[rj * 3 + 2];
        const float r = d0 * d0 + d1 * d1 + d2 * d2;
        const float nrm = sqrtf(r + 1e-8f) + norm_constant;
        const float dd = t == 0 ? d0 : (t == 1 ? d1 : d2);
        const float phi = s_att[jj];
        agg += use_tanh ? (dd / nrm) * tanhf(phi) * range : (dd / nrm) * phi;
      }
    }
    __syncthreads();
  }
  if (GCL) out[ri * H + t] = agg / inv_norm_div;
  else if (t < 3) out[ri * 3 + t] = (t == 0 ? xi0 : (t == 1 ? xi1 : xi2)) + agg / inv_norm_div;
}

static float norm_div(const FwdCtx& c) {
  // unsorted_segment_sum (egnn_new.py:269-289): 'sum' divides by normalization_factor, 'mean' by the
  // number of edges listed for the row, which is N for the dense edge list of en_dynamics.py:131-136
  return c.cfg->aggregation_mean ? (float)c.N : c.cfg->normalization_factor;
}

int fp32_gcl(const FwdCtx& c, int si, float* h, const float* x, const float* x0) {
  const SubLayer& S = c.L->subs[si];
  auto F = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  float* ab = reinterpret_cast<float*>(c.ws + c.W.ab);
  float* agg = reinterpret_cast<float*>(c.ws + c.W.agg);
  float* hid = reinterpret_cast<float*>(c.ws + c.W.hid);
  int rc;
  // A = h W1a^T + b1 ; B = h W1b^T   (packed b1 image is [b1 | 0], so the bias lands on the A half only)
  if ((rc = linear(c, h, H, H, nullptr, 0, 0, F(S.w1abT), 2 * H, F(S.b1), ab, 2 * H, 0, nullptr))) return rc;
  edge_fp32_k<true><<<c.B * c.N, 256, 0, c.stream>>>(ab, x, x0, F(S.wr), F(S.wd), F(S.w2T), F(S.b2), F(S.wa),
                                                    F(S.ba), c.sizes, c.N, c.cfg->attention, 0, 0.f, 0.f,
                                                    norm_div(c), agg);
  HD_CHECK_LAUNCH();
  if ((rc = linear(c, h, H, H, agg, H, H, F(S.v1T), H, F(S.c1), hid, H, 1, nullptr))) return rc;
  if ((rc = linear(c, hid, H, H, nullptr, 0, 0, F(S.v2T), H, F(S.c2), h, H, 2, h))) return rc;
  return HD_OK;
}

int fp32_equiv(const FwdCtx& c, int si, const float* h, const float* x, const float* x0, float* x_out) {
  const SubLayer& S = c.L->subs[si];
  auto F = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  float* ab = reinterpret_cast<float*>(c.ws + c.W.ab);
  int rc;
  if ((rc = linear(c, h, H, H, nullptr, 0, 0, F(S.w1abT), 2 * H, F(S.b1), ab, 2 * H, 0, nullptr))) return rc;
  const float range = c.cfg->coords_range / (float)c.cfg->n_layers;
  edge_fp32_k<false><<<c.B * c.N, 256, 0, c.stream>>>(ab, x, x0, F(S.wr), F(S.wd), F(S.w2T), F(S.b2), F(S.wa),
                                                     F(S.ba), c.sizes, c.N, 0, c.cfg->tanh, range,
                                                     c.cfg->norm_constant, norm_div(c), x_out);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

int fp32_edge_only(const FwdCtx& c, int si, const float* x, const float* x0) {
  const SubLayer& S = c.L->subs[si];
  auto F = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  const float* ab = reinterpret_cast<const float*>(c.ws + c.W.ab);
  if (S.is_gcl) {
    edge_fp32_k<true><<<c.B * c.N, 256, 0, c.stream>>>(ab, x, x0, F(S.wr), F(S.wd), F(S.w2T), F(S.b2), F(S.wa),
                                                      F(S.ba), c.sizes, c.N, c.cfg->attention, 0, 0.f, 0.f,
                                                      norm_div(c), reinterpret_cast<float*>(c.ws + c.W.agg));
  } else {
    edge_fp32_k<false><<<c.B * c.N, 256, 0, c.stream>>>(
        ab, x, x0, F(S.wr), F(S.wd), F(S.w2T), F(S.b2), F(S.wa), F(S.ba), c.sizes, c.N, 0, c.cfg->tanh,
        c.cfg->coords_range / (float)c.cfg->n_layers, c.cfg->norm_constant, norm_div(c),
        reinterpret_cast<float*>(c.ws + c.W.x2));
  }
  HD_CHECK_LAUNCH();
  return HD_OK;
}

int linear_fp32(const FwdCtx& c, const float* X1, int ld1, int K1, const float* X2, int ld2, int K2, const float* WT,
                int n_out, const float* bias, float* Y, int ldy, int mode, const float* resid) {
  return linear(c, X1, ld1, K1, X2, ld2, K2, WT, n_out, bias, Y, ldy, mode, resid);
}

}  // namespace hd
