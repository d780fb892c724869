// hd_egcl.cu - the stage-2 equivariant layer E_GCL (reference ROOT models/egnn/gcl.py:9-209, as built by
// models/edge_denoise.py:35-43) on the dense edge list of edge_denoise.py:506-524: SURVEY.md 8f-3, first correct CUDA
// path (fp32 CUDA-core arithmetic; not yet on the tensor cores).
//
// Unlike the coarse-grained GCL, an edge carries a feature VECTOR (edges_in_d = hidden_nf for gcl_full_*), the messages
// are aggregated over `col` (gcl.py:121), and the layer also updates the edge features (edge_model :109-115).  The first
// Linear of mes_mlp is split per input block, W1 . [h_row, h_col, radial, e] = A_row + B_col + radial * w_r + We . e, so
// the [E, 2H+1+De] concatenation of gcl.py:93-97 never exists; the same for edge_mlp.0 on [m, radial, e].
//
// Two edge sources: the dense list (row = col = null: every (b, i, j) of B molecules padded to N nodes, masks from
// `sizes`; gcl_full_* of sample_AR, edge_denoise.py:293-294) with a deterministic per-column reduction, and an explicit
// (row, col) list with optional per-edge / per-node float masks (gcl_focal_* / gcl_edge / gcl_denoise on the search
// edges, edge_denoise.py:309-310, :347, :398) reduced with atomics.
//
//   gemm_k         Y = epilogue([X1 | X2] . W^T + bias + s * w_s): 64x64 tiles, fp32 FFMA, nn.Linear layout [out][in]
//   mes1_k         m1 = SiLU(A_row + B_col + radial * w_r + P)                             (gcl.py:98, first layer)
//   att_k          m = m2 * sigmoid(m2 . wa + ba) * edge_mask                              (:100-106)
//   node_reduce_k  dense: agg_j = sum_i m_ij (:121); x_j' = (x_j + sum_i coord_diff_ij * tanh(c1_ij . wc) * range
//                  * mask) * node_mask (:131-154, :193-194)
//   scatter_k / x_finish_k   the same for an explicit list (atomicAdd into zeroed accumulators)
#include "hd_common.cuh"

namespace hd {
int linear_tc_rows(cudaStream_t st, int rows, const float* X1, int ld1, int K1, const float* X2, int ld2, int K2,
                   const void* w_hi, const void* w_lo, int n_out, const float* bias, float* Y, int ldy, int mode,
                   const float* resid, const float* s, const float* ws, const float* ab, const int32_t* sizes, int N,
                   bool strict);
int make_image128(const float* src, int ld, int row0, int col0, int n_out, int K, int k_at, int k_total, void* hi, void* lo,
                  cudaStream_t st);
namespace egcl {

struct Offsets {   // float offsets into the flat state_dict-order parameter buffer (-1: absent)
  int64_t mes0_w, mes0_b, mes2_w, mes2_b, edge0_w, edge0_b, edge2_w, edge2_b, node0_w, node0_b, node2_w, node2_b,
      coord0_w, coord0_b, coord2_w, att_w, att_b, total;
};

static Offsets offsets(const hd_egcl_config& c) {
  Offsets o{};
  const int64_t Hh = c.hidden_nf, De = c.edges_in_d;
  int64_t s = 0;
  auto take = [&](int64_t n) { int64_t r = s; s += n; return r; };
  o.mes0_w = take(Hh * (2 * Hh + 1 + De)); o.mes0_b = take(Hh);
  o.mes2_w = take(Hh * Hh); o.mes2_b = take(Hh);
  if (c.edge_update) {
    o.edge0_w = take(Hh * (Hh + 1 + De)); o.edge0_b = take(Hh);
    o.edge2_w = take(Hh * Hh); o.edge2_b = take(Hh);
  } else {
    o.edge0_w = o.edge0_b = o.edge2_w = o.edge2_b = -1;
  }
  o.node0_w = take(Hh * 2 * Hh); o.node0_b = take(Hh);
  o.node2_w = take(Hh * Hh); o.node2_b = take(Hh);
  o.coord0_w = take(Hh * Hh); o.coord0_b = take(Hh);
  o.coord2_w = take(Hh);
  if (c.attention) { o.att_w = take(Hh); o.att_b = take(1); } else { o.att_w = o.att_b = -1; }
  o.total = s;
  return o;
}

// Packed image for the tensor-core path (dense list, hidden_nf = edges_in_d = 256): bf16 hi / lo operand images of every
// Linear in 128-row output tiles (hd_node.cu) + the fp32 vectors its epilogues read.  Byte offsets, 256-aligned.
struct Pack {
  int64_t ab_hi, ab_lo, ab_bias;   // [2H out][H k]: mes_mlp.0 columns [0, H) | [H, 2H); bias [b | 0]
  int64_t we_hi, we_lo, w_r;       // [H][De]: mes_mlp.0 columns [2H+1, ...); w_r = column 2H
  int64_t w2_hi, w2_lo;            // mes_mlp.2
  int64_t wc_hi, wc_lo;            // coord_mlp.0
  int64_t u1_hi, u1_lo, u_r;       // [H][H + De]: edge_mlp.0 columns [0, H) | [H+1, ...); u_r = column H
  int64_t u2_hi, u2_lo;            // edge_mlp.2
  int64_t v1_hi, v1_lo;            // node_mlp.0 [H][2H]
  int64_t v2_hi, v2_lo;            // node_mlp.2
  int64_t total;
};
static bool tc_shape(const hd_egcl_config& c) { return c.hidden_nf == 256 && c.edges_in_d == 256; }
static Pack pack_layout(const hd_egcl_config& c) {
  Pack k{};
  const int64_t Hh = c.hidden_nf, De = c.edges_in_d;
  int64_t p = 0;
  auto put = [&](int64_t bytes) { int64_t r = p; p = (p + bytes + 255) & ~int64_t(255); return r; };
  k.ab_hi = put(2 * Hh * Hh * 2); k.ab_lo = put(2 * Hh * Hh * 2); k.ab_bias = put(2 * Hh * 4);
  k.we_hi = put(Hh * De * 2); k.we_lo = put(Hh * De * 2); k.w_r = put(Hh * 4);
  k.w2_hi = put(Hh * Hh * 2); k.w2_lo = put(Hh * Hh * 2);
  k.wc_hi = put(Hh * Hh * 2); k.wc_lo = put(Hh * Hh * 2);
  k.u1_hi = put(Hh * (Hh + De) * 2); k.u1_lo = put(Hh * (Hh + De) * 2); k.u_r = put(Hh * 4);
  k.u2_hi = put(Hh * Hh * 2); k.u2_lo = put(Hh * Hh * 2);
  k.v1_hi = put(Hh * 2 * Hh * 2); k.v1_lo = put(Hh * 2 * Hh * 2);
  k.v2_hi = put(Hh * Hh * 2); k.v2_lo = put(Hh * Hh * 2);
  k.total = p;
  return k;
}

struct Work {   // byte offsets, 256-aligned
  int64_t ab, p, m1, m, c1, agg, hid, e1, radial, xacc, total;
};
static Work work(const hd_egcl_config& c, int64_t n_nodes, int64_t E) {
  Work w{};
  const int64_t Hh = c.hidden_nf;
  int64_t p = 0;
  auto put = [&](int64_t bytes) { int64_t r = p; p = (p + bytes + 255) & ~int64_t(255); return r; };
  w.ab = put(n_nodes * 2 * Hh * 4);
  w.p = put(E * Hh * 4);
  w.m1 = put(E * Hh * 4);
  w.m = put(E * Hh * 4);
  w.c1 = put(E * Hh * 4);
  w.agg = put(n_nodes * Hh * 4);
  w.hid = put(n_nodes * Hh * 4);
  w.e1 = put(c.edge_update ? E * Hh * 4 : 0);
  w.radial = put(E * 4);
  w.xacc = put(n_nodes * 3 * 4);
  w.total = p;
  return w;
}

// Where the edges come from.  Dense (row == null): edge e = (b, i, j), row = b*N + i, col = b*N + j, live iff
// i, j < n_b and i != j (the prefix node masks / off-diagonal edge masks of the sampler's batches).  Explicit list:
// row[e], col[e], multiplier emask[e] (null: 1).
struct EdgeSrc {
  const int32_t *row, *col;       // explicit list, 32-bit indices ...
  const float* emask;
  const int32_t* sizes;
  int N;
  const int64_t *row64, *col64;   // ... or 64-bit ones (torch's edge_index as it is)
};
__device__ __forceinline__ float edge_get(const EdgeSrc& s, int64_t e, int& row, int& col) {
  if (s.row64) {
    row = (int)s.row64[e];
    col = (int)s.col64[e];
    return s.emask ? s.emask[e] : 1.f;
  }
  if (s.row) {
    row = s.row[e];
    col = s.col[e];
    return s.emask ? s.emask[e] : 1.f;
  }
  const int j = (int)(e % s.N);
  const int64_t r = e / s.N;
  const int i = (int)(r % s.N), b = (int)(r / s.N);
  row = b * s.N + i;
  col = b * s.N + j;
  const int n = s.sizes[b];
  return (i < n && j < n && i != j) ? 1.f : 0.f;
}
// node mask of row m: dense = prefix mask from sizes; list = nmask[m] (null: 1)
struct NodeSrc {
  const float* nmask;
  const int32_t* sizes;
  int N;
};
__device__ __forceinline__ float node_get(const NodeSrc& s, int m) {
  if (s.sizes) return (m % s.N) < s.sizes[m / s.N] ? 1.f : 0.f;
  return s.nmask ? s.nmask[m] : 1.f;
}

// Y[M, Nout] = act([X1 | X2] . W^T + bias + s[m] * ws) (+ resid) (* rowmask), W = nn.Linear weight [Nout][ldw] whose columns
// [c1, c1+K1) multiply X1 and [c2, c2+K2) multiply X2, ws = column cs of W (rank-1 term, s may be null).
struct GemmArgs {
  const float *X1, *X2, *W, *bias, *s, *resid;
  float* Y;
  int M, Nout, K1, K2, ld1, ld2, ldw, c1, c2, cs, ldy, act;   // act: 0 none, 1 SiLU
  int mask_mode;          // 0 none, 1 node rows (NodeSrc), 2 edge rows (EdgeSrc, applied twice: gcl.py:113-115, :196-197)
  int mask_twice;
  EdgeSrc es;
  NodeSrc ns;
};

__global__ void __launch_bounds__(256) gemm_k(const GemmArgs a) {
  __shared__ float sx[16][64 + 1], sw[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 x 4 outputs each
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  float acc[4][4] = {};
  const int K = a.K1 + a.K2;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
      const int r = idx >> 4, k = (idx & 15) + k0;
      float xv = 0.f, wv = 0.f;
      if (k < K) {
        const int m = m0 + r, n = n0 + r;
        if (m < a.M) xv = k < a.K1 ? a.X1[(int64_t)m * a.ld1 + k] : a.X2[(int64_t)m * a.ld2 + (k - a.K1)];
        if (n < a.Nout) wv = a.W[(int64_t)n * a.ldw + (k < a.K1 ? a.c1 + k : a.c2 + (k - a.K1))];
      }
      sx[idx & 15][r] = xv;
      sw[idx & 15][r] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float xr[4], wr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xr[i] = sx[k][ty * 4 + i];
        wr[i] = sw[k][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    float mk = 1.f;
    if (a.mask_mode == 1) mk = node_get(a.ns, m);
    if (a.mask_mode == 2) {
      int row, col;
      mk = edge_get(a.es, m, row, col);
      if (a.mask_twice) mk *= mk;
    }
    const float sv = a.s ? a.s[m] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.Nout) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[n];
      if (a.s) v = fmaf(sv, a.W[(int64_t)n * a.ldw + a.cs], v);
      if (a.act == 1) v = silu_acc(v);
      if (a.resid) v += a.resid[(int64_t)m * a.ldy + n];
      a.Y[(int64_t)m * a.ldy + n] = a.mask_mode ? v * mk : v;
    }
  }
}

// The same product for FEW rows (the explicit edge lists and the prediction heads of sample_AR: tens to hundreds of
// rows, where the 64 x 64 tiles above leave the GPU to 4-8 CTAs that each walk all of K): a CTA owns 32 rows x 8 output
// columns, a warp one column; the lanes split K (coalesced weight reads, the 32 x 128 activation chunk in shared memory)
// and a shuffle transpose-reduction leaves lane m with the sum of row m.
constexpr int GS_ROWS = 32, GS_COLS = 8, GS_KC = 128;
__global__ void __launch_bounds__(32 * GS_COLS) gemm_small_k(const GemmArgs a) {
  __shared__ float sx[GS_ROWS][GS_KC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * GS_ROWS, n = blockIdx.y * GS_COLS + warp;
  const bool col_ok = n < a.Nout;
  const int K = a.K1 + a.K2;
  float acc[GS_ROWS];
#pragma unroll
  for (int r = 0; r < GS_ROWS; ++r) acc[r] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GS_KC) {
    for (int idx = threadIdx.x; idx < GS_ROWS * GS_KC; idx += 32 * GS_COLS) {
      const int r = idx / GS_KC, k = k0 + idx % GS_KC, m = m0 + r;
      float v = 0.f;
      if (m < a.M && k < K) v = k < a.K1 ? a.X1[(int64_t)m * a.ld1 + k] : a.X2[(int64_t)m * a.ld2 + (k - a.K1)];
      sx[r][idx % GS_KC] = v;
    }
    __syncthreads();
    float w[GS_KC / 32];
#pragma unroll
    for (int j = 0; j < GS_KC / 32; ++j) {
      const int k = k0 + lane + 32 * j;
      w[j] = (col_ok && k < K) ? a.W[(int64_t)n * a.ldw + (k < a.K1 ? a.c1 + k : a.c2 + (k - a.K1))] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < GS_ROWS; ++r)
#pragma unroll
      for (int j = 0; j < GS_KC / 32; ++j) acc[r] = fmaf(sx[r][lane + 32 * j], w[j], acc[r]);
    __syncthreads();
  }
  // transpose-reduction: after the five halving steps lane m holds sum over lanes of acc[m]
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int r = 0; r < step; ++r) {
      const float send = up ? acc[r] : acc[r + step], keep = up ? acc[r + step] : acc[r];
      acc[r] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  const int m = m0 + lane;   // lane's row: bit k of the lane index selected the upper half at the step of size 2^k
  if (!col_ok || m >= a.M) return;
  float v = acc[0];
  if (a.bias) v += a.bias[n];
  if (a.s) v = fmaf(a.s[m], a.W[(int64_t)n * a.ldw + a.cs], v);
  if (a.act == 1) v = silu_acc(v);
  if (a.resid) v += a.resid[(int64_t)m * a.ldy + n];
  if (a.mask_mode == 1) v *= node_get(a.ns, m);
  if (a.mask_mode == 2) {
    int row, col;
    float mk = edge_get(a.es, m, row, col);
    if (a.mask_twice) mk *= mk;
    v *= mk;
  }
  a.Y[(int64_t)m * a.ldy + n] = v;
}

// per edge: radial, m1 = SiLU(A_row + B_col + radial * w_r + (P or sum_d e_d * w_e,d))     one CTA of 128 threads per edge
__global__ void mes1_k(const float* __restrict__ ab, const float* __restrict__ pe, const float* __restrict__ edge_attr,
                       int De, const float* __restrict__ x, const float* __restrict__ w0, int ldw, int col_r,
                       const EdgeSrc es, int Hh, float* __restrict__ radial, float* __restrict__ m1) {
  const int64_t e = blockIdx.x;
  int row, col;
  edge_get(es, e, row, col);
  const float d0 = x[row * 3] - x[col * 3], d1 = x[row * 3 + 1] - x[col * 3 + 1], d2 = x[row * 3 + 2] - x[col * 3 + 2];
  const float r = d0 * d0 + d1 * d1 + d2 * d2;
  if (threadIdx.x == 0) radial[e] = r;
  for (int k = threadIdx.x; k < Hh; k += blockDim.x) {
    float v = ab[(int64_t)row * 2 * Hh + k] + ab[(int64_t)col * 2 * Hh + Hh + k];
    v = fmaf(r, w0[(int64_t)k * ldw + col_r], v);
    if (pe) {
      v += pe[e * Hh + k];
    } else if (edge_attr) {
      for (int d = 0; d < De; ++d) v = fmaf(edge_attr[e * De + d], w0[(int64_t)k * ldw + col_r + 1 + d], v);
    } else {
      v = fmaf(r, w0[(int64_t)k * ldw + col_r + 1], v);   // edge_attr == NULL, De == 1: the edge feature IS the squared distance
    }
    m1[e * Hh + k] = silu_acc(v);
  }
}

// per edge (one warp): m = m2 * sigmoid(m2 . wa + ba) * edge_mask   (attention == 0: only the mask)
__global__ void att_k(float* __restrict__ m, const float* __restrict__ wa, const float* __restrict__ ba, int attention,
                      const EdgeSrc es, int Hh, int64_t E) {
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  int row, col;
  const float mk = edge_get(es, e, row, col);
  float att = 1.f;
  if (attention) {
    float dot = 0.f;
    for (int k = lane; k < Hh; k += 32) dot = fmaf(m[e * Hh + k], wa[k], dot);
    dot = warp_sum(dot);
    att = sigmoid_acc(dot + ba[0]);
  }
  const float sc = att * mk;
  if (!attention && mk == 1.f) return;
  for (int k = lane; k < Hh; k += 32) m[e * Hh + k] *= sc;
}

__device__ __forceinline__ float trans_scale(float phi, int use_tanh, float range) {
  return use_tanh ? tanhf(phi) * range : phi;
}

// dense list, per node j (one CTA): agg_j = sum_i m_ij ; x_j' = (x_j + sum_i trans_ij) * node_mask
__global__ void node_reduce_k(const float* __restrict__ m, const float* __restrict__ c1, const float* __restrict__ wc,
                              const float* __restrict__ x, const int32_t* __restrict__ sizes, int N, int Hh, int use_tanh,
                              float range, float* __restrict__ agg, float* __restrict__ x_out) {
  extern __shared__ float s_phi[];   // [N]
  const int node = blockIdx.x, b = node / N, j = node % N, n = sizes[b], tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  // phi_ij = c1_ij . wc for every sender i of this column (a warp per edge)
  for (int i = warp; i < N; i += nwarp) {
    const int64_t e = ((int64_t)b * N + i) * N + j;
    float dot = 0.f;
    for (int k = lane; k < Hh; k += 32) dot = fmaf(c1[e * Hh + k], wc[k], dot);
    dot = warp_sum(dot);
    if (lane == 0) s_phi[i] = dot;
  }
  __syncthreads();
  for (int k = tid; k < Hh; k += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < N; ++i) s += m[(((int64_t)b * N + i) * N + j) * Hh + k];   // masked edges contribute 0
    agg[(int64_t)node * Hh + k] = s;
  }
  if (tid < 3) {
    float s = 0.f;
    const float xj = x[node * 3 + tid];
    for (int i = 0; i < N; ++i) {
      if (!(i < n && j < n && i != j)) continue;
      const int row = b * N + i;
      const float d0 = x[row * 3] - x[node * 3], d1 = x[row * 3 + 1] - x[node * 3 + 1], d2 = x[row * 3 + 2] - x[node * 3 + 2];
      const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2 + 1e-8f) + 1.0f;                 // gcl.py:205-207
      const float cd = (tid == 0 ? d0 : (tid == 1 ? d1 : d2)) / nrm;
      s += cd * trans_scale(s_phi[i], use_tanh, range);
    }
    x_out[node * 3 + tid] = (xj + s) * (j < n ? 1.f : 0.f);
  }
}

// explicit list, per edge (one warp): agg[col] += m_e ; xacc[col] += coord_diff_e * tanh(c1_e . wc) * range * mask
__global__ void scatter_k(const float* __restrict__ m, const float* __restrict__ c1, const float* __restrict__ wc,
                          const float* __restrict__ x, const EdgeSrc es, int Hh, int use_tanh, float range, int64_t E,
                          float* __restrict__ agg, float* __restrict__ xacc) {
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  int row, col;
  const float mk = edge_get(es, e, row, col);
  float dot = 0.f;
  for (int k = lane; k < Hh; k += 32) {
    dot = fmaf(c1[e * Hh + k], wc[k], dot);
    atomicAdd(&agg[(int64_t)col * Hh + k], m[e * Hh + k]);
  }
  dot = warp_sum(dot);
  if (lane < 3) {
    const float d0 = x[row * 3] - x[col * 3], d1 = x[row * 3 + 1] - x[col * 3 + 1], d2 = x[row * 3 + 2] - x[col * 3 + 2];
    const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2 + 1e-8f) + 1.0f;
    const float cd = (lane == 0 ? d0 : (lane == 1 ? d1 : d2)) / nrm;
    atomicAdd(&xacc[col * 3 + lane], cd * trans_scale(dot, use_tanh, range) * mk);
  }
}
__global__ void x_finish_k(const float* __restrict__ x, const float* __restrict__ xacc, const NodeSrc ns, int count,
                           float* __restrict__ x_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < count) x_out[idx] = (x[idx] + xacc[idx]) * node_get(ns, idx / 3);
}

// dst[o] = src[o * ld + col] (o < n); col < 0: dst[o] = o < n_src ? src[o] : 0   (bias image [b | 0])
__global__ void vec_k(const float* __restrict__ src, int ld, int col, int n, int n_src, float* __restrict__ dst) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  dst[o] = col >= 0 ? src[(int64_t)o * ld + col] : (o < n_src ? src[o] : 0.f);
}
__global__ void radial_dense_k(const float* __restrict__ x, int N, int64_t E, float* __restrict__ radial) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t row = e / N, col = (e / ((int64_t)N * N)) * N + e % N;
  const float d0 = x[row * 3] - x[col * 3], d1 = x[row * 3 + 1] - x[col * 3 + 1], d2 = x[row * 3 + 2] - x[col * 3 + 2];
  radial[e] = d0 * d0 + d1 * d1 + d2 * d2;
}

static int gemm(cudaStream_t st, const GemmArgs& a) {
  if (a.M < 1) return HD_OK;
  if (a.M <= 2048) {   // few rows: many small CTAs instead of a handful of 64 x 64 tiles
    dim3 grid((a.M + GS_ROWS - 1) / GS_ROWS, (a.Nout + GS_COLS - 1) / GS_COLS);
    gemm_small_k<<<grid, 32 * GS_COLS, 0, st>>>(a);
    HD_CHECK_LAUNCH();
    return HD_OK;
  }
  dim3 grid((a.M + 63) / 64, (a.Nout + 63) / 64);
  gemm_k<<<grid, 256, 0, st>>>(a);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

}  // namespace egcl
}  // namespace hd

using namespace hd;

extern "C" {

HD_API int64_t hd_egcl_weight_count(const hd_egcl_config* cfg) {
  if (!cfg || cfg->hidden_nf < 1 || cfg->edges_in_d < 1) return HD_E_INVALID;
  return egcl::offsets(*cfg).total;
}

HD_API int64_t hd_egcl_workspace_bytes(const hd_egcl_config* cfg, int64_t n_nodes, int64_t n_edges) {
  if (!cfg || n_nodes < 1 || n_edges < 0) return HD_E_INVALID;
  return egcl::work(*cfg, n_nodes, n_edges).total;
}

HD_API int32_t hd_linear_forward(const float* x, int64_t rows, int32_t in_nf, const float* weight, const float* bias,
                                 int32_t out_nf, int32_t act, float* y, hd_stream_t stream) {
  if (!x || !weight || !y || rows < 0 || in_nf < 1 || out_nf < 1 || act < 0 || act > 1 || rows > 0x7fffffff / 4) {
    set_error("hd_linear_forward: bad argument");
    return HD_E_INVALID;
  }
  egcl::GemmArgs g{};
  g.X1 = x; g.X2 = x; g.K1 = in_nf; g.ld1 = in_nf; g.ld2 = in_nf; g.W = weight; g.ldw = in_nf; g.bias = bias;
  g.M = (int)rows; g.Nout = out_nf; g.ldy = out_nf; g.Y = y; g.act = act;
  return egcl::gemm(static_cast<cudaStream_t>(stream), g);
}

HD_API int64_t hd_egcl_packed_bytes(const hd_egcl_config* cfg) {
  if (!cfg) return HD_E_INVALID;
  return egcl::tc_shape(*cfg) ? egcl::pack_layout(*cfg).total : 0;
}

HD_API int32_t hd_egcl_pack_weights(const hd_egcl_config* cfg, const float* w, void* packed, hd_stream_t stream) {
  if (!cfg || !w || !packed || !egcl::tc_shape(*cfg)) {
    set_error("hd_egcl_pack_weights: the tensor-core path needs hidden_nf = edges_in_d = 256");
    return HD_E_INVALID;
  }
  const egcl::Offsets o = egcl::offsets(*cfg);
  const egcl::Pack K = egcl::pack_layout(*cfg);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* P = static_cast<char*>(packed);
  auto PF = [&](int64_t off) { return reinterpret_cast<float*>(P + off); };
  const int Hh = cfg->hidden_nf, De = cfg->edges_in_d, ld_mes = 2 * Hh + 1 + De, ld_edge = Hh + 1 + De;
  int rc;
  // A | B: outputs [0, H) from mes_mlp.0 columns [0, H), outputs [H, 2H) from columns [H, 2H)
  for (int part = 0; part < 2; ++part)
    for (int t = 0; t < Hh / 128; ++t) {
      const int64_t off = (int64_t)(part * (Hh / 128) + t) * 128 * Hh * 2;
      if ((rc = make_image128(w + o.mes0_w, ld_mes, t * 128, part * Hh, 128, Hh, 0, Hh, P + K.ab_hi + off, P + K.ab_lo + off, st)))
        return rc;
    }
  egcl::vec_k<<<(2 * Hh + 255) / 256, 256, 0, st>>>(w + o.mes0_b, 0, -1, 2 * Hh, Hh, PF(K.ab_bias));
  if ((rc = make_image128(w + o.mes0_w, ld_mes, 0, 2 * Hh + 1, Hh, De, 0, De, P + K.we_hi, P + K.we_lo, st))) return rc;
  egcl::vec_k<<<(Hh + 255) / 256, 256, 0, st>>>(w + o.mes0_w, ld_mes, 2 * Hh, Hh, Hh, PF(K.w_r));
  if ((rc = make_image128(w + o.mes2_w, Hh, 0, 0, Hh, Hh, 0, Hh, P + K.w2_hi, P + K.w2_lo, st))) return rc;
  if ((rc = make_image128(w + o.coord0_w, Hh, 0, 0, Hh, Hh, 0, Hh, P + K.wc_hi, P + K.wc_lo, st))) return rc;
  if (cfg->edge_update) {
    if ((rc = make_image128(w + o.edge0_w, ld_edge, 0, 0, Hh, Hh, 0, Hh + De, P + K.u1_hi, P + K.u1_lo, st))) return rc;
    if ((rc = make_image128(w + o.edge0_w, ld_edge, 0, Hh + 1, Hh, De, Hh, Hh + De, P + K.u1_hi, P + K.u1_lo, st))) return rc;
    egcl::vec_k<<<(Hh + 255) / 256, 256, 0, st>>>(w + o.edge0_w, ld_edge, Hh, Hh, Hh, PF(K.u_r));
    if ((rc = make_image128(w + o.edge2_w, Hh, 0, 0, Hh, Hh, 0, Hh, P + K.u2_hi, P + K.u2_lo, st))) return rc;
  }
  if ((rc = make_image128(w + o.node0_w, 2 * Hh, 0, 0, Hh, 2 * Hh, 0, 2 * Hh, P + K.v1_hi, P + K.v1_lo, st))) return rc;
  if ((rc = make_image128(w + o.node2_w, Hh, 0, 0, Hh, Hh, 0, Hh, P + K.v2_hi, P + K.v2_lo, st))) return rc;
  HD_CHECK_LAUNCH();
  return HD_OK;
}

// the dense list on the tensor cores: every Linear through lin::linear_tc_k (tcgen05, bf16x3 split operands in strict
// mode), the per-edge glue folded into its epilogues (hd_node.cu modes 4-6)
static int egcl_dense_tc(const hd_egcl_config* cfg, const float* w, const char* P, const float* h, const float* x,
                         const float* edge_attr, const int32_t* sizes, int B, int N, float* h_out, float* x_out,
                         float* edge_out, char* ws, cudaStream_t st, bool strict) {
  const egcl::Offsets o = egcl::offsets(*cfg);
  const egcl::Pack K = egcl::pack_layout(*cfg);
  const int Hh = cfg->hidden_nf, De = cfg->edges_in_d, BN = B * N, E = BN * N;
  const egcl::Work W = egcl::work(*cfg, BN, E);
  auto WF = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  auto PF = [&](int64_t off) { return reinterpret_cast<const float*>(P + off); };
  float* radial = WF(W.radial);
  int rc;
  egcl::radial_dense_k<<<(E + 255) / 256, 256, 0, st>>>(x, N, E, radial);
  HD_CHECK_LAUNCH();
  // A|B = h . [W1a | W1b]^T + [b1 | 0]
  if ((rc = linear_tc_rows(st, BN, h, Hh, Hh, nullptr, 0, 0, P + K.ab_hi, P + K.ab_lo, 2 * Hh, PF(K.ab_bias), WF(W.ab), 2 * Hh, 0,
                           nullptr, nullptr, nullptr, nullptr, sizes, N, strict)))
    return rc;
  // m1 = SiLU(e . We^T + radial * w_r + A_row + B_col)
  if ((rc = linear_tc_rows(st, E, edge_attr, De, De, nullptr, 0, 0, P + K.we_hi, P + K.we_lo, Hh, nullptr, WF(W.m1), Hh, 6,
                           nullptr, radial, PF(K.w_r), WF(W.ab), sizes, N, strict)))
    return rc;
  // m2 = SiLU(m1 . W2^T + b2); attention gate and edge mask
  if ((rc = linear_tc_rows(st, E, WF(W.m1), Hh, Hh, nullptr, 0, 0, P + K.w2_hi, P + K.w2_lo, Hh, w + o.mes2_b, WF(W.m), Hh, 1,
                           nullptr, nullptr, nullptr, nullptr, sizes, N, strict)))
    return rc;
  const egcl::EdgeSrc es{nullptr, nullptr, nullptr, sizes, N, nullptr, nullptr};
  egcl::att_k<<<(E + 7) / 8, 256, 0, st>>>(WF(W.m), cfg->attention ? w + o.att_w : nullptr,
                                       cfg->attention ? w + o.att_b : nullptr, cfg->attention, es, Hh, E);
  HD_CHECK_LAUNCH();
  // c1 = SiLU(m . Wc0^T + bc0); coordinate update and aggregation over col
  if ((rc = linear_tc_rows(st, E, WF(W.m), Hh, Hh, nullptr, 0, 0, P + K.wc_hi, P + K.wc_lo, Hh, w + o.coord0_b, WF(W.c1), Hh, 1,
                           nullptr, nullptr, nullptr, nullptr, sizes, N, strict)))
    return rc;
  egcl::node_reduce_k<<<BN, 256, N * sizeof(float), st>>>(WF(W.m), WF(W.c1), w + o.coord2_w, x, sizes, N, Hh, cfg->tanh,
                                                        cfg->coords_range, WF(W.agg), x_out);
  HD_CHECK_LAUNCH();
  if (cfg->edge_update) {
    // e1 = SiLU([m | e] . U1^T + radial * u_r + d1); e' = (e1 . U2^T + d2) * edge_mask
    if ((rc = linear_tc_rows(st, E, WF(W.m), Hh, Hh, edge_attr, De, De, P + K.u1_hi, P + K.u1_lo, Hh, w + o.edge0_b, WF(W.e1),
                             Hh, 4, nullptr, radial, PF(K.u_r), nullptr, sizes, N, strict)))
      return rc;
    if ((rc = linear_tc_rows(st, E, WF(W.e1), Hh, Hh, nullptr, 0, 0, P + K.u2_hi, P + K.u2_lo, Hh, w + o.edge2_b, edge_out, Hh, 5,
                             nullptr, nullptr, nullptr, nullptr, sizes, N, strict)))
      return rc;
  }
  // hid = SiLU([h | agg] . V1^T + c1); h' = (h + hid . V2^T + c2) * node_mask
  if ((rc = linear_tc_rows(st, BN, h, Hh, Hh, WF(W.agg), Hh, Hh, P + K.v1_hi, P + K.v1_lo, Hh, w + o.node0_b, WF(W.hid), Hh, 1,
                           nullptr, nullptr, nullptr, nullptr, sizes, N, strict)))
    return rc;
  return linear_tc_rows(st, BN, WF(W.hid), Hh, Hh, nullptr, 0, 0, P + K.v2_hi, P + K.v2_lo, Hh, w + o.node2_b, h_out, Hh, 2, h,
                        nullptr, nullptr, nullptr, sizes, N, strict);
}

HD_API int32_t hd_egcl_forward(const hd_egcl_config* cfg, const float* w, const void* packed, const float* h,
                               const float* x, const float* edge_attr, const void* row, const void* col,
                               int32_t index_bits, const float* edge_mask, const float* node_mask, const int32_t* sizes,
                               int32_t B, int32_t N, int64_t n_nodes, int64_t n_edges, float* h_out, float* x_out,
                               float* edge_out, void* workspace, int32_t engine, hd_stream_t stream) {
  if (!cfg || !w || !h || !x || !h_out || !x_out || !workspace) {
    set_error("null argument");
    return HD_E_INVALID;
  }
  const bool dense = row == nullptr && (col != nullptr || sizes != nullptr || n_edges != 0);   // an EMPTY list has no row either
  if (dense) {
    if (col || !sizes || B < 1 || N < 1 || N > 1024) {
      set_error("dense edge list needs sizes, B >= 1, 1 <= N <= 1024 and no col");
      return HD_E_INVALID;
    }
    n_nodes = (int64_t)B * N;
    n_edges = n_nodes * N;
  } else if ((n_edges > 0 && (!row || !col)) || n_nodes < 1 || n_edges < 0) {
    set_error("explicit edge list needs row, col, n_nodes >= 1, n_edges >= 0");
    return HD_E_INVALID;
  }
  if (index_bits != 32 && index_bits != 64) {
    set_error("index_bits must be 32 or 64");
    return HD_E_INVALID;
  }
  // edge_attr == NULL with one edge feature and no edge update: the feature is |x_row - x_col|^2, computed here
  // (what edge_denoise.py:345-347 / :396-398 pass to gcl_edge / gcl_denoise)
  const bool radial_attr = !edge_attr && cfg->edges_in_d == 1 && !cfg->edge_update;
  if (n_edges > 0 && ((!edge_attr && !radial_attr) || (cfg->edge_update && !edge_out))) {
    set_error("null edge_attr / edge_out");
    return HD_E_INVALID;
  }
  if (cfg->hidden_nf < 32 || cfg->hidden_nf > 1024 || cfg->edges_in_d < 1) {
    set_error("unsupported shape hidden_nf=%d edges_in_d=%d", cfg->hidden_nf, cfg->edges_in_d);
    return HD_E_INVALID;
  }
  if (n_edges > 0x7fffffff / 4 || n_nodes > 0x7fffffff / 4) {
    set_error("too many edges (%lld) or nodes (%lld)", (long long)n_edges, (long long)n_nodes);
    return HD_E_INVALID;
  }
  const egcl::Offsets o = egcl::offsets(*cfg);
  const egcl::Work W = egcl::work(*cfg, n_nodes, n_edges);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  if (engine != HD_ENGINE_FP32) {
    if (engine != HD_ENGINE_TC_STRICT && engine != HD_ENGINE_TC_FAST) {
      set_error("unknown engine %d", engine);
      return HD_E_INVALID;
    }
    if (!dense || !packed || !egcl::tc_shape(*cfg)) {
      set_error("the tensor-core engines run the dense list with hidden_nf = edges_in_d = 256 and a packed weight image");
      return HD_E_UNSUPPORTED;
    }
    return egcl_dense_tc(cfg, w, static_cast<const char*>(packed), h, x, edge_attr, sizes, B, N, h_out, x_out, edge_out, ws,
                         st, engine == HD_ENGINE_TC_STRICT);
  }
  auto WF = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  const int Hh = cfg->hidden_nf, De = cfg->edges_in_d;
  const int BN = (int)n_nodes, E = (int)n_edges;
  const bool wide = !dense && index_bits == 64;
  const egcl::EdgeSrc es{wide ? nullptr : static_cast<const int32_t*>(row), wide ? nullptr : static_cast<const int32_t*>(col),
                         edge_mask, dense ? sizes : nullptr, N, wide ? static_cast<const int64_t*>(row) : nullptr,
                         wide ? static_cast<const int64_t*>(col) : nullptr};
  const egcl::NodeSrc ns{node_mask, dense ? sizes : nullptr, N};
  const bool has_emask = dense || edge_mask != nullptr;
  const bool has_nmask = dense || node_mask != nullptr;
  float* radial = WF(W.radial);
  const int ld_mes = 2 * Hh + 1 + De, ld_edge = Hh + 1 + De;
  int rc;
  // 1. A|B = h . [W1[:, :H] | W1[:, H:2H]]^T (+ b1 on the A half): two GEMMs on the node rows
  egcl::GemmArgs g{};
  g.X1 = h; g.X2 = h; g.K1 = Hh; g.K2 = 0; g.ld1 = Hh; g.ld2 = Hh; g.W = w + o.mes0_w; g.ldw = ld_mes; g.M = BN; g.Nout = Hh;
  g.ldy = 2 * Hh;
  g.c1 = 0; g.bias = w + o.mes0_b; g.Y = WF(W.ab);
  if ((rc = egcl::gemm(st, g))) return rc;
  g.c1 = Hh; g.bias = nullptr; g.Y = WF(W.ab) + Hh;
  if ((rc = egcl::gemm(st, g))) return rc;
  if (E > 0) {
    // 2. P = edge_attr . W1[:, 2H+1:]^T on the edge rows (narrow edge features are folded into mes1_k instead)
    const bool use_p = De > 8;
    if (use_p) {
      g = egcl::GemmArgs{};
      g.X1 = edge_attr; g.X2 = edge_attr; g.K1 = De; g.ld1 = De; g.ld2 = De; g.W = w + o.mes0_w; g.ldw = ld_mes; g.c1 = 2 * Hh + 1;
      g.M = E; g.Nout = Hh; g.ldy = Hh; g.Y = WF(W.p);
      if ((rc = egcl::gemm(st, g))) return rc;
    }
    // 3. m1 = SiLU(A_row + B_col + radial * w_r + P)
    egcl::mes1_k<<<E, 128, 0, st>>>(WF(W.ab), use_p ? WF(W.p) : nullptr, edge_attr, De, x, w + o.mes0_w, ld_mes, 2 * Hh, es,
                                    Hh, radial, WF(W.m1));
    HD_CHECK_LAUNCH();
    // 4. m2 = SiLU(m1 . W2^T + b2)
    g = egcl::GemmArgs{};
    g.X1 = WF(W.m1); g.X2 = g.X1; g.K1 = Hh; g.ld1 = Hh; g.ld2 = Hh; g.W = w + o.mes2_w; g.ldw = Hh; g.bias = w + o.mes2_b;
    g.M = E; g.Nout = Hh; g.ldy = Hh; g.Y = WF(W.m); g.act = 1;
    if ((rc = egcl::gemm(st, g))) return rc;
    // 5. attention gate and edge mask
    if (cfg->attention || has_emask) {
      egcl::att_k<<<(E + 7) / 8, 256, 0, st>>>(WF(W.m), cfg->attention ? w + o.att_w : nullptr,
                                           cfg->attention ? w + o.att_b : nullptr, cfg->attention, es, Hh, E);
      HD_CHECK_LAUNCH();
    }
    // 6. c1 = SiLU(m . Wc0^T + bc0)
    g.X1 = WF(W.m); g.X2 = g.X1; g.W = w + o.coord0_w; g.bias = w + o.coord0_b; g.Y = WF(W.c1);
    if ((rc = egcl::gemm(st, g))) return rc;
  }
  // 7. + 8. coordinate update and message aggregation over `col`
  if (dense) {
    egcl::node_reduce_k<<<BN, 256, N * sizeof(float), st>>>(WF(W.m), WF(W.c1), w + o.coord2_w, x, sizes, N, Hh, cfg->tanh,
                                                          cfg->coords_range, WF(W.agg), x_out);
    HD_CHECK_LAUNCH();
  } else {
    HD_CHECK_CUDA(cudaMemsetAsync(WF(W.agg), 0, (size_t)BN * Hh * 4, st));
    HD_CHECK_CUDA(cudaMemsetAsync(WF(W.xacc), 0, (size_t)BN * 3 * 4, st));
    if (E > 0) {
      egcl::scatter_k<<<(E + 7) / 8, 256, 0, st>>>(WF(W.m), WF(W.c1), w + o.coord2_w, x, es, Hh, cfg->tanh, cfg->coords_range,
                                               E, WF(W.agg), WF(W.xacc));
      HD_CHECK_LAUNCH();
    }
    egcl::x_finish_k<<<(BN * 3 + 255) / 256, 256, 0, st>>>(x, WF(W.xacc), ns, BN * 3, x_out);
    HD_CHECK_LAUNCH();
  }
  // 11. + 12. edge update: e' = (SiLU([m | radial | e] . U1^T + d1) . U2^T + d2) * edge_mask * edge_mask
  if (cfg->edge_update && E > 0) {
    g = egcl::GemmArgs{};
    g.X1 = WF(W.m); g.K1 = Hh; g.ld1 = Hh; g.c1 = 0;
    g.X2 = edge_attr; g.K2 = De; g.ld2 = De; g.c2 = Hh + 1;
    g.s = radial; g.cs = Hh;
    g.W = w + o.edge0_w; g.ldw = ld_edge; g.bias = w + o.edge0_b; g.M = E; g.Nout = Hh; g.ldy = Hh; g.Y = WF(W.e1); g.act = 1;
    if ((rc = egcl::gemm(st, g))) return rc;
    g = egcl::GemmArgs{};
    g.X1 = WF(W.e1); g.X2 = g.X1; g.K1 = Hh; g.ld1 = Hh; g.ld2 = Hh; g.W = w + o.edge2_w; g.ldw = Hh; g.bias = w + o.edge2_b;
    g.M = E; g.Nout = Hh; g.ldy = Hh; g.Y = edge_out; g.mask_mode = has_emask ? 2 : 0; g.mask_twice = 1; g.es = es;
    if ((rc = egcl::gemm(st, g))) return rc;
  }
  // 9. hid = SiLU([h | agg] . V1^T + c1)
  g = egcl::GemmArgs{};
  g.X1 = h; g.K1 = Hh; g.ld1 = Hh; g.c1 = 0; g.X2 = WF(W.agg); g.K2 = Hh; g.ld2 = Hh; g.c2 = Hh;
  g.W = w + o.node0_w; g.ldw = 2 * Hh; g.bias = w + o.node0_b; g.M = BN; g.Nout = Hh; g.ldy = Hh; g.Y = WF(W.hid); g.act = 1;
  if ((rc = egcl::gemm(st, g))) return rc;
  // 10. h' = (h + hid . V2^T + c2) * node_mask
  g = egcl::GemmArgs{};
  g.X1 = WF(W.hid); g.X2 = g.X1; g.K1 = Hh; g.ld1 = Hh; g.ld2 = Hh; g.W = w + o.node2_w; g.ldw = Hh; g.bias = w + o.node2_b;
  g.M = BN; g.Nout = Hh; g.ldy = Hh; g.Y = h_out; g.resid = h; g.mask_mode = has_nmask ? 1 : 0; g.ns = ns;
  if ((rc = egcl::gemm(st, g))) return rc;
  return HD_OK;
}

}  // extern "C"
