// hd_layout.cu - packed-weight and workspace layouts, error string, weight packing kernels.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include <cuda_bf16.h>

#include "hd_common.cuh"

namespace hd {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("HD_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}

static int64_t align256(int64_t v) { return (v + 255) & ~int64_t(255); }

bool make_layout(const hd_config& c, Layout* out) {
  if (c.hidden_nf != H) {
    set_error("hidden_nf=%d unsupported: this build is specialised for hidden_nf=%d", c.hidden_nf, H);
    return false;
  }
  if (c.n_layers < 1 || c.n_layers > 64 || c.inv_sublayers < 1 || c.inv_sublayers > 8 || c.in_node_nf < 1 ||
      c.in_node_nf > 64) {
    set_error("bad config: n_layers=%d inv_sublayers=%d in_node_nf=%d", c.n_layers, c.inv_sublayers, c.in_node_nf);
    return false;
  }
  if (!(c.normalization_factor > 0.f)) {
    set_error("normalization_factor must be > 0");
    return false;
  }
  Layout& L = *out;
  L.subs.clear();
  const int64_t Fi = c.in_node_nf;
  int64_t s = 0;  // flat cursor (floats)
  auto take = [&](int64_t n) {
    int64_t r = s;
    s += n;
    return r;
  };
  L.s_emb_w = take(H * Fi);
  L.s_emb_b = take(H);
  L.s_out_w = take(Fi * H);
  L.s_out_b = take(Fi);
  int64_t p = 0;  // packed cursor (bytes)
  auto put = [&](int64_t bytes) {
    int64_t r = p;
    p = align256(p + bytes);
    return r;
  };
  L.fuse_tmp = put((int64_t)2 * H * 2 * H * 4);
  L.emb_wT = put(Fi * H * 4);
  L.emb_b = put(H * 4);
  L.emb_abT = put(Fi * 2 * H * 4);
  L.emb_abb = put(2 * H * 4);
  L.out_w = put(Fi * H * 4);
  L.out_b = put(Fi * 4);
  for (int b = 0; b < c.n_layers; ++b) {
    for (int k = 0; k <= c.inv_sublayers; ++k) {
      SubLayer S{};
      S.is_gcl = k < c.inv_sublayers;
      S.s_w1 = take((int64_t)H * (2 * H + 2));
      S.s_b1 = take(H);
      S.s_w2 = take((int64_t)H * H);
      S.s_b2 = take(H);
      S.s_v1 = S.s_c1 = S.s_v2 = S.s_c2 = S.s_ba = -1;
      if (S.is_gcl) {
        S.s_v1 = take((int64_t)H * 2 * H);
        S.s_c1 = take(H);
        S.s_v2 = take((int64_t)H * H);
        S.s_c2 = take(H);
        if (c.attention) {
          S.s_wa = take(H);
          S.s_ba = take(1);
        } else {
          S.s_wa = -1;
        }
      } else {
        S.s_wa = take(H);
      }
      S.w1abT = put((int64_t)H * 2 * H * 4);
      S.b1 = put(2 * H * 4);  // [b1 | 0]
      S.wr = put(H * 4);
      S.wd = put(H * 4);
      S.w2T = put((int64_t)H * H * 4);
      S.b2 = put(H * 4);
      S.wa = put(H * 4);
      S.ba = put(16);
      S.b1s = put(2 * H * 4);
      S.wrs = put(H * 4);
      S.wds = put(H * 4);
      S.b2s = put(H * 4);
      if (S.is_gcl) {
        S.v1T = put((int64_t)2 * H * H * 4);
        S.c1 = put(H * 4);
        S.v2T = put((int64_t)H * H * 4);
        S.c2 = put(H * 4);
      } else {
        S.v1T = S.c1 = S.v2T = S.c2 = -1;
      }
      S.w2_hi = put((int64_t)H * H * 2);
      S.w2_lo = put((int64_t)H * H * 2);
      S.w1ab_hi = put((int64_t)2 * H * H * 2);
      S.w1ab_lo = put((int64_t)2 * H * H * 2);
      if (S.is_gcl) {
        S.v1_hi = put((int64_t)2 * H * H * 2);
        S.v1_lo = put((int64_t)2 * H * H * 2);
        S.v2_hi = put((int64_t)H * H * 2);
        S.v2_lo = put((int64_t)H * H * 2);
        // every GCL is followed by another sub-layer of its block (a GCL or the EquivariantUpdate)
        S.fuse_next = true;
        S.v2w_hi = put((int64_t)H * H * 2);
        S.v2w_lo = put((int64_t)H * H * 2);
        S.m_hi = put((int64_t)2 * H * 2 * H * 2);
        S.m_lo = put((int64_t)2 * H * 2 * H * 2);
        S.bm = put(2 * H * 4);
        S.fuse_next2 = k == c.inv_sublayers - 1 && b + 1 < c.n_layers;
        if (S.fuse_next2) {
          S.m2_hi = put((int64_t)2 * H * 2 * H * 2);
          S.m2_lo = put((int64_t)2 * H * 2 * H * 2);
          S.bm2 = put(2 * H * 4);
        } else {
          S.m2_hi = S.m2_lo = S.bm2 = -1;
        }
      } else {
        S.v1_hi = S.v1_lo = S.v2_hi = S.v2_lo = -1;
        S.fuse_next = S.fuse_next2 = false;
        S.v2w_hi = S.v2w_lo = S.m_hi = S.m_lo = S.bm = S.m2_hi = S.m2_lo = S.bm2 = -1;
      }
      L.subs.push_back(S);
    }
  }
  L.total_bytes = p;
  L.flat_count = s;
  return true;
}

Workspace make_workspace(const hd_config& c, int B, int N) {
  Workspace W{};
  const int64_t BN = (int64_t)B * N, Fi = c.in_node_nf;
  int64_t p = 0;
  auto put = [&](int64_t bytes) {
    int64_t r = p;
    p = align256(p + bytes);
    return r;
  };
  W.h = put(BN * H * 4);
  W.ab = put(BN * 2 * H * 4);
  W.ab2 = put(BN * 2 * H * 4);
  W.agg = put(BN * H * 4);
  W.hid = put(BN * H * 4);
  W.h2 = put(BN * H * 4);
  W.x = put(BN * 3 * 4);
  W.x2 = put(BN * 3 * 4);
  W.x0 = put(BN * 3 * 4);
  W.hin = put(BN * Fi * 4);
  W.hout = put(BN * Fi * 4);
  W.eps_raw = put(BN * (3 + Fi) * 4);
  W.nanflag = put(256);
  W.state = put(256);
  W.row_off = put(((int64_t)B + 1) * 4);
  W.node_off = put(((int64_t)B + 1) * 4);
  W.total_bytes = p;
  return W;
}

// ---------------------------------------------------------------------------------------
// packing kernels
// ---------------------------------------------------------------------------------------
// dst[k*n_out + o] = src[o*ld + col0 + k]   (k < n_k, o < n_out)
__global__ void transpose_k(const float* __restrict__ src, int ld, int col0, int n_out, int n_k,
                            float* __restrict__ dst, int dst_ld, int dst_col0) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_out * n_k) return;
  int k = idx / n_out, o = idx % n_out;
  dst[(int64_t)k * dst_ld + dst_col0 + o] = src[(int64_t)o * ld + col0 + k];
}
// dst[o] = src[o*ld + col]
__global__ void column_k(const float* __restrict__ src, int ld, int col, int n, float* __restrict__ dst) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n) dst[o] = src[(int64_t)o * ld + col];
}
__global__ void copy_k(const float* __restrict__ src, int n, float* __restrict__ dst) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n) dst[o] = src ? src[o] : 0.f;
}
// dst[o] = scale * src[o]
__global__ void scale_k(const float* __restrict__ src, int n, float scale, float* __restrict__ dst) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n) dst[o] = scale * src[o];
}
// Fused pre-projection of the sub-layer that follows a GCL:  A|B(h') with h' = h + V2 hid + c2 equals
//   [h | hid] . [W1ab | W1ab V2]^T + (b1ab + W1ab c2).
// tmp[o][k] (o < 2H outputs, k < 2H inputs): k < H -> W1ab[o][k], k >= H -> sum_j W1ab[o][j] * V2[j][k-H];  bm[o] likewise.
// W1ab[o][j] = W1n[o % H][(o / H) * H + j] (outputs 0..H-1 act on h_i, H..2H-1 on h_j); fp64 accumulation.
__global__ void fuse_k(const float* __restrict__ w1n, int ld1, const float* __restrict__ b1n,
                       const float* __restrict__ v2, const float* __restrict__ c2, float* __restrict__ tmp,
                       float* __restrict__ bm, float scale) {
  const int o = blockIdx.x, k = threadIdx.x;   // grid 2H, block H
  const float* wrow = w1n + (int64_t)(o % H) * ld1 + (o / H) * H;
  double acc = 0.0;
  for (int j = 0; j < H; ++j) acc += (double)wrow[j] * (double)v2[(int64_t)j * H + k];
  tmp[(int64_t)o * 2 * H + k] = wrow[k];
  tmp[(int64_t)o * 2 * H + H + k] = (float)acc;
  if (k == 0) {
    double b = o < H ? (double)b1n[o] : 0.0;
    for (int j = 0; j < H; ++j) b += (double)wrow[j] * (double)c2[j];
    bm[o] = scale * (float)b;
  }
}
// Embedding folded into the first pre-projection:  A|B(W_emb in + b_emb) = (W1ab W_emb) in + (W1ab b_emb + [b1 | 0]).
// outT[f][o] = scale * sum_j W1ab[o][j] * W_emb[j][f],  outb[o] = scale * (sum_j W1ab[o][j] * b_emb[j] + (o < H ? b1[o] : 0))
__global__ void fuse_embed_k(const float* __restrict__ w1, int ld1, const float* __restrict__ b1,
                             const float* __restrict__ emb_w, const float* __restrict__ emb_b, int Fi,
                             float* __restrict__ outT, float* __restrict__ outb, float scale) {
  const int o = blockIdx.x, f = threadIdx.x;   // grid 2H, block Fi + 1 (thread Fi: the bias)
  const float* wrow = w1 + (int64_t)(o % H) * ld1 + (o / H) * H;
  double acc = 0.0;
  if (f < Fi) {
    for (int j = 0; j < H; ++j) acc += (double)wrow[j] * (double)emb_w[(int64_t)j * Fi + f];
    outT[(int64_t)f * 2 * H + o] = (float)((double)scale * acc);
  } else {
    acc = o < H ? (double)b1[o] : 0.0;
    for (int j = 0; j < H; ++j) acc += (double)wrow[j] * (double)emb_b[j];
    outb[o] = (float)((double)scale * acc);
  }
}
// bf16 hi/lo operand image of rows [row0, row0+rows) x K columns [col0, col0+K) of src (ld):
// img[kg][r][e] (kg < K/8, r < rows, e < 8)
__global__ void image_k(const float* __restrict__ src, int ld, int row0, int col0, int rows, int K,
                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float scale = 1.0f) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * K) return;
  int e = idx & 7, r = (idx >> 3) % rows, kg = (idx >> 3) / rows;
  float v = scale * src[(int64_t)(row0 + r) * ld + col0 + kg * 8 + e];
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// image of W[row0 .. row0+n_out) x K columns [col0, col0+K) in 128-row output tiles, written at K offset `k_at` of an
// image whose full K is `k_total` (hd_node.cu weight layout img[tile][kg < k_total/8][128][8]); the stage-2 layer's packer
int make_image128(const float* src, int ld, int row0, int col0, int n_out, int K, int k_at, int k_total, void* hi, void* lo,
                  cudaStream_t st) {
  for (int t = 0; t < n_out / 128; ++t) {
    const int64_t o = ((int64_t)t * (k_total / 8) + k_at / 8) * 128 * 8;
    image_k<<<(128 * K + 255) / 256, 256, 0, st>>>(src, ld, row0 + t * 128, col0, 128, K,
                                                  reinterpret_cast<__nv_bfloat16*>(hi) + o,
                                                  reinterpret_cast<__nv_bfloat16*>(lo) + o);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("make_image128: %s", cudaGetErrorString(e));
    return HD_E_CUDA;
  }
  return HD_OK;
}

int pack_weights(const hd_config& c, const Layout& L, const float* w, char* P, cudaStream_t st) {
  const int T = 256;
  const float NEG_LOG2E = -1.4426950408889634f;
  auto grid = [&](int64_t n) { return (unsigned)((n + T - 1) / T); };
  auto F = [&](int64_t off) { return reinterpret_cast<float*>(P + off); };
  auto BF = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(P + off); };
  const int Fi = c.in_node_nf;
  transpose_k<<<grid(H * Fi), T, 0, st>>>(w + L.s_emb_w, Fi, 0, H, Fi, F(L.emb_wT), H, 0);
  copy_k<<<grid(H), T, 0, st>>>(w + L.s_emb_b, H, F(L.emb_b));
  copy_k<<<grid(Fi * H), T, 0, st>>>(w + L.s_out_w, Fi * H, F(L.out_w));
  copy_k<<<grid(Fi), T, 0, st>>>(w + L.s_out_b, Fi, F(L.out_b));
  fuse_embed_k<<<2 * H, Fi + 1, 0, st>>>(w + L.subs[0].s_w1, 2 * H + 2, w + L.subs[0].s_b1, w + L.s_emb_w, w + L.s_emb_b,
                                        Fi, F(L.emb_abT), F(L.emb_abb), NEG_LOG2E);
  for (size_t si = 0; si < L.subs.size(); ++si) {
    const SubLayer& S = L.subs[si];
    const int ld1 = 2 * H + 2;
    transpose_k<<<grid(H * H), T, 0, st>>>(w + S.s_w1, ld1, 0, H, H, F(S.w1abT), 2 * H, 0);
    transpose_k<<<grid(H * H), T, 0, st>>>(w + S.s_w1, ld1, H, H, H, F(S.w1abT), 2 * H, H);
    copy_k<<<grid(H), T, 0, st>>>(w + S.s_b1, H, F(S.b1));
    copy_k<<<grid(H), T, 0, st>>>(nullptr, H, F(S.b1) + H);
    column_k<<<grid(H), T, 0, st>>>(w + S.s_w1, ld1, 2 * H, H, F(S.wr));
    column_k<<<grid(H), T, 0, st>>>(w + S.s_w1, ld1, 2 * H + 1, H, F(S.wd));
    transpose_k<<<grid(H * H), T, 0, st>>>(w + S.s_w2, H, 0, H, H, F(S.w2T), H, 0);
    copy_k<<<grid(H), T, 0, st>>>(w + S.s_b2, H, F(S.b2));
    copy_k<<<grid(H), T, 0, st>>>(S.s_wa >= 0 ? w + S.s_wa : nullptr, H, F(S.wa));
    copy_k<<<1, T, 0, st>>>(S.s_ba >= 0 ? w + S.s_ba : nullptr, 1, F(S.ba));
    // tensor-core images: W2 as two N-halves of 128 rows, K = H
    for (int half = 0; half < 2; ++half) {
      const int64_t o = (int64_t)half * (H / 2) * H;  // elements
      image_k<<<grid(H / 2 * H), T, 0, st>>>(w + S.s_w2, H, half * (H / 2), 0, H / 2, H, BF(S.w2_hi) + o,
                                             BF(S.w2_lo) + o);
    }
    // node-GEMM operand images (hd_node.cu), img[tile][kg][tile rows][8].  W1ab in 128-row output tiles:
    // outputs 0..255 = h_i part (W1 cols 0..H), outputs 256..511 = h_j part (W1 cols H..2H)
    for (int t = 0; t < 2 * H / 128; ++t) {
      const int64_t o = (int64_t)t * 128 * H;
      const int part = t / (H / 128), row0 = (t % (H / 128)) * 128;
      image_k<<<grid(128 * H), T, 0, st>>>(w + S.s_w1, ld1, row0, part * H, 128, H, BF(S.w1ab_hi) + o,
                                           BF(S.w1ab_lo) + o, NEG_LOG2E);
    }
    // scaled fp32 copies for the tensor-core engines (stream order: after the unscaled images above)
    scale_k<<<grid(2 * H), T, 0, st>>>(F(S.b1), 2 * H, NEG_LOG2E, F(S.b1s));
    scale_k<<<grid(H), T, 0, st>>>(F(S.wr), H, NEG_LOG2E, F(S.wrs));
    scale_k<<<grid(H), T, 0, st>>>(F(S.wd), H, NEG_LOG2E, F(S.wds));
    scale_k<<<grid(H), T, 0, st>>>(F(S.b2), H, NEG_LOG2E, F(S.b2s));
    if (S.is_gcl) {
      transpose_k<<<grid(2 * H * H), T, 0, st>>>(w + S.s_v1, 2 * H, 0, H, 2 * H, F(S.v1T), H, 0);
      copy_k<<<grid(H), T, 0, st>>>(w + S.s_c1, H, F(S.c1));
      transpose_k<<<grid(H * H), T, 0, st>>>(w + S.s_v2, H, 0, H, H, F(S.v2T), H, 0);
      copy_k<<<grid(H), T, 0, st>>>(w + S.s_c2, H, F(S.c2));
      for (int t = 0; t < H / 64; ++t) {
        const int64_t o1 = (int64_t)t * 64 * 2 * H, o2 = (int64_t)t * 64 * H;
        image_k<<<grid(64 * 2 * H), T, 0, st>>>(w + S.s_v1, 2 * H, t * 64, 0, 64, 2 * H, BF(S.v1_hi) + o1,
                                                BF(S.v1_lo) + o1);
        image_k<<<grid(64 * H), T, 0, st>>>(w + S.s_v2, H, t * 64, 0, 64, H, BF(S.v2_hi) + o2, BF(S.v2_lo) + o2);
      }
      if (S.fuse_next) {
        const SubLayer& Nx = L.subs[si + 1];   // next sub-layer of the same block
        for (int t = 0; t < H / 128; ++t) {
          const int64_t o = (int64_t)t * 128 * H;
          image_k<<<grid(128 * H), T, 0, st>>>(w + S.s_v2, H, t * 128, 0, 128, H, BF(S.v2w_hi) + o, BF(S.v2w_lo) + o);
        }
        fuse_k<<<2 * H, H, 0, st>>>(w + Nx.s_w1, ld1, w + Nx.s_b1, w + S.s_v2, w + S.s_c2, F(L.fuse_tmp), F(S.bm),
                                    NEG_LOG2E);
        for (int t = 0; t < 2 * H / 128; ++t) {
          const int64_t o = (int64_t)t * 128 * 2 * H;
          image_k<<<grid(128 * 2 * H), T, 0, st>>>(F(L.fuse_tmp), 2 * H, t * 128, 0, 128, 2 * H, BF(S.m_hi) + o,
                                                   BF(S.m_lo) + o, NEG_LOG2E);
        }
      }
      if (S.fuse_next2) {
        const SubLayer& Nx = L.subs[si + 2];   // first sub-layer of the next block (si + 1 is this block's EquivariantUpdate)
        fuse_k<<<2 * H, H, 0, st>>>(w + Nx.s_w1, ld1, w + Nx.s_b1, w + S.s_v2, w + S.s_c2, F(L.fuse_tmp), F(S.bm2),
                                    NEG_LOG2E);
        for (int t = 0; t < 2 * H / 128; ++t) {
          const int64_t o = (int64_t)t * 128 * 2 * H;
          image_k<<<grid(128 * 2 * H), T, 0, st>>>(F(L.fuse_tmp), 2 * H, t * 128, 0, 128, 2 * H, BF(S.m2_hi) + o,
                                                   BF(S.m2_lo) + o, NEG_LOG2E);
        }
      }
    }
  }
  HD_CHECK_LAUNCH();
  return HD_OK;
}

}  // namespace hd
