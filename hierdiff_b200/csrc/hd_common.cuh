// hd_common.cuh - shared declarations of the native library (config, packed-weight layout,
// workspace layout, device helpers).  Host+device header; no torch types anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/hierdiff_b200.h"

namespace hd {

constexpr int H = 256;  // hidden_nf this build is specialised for

void set_error(const char* fmt, ...);
void count_launch();

#define HD_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      hd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return HD_E_CUDA;                                                                  \
    }                                                                                    \
  } while (0)

#define HD_CHECK_LAUNCH()                                                                \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      hd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return HD_E_CUDA;                                                                  \
    }                                                                                    \
    hd::count_launch();                                                                  \
  } while (0)

// ---------------------------------------------------------------------------------------
// Packed weight image.  All offsets are BYTES from the start of the packed buffer and are
// multiples of 256.  One `SubLayer` per edge-MLP sub-layer: inv_sublayers GCLs then one
// EquivariantUpdate per block.
// ---------------------------------------------------------------------------------------
struct SubLayer {
  bool is_gcl;
  // sources: float offsets into the flat state_dict-order buffer (-1 = absent)
  int64_t s_w1, s_b1, s_w2, s_b2, s_v1, s_c1, s_v2, s_c2, s_wa, s_ba;
  // fp32 images
  int64_t w1abT;  // [H][2H]  : w1abT[k][o] = W1[o % H][ (o / H) * H + k ]  (o<H: h_i part, o>=H: h_j part)
  int64_t b1;     // [2H]     : [b1 | 0] (bias of the fused A|B pre-projection)
  int64_t wr;     // [H]      : W1[:, 2H]   (weight of |x_i-x_j|^2)
  int64_t wd;     // [H]      : W1[:, 2H+1] (weight of |x0_i-x0_j|^2)
  int64_t w2T;    // [H][H]   : w2T[k][o] = W2[o][k]
  int64_t b2;     // [H]
  int64_t wa;     // [H]      : att_mlp.0.weight (GCL) or coord_mlp.4.weight (equiv)
  int64_t ba;     // [1] (+pad): att_mlp.0.bias (GCL) or 0
  // tensor-core engines evaluate SiLU(v) = v / (1 + 2^t) on t = -log2(e) * v directly: these copies carry the factor,
  // so the kernels save the multiply in front of every ex2 (the A|B pre-projection image w1ab_* is scaled likewise)
  int64_t b1s;    // [2H]     : -log2(e) * [b1 | 0]
  int64_t wrs;    // [H]      : -log2(e) * wr
  int64_t wds;    // [H]      : -log2(e) * wd
  int64_t b2s;    // [H]      : -log2(e) * b2
  int64_t v1T;    // [2H][H]  : GCL only, node_mlp.0 transposed ([h | agg] input order)
  int64_t c1;     // [H]
  int64_t v2T;    // [H][H]
  int64_t c2;     // [H]
  // tcgen05 operand images (bf16, canonical K-major no-swizzle core-matrix order):
  //   img[half][kg][n_local][8] with half<2, kg<H/8, n_local<H/2: element = W[half*H/2+n_local][kg*8+e]
  int64_t w2_hi, w2_lo;        // edge_mlp.2 / coord_mlp.2
  // node-GEMM images (hd_node.cu): img[tile][kg < K/8][tile rows][8], 128-row tiles for w1ab, 64-row for v1/v2
  int64_t w1ab_hi, w1ab_lo;    // [2H out][H k] node pre-projection (4 tiles)
  int64_t v1_hi, v1_lo;        // [H out][2H k]  (GCL) node_mlp.0 (4 tiles, K = 2H)
  int64_t v2_hi, v2_lo;        // [H out][H k]   (GCL) node_mlp.2 (4 tiles)
  // fused launch "node_mlp.2 + next pre-projection" (GCL followed by another sub-layer of the same block), 128-row tiles:
  int64_t v2w_hi, v2w_lo;      // [H out][H k]    node_mlp.2 again, 2 tiles
  int64_t m_hi, m_lo;          // [2H out][2H k]  -log2(e) * [W1ab_next | W1ab_next . V2]: applied to [h | hid] it gives the
                               //                 next sub-layer's A|B of h' = h + V2 hid + c2 without waiting for h'
  int64_t bm;                  // [2H] fp32       -log2(e) * ([b1_next | 0] + W1ab_next . c2)
  bool fuse_next;              // the three fields above are populated
  // last GCL of a block that has a successor block: the same pre-multiplied image for the FIRST sub-layer of the next
  // block (h does not change across the EquivariantUpdate in between), so that its pre-projection rides in this launch too
  int64_t m2_hi, m2_lo, bm2;
  bool fuse_next2;
};

struct Layout {
  int64_t s_emb_w, s_emb_b, s_out_w, s_out_b;  // flat offsets
  int64_t fuse_tmp;  // [2H][2H] fp32 scratch of pack_weights
  int64_t emb_wT;  // [Fi][H]
  int64_t emb_b;   // [H]
  // embedding folded into the first sub-layer's pre-projection (tensor-core engines, sampling loop):
  //   A|B(embedding(in)) = in . emb_abT + emb_abb,  emb_abT = -log2(e) * (W1ab . W_emb)^T, emb_abb = -log2(e) * (W1ab . b_emb + [b1 | 0])
  int64_t emb_abT; // [Fi][2H]
  int64_t emb_abb; // [2H]
  int64_t out_w;   // [Fi][H] (as in the state_dict)
  int64_t out_b;   // [Fi] (+pad)
  std::vector<SubLayer> subs;
  int64_t total_bytes;
  int64_t flat_count;
};

// returns false (and sets the error) when cfg is not supported by this build
bool make_layout(const hd_config& cfg, Layout* out);

// ---------------------------------------------------------------------------------------
// Workspace (bytes offsets, 256-aligned)
// ---------------------------------------------------------------------------------------
struct Workspace {
  int64_t h;      // [BN][H]
  int64_t ab;     // [BN][2H]   node pre-projection: A_i (bias folded) | B_j
  int64_t ab2;    // [BN][2H]   same, for the first sub-layer of the NEXT block (written one edge kernel ahead)
  int64_t agg;    // [BN][H]
  int64_t hid;    // [BN][H]
  int64_t h2;     // [BN][H]    ping-pong partner of h (the fused node launch must not update h in place)
  int64_t x;      // [BN][3]
  int64_t x2;     // [BN][3]    ping-pong for the coordinate update
  int64_t x0;     // [BN][3]    EGNN-entry coordinates
  int64_t hin;    // [BN][Fi]
  int64_t hout;   // [BN][Fi]
  int64_t eps_raw;// [BN][3+F]
  int64_t nanflag;// [1] int32 per forward
  int64_t state;  // [4] int32: sampling-loop state {step (advanced by the tail kernel), step of the running forward
                  // (latched by out_vel_k), NaN flag of even steps, NaN flag of odd steps}
  int64_t node_off;// [B+1] int32: prefix of n_b (ragged node rows of the sampling path, tensor-core engines)
  int64_t row_off;// [B+1] int32: prefix of n_b * pad8(n_b) (edge rows per molecule, tensor-core engines)
  int64_t total_bytes;
};
Workspace make_workspace(const hd_config& cfg, int B, int N);

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the per-step chain is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may start (barrier / TMEM set-up, weight copies)
// while the previous kernel drains, and must call pdl_wait() before touching anything an earlier kernel wrote
// and before their first global write.  pdl_wait() returns when the previous grid has completed and flushed;
// completion is transitive because every kernel of the chain executes it in every CTA.  HD_NO_PDL=1 disables.
// ---------------------------------------------------------------------------------------
bool pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float sigmoid_acc(float v) { return 1.0f / (1.0f + expf(-v)); }
// SiLU as the reference evaluates it, v * sigmoid(v) (ATen: x / (1 + exp(-x)))
__device__ __forceinline__ float silu_acc(float v) { return v / (1.0f + expf(-v)); }

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

// Edge rows of a molecule with n nodes in the tensor-core edge kernel's row space (row_off is its prefix sum):
// receivers are taken two at a time and, per pair, the rows run over blocks of 8 senders - 8 rows (receiver 2p, senders
// 8jb..8jb+7) followed by 8 rows (receiver 2p+1, the SAME senders) - so that the two rows an operand-producer thread
// feeds share their sender and its B_j row is fetched once.  ceil(n/2) pairs x 2 x pad8(n) rows (odd n: the last
// pair's second half is dead).
__host__ __device__ inline int edge_rows(int n) { return ((n + 1) >> 1) * 2 * ((n + 7) & ~7); }

// engine entry points (hd_fp32.cu / hd_tc.cu) ------------------------------------------------
struct FwdCtx {
  const hd_config* cfg;
  const Layout* L;
  const char* packed;
  char* ws;
  Workspace W;
  const int32_t* sizes;
  int B, N;
  cudaStream_t stream;
  const int32_t* node_off = nullptr;   // ragged node rows: [B+1] prefix of n_b; row of node (b,i) = node_off[b] + i
  int rows_bound = 0;                  // ragged node rows: the caller's bound on sum(sizes) (0: B*N); sizes the node-GEMM grids
  mutable bool planned = false;  // ws.row_off holds the edge-row prefix for `sizes`
  mutable bool ab_ready = false; // ws.ab already holds the next sub-layer's A|B operands (fused node launch)
  mutable bool ab2_ready = false;// ws.ab2 already holds the A|B operands of the next block's first sub-layer
  bool x_prezeroed = false;      // padded rows of ws.x / ws.x2 are already 0 (hd_dynamics_forward): the coordinate
                                 // update then needs no memset of its output
};

// GCL sub-layer `si` (index into L.subs) in place on ctx.ws h; reads x (ws.x) and x0.
int fp32_gcl(const FwdCtx& c, int si, float* h, const float* x, const float* x0);
int fp32_equiv(const FwdCtx& c, int si, const float* h, const float* x, const float* x0, float* x_out);
int tc_gcl(const FwdCtx& c, int si, float* h, const float* x, const float* x0, int engine);
// same, but when `h_out` != nullptr and the sub-layer has a fused successor: h_out receives the updated features
// (h is left untouched) and the successor's A|B operands are produced in the same launch (c.ab_ready is set)
// use_ab2: read this sub-layer's own A|B operands from ws.ab2 (first sub-layer of a block, see ab2_ready)
int tc_gcl_fused(const FwdCtx& c, int si, const float* h, float* h_out, const float* x, const float* x0, int engine,
                 bool use_ab2 = false);
int tc_equiv(const FwdCtx& c, int si, const float* h, const float* x, const float* x0, float* x_out, int engine);
// edge kernel alone on ws.ab (profiling hook); output into ws.agg (GCL) / ws.x2 (equiv)
int fp32_edge_only(const FwdCtx& c, int si, const float* x, const float* x0);
int tc_edge_only(const FwdCtx& c, int si, const float* x, const float* x0, int engine);
bool tc_available();

}  // namespace hd
