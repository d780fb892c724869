// hd_tc.cu - tensor-core engines (HD_ENGINE_TC_STRICT / HD_ENGINE_TC_FAST): the fused edge kernel.
//
// One persistent, warp-specialised kernel per edge-MLP sub-layer (GCL or EquivariantUpdate):
//
//   rows     : the real (i,j) node pairs of every molecule, j padded to a multiple of 8 per receiver i;
//              every CTA owns a contiguous range of whole receivers (no cross-CTA reduction, no atomics)
//   producer : 8 warps build the MMA A operand on the fly, m1_ij = SiLU(A_i + B_j + r_ij*wr + d0_ij*wd)
//              (layer 1 of the edge MLP split per node), as bf16 (hi [+ lo]) core matrices in a
//              shared-memory ring, K = 64 per stage (two stages, 64 KB)
//   MMA      : 1 thread issues tcgen05.mma (kind::f16, M=128 per CTA, N=256, K=16) against W2 resident in
//              shared memory; strict mode runs 3 passes (hi*hi + hi*lo + lo*hi) into the same fp32 TMEM
//              accumulator; cta_group::2 pairs two SMs so each holds one 128-row half of W2 (hi+lo = 128 KB)
//   epilogue : 8 warps read the accumulator from TMEM (thread = edge row, half of the columns): + b2, SiLU,
//              attention / coord dot product (halves exchanged through shared memory), mask, then a shuffle
//              transpose-reduction over the 8 rows of a group and a per-receiver running sum -> agg_i (GCL) or
//              x_i + sum_j trans_ij (EquivariantUpdate)
//   metadata : 1 warp prepares the row descriptors (receiver, sender, |x_i-x_j|^2, unit vector, flags) of tile t+1
//
// Accumulators are double buffered in TMEM (2 x 256 columns), so tile t's epilogue overlaps tile t+1's MMAs
// and tile t+2's operand generation.
#include <cuda_bf16.h>

#include "hd_common.cuh"
#include "hd_ptx.cuh"

namespace hd {

namespace tc {

// per-role cycle accounting for scripts/edge_timing.cu (compiled out of the library).  The same harness builds the
// ablation variants -DHD_EXP_NO_PROD_SILU / NO_EPI_SILU / NO_LO / NO_PASS2 / NO_LDG (results: profiles/r1_notes.md);
// none of these macros is ever defined for the library.
#ifdef HD_PHASE_TIMING
__device__ long long g_acc[2][3][16];   // [CTA 10 | CTA 11][role][slot]
// accumulate in registers (a global read-modify-write per stamp would stall the warp ~350 cycles), flush once
#define HD_T0() long long _t0 = clock64(); long long _acc[6] = {0, 0, 0, 0, 0, 0}
#define HD_ACC(role, slot, cond) do { long long _t1 = clock64(); _acc[slot] += _t1 - _t0; _t0 = _t1; } while (0)
#define HD_FLUSH(role, cond) do { if ((cond) && (blockIdx.x >> 1) == 5) for (int _i = 0; _i < 6; ++_i) g_acc[blockIdx.x & 1][role][_i] += _acc[_i]; } while (0)
#else
#define HD_T0() do { } while (0)
#define HD_ACC(role, slot, cond) do { } while (0)
#define HD_FLUSH(role, cond) do { } while (0)
#endif

// event timeline of one CTA (scripts/edge_timing.cu -DHD_TIMELINE): clock64 stamps, slot -> first time the event happened
#ifdef HD_TIMELINE
__device__ long long g_tl[2][128];
__device__ unsigned long long g_span[160][2];   // globaltimer at entry / exit of every CTA
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define HD_STAMP(slot, cond) do { if ((cond) && (blockIdx.x >> 1) == 5) g_tl[blockIdx.x & 1][slot] = clock64(); \
    if ((cond) && ((slot) == 0 || (slot) == 4)) g_span[blockIdx.x][(slot) == 4] = gtimer(); } while (0)
#else
#define HD_STAMP(slot, cond) do { } while (0)
#endif

constexpr int TILE_M = 128;           // edge rows per CTA tile
#ifndef HD_KCH
#define HD_KCH 64
#endif
constexpr int KCH = HD_KCH;           // K columns per operand stage (one producer -> MMA hand-off)
constexpr int NCH = H / KCH;          // stages per tile
constexpr int NSTAGE = 128 / KCH;     // operand ring depth: 128 K-columns in flight
constexpr int HSPS = KCH / 16;        // 16-column half steps per stage
constexpr int A_KG = 2048;            // bytes between K-adjacent core matrices of A (128 rows x 16 B)
constexpr int A_HALF = (KCH / 8) * A_KG;   // one stage of hi (or lo): KCH/8 core-matrix columns
constexpr int W_KG = 2048;            // bytes between K-adjacent core matrices of a 128-row W half image
constexpr int W_HALF = (H / 8) * W_KG;  // 64 KB: one 128-row half image (hi or lo)
constexpr int PMETA_BUFS = 2;         // producer-side row metadata: tiles t, t+1
constexpr int EMETA_BUFS = 4;         // epilogue-side row metadata / group table: tiles t-2 .. t+1
constexpr int MAX_B = 255;            // molecules whose row_off table is staged in shared memory (larger: read from L2)
constexpr int MAX_MOL = 4096;         // molecules per launch
constexpr int NPW = 8;                // producer warps
constexpr int NEW = 8;                // epilogue warps
constexpr int PROD_THREADS = 32 * NPW, EPI_THREADS = 32 * NEW;
#ifndef HD_EXP_IDLE_WARPS
#define HD_EXP_IDLE_WARPS 0   // timing experiment only: extra idle warps (what a lower register cap alone costs)
#endif
constexpr int NTHREADS = PROD_THREADS + EPI_THREADS + 64 + 32 * HD_EXP_IDLE_WARPS;   // + MMA/alloc warp + metadata warp
#ifdef HD_EXP_NO_FILL_HELP
constexpr bool FILL_HELP = false;
#else
constexpr bool FILL_HELP = true;    // the epilogue warps build half of the first tile's operand chunks (pipeline fill)
#endif

struct PMeta {        // what the operand producers need of an edge row
  float r, d0;        // |x_i-x_j|^2, |x0_i-x0_j|^2
  int recv, send;     // flat node rows b*N+i, b*N+j
};
struct EMeta {        // what the epilogue needs of an edge row
  float cd0, cd1, cd2;  // (x_i-x_j)/(sqrt(r+1e-8)+norm_constant)  (EquivariantUpdate only)
  int flags;          // bit0: row is a real edge slot (j<n, inside this CTA's range); bit1: j==i
};

struct Params {
  const float* a_img;  // [H/16][BN][16]  A_i (+b1), K-chunk-major so that 8 consecutive nodes x 16 columns are 512 B
  const float* b_img;  // [H/16][BN][16]  B_j
  int64_t kc_stride;   // BN*16 floats between 16-column chunks
  const float* x;      // [BN][3] block-entry coordinates
  const float* x0;     // [BN][3] EGNN-entry coordinates
  const int32_t* sizes;
  const int32_t* row_off;  // [B+1] prefix of n_b * pad8(n_b)
  const int32_t* node_off; // [B+1] prefix of n_b when the node rows are ragged, else null (row of (b,i) = b*N + i)
  const void* w_hi;    // bf16 images [2][32][128][8]
  const void* w_lo;
  const float *b2, *wa, *ba, *wr, *wd;
  float* out;          // GCL: agg [BN][H]; EQUIV: x_out [BN][3]
  int B, N;
  int attention, use_tanh;
  float range, norm_constant, norm_div;
};

template <bool STRICT, int CG>
struct Smem {
#ifdef HD_EXP_WLO_ALIAS   // timing experiment only (wrong numerics): no resident W2 lo image, the lo pass re-reads W2 hi
  static constexpr int W_PARTS = 1;
#else
  static constexpr int W_PARTS = STRICT ? 2 : 1;
#endif
  static constexpr int W_BYTES = (CG == 2 ? 1 : 2) * W_HALF * W_PARTS;
  static constexpr int STAGE = A_HALF * (STRICT ? 2 : 1);
  static constexpr int OFF_W = 0;
  static constexpr int OFF_A = OFF_W + W_BYTES;
  static constexpr int OFF_SCR = OFF_A + NSTAGE * STAGE;           // [16][H] fp32
  static constexpr int OFF_VEC = OFF_SCR + 16 * H * 4;             // b2, wa, wr, wd
  static constexpr int OFF_ROW = OFF_VEC + 4 * H * 4;              // row_off [MAX_B+1]
  static constexpr int OFF_PMETA = OFF_ROW + (MAX_B + 1) * 4;
  static constexpr int OFF_EMETA = OFF_PMETA + PMETA_BUFS * TILE_M * (int)sizeof(PMeta);
  static constexpr int OFF_GRP = OFF_EMETA + EMETA_BUFS * TILE_M * (int)sizeof(EMeta);   // int2 [EMETA_BUFS][16]
  static constexpr int OFF_DOT = OFF_GRP + EMETA_BUFS * 16 * 8;    // [2][TILE_M] partial attention dots
  static constexpr int OFF_BAR = OFF_DOT + 2 * TILE_M * 4;
  // barriers: full[NSTAGE], empty[NSTAGE], acc_full[2], acc_empty[2], w_local, w_ready, meta, tile_start, meta0,
  // helper_meta ; then tmem ptr
  static constexpr int NBAR = 2 * NSTAGE + 10;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
  static_assert(TOTAL <= 232448, "shared memory budget (227 KB per CTA)");
};

// largest b with row_off[b] <= R  (row_off[0] = 0, row_off[B] = total > R)
__device__ __forceinline__ int find_mol(const int* row_off, int B, int R) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (row_off[mid] <= R) lo = mid; else hi = mid;
  }
  return lo;
}
// find_mol by a whole warp on a table in global memory: 32 probes per round (one L2 round trip each) instead of
// one; three rounds for 4096 molecules.  Result is warp-uniform.
__device__ __forceinline__ int find_mol_warp(const int* row_off, int B, int R, int lane) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int step = (hi - lo + 31) >> 5;
    const int idx = lo + lane * step;
    const bool le = idx < hi && row_off[idx] <= R;          // monotone in lane; lane 0 always true
    const int k = 31 - __clz(__ballot_sync(0xffffffffu, le));
    lo += k * step;
    hi = min(lo + step, hi);
  }
  return lo;
}
// first receiver-pair boundary >= S (rows of a molecule: hd_common.cuh edge_rows)
__device__ __forceinline__ int align_recv(const int* row_off, const int32_t* sizes, int B, int S) {
  const int total = row_off[B];
  if (S >= total) return total;
  const int b = find_mol(row_off, B, S);
  const int n = sizes[b], stride = 2 * ((n + 7) & ~7);
  const int local = S - row_off[b];
  return row_off[b] + ((local + stride - 1) / stride) * stride;
}

// SiLU in the scaled domain the kernel works in.  Operands arrive as t = -log2(e) * v (the factor is folded into the
// packed weights, hd_layout.cu), so  SiLU(v) = v / (1 + 2^t) = -ln2 * t / (1 + 2^t).
//   strict: returns t / (1 + 2^t)            = SiLU(v) / K_OUT, K_OUT = -ln 2       (ex2.approx, rcp.approx, ~2 ulp)
//   fast  : returns c*t * (1 + tanh(c*t)), c = -ln2 / 2   = SiLU(v) / K_OUT, K_OUT = -1 / ln 2... see silu_kout()
template <bool STRICT>
__device__ __forceinline__ float silu_scaled(float t) {
  if constexpr (STRICT) {
    return t * ptx::rcp_approx(1.0f + ptx::ex2_approx(t));
  } else {
    const float hv = -0.34657359027997264f * t;   // v / 2
    return fmaf(hv, ptx::tanh_approx(hv), hv);    // = SiLU(v)
  }
}
// Strict mode, two elements at once: 1/(1+2^ta) and 1/(1+2^tb) from ONE reciprocal, r = 1/((1+2^ta)(1+2^tb)), as
// r*(1+2^tb) and r*(1+2^ta) - 3 MUFU operations per pair instead of 4 (the kernel's co-bound is the MUFU pipe).
// The exponent is clamped at 63 so that the product stays finite (2^126): beyond it SiLU(v) = v/(1+e^-v) is below
// 5e-18 in magnitude (v < -43.6) and the clamped evaluation t * 2^-63 differs from it by less than that.
__device__ __forceinline__ void silu_scaled_pair(float ta, float tb, float& ya, float& yb) {
  const float da = 1.0f + ptx::ex2_approx(fminf(ta, 63.0f));
  const float db = 1.0f + ptx::ex2_approx(fminf(tb, 63.0f));
  const float r = ptx::rcp_approx(da * db);
  ya = ta * (r * db);
  yb = tb * (r * da);
}
// SiLU(v) = silu_kout() * silu_scaled(t)
template <bool STRICT>
__host__ __device__ constexpr float silu_kout() { return STRICT ? -0.6931471805599453f : 1.0f; }
// t2 = -log2(e) * (W2 . SiLU + b2) = silu_kacc() * (W2 . silu_scaled) + b2s
template <bool STRICT>
__host__ __device__ constexpr float silu_kacc() { return STRICT ? 1.0f : -1.4426950408889634f; }

// WIDE: the variant for batches whose row table does not fit the shared-memory slot (B > MAX_B) and / or whose node
// rows are ragged (p.node_off); the default instantiation carries neither branch.
template <bool GCL, bool STRICT, int CG, bool WIDE = false>
__global__ void __launch_bounds__(NTHREADS, 1) edge_tc_k(const Params p) {
  using S = Smem<STRICT, CG>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0;

  float* s_scr = reinterpret_cast<float*>(smem + S::OFF_SCR);
  float* s_b2 = reinterpret_cast<float*>(smem + S::OFF_VEC);
  float* s_wa = s_b2 + H;
  float* s_wr = s_wa + H;
  float* s_wd = s_wr + H;
  int* s_row = reinterpret_cast<int*>(smem + S::OFF_ROW);
  PMeta* s_pmeta = reinterpret_cast<PMeta*>(smem + S::OFF_PMETA);
  EMeta* s_emeta = reinterpret_cast<EMeta*>(smem + S::OFF_EMETA);
  int2* s_grp = reinterpret_cast<int2*>(smem + S::OFF_GRP);   // {receiver row, group is inside this CTA's range}
  float* s_dot = reinterpret_cast<float*>(smem + S::OFF_DOT);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + S::OFF_TMEM);
  const uint32_t bar0 = sbase + S::OFF_BAR;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto bar_accf = [&](int a) { return bar0 + 8u * (2 * NSTAGE + a); };
  auto bar_acce = [&](int a) { return bar0 + 8u * (2 * NSTAGE + 2 + a); };
  const uint32_t bar_wl = bar0 + 8u * (2 * NSTAGE + 4), bar_wr = bar0 + 8u * (2 * NSTAGE + 5);
  const uint32_t bar_meta = bar0 + 8u * (2 * NSTAGE + 6), bar_tstart = bar0 + 8u * (2 * NSTAGE + 7);
  // one-shot pair for the helpers of the first tile: tile 0's metadata is ready / every helper warp has read it
  const uint32_t bar_meta0 = bar0 + 8u * (2 * NSTAGE + 8), bar_hmeta = bar0 + 8u * (2 * NSTAGE + 9);
  constexpr int MMA_WARP = NPW + NEW, META_WARP = MMA_WARP + 1;

  // ---- one-time setup --------------------------------------------------------------------------
  // Everything up to pdl_wait() touches only constants (weights) and on-chip state, so it overlaps the previous
  // kernel's tail: barrier init, the resident W2 image (TMA bulk copies, 64-128 KB), the bias / weight vectors, TMEM.
  pdl_trigger();
  HD_STAMP(0, tid == 0);
  for (int k = tid; k < H; k += NTHREADS) {
    s_b2[k] = p.b2[k];
    s_wa[k] = p.wa[k];
    s_wr[k] = p.wr[k];
    s_wd[k] = p.wd[k];
  }
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      ptx::mbar_init(bar_full(s), NPW * CG);   // one elected arrive per producer warp of each CTA
      ptx::mbar_init(bar_empty(s), 1);         // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_accf(a), 1);          // tcgen05.commit
      ptx::mbar_init(bar_acce(a), NEW * CG);   // one elected arrive per epilogue warp of each CTA
    }
    ptx::mbar_init(bar_wl, 1);
    ptx::mbar_init(bar_meta, 1);
    ptx::mbar_init(bar_tstart, NPW);
    ptx::mbar_init(bar_meta0, 1);
    ptx::mbar_init(bar_hmeta, NEW);
    ptx::mbar_init(bar_wr, CG);
    ptx::fence_mbar_init();
    // resident W2 image(s): this CTA's 128-row half (CG=2) or both halves (CG=1)
    ptx::mbar_expect_tx(bar_wl, S::W_BYTES);
    const int nhalf = CG == 2 ? 1 : 2;
    for (int hf = 0; hf < nhalf; ++hf) {
      const int src_half = CG == 2 ? (int)rank : hf;
      for (int part = 0; part < S::W_PARTS; ++part) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(part ? p.w_lo : p.w_hi) + (size_t)src_half * W_HALF;
        const uint32_t dst = sbase + S::OFF_W + (hf * S::W_PARTS + part) * W_HALF;
        for (int off = 0; off < W_HALF; off += 16384) ptx::bulk_g2s(dst + off, src + off, 16384, bar_wl);
      }
    }
  }
  if (warp == MMA_WARP) ptx::tmem_alloc<CG>(sbase + S::OFF_TMEM, 512);
  pdl_wait();   // row_off, sizes, x, the A|B operands: written by earlier kernels of the chain
  const bool big_b = WIDE && p.B > MAX_B;   // table does not fit the shared-memory slot: searched in global memory
  if (!big_b) for (int k = tid; k <= p.B; k += NTHREADS) s_row[k] = p.row_off[k];
  // ---- this pair's rows: three lanes find the receiver-pair-aligned boundaries of the two CTAs' ranges -------------
  // (one search each, side by side; every thread doing all four searches itself cost 2 us of set-up per launch)
  const int nCTA = gridDim.x;
  const int* row_tab = big_b ? p.row_off : s_row;
  int* s_rng = reinterpret_cast<int*>(s_dot);   // [3] boundaries; the dot scratch is not in use yet
  __syncthreads();                              // s_row complete
  if (tid < 3) {
    const int total = row_tab[p.B];
    const int rpc = (total + nCTA - 1) / nCTA;
    s_rng[tid] = align_recv(row_tab, p.sizes, p.B, min(((CG == 2 ? (int)(blockIdx.x & ~1u) : (int)blockIdx.x) + tid) * rpc, total));
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_relaxed(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  HD_STAMP(1, tid == 0);
  const int row_begin = s_rng[CG == 2 ? (int)(blockIdx.x & 1u) : 0], row_end = s_rng[(CG == 2 ? (int)(blockIdx.x & 1u) : 0) + 1];
  int ntiles = (row_end - row_begin + TILE_M - 1) / TILE_M;
  if constexpr (CG == 2) {
    const int pa = s_rng[(blockIdx.x & 1u) ^ 1u], pe = s_rng[((blockIdx.x & 1u) ^ 1u) + 1];
    ntiles = max(ntiles, (pe - pa + TILE_M - 1) / TILE_M);
  }
  __syncthreads();                              // every thread has read s_rng before the epilogue reuses the slot

  HD_STAMP(2, tid == 0);
  // ---- operand production (the producer warps; for the first tile also the epilogue warps, see below) ----------
  // warp pw builds rows [16 pw, 16 pw + 16) of an operand stage: lane -> (row in 8-group, 4 columns of a 16-column half)
  const int pw = warp < NPW ? warp : warp - NPW;
  const int rsub = (lane >> 1) & 7;
  const int qsub = 2 * (lane >> 4) + (lane & 1);
  const uint32_t kcs = (uint32_t)p.kc_stride;   // floats between 16-column chunks; b_img = a_img + 16 chunks
  // per-tile state of the 2 rows this thread feeds (16*pw + 8*rb + rsub): receivers 2p and 2p+1 of a pair with the
  // SAME sender (hd_common.cuh edge_rows), so one B_j load serves both; an 8-row group is live or dead as a whole, and
  // the second row is never live without the first
  struct RowState {
    float rr[2], dd[2];
    uint32_t oa[2], ob;      // element offsets of A_i (per row) / B_j (shared) in chunk 0
    bool ok[2];
  };
  auto read_meta = [&](int t, RowState& r) {
    const PMeta* meta = s_pmeta + (t % PMETA_BUFS) * TILE_M;
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
      const PMeta m = meta[16 * pw + 8 * rb + rsub];
      r.rr[rb] = m.r;
      r.dd[rb] = m.d0;
      r.ok[rb] = m.recv >= 0;
      r.oa[rb] = (r.ok[rb] ? (uint32_t)m.recv : 0u) * 16u + 4u * qsub;
      if (rb == 0) r.ob = (r.ok[0] ? (uint32_t)m.send : 0u) * 16u + 4u * qsub + 16u * kcs;
    }
  };
  // v = {A of row 0, A of row 1, B of both}
  auto load_half = [&](float4 (&v)[3], const RowState& r, int hs) {
    const uint32_t o = hs * kcs;
#ifdef HD_EXP_HALF_PROD   // timing experiment only (wrong numerics): producers build every other half step
    if (hs & 1) return;
#endif
#ifdef HD_EXP_NO_PROD     // timing experiment only: producers build nothing, only the hand-offs remain
    return;
#endif
#ifndef HD_EXP_NO_LDG
    if (r.ok[0]) {
      v[0] = __ldg(reinterpret_cast<const float4*>(p.a_img + (o + r.oa[0])));
      v[2] = __ldg(reinterpret_cast<const float4*>(p.a_img + (o + r.ob)));
    }
    if (r.ok[1]) v[1] = __ldg(reinterpret_cast<const float4*>(p.a_img + (o + r.oa[1])));
#endif
  };
  // one 16-column half stage of this thread's 2 rows: 8 independent SiLU chains, written phase by phase so the
  // MUFU latencies of the chains overlap
  auto half_step = [&](const float4 (&v)[3], const RowState& r, int hs, int s) {
#ifdef HD_EXP_HALF_PROD
    if (hs & 1) return;
#endif
#ifdef HD_EXP_NO_PROD
    return;
#endif
    const int ph = hs % HSPS;
    const int k0 = 16 * hs + 4 * qsub;
    const float4 w_r = *reinterpret_cast<const float4*>(s_wr + k0);
    const float4 w_d = *reinterpret_cast<const float4*>(s_wd + k0);
    uint8_t* stage = smem + S::OFF_A + s * S::STAGE + (2 * ph + (lane >> 4)) * A_KG + (lane & 1) * 8 +
                     (16 * pw + rsub) * 16;
    if (!(r.ok[0] || r.ok[1])) return;   // rows outside this CTA's range keep stale operand data; their
                                         // accumulator rows are never read (group table)
    float pre[8], m[8];
#if !defined(HD_EXP_NO_F32X2) && !defined(HD_EXP_NO_PAIR_RCP) && !defined(HD_EXP_NO_PROD_SILU)
    if constexpr (STRICT) {
      // packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2): the same IEEE operations in the same order, half the issue
      // slots.  q[2*rb], q[2*rb+1] = elements (0,1), (2,3) of row rb.
      float2 q[4], d[4];
      const float2 one2 = make_float2(1.0f, 1.0f);
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
        const float4 a = v[rb], b = v[2];
        const float2 rr2 = make_float2(r.rr[rb], r.rr[rb]), dd2 = make_float2(r.dd[rb], r.dd[rb]);
        q[2 * rb] = ptx::fma2(dd2, make_float2(w_d.x, w_d.y),
                              ptx::fma2(rr2, make_float2(w_r.x, w_r.y), ptx::add2(make_float2(a.x, a.y), make_float2(b.x, b.y))));
        q[2 * rb + 1] = ptx::fma2(dd2, make_float2(w_d.z, w_d.w),
                                  ptx::fma2(rr2, make_float2(w_r.z, w_r.w), ptx::add2(make_float2(a.z, a.w), make_float2(b.z, b.w))));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        d[k] = make_float2(ptx::ex2_approx(fminf(q[k].x, 63.0f)), ptx::ex2_approx(fminf(q[k].y, 63.0f)));
#pragma unroll
      for (int k = 0; k < 4; ++k) d[k] = ptx::add2(d[k], one2);
      // silu_scaled_pair across the two packed halves of a row: elements (0,2) and (1,3) share a reciprocal
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
        const float2 pr = ptx::mul2(d[2 * rb], d[2 * rb + 1]);
        const float2 rc = make_float2(ptx::rcp_approx(pr.x), ptx::rcp_approx(pr.y));
        const float2 ma = ptx::mul2(q[2 * rb], ptx::mul2(rc, d[2 * rb + 1]));       // SiLU / (-ln 2)
        const float2 mb = ptx::mul2(q[2 * rb + 1], ptx::mul2(rc, d[2 * rb]));
        m[4 * rb] = ma.x; m[4 * rb + 1] = ma.y; m[4 * rb + 2] = mb.x; m[4 * rb + 3] = mb.y;
      }
    } else
#endif
    {
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
      const float4 a = v[rb], b = v[2];
      pre[4 * rb + 0] = fmaf(r.dd[rb], w_d.x, fmaf(r.rr[rb], w_r.x, a.x + b.x));
      pre[4 * rb + 1] = fmaf(r.dd[rb], w_d.y, fmaf(r.rr[rb], w_r.y, a.y + b.y));
      pre[4 * rb + 2] = fmaf(r.dd[rb], w_d.z, fmaf(r.rr[rb], w_r.z, a.z + b.z));
      pre[4 * rb + 3] = fmaf(r.dd[rb], w_d.w, fmaf(r.rr[rb], w_r.w, a.w + b.w));
    }
#ifdef HD_EXP_NO_PROD_SILU
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = pre[k];
#else
    // pre[] is t = -log2(e) * (A_i + B_j + r*wr + d*wd): A|B, wr, wd carry the factor (hd_layout.cu)
    if constexpr (STRICT) {
#ifdef HD_EXP_NO_PAIR_RCP
      float e[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) e[k] = ptx::ex2_approx(pre[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) e[k] = ptx::rcp_approx(1.0f + e[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = pre[k] * e[k];            // SiLU / (-ln 2)
#else
      // silu_scaled_pair, phase by phase over the 8 elements (4 pairs) so the MUFU latencies overlap
      float d[8], rc[4];
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = ptx::ex2_approx(fminf(pre[k], 63.0f));
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = 1.0f + d[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) rc[k] = ptx::rcp_approx(d[2 * k] * d[2 * k + 1]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        m[2 * k] = pre[2 * k] * (rc[k] * d[2 * k + 1]);            // SiLU / (-ln 2)
        m[2 * k + 1] = pre[2 * k + 1] * (rc[k] * d[2 * k]);
      }
#endif
    } else {
      float hv[8], th[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) hv[k] = -0.34657359027997264f * pre[k];   // v / 2
#pragma unroll
      for (int k = 0; k < 8; ++k) th[k] = ptx::tanh_approx(hv[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = fmaf(hv[k], th[k], hv[k]);        // SiLU
    }
#endif
    }
    (void)pre;
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
      if (!r.ok[rb]) continue;
      const float m0 = m[4 * rb], m1 = m[4 * rb + 1], m2 = m[4 * rb + 2], m3 = m[4 * rb + 3];
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(m0, m1), h23 = __floats2bfloat162_rn(m2, m3);
      uint2 hi;
      hi.x = *reinterpret_cast<const uint32_t*>(&h01);
      hi.y = *reinterpret_cast<const uint32_t*>(&h23);
      uint8_t* dst = stage + 8 * rb * 16;
      *reinterpret_cast<uint2*>(dst) = hi;
#ifndef HD_EXP_NO_LO
      if constexpr (STRICT) {
#ifndef HD_EXP_NO_F32X2
        const float2 r01 = ptx::add2(make_float2(m0, m1), make_float2(-__uint_as_float(hi.x << 16),
                                                                      -__uint_as_float(hi.x & 0xffff0000u)));
        const float2 r23 = ptx::add2(make_float2(m2, m3), make_float2(-__uint_as_float(hi.y << 16),
                                                                      -__uint_as_float(hi.y & 0xffff0000u)));
        const float l0 = r01.x, l1 = r01.y, l2 = r23.x, l3 = r23.y;
#else
        const float l0 = m0 - __uint_as_float(hi.x << 16), l1 = m1 - __uint_as_float(hi.x & 0xffff0000u);
        const float l2 = m2 - __uint_as_float(hi.y << 16), l3 = m3 - __uint_as_float(hi.y & 0xffff0000u);
#endif
        const __nv_bfloat162 g01 = __floats2bfloat162_rn(l0, l1), g23 = __floats2bfloat162_rn(l2, l3);
        uint2 lo;
        lo.x = *reinterpret_cast<const uint32_t*>(&g01);
        lo.y = *reinterpret_cast<const uint32_t*>(&g23);
        *reinterpret_cast<uint2*>(dst + A_HALF) = lo;
      }
#endif
    }
  };
  auto publish = [&](int s) {
#ifndef HD_EXP_NO_PROXY_FENCE   // timing experiment only
    ptx::fence_async_smem();
#endif
    __syncwarp();
    if (lane == 0) {
      if (CG == 2 && rank != 0) ptx::mbar_arrive_cluster_relaxed(bar_full(s), 0);
      else ptx::mbar_arrive_relaxed(bar_full(s));
    }
  };
  // Operand chunks c_first, c_first + c_step, ... of the tiles [t_begin, t_end).  The producer warps run it once over
  // all tiles (every chunk, except that they leave the odd chunks of tile 0 to the helpers); HELPER = the epilogue
  // warps, idle until the first accumulator exists, building the odd chunks of tile 0 - the pipeline fill
  // (producing the first tile) then takes about half as long.
  auto produce = [&](int t_begin, int t_end, const bool helper) {
    HD_T0();
    RowState cur, nxt;
    float4 va[3], vb[3];   // {A row 0, A row 1, B} of the two half stages in flight
#pragma unroll
    for (int k = 0; k < 3; ++k) va[k] = vb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto c_first_of = [&](int t) { return (helper && t == 0) ? 1 : 0; };
    auto c_step_of = [&](int t) { return (FILL_HELP && t == 0) ? 2 : 1; };
    auto tile_started = [&]() {   // this warp holds the tile's metadata in registers: its slot may be recycled
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(helper ? bar_hmeta : bar_tstart);
    };
    if (t_end > t_begin) {
      if (helper) ptx::mbar_wait(bar_meta0, 0); else ptx::mbar_wait(bar_meta, t_begin & 1);
      read_meta(t_begin, cur);
      tile_started();
      load_half(va, cur, c_first_of(t_begin) * HSPS);
    }
    nxt = cur;
    int pending = -1;   // operand stage written but not yet published (published one half step late, when the
                        // fence no longer has to wait for its stores)
    for (int t = t_begin; t < t_end; ++t) {
      HD_ACC(0, 0, tid == 0);
      const int c_step = c_step_of(t);
#pragma unroll 1
      for (int c = c_first_of(t); c < NCH; c += c_step) {
        const int gc = t * NCH + c, st = gc % NSTAGE;
#pragma unroll
        for (int pp = 0; pp < HSPS / 2; ++pp) {
          const int hs0 = c * HSPS + 2 * pp;
          load_half(vb, cur, hs0 + 1);
          HD_ACC(0, 5, tid == 0);   // issue loads
          if (pp == 0 && pending == st) {   // (split first tile) this role's previous chunk sits in the very stage it
            publish(pending);               // needs next: the MMA can only free it once it has been published
            pending = -1;
          }
          if (pp == 0) ptx::mbar_wait(bar_empty(st), ((gc / NSTAGE) & 1) ^ 1);
          HD_ACC(0, 1, tid == 0);   // wait for a free operand stage
          half_step(va, cur, hs0, st);
          HD_ACC(0, 2, tid == 0);   // half steps
          if (pp == 0 && pending >= 0) publish(pending);
          HD_ACC(0, 3, tid == 0);   // publish
          if (pp + 1 < HSPS / 2) {
            load_half(va, cur, hs0 + 2);
          } else if (c + c_step < NCH) {   // first half step of this role's next chunk
            load_half(va, cur, (c + c_step) * HSPS);
          } else if (t + 1 < t_end) {      // first operands of the next tile: the tile boundary costs no load latency
            ptx::mbar_wait(bar_meta, (t + 1) & 1);
            read_meta(t + 1, nxt);
            load_half(va, nxt, c_first_of(t + 1) * HSPS);
          }
          HD_ACC(0, 5, tid == 0);
          half_step(vb, cur, hs0 + 1, st);
          HD_ACC(0, 2, tid == 0);
        }
        pending = st;
      }
      HD_STAMP(16 + t, tid == 0 && t < 16);
      if (t + 1 < t_end) {
        cur = nxt;
        tile_started();
      }
    }
    if (pending >= 0) publish(pending);
    HD_FLUSH(0, tid == 0);
  };

  if (warp < NPW) {
    // =========================== producers ===========================
    produce(0, ntiles, false);
  } else if (warp < MMA_WARP) {
    // =========================== epilogue ===========================
    // warp -> TMEM lane quarter q (= warp % 4) and column half; thread = edge row, 128 of the 256 columns
    const int ew = warp - NPW, q = ew & 3, half = ew >> 2, etid = tid - PROD_THREADS;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16) + 128u * half;
    const float* b2h = s_b2 + 128 * half;
    const float* wah = s_wa + 128 * half;
    // running sums of the two receivers of the current pair: even 8-row groups of a tile belong to receiver 2p, odd
    // groups to receiver 2p+1 (hd_common.cuh edge_rows; tiles start on a 16-row boundary of the pair-aligned range)
    float carry[2] = {0.f, 0.f};
    int cur_recv[2] = {-1, -1};
    const float ba = GCL ? p.ba[0] : 0.f;
    if (FILL_HELP && ntiles > 0) produce(0, 1, true);   // pipeline fill: build the odd operand chunks of tile 0
    auto flush = [&](int hh) {
      const int rcv = cur_recv[hh];
      if (rcv < 0) return;
      if (GCL) p.out[(int64_t)rcv * H + etid] = (silu_kout<STRICT>() * carry[hh]) / p.norm_div;
      else if (etid < 3) p.out[(int64_t)rcv * 3 + etid] = p.x[(int64_t)rcv * 3 + etid] + carry[hh] / p.norm_div;
    };
    HD_T0();
    for (int t = 0; t < ntiles; ++t) {
      const int as = t & 1;
      ptx::mbar_wait(bar_accf(as), (t >> 1) & 1);
      ptx::tc_fence_after();
      HD_ACC(1, 0, etid == 0);   // wait for the accumulator
      HD_STAMP(32 + t, etid == 0 && t < 16);
      const EMeta mine = s_emeta[(t % EMETA_BUFS) * TILE_M + 32 * q + lane];
      const bool live = (mine.flags & 3) == 1;   // real, off-diagonal edge
      const bool warp_live = __any_sync(0xffffffffu, mine.flags != 0);   // any row of this lane quarter in range
      const uint32_t acc = lane_base + 256u * as;
      float v[32];
      float dot = 0.f;
      float2 dot2 = make_float2(0.f, 0.f);   // packed partial sums of the attention / coordinate dot product
#pragma unroll 1
      for (int cc = 0; cc < (warp_live ? 4 : 0); ++cc) {
        ptx::tmem_ld32(acc + 32 * cc, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 bb = *reinterpret_cast<const float4*>(b2h + 32 * cc + 4 * k4);
          const float4 ww = *reinterpret_cast<const float4*>(wah + 32 * cc + 4 * k4);
#ifdef HD_EXP_NO_EPI_SILU
          const float m0 = v[4 * k4 + 0] + bb.x, m1 = v[4 * k4 + 1] + bb.y, m2 = v[4 * k4 + 2] + bb.z, m3 = v[4 * k4 + 3] + bb.w;
#else
          // bb = -log2(e) * b2; KACC undoes the producer's output scale: t2 = -log2(e) * (W2 . SiLU + b2)
          constexpr float KACC = silu_kacc<STRICT>();
          float m0, m1, m2, m3;
#if !defined(HD_EXP_NO_F32X2) && !defined(HD_EXP_NO_PAIR_RCP)
          if constexpr (STRICT) {
            // packed fp32x2 (FADD2 / FMUL2 / FFMA2); KACC == 1 in strict mode.  Elements (0,2) and (1,3) share a
            // reciprocal (silu_scaled_pair across the two packed halves)
            const float2 ta = ptx::add2(make_float2(v[4 * k4 + 0], v[4 * k4 + 1]), make_float2(bb.x, bb.y));
            const float2 tb = ptx::add2(make_float2(v[4 * k4 + 2], v[4 * k4 + 3]), make_float2(bb.z, bb.w));
            const float2 one2 = make_float2(1.0f, 1.0f);
            const float2 da = ptx::add2(make_float2(ptx::ex2_approx(fminf(ta.x, 63.0f)), ptx::ex2_approx(fminf(ta.y, 63.0f))), one2);
            const float2 db = ptx::add2(make_float2(ptx::ex2_approx(fminf(tb.x, 63.0f)), ptx::ex2_approx(fminf(tb.y, 63.0f))), one2);
            const float2 pr = ptx::mul2(da, db);
            const float2 rc = make_float2(ptx::rcp_approx(pr.x), ptx::rcp_approx(pr.y));
            const float2 ma = ptx::mul2(ta, ptx::mul2(rc, db));
            const float2 mb = ptx::mul2(tb, ptx::mul2(rc, da));
            dot2 = ptx::fma2(ma, make_float2(ww.x, ww.y), dot2);
            dot2 = ptx::fma2(mb, make_float2(ww.z, ww.w), dot2);
            v[4 * k4 + 0] = ma.x; v[4 * k4 + 1] = ma.y; v[4 * k4 + 2] = mb.x; v[4 * k4 + 3] = mb.y;
            continue;
          }
#endif
#ifndef HD_EXP_NO_PAIR_RCP
          if constexpr (STRICT) {
            silu_scaled_pair(fmaf(KACC, v[4 * k4 + 0], bb.x), fmaf(KACC, v[4 * k4 + 1], bb.y), m0, m1);
            silu_scaled_pair(fmaf(KACC, v[4 * k4 + 2], bb.z), fmaf(KACC, v[4 * k4 + 3], bb.w), m2, m3);
          } else
#endif
          {
            m0 = silu_scaled<STRICT>(fmaf(KACC, v[4 * k4 + 0], bb.x));
            m1 = silu_scaled<STRICT>(fmaf(KACC, v[4 * k4 + 1], bb.y));
            m2 = silu_scaled<STRICT>(fmaf(KACC, v[4 * k4 + 2], bb.z));
            m3 = silu_scaled<STRICT>(fmaf(KACC, v[4 * k4 + 3], bb.w));
          }
#endif
          dot = fmaf(m0, ww.x, dot);
          dot = fmaf(m1, ww.y, dot);
          dot = fmaf(m2, ww.z, dot);
          dot = fmaf(m3, ww.w, dot);
          v[4 * k4 + 0] = m0; v[4 * k4 + 1] = m1; v[4 * k4 + 2] = m2; v[4 * k4 + 3] = m3;
        }
        if (GCL) ptx::tmem_st32(acc + 32 * cc, v);
      }
      HD_ACC(1, 1, etid == 0);   // pass 1
      // full-row dot product: exchange the two column halves
      dot += dot2.x + dot2.y;
      s_dot[half * TILE_M + 32 * q + lane] = dot;
      ptx::named_bar_sync(4, EPI_THREADS);
      HD_ACC(1, 2, etid == 0);   // exchange barrier
      dot = silu_kout<STRICT>() * (s_dot[32 * q + lane] + s_dot[TILE_M + 32 * q + lane]);   // back to the true scale
      if (GCL) {
        ptx::tmem_wait_st();
        float att = 1.0f;
        if (p.attention) att = STRICT ? 1.0f / (1.0f + __expf(-(dot + ba))) : fmaf(0.5f, ptx::tanh_approx(0.5f * (dot + ba)), 0.5f);
        const float scale = live ? att : 0.f;
        if (t > 0) ptx::named_bar_sync(2, EPI_THREADS);   // previous tile's combine has finished reading the scratch
        const int c4 = (lane & 4) ? 16 : 0, c2 = (lane & 2) ? 8 : 0, c1 = (lane & 1) ? 4 : 0;
#pragma unroll 1
#ifdef HD_EXP_NO_PASS2
        for (int cc = 0; cc < 0; ++cc) {
#else
        for (int cc = 0; cc < (warp_live ? 4 : 0); ++cc) {
#endif
          ptx::tmem_ld32(acc + 32 * cc, v);
          ptx::tmem_wait_ld();
          float f[16], g[8], hsum[4];
#ifndef HD_EXP_NO_F32X2
          // the same transpose-reduction with the multiplies and adds packed two by two (FMUL2 / FADD2)
          const float2 scale2 = make_float2(scale, scale);
          const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
#pragma unroll
          for (int k = 0; k < 16; k += 2) {
            const float2 lo_v = ptx::mul2(make_float2(v[k], v[k + 1]), scale2);
            const float2 hi_v = ptx::mul2(make_float2(v[k + 16], v[k + 17]), scale2);
            const float2 send = up4 ? lo_v : hi_v, keep = up4 ? hi_v : lo_v;
            const float2 got = make_float2(__shfl_xor_sync(0xffffffffu, send.x, 4), __shfl_xor_sync(0xffffffffu, send.y, 4));
            const float2 sum = ptx::add2(keep, got);
            f[k] = sum.x; f[k + 1] = sum.y;
          }
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            const float2 a = make_float2(f[k], f[k + 1]), b = make_float2(f[k + 8], f[k + 9]);
            const float2 send = up2 ? a : b, keep = up2 ? b : a;
            const float2 got = make_float2(__shfl_xor_sync(0xffffffffu, send.x, 2), __shfl_xor_sync(0xffffffffu, send.y, 2));
            const float2 sum = ptx::add2(keep, got);
            g[k] = sum.x; g[k + 1] = sum.y;
          }
#pragma unroll
          for (int k = 0; k < 4; k += 2) {
            const float2 a = make_float2(g[k], g[k + 1]), b = make_float2(g[k + 4], g[k + 5]);
            const float2 send = up1 ? a : b, keep = up1 ? b : a;
            const float2 got = make_float2(__shfl_xor_sync(0xffffffffu, send.x, 1), __shfl_xor_sync(0xffffffffu, send.y, 1));
            const float2 sum = ptx::add2(keep, got);
            hsum[k] = sum.x; hsum[k + 1] = sum.y;
          }
#else
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float lo_v = v[k] * scale, hi_v = v[k + 16] * scale;
            const float send = (lane & 4) ? lo_v : hi_v;
            const float keep = (lane & 4) ? hi_v : lo_v;
            f[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float send = (lane & 2) ? f[k] : f[k + 8];
            const float keep = (lane & 2) ? f[k + 8] : f[k];
            g[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float send = (lane & 1) ? g[k] : g[k + 4];
            const float keep = (lane & 1) ? g[k + 4] : g[k];
            hsum[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          }
#endif
          float4 o;
          o.x = hsum[0]; o.y = hsum[1]; o.z = hsum[2]; o.w = hsum[3];
          *reinterpret_cast<float4*>(s_scr + (4 * q + (lane >> 3)) * H + 128 * half + 32 * cc + c4 + c2 + c1) = o;
        }
      } else {
        float tv = p.use_tanh ? (STRICT ? tanhf(dot) : ptx::tanh_approx(dot)) * p.range : dot;
        if (!live) tv = 0.f;
        float t0 = mine.cd0 * tv, t1 = mine.cd1 * tv, t2 = mine.cd2 * tv;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          t0 += __shfl_xor_sync(0xffffffffu, t0, o);
          t1 += __shfl_xor_sync(0xffffffffu, t1, o);
          t2 += __shfl_xor_sync(0xffffffffu, t2, o);
        }
        if (t > 0) ptx::named_bar_sync(2, EPI_THREADS);
        if (half == 0 && (lane & 7) == 0) {
          float* dst = s_scr + (4 * q + (lane >> 3)) * H;
          dst[0] = t0; dst[1] = t1; dst[2] = t2;
        }
      }
      HD_ACC(1, 3, etid == 0);   // pass 2 (incl. waiting for the previous combine)
      ptx::tc_fence_before();
      ptx::named_bar_sync(3, EPI_THREADS);   // scratch complete
      HD_ACC(1, 4, etid == 0);   // scratch barrier
      // per-receiver running sums: the groups of a receiver are consecutive among those of its parity; thread = output column
      if (GCL || etid < 3) {
        int2 gi[16];
        float sv[16];
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          gi[g] = s_grp[(t % EMETA_BUFS) * 16 + g];
          sv[g] = s_scr[g * H + etid];
        }
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          if (!gi[g].y) continue;                 // group outside this CTA's range / dead half of an odd pair
          const int hh = g & 1;
          if (gi[g].x != cur_recv[hh]) {
            flush(hh);
            cur_recv[hh] = gi[g].x;
            carry[hh] = 0.f;
          }
          carry[hh] += sv[g];
        }
      }
      // accumulator stage drained and this tile's group table no longer needed: hand both back (the producers
      // recycle the group-table slot of tile t when they prepare tile t+4, which the MMA warp gates on this arrive)
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && rank != 0) ptx::mbar_arrive_cluster_relaxed(bar_acce(as), 0);
        else ptx::mbar_arrive_relaxed(bar_acce(as));
      }
      HD_ACC(1, 5, etid == 0);   // combine + release
      HD_STAMP(48 + t, etid == 0 && t < 16);
    }
    flush(0);
    flush(1);
    HD_FLUSH(1, etid == 0);
  } else if (warp == META_WARP) {
    // =========================== row metadata ===========================
    // one warp prepares the per-row metadata of tile t+1 while the producers work on tile t
    // big_b: the row table lives in global memory, and the little L1 left beside 227 KB of shared memory is swept by
    // the producers' gathers, so a per-row binary search costs ten L2 round trips.  Instead the warp keeps a window
    // of 32 consecutive molecules in registers (a 128-row tile overlaps at most 17: every molecule has >= 8 rows),
    // anchored by one cooperative search for the first tile and carried from tile to tile after that; a row finds
    // its molecule with five shuffles.
    int win_b0 = -1, win_first = 0, win_n = 0, win_node0 = 0;
    for (int t = 0; t < ntiles; ++t) {
      if (t > 0) ptx::mbar_wait(bar_tstart, (t - 1) & 1);   // producers hold tile t-1's metadata in registers
      if (FILL_HELP && t == PMETA_BUFS) ptx::mbar_wait(bar_hmeta, 0);   // ... and so do the helpers of tile 0 (its slot is recycled now)
      if (big_b && row_begin + t * TILE_M < row_end) {
        if (win_b0 < 0) win_b0 = find_mol_warp(p.row_off, p.B, row_begin, lane);
        const int wb = min(win_b0 + lane, p.B - 1);
        win_first = win_b0 + lane < p.B ? p.row_off[wb] : 0x7fffffff;
        win_n = p.sizes[wb];
        win_node0 = p.node_off ? p.node_off[wb] : wb * p.N;
      }
      // the four rows of a lane (32 rq + lane) go through the dependent steps side by side - table search, sizes,
      // coordinates - so the tile costs one chain of latencies, not four (2.5 us -> < 1 us ahead of the first tile)
      constexpr int RQ = TILE_M / 32;
      int Rr[RQ], first[RQ], nn[RQ], node0[RQ];
#pragma unroll
      for (int rq = 0; rq < RQ; ++rq) Rr[rq] = row_begin + t * TILE_M + 32 * rq + lane;
      if (big_b) {             // all lanes take part in the shuffles, in range or not
#pragma unroll
        for (int rq = 0; rq < RQ; ++rq) {
          int wk = 0;
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1) {
            const int v = __shfl_sync(0xffffffffu, win_first, wk + step);
            if (v <= Rr[rq]) wk += step;
          }
          first[rq] = __shfl_sync(0xffffffffu, win_first, wk);
          nn[rq] = __shfl_sync(0xffffffffu, win_n, wk);
          node0[rq] = __shfl_sync(0xffffffffu, win_node0, wk);
          if (rq == RQ - 1) {   // the next tile's window starts at the molecule of this tile's last row
            const int last = __shfl_sync(0xffffffffu, wk, 31);
            win_b0 = min(win_b0 + last, p.B - 1);
          }
        }
      } else {
        // largest b < B with s_row[b] <= R, eight probes (B <= MAX_B = 255); rows past the table end up at B - 1
        int lo[RQ];
#pragma unroll
        for (int rq = 0; rq < RQ; ++rq) lo[rq] = 0;
#pragma unroll
        for (int step = 128; step >= 1; step >>= 1) {
#pragma unroll
          for (int rq = 0; rq < RQ; ++rq) {
            const int c = lo[rq] + step;
            if (c < p.B && s_row[c] <= Rr[rq]) lo[rq] = c;
          }
        }
#pragma unroll
        for (int rq = 0; rq < RQ; ++rq) {
          first[rq] = s_row[lo[rq]];
          nn[rq] = p.sizes[lo[rq]];
          node0[rq] = (WIDE && p.node_off) ? p.node_off[lo[rq]] : lo[rq] * p.N;
        }
      }
      // rows of a molecule (edge_rows): pair p of receivers, block jb of 8 senders, half hh (receiver 2p + hh)
      PMeta pm[RQ];
      EMeta em[RQ];
      float xi[RQ][3], xj[RQ][3], oi[RQ][3], oj[RQ][3];
#pragma unroll
      for (int rq = 0; rq < RQ; ++rq) {
        const int n = nn[rq], npad = max((n + 7) & ~7, 8);
        const int local = max(Rr[rq] - first[rq], 0);
        const int pr = local / (2 * npad), rem = local - pr * 2 * npad;
        const int i = 2 * pr + ((rem >> 3) & 1), j = ((rem >> 4) << 3) + (rem & 7);
        const bool recv_ok = Rr[rq] < row_end && i < n;   // else: outside this CTA's range / the dead second half of an
                                                           // odd molecule's last pair
        const bool edge = recv_ok && j < n;
        pm[rq].recv = recv_ok ? node0[rq] + i : -1;
        // padding slot (j >= n): any finite B row, the same for both halves of the pair
        pm[rq].send = recv_ok ? node0[rq] + (edge ? j : 2 * pr) : -1;
        em[rq].flags = edge ? (1 | (j == i ? 2 : 0)) : (recv_ok ? 4 : 0);   // 4: padding slot inside a real receiver's
                                                                            // group (operand row finite, masked later)
        const int64_t ri = edge ? pm[rq].recv : 0, rj = edge ? pm[rq].send : 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          xi[rq][k] = p.x[3 * ri + k];
          xj[rq][k] = p.x[3 * rj + k];
          oi[rq][k] = p.x0[3 * ri + k];
          oj[rq][k] = p.x0[3 * rj + k];
        }
      }
#pragma unroll
      for (int rq = 0; rq < RQ; ++rq) {
        const int row = 32 * rq + lane;
        const bool edge = (em[rq].flags & 1) != 0;
        const float d0 = xi[rq][0] - xj[rq][0], d1 = xi[rq][1] - xj[rq][1], d2 = xi[rq][2] - xj[rq][2];
        const float e0 = oi[rq][0] - oj[rq][0], e1 = oi[rq][1] - oj[rq][1], e2 = oi[rq][2] - oj[rq][2];
        const float r = d0 * d0 + d1 * d1 + d2 * d2;
        pm[rq].r = edge ? r : 0.f;
        pm[rq].d0 = edge ? e0 * e0 + e1 * e1 + e2 * e2 : 0.f;
        em[rq].cd0 = em[rq].cd1 = em[rq].cd2 = 0.f;
        if (!GCL && edge) {
          const float nrm = sqrtf(r + 1e-8f) + p.norm_constant;
          em[rq].cd0 = d0 / nrm;
          em[rq].cd1 = d1 / nrm;
          em[rq].cd2 = d2 / nrm;
        }
        s_pmeta[(t % PMETA_BUFS) * TILE_M + row] = pm[rq];
        s_emeta[(t % EMETA_BUFS) * TILE_M + row] = em[rq];
        if ((row & 7) == 0) s_grp[(t % EMETA_BUFS) * 16 + (row >> 3)] = make_int2(pm[rq].recv, em[rq].flags != 0);
      }
      __syncwarp();
      HD_STAMP(64 + t, lane == 0 && t < 16);
      if (lane == 0) {
        ptx::mbar_arrive(bar_meta);
        if (FILL_HELP && t == 0) ptx::mbar_arrive(bar_meta0);
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      ptx::mbar_wait(bar_wl, 0);
      if (CG == 2 && rank != 0) ptx::mbar_arrive_cluster(bar_wr, 0);
      else ptx::mbar_arrive(bar_wr);
      if (rank == 0) {
        ptx::mbar_wait(bar_wr, 0);
        HD_STAMP(3, true);
        constexpr uint32_t IDESC = CG == 2 ? ptx::idesc_bf16(256, 256) : ptx::idesc_bf16(128, 128);
        HD_T0();
        for (int t = 0; t < ntiles; ++t) {
          const int as = t & 1;
          ptx::mbar_wait(bar_acce(as), ((t >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          HD_ACC(2, 0, true);   // wait for a free accumulator
          for (int c = 0; c < NCH; ++c) {
            const int gc = t * NCH + c, s = gc % NSTAGE;
            ptx::mbar_wait(bar_full(s), (gc / NSTAGE) & 1);
            ptx::tc_fence_after();
            HD_ACC(2, 1, true);   // wait for operands
            const uint32_t a_hi = sbase + S::OFF_A + s * S::STAGE;
#pragma unroll
            for (int ks = 0; ks < KCH / 16; ++ks) {
              const uint32_t acc_on = (c | ks) ? 1u : 0u;
              const uint32_t kgw = (c * (KCH / 8) + ks * 2) * W_KG;
              const uint64_t da_hi = ptx::smem_desc(a_hi + ks * 2 * A_KG, A_KG, 128);
              const uint64_t da_lo = ptx::smem_desc(a_hi + A_HALF + ks * 2 * A_KG, A_KG, 128);
              if constexpr (CG == 2) {
                const uint32_t d = tmem + 256u * as;
                const uint64_t db_hi = ptx::smem_desc(sbase + S::OFF_W + kgw, W_KG, 128);
                ptx::mma_bf16<2>(d, da_hi, db_hi, IDESC, acc_on);
                if constexpr (STRICT) {
                  const uint64_t db_lo = ptx::smem_desc(sbase + S::OFF_W + (S::W_PARTS - 1) * W_HALF + kgw, W_KG, 128);
                  ptx::mma_bf16<2>(d, da_hi, db_lo, IDESC, 1u);
                  ptx::mma_bf16<2>(d, da_lo, db_hi, IDESC, 1u);
                }
              } else {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  const uint32_t d = tmem + 256u * as + 128u * hf;
                  const uint32_t wb = sbase + S::OFF_W + hf * S::W_PARTS * W_HALF;
                  const uint64_t db_hi = ptx::smem_desc(wb + kgw, W_KG, 128);
                  ptx::mma_bf16<1>(d, da_hi, db_hi, IDESC, acc_on);
                  if constexpr (STRICT) {
                    const uint64_t db_lo = ptx::smem_desc(wb + (S::W_PARTS - 1) * W_HALF + kgw, W_KG, 128);
                    ptx::mma_bf16<1>(d, da_hi, db_lo, IDESC, 1u);
                    ptx::mma_bf16<1>(d, da_lo, db_hi, IDESC, 1u);
                  }
                }
              }
            }
            ptx::mma_commit<CG>(bar_empty(s));     // operand stage consumed (both CTAs)
            HD_ACC(2, 2, true);   // issue
          }
          ptx::mma_commit<CG>(bar_accf(as));       // accumulator complete (both CTAs)
          HD_STAMP(80 + t, t < 16);
        }
        HD_FLUSH(2, true);
      }
    }
    __syncwarp();
  }

  // ---- teardown ----------------------------------------------------------------------------------
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_relaxed(); else __syncthreads();   // the peer may still arrive on our barriers
  if (warp == MMA_WARP) ptx::tmem_dealloc<CG>(tmem, 512);
  HD_STAMP(4, tid == 0);
}

__global__ void plan_k(const int32_t* __restrict__ sizes, int B, int32_t* __restrict__ row_off) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int acc = 0;
  row_off[0] = 0;
  for (int b = 0; b < B; ++b) {
    const int n = sizes[b];
    acc += edge_rows(n);
    row_off[b + 1] = acc;
  }
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <bool GCL, bool STRICT, int CG, bool WIDE = false>
static int launch_edge(const Params& p, cudaStream_t st) {
  using S = Smem<STRICT, CG>;
  static bool configured = false;
  auto kern = edge_tc_k<GCL, STRICT, CG, WIDE>;
  if (!configured) {
    HD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
#ifdef HD_EXP_CARVEOUT
    HD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, HD_EXP_CARVEOUT));
#endif
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  int grid = sm_count();
  if (CG == 2) grid &= ~1;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  HD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
  return HD_OK;
}

}  // namespace tc

bool tc_available() { return true; }

static int ensure_plan(const FwdCtx& c) {
  if (c.B > tc::MAX_MOL) {
    set_error("tensor-core engine supports at most %d molecules per call (got %d)", tc::MAX_MOL, c.B);
    return HD_E_INVALID;
  }
  if (!c.planned) {
    HD_CHECK_CUDA(launch_pdl(tc::plan_k, dim3(1), dim3(32), 0, c.stream, c.sizes, c.B,
                            reinterpret_cast<int32_t*>(c.ws + c.W.row_off)));
    count_launch();
    c.planned = true;
  }
  return HD_OK;
}

static int edge_launch(const FwdCtx& c, int si, const float* x, const float* x0, float* out, int engine,
                       bool use_ab2 = false) {
  const SubLayer& S = c.L->subs[si];
  auto F = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  int rc = ensure_plan(c);
  if (rc) return rc;
#ifdef HD_EXP_SKIP_EDGE   // timing experiment only (scripts/step_ablation.sh)
  return HD_OK;
#endif
  tc::Params p{};
  p.a_img = reinterpret_cast<const float*>(c.ws + (use_ab2 ? c.W.ab2 : c.W.ab));
  p.kc_stride = (int64_t)c.B * c.N * 16;
  p.b_img = p.a_img + (H / 16) * p.kc_stride;
  p.x = x;
  p.x0 = x0;
  p.sizes = c.sizes;
  p.row_off = reinterpret_cast<const int32_t*>(c.ws + c.W.row_off);
  p.node_off = c.node_off;
  p.w_hi = c.packed + S.w2_hi;
  p.w_lo = c.packed + S.w2_lo;
  p.b2 = F(S.b2s);   // the -log2(e)-scaled copies (silu_scaled)
  p.wa = F(S.wa);
  p.ba = F(S.ba);
  p.wr = F(S.wrs);
  p.wd = F(S.wds);
  p.out = out;
  p.B = c.B;
  p.N = c.N;
  p.attention = c.cfg->attention;
  p.use_tanh = c.cfg->tanh;
  p.range = c.cfg->coords_range / (float)c.cfg->n_layers;
  p.norm_constant = c.cfg->norm_constant;
  p.norm_div = c.cfg->aggregation_mean ? (float)c.N : c.cfg->normalization_factor;
  const bool strict = engine == HD_ENGINE_TC_STRICT;
  const bool wide = c.B > tc::MAX_B || c.node_off != nullptr;
  if (S.is_gcl) {
    if (wide)
      return strict ? tc::launch_edge<true, true, 2, true>(p, c.stream) : tc::launch_edge<true, false, 2, true>(p, c.stream);
    return strict ? tc::launch_edge<true, true, 2>(p, c.stream) : tc::launch_edge<true, false, 2>(p, c.stream);
  }
  if (!c.x_prezeroed)   // padded rows: x*mask = 0 (the kernel only writes the rows of real receivers)
    HD_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * 3 * c.B * c.N, c.stream));
  if (wide)
    return strict ? tc::launch_edge<false, true, 2, true>(p, c.stream) : tc::launch_edge<false, false, 2, true>(p, c.stream);
  return strict ? tc::launch_edge<false, true, 2>(p, c.stream) : tc::launch_edge<false, false, 2>(p, c.stream);
}

int linear_tc(const FwdCtx& c, const float* X1, int ld1, int K1, const float* X2, int ld2, int K2, const void* w_hi,
              const void* w_lo, int n_out, int tile_n, const float* bias, float* Y, int ldy, int mode,
              const float* resid, bool strict);

// A = -log2(e) (h W1a^T + b1) ; B = -log2(e) h W1b^T  (scaled images: silu_scaled; the bias image is [b1 | 0], so the
// bias lands on the A half only)
static int preproject(const FwdCtx& c, const SubLayer& S, const float* h, bool strict) {
  float* ab = reinterpret_cast<float*>(c.ws + c.W.ab);
  return linear_tc(c, h, H, H, nullptr, 0, 0, c.packed + S.w1ab_hi, c.packed + S.w1ab_lo, 2 * H, 128,
                   reinterpret_cast<const float*>(c.packed + S.b1s), ab, 2 * H, 3, nullptr, strict);
}

int linear_tc_v2_and_preproject(const FwdCtx& c, const float* hid, const float* h, float* h_out, const void* v2_hi,
                                const void* v2_lo, const float* c2, const void* m_hi, const void* m_lo,
                                const float* bm, float* ab, bool strict, const void* m2_hi, const void* m2_lo,
                                const float* bm2, float* ab2);

int tc_gcl_fused(const FwdCtx& c, int si, const float* h, float* h_out, const float* x, const float* x0, int engine,
                 bool use_ab2) {
  const SubLayer& S = c.L->subs[si];
  auto F = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  float* ab = reinterpret_cast<float*>(c.ws + c.W.ab);
  float* agg = reinterpret_cast<float*>(c.ws + c.W.agg);
  float* hid = reinterpret_cast<float*>(c.ws + c.W.hid);
  const bool strict = engine == HD_ENGINE_TC_STRICT;
  int rc;
  if (use_ab2) {
    c.ab2_ready = false;   // consumed by the edge kernel below
  } else {
    if (!c.ab_ready && (rc = preproject(c, S, h, strict))) return rc;
    c.ab_ready = false;
  }
  if ((rc = edge_launch(c, si, x, x0, agg, engine, use_ab2))) return rc;
  if ((rc = linear_tc(c, h, H, H, agg, H, H, c.packed + S.v1_hi, c.packed + S.v1_lo, H, 64, F(S.c1), hid, H, 1, nullptr,
                      strict)))
    return rc;
  // node_mlp.2 -> h_out, the successor's A|B from [h | hid] in the same launch (the edge kernel has consumed ab) and,
  // after the last GCL of a block, the A|B of the next block's first sub-layer as well (h is not touched in between)
  float* ab2 = reinterpret_cast<float*>(c.ws + c.W.ab2);
  const bool two = S.fuse_next2;
  if ((rc = linear_tc_v2_and_preproject(c, hid, h, h_out, c.packed + S.v2w_hi, c.packed + S.v2w_lo, F(S.c2),
                                        c.packed + S.m_hi, c.packed + S.m_lo, F(S.bm), ab, strict,
                                        two ? c.packed + S.m2_hi : nullptr, two ? c.packed + S.m2_lo : nullptr,
                                        two ? F(S.bm2) : nullptr, ab2)))
    return rc;
  c.ab_ready = true;
  if (two) c.ab2_ready = true;
  return HD_OK;
}

int tc_gcl(const FwdCtx& c, int si, float* h, const float* x, const float* x0, int engine) {
  const SubLayer& S = c.L->subs[si];
  auto F = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  float* agg = reinterpret_cast<float*>(c.ws + c.W.agg);
  float* hid = reinterpret_cast<float*>(c.ws + c.W.hid);
  const bool strict = engine == HD_ENGINE_TC_STRICT;
  int rc;
  if ((rc = preproject(c, S, h, strict))) return rc;
  if ((rc = edge_launch(c, si, x, x0, agg, engine))) return rc;
  // node_model (egnn_new.py:52-62): h = (h + node_mlp([h, agg])) * node_mask
  if ((rc = linear_tc(c, h, H, H, agg, H, H, c.packed + S.v1_hi, c.packed + S.v1_lo, H, 64, F(S.c1), hid, H, 1, nullptr,
                      strict)))
    return rc;
  return linear_tc(c, hid, H, H, nullptr, 0, 0, c.packed + S.v2_hi, c.packed + S.v2_lo, H, 64, F(S.c2), h, H, 2, h,
                   strict);
}

int tc_equiv(const FwdCtx& c, int si, const float* h, const float* x, const float* x0, float* x_out, int engine) {
  const SubLayer& S = c.L->subs[si];
  int rc;
  if (!c.ab_ready && (rc = preproject(c, S, h, engine == HD_ENGINE_TC_STRICT))) return rc;
  c.ab_ready = false;
  return edge_launch(c, si, x, x0, x_out, engine);
}

int tc_edge_only(const FwdCtx& c, int si, const float* x, const float* x0, int engine) {
  const SubLayer& S = c.L->subs[si];
  float* out = reinterpret_cast<float*>(c.ws + (S.is_gcl ? c.W.agg : c.W.x2));
  return edge_launch(c, si, x, x0, out, engine);
}

}  // namespace hd
