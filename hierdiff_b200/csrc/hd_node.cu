// hd_node.cu - tensor-core node GEMMs of the tensor-core engines: every per-node nn.Linear of a sub-layer
// (the A|B pre-projection of edge_mlp.0 / coord_mlp.0, node_mlp.0, node_mlp.2; egnn_new.py:52-62 and SURVEY.md 7
// "layer-1 split") as   Y[r, o] = epilogue(bias[o] + sum_k [X1 | X2][r, k] * W[o, k]).
//
//   tile     : 128 node rows x 64 or 128 output columns per CTA, K streamed in chunks of 64 through a 4-stage ring
//   producer : 8 warps read the fp32 activations (L2-resident, written by the previous kernel), split them into
//              bf16 hi (+ lo in strict mode) and store them as canonical K-major core matrices (the MMA A operand)
//   weights  : 1 thread bulk-copies (TMA 1-D) the pre-built bf16 hi/lo image chunk of W for this column tile
//   MMA      : 1 thread issues tcgen05.mma kind::f16 M=128 N=64 K=16; strict = hi*hi + hi*lo + lo*hi into one fp32
//              TMEM accumulator (64 columns)
//   epilogue : the 8 producer warps read the accumulator (thread = row, half of the columns): + bias, SiLU;
//              the tile is staged in shared memory and written out coalesced, (resid + v) * node_mask applied
//              on the way out with the residual rows prefetched before the accumulator wait
//
// Algorithmic HBM bytes: rows*(K + n_out)*4 activations + n_out*K*2(*2) weights; everything is L2-resident at the
// sizes of the sampling path (h is 2.6 MB at B=64, N=40), so the kernel is bound by L2->SM latency/bandwidth.
#include <cuda_bf16.h>

#include "hd_common.cuh"
#include "hd_ptx.cuh"

namespace hd {
namespace lin {

// phase stamps for scripts/node_timing.cu (compiled out of the library)
#ifdef HD_PHASE_TIMING
__device__ unsigned long long g_phase[64];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define HD_STAMP(slot, cond) do { if ((cond) && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0) g_phase[slot] = gtimer(); } while (0)
#else
#define HD_STAMP(slot, cond) do { } while (0)
#endif

constexpr int TM = 128;                  // rows per CTA
constexpr int KC = 64;                   // K per stage
constexpr int A_KG = TM * 16 + 16;       // bytes between K-adjacent core matrices of the A stage image (+16: the
                                         // 8-byte producer stores of one row then spread over all banks)
constexpr int A_PART = (KC / 8) * A_KG;  // ~16 KB: hi (or lo) of one stage
constexpr int NPROD = 8;                 // producer / epilogue warps
constexpr int NTHREADS = 32 * (NPROD + 2);   // + MMA warp + weight-copy warp

// NT = output columns per CTA (64 or 128)
template <bool STRICT, int NT>
struct Smem {
  static constexpr int NP = STRICT ? 2 : 1;
  static constexpr int NSTG = (STRICT && NT > 64) ? 3 : 4;   // ring depth (227 KB shared memory per CTA)
  static constexpr int W_KG = NT * 16;            // bytes between K-adjacent core matrices of the weight image
  static constexpr int W_PART = (KC / 8) * W_KG;  // hi (or lo) weight chunk of one stage
  static constexpr int OT_LD = NT + 4;            // padded pitch (floats) of the output staging tile
  static constexpr int STAGE = (A_PART + W_PART) * NP;   // [A_hi][A_lo][W_hi][W_lo]; the ring is reused as the
                                                         // [TM][OT_LD] fp32 output staging tile after the last MMA
  static constexpr int OFF_BIAS = NSTG * STAGE;          // [NT] fp32
  static constexpr int OFF_BAR = OFF_BIAS + NT * 4;      // full[NSTG], empty[NSTG], acc_full
  static constexpr int NBAR = 2 * NSTG + 1;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
  static_assert(TM * OT_LD * 4 <= NSTG * STAGE, "output staging tile must fit in the operand ring");
};

struct Params {
  const float* X1;
  const float* X2;
  int ld1, K1, ld2, K2;   // K1, K2 multiples of KC
  const void* w_hi;       // bf16 images [n_out/64][K/8][64][8]
  const void* w_lo;
  const float* bias;      // [n_out] or null
  float* Y;
  int grid_rows;          // rows the launch covers (<= rows: the caller's bound on the real rows when ragged)
  int ldy, rows, mode;    // mode 0: store, 1: SiLU, 2: (resid + v) * node_mask, 3: store K-chunk-major
                          // Y[col/16][row][col%16] (the edge kernel's A|B operand layout); the stage-2 layer
                          // (hd_egcl.cu), rows = dense edges e = (b, i, j): 4: SiLU(v + s[row] * ws[col]),
                          // 5: v * edge_mask(row), 6: SiLU(v + s[row] * ws[col] + ab[b*N+i][col] + ab[b*N+j][H + col])
  const float* s;         // modes 4, 6: per-row scalar (|x_i - x_j|^2)
  const float* ws;        // modes 4, 6: its weight column [n_out]
  const float* ab;        // mode 6: [nodes][2 * n_out] per-node halves of the first message Linear
  const float* resid;
  const int32_t* sizes;
  int N;
  int B, ragged;          // ragged node rows (hd_api.cu): the first sum(sizes) rows are real, all of them unmasked;
                          // the grid covers the padded worst case and row tiles beyond the count retire at once
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <bool STRICT>
__device__ __forceinline__ float silu_node(float v) {
  if constexpr (STRICT) {   // ex2.approx / rcp.approx, ~2 ulp each (same evaluation as the edge kernel)
    return v * ptx::rcp_approx(1.0f + ptx::ex2_approx(-1.4426950408889634f * v));
  } else {
    const float hv = 0.5f * v;
    return fmaf(hv, ptx::tanh_approx(hv), hv);
  }
}

// Two independent GEMMs over the same row tiles may share one launch (column tiles [0, tiles_a) belong to problem a,
// the rest to problem b): node_mlp.2 and the next sub-layer's pre-projection run side by side that way.
struct Params2 {
  Params a, b, c;          // column tiles [0, tiles_a) -> a, [tiles_a, tiles_ab) -> b, the rest -> c
  int tiles_a, tiles_ab;
};

// EXT: the stage-2 epilogue modes 4-6 (hd_egcl.cu); the coarse-grained path instantiates the kernel without them
template <bool STRICT, int NT, bool RAGGED = false, bool EXT = false>
__global__ void __launch_bounds__(NTHREADS, 1) linear_tc_k(const Params2 pp) {
  using S = Smem<STRICT, NT>;
  const int which = (int)blockIdx.y < pp.tiles_a ? 0 : ((int)blockIdx.y < pp.tiles_ab ? 1 : 2);
  const Params p = which == 0 ? pp.a : (which == 1 ? pp.b : pp.c);
  constexpr int W_KG = S::W_KG, W_PART = S::W_PART, OT_LD = S::OT_LD, NSTG = S::NSTG;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t bar0 = sbase + S::OFF_BAR;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (NSTG + s); };
  const uint32_t bar_acc = bar0 + 8u * (2 * NSTG);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + S::OFF_TMEM);
  float* s_bias = reinterpret_cast<float*>(smem + S::OFF_BIAS);

  const int K = p.K1 + p.K2, nch = K / KC, nch1 = p.K1 / KC;
  const int row0 = blockIdx.x * TM;
  const int ct = (int)blockIdx.y - (which == 0 ? 0 : (which == 1 ? pp.tiles_a : pp.tiles_ab));
  HD_STAMP(0, tid == 0);
  pdl_trigger();   // the next kernel's CTAs may be scheduled as soon as resources free up (they wait for our completion)

  if (tid == 0) {
    for (int s = 0; s < NSTG; ++s) {
      ptx::mbar_init(bar_full(s), NPROD + 1);   // producer warps + the weight copy's expect_tx arrive
      ptx::mbar_init(bar_empty(s), 1);          // tcgen05.commit
    }
    ptx::mbar_init(bar_acc, 1);
    ptx::fence_mbar_init();
  }
  if (tid < NT) s_bias[tid] = p.bias ? p.bias[ct * NT + tid] : 0.f;
  int* s_cnt = nullptr;
  if constexpr (RAGGED) {   // `sizes` is an input of the forward (no kernel of the chain writes it): summed ahead
    __shared__ int s_cnt_buf[NTHREADS / 32];
    s_cnt = s_cnt_buf;
    int part = 0;           // of the dependency wait, like the rest of the set-up
    for (int k = tid; k < p.B; k += NTHREADS) part += __ldg(p.sizes + k);
    part = warp_sum_int(part);
    if (lane == 0) s_cnt[warp] = part;
  }
  if (warp == NPROD) ptx::tmem_alloc<1>(sbase + S::OFF_TMEM, NT);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;
  HD_STAMP(1, tid == 0);
  int live = p.rows;
  if constexpr (RAGGED) {
    live = 0;
#pragma unroll
    for (int w = 0; w < NTHREADS / 32; ++w) live += s_cnt[w];
  }

  if (row0 >= live) {
    // no real row in this tile
  } else if (warp < NPROD) {
    // =========================== producers ===========================
    // warp w converts rows [16w, 16w+16) of every K chunk; a warp-wide 16-byte load covers two full 256-byte row
    // segments (fully coalesced), lane -> (row parity, 4 consecutive k)
    const int l16 = lane & 15, rpar = lane >> 4;
    auto load = [&](float4 (&v)[8], int c) {
      const bool first = c < nch1;
      const float* base = first ? p.X1 : p.X2;
      const int ld = first ? p.ld1 : p.ld2;
      const int col = (first ? c : c - nch1) * KC + 4 * l16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row0 + 16 * warp + 2 * i + rpar;
        v[i] = row < live ? __ldg(reinterpret_cast<const float4*>(base + (int64_t)row * ld + col))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto convert_store = [&](const float4 (&v)[8], int s) {
      uint8_t* dst0 = smem + s * S::STAGE + (l16 >> 1) * A_KG + (16 * warp + rpar) * 16 + (l16 & 1) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = v[i];
        uint2 hi;
        hi.x = pack_bf16(a.x, a.y);
        hi.y = pack_bf16(a.z, a.w);
        *reinterpret_cast<uint2*>(dst0 + 32 * i) = hi;
        if constexpr (STRICT) {
          uint2 lo;
          lo.x = pack_bf16(a.x - __uint_as_float(hi.x << 16), a.y - __uint_as_float(hi.x & 0xffff0000u));
          lo.y = pack_bf16(a.z - __uint_as_float(hi.y << 16), a.w - __uint_as_float(hi.y & 0xffff0000u));
          *reinterpret_cast<uint2*>(dst0 + 32 * i + A_PART) = lo;
        }
      }
    };
    auto publish = [&](int s) {
      ptx::fence_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_relaxed(bar_full(s));
    };
    // set-up, the weight copies (constant data) and the bias run ahead of the previous kernel's completion; the
    // activations it produced are only read from here on, and every global write of this kernel comes later
    pdl_wait();
    // four chunks of loads in flight per thread (the whole ring)
    float4 v0[8], v1[8], v2[8], v3[8];
    load(v0, 0);
    if (nch > 1) load(v1, 1);
    if (nch > 2) load(v2, 2);
    if (nch > 3) load(v3, 3);
#pragma unroll 1
    for (int c = 0; c < nch; c += 4) {
      auto step = [&](float4 (&v)[8], int cc) {
        if (cc >= nch) return;
        const int s = cc % NSTG;
        ptx::mbar_wait(bar_empty(s), ((cc / NSTG) & 1) ^ 1);
        convert_store(v, s);
        if (cc + 4 < nch) load(v, cc + 4);
        publish(s);
      };
      step(v0, c);
      step(v1, c + 1);
      step(v2, c + 2);
      step(v3, c + 3);
    }
    HD_STAMP(6, tid == 0);
    // residual rows, fetched coalesced while the last MMAs run: pass q covers rows [RPP*q, RPP*q + RPP)
    constexpr int TPR = NT / 4, RPP = 32 * NPROD / TPR, NPASS = TM / RPP;   // threads per row, rows per pass
    const int orow = tid / TPR, oc4 = (tid % TPR) * 4;
    float4 rr[NPASS];
    if (p.mode == 2) {
#pragma unroll
      for (int q = 0; q < NPASS; ++q) {
        const int row = row0 + RPP * q + orow;
        rr[q] = row < live ? *reinterpret_cast<const float4*>(p.resid + (int64_t)row * p.ldy + ct * NT + oc4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    ptx::mbar_wait(bar_acc, 0);
    ptx::tc_fence_after();
    HD_STAMP(7, tid == 0);
    {
      // thread = accumulator row (TMEM lane), warp -> lane quarter (warp % 4) and column half (warp / 4)
      const int lq = warp & 3, ch = warp >> 2;
      float* trow = reinterpret_cast<float*>(smem) + (32 * lq + lane) * OT_LD + (NT / 2) * ch;
#pragma unroll 1
      for (int part = 0; part < NT / 64; ++part) {
        float v[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(32 * lq) << 16) + (NT / 2) * ch + 32 * part, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + (NT / 2) * ch + 32 * part + 4 * k4);
          float4 o = make_float4(v[4 * k4] + bb.x, v[4 * k4 + 1] + bb.y, v[4 * k4 + 2] + bb.z, v[4 * k4 + 3] + bb.w);
          if (p.mode == 1) {
            o.x = silu_node<STRICT>(o.x); o.y = silu_node<STRICT>(o.y);
            o.z = silu_node<STRICT>(o.z); o.w = silu_node<STRICT>(o.w);
          }
          *reinterpret_cast<float4*>(trow + 32 * part + 4 * k4) = o;
        }
      }
    }
    ptx::named_bar_sync(1, 32 * NPROD);
    HD_STAMP(8, tid == 0);
#pragma unroll
    for (int q = 0; q < NPASS; ++q) {
      const int row = row0 + RPP * q + orow;
      if (row >= live) continue;
      float4 o = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(smem) + (RPP * q + orow) * OT_LD + oc4);
      if (p.mode == 2) {
        if (RAGGED ? row < live : (row % p.N) < p.sizes[row / p.N]) {
          o.x += rr[q].x; o.y += rr[q].y; o.z += rr[q].z; o.w += rr[q].w;
        } else {
          o = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      const int col = ct * NT + oc4;
      if (EXT && p.mode >= 4) {
        if (p.mode == 5) {
          const int j = row % p.N, i = (row / p.N) % p.N, n = p.sizes[row / (p.N * p.N)];
          const float mk = (i < n && j < n && i != j) ? 1.f : 0.f;
          o.x *= mk; o.y *= mk; o.z *= mk; o.w *= mk;
        } else {
          const float sv = p.s[row];
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.ws + col));
          o.x = fmaf(sv, w4.x, o.x); o.y = fmaf(sv, w4.y, o.y); o.z = fmaf(sv, w4.z, o.z); o.w = fmaf(sv, w4.w, o.w);
          if (p.mode == 6) {
            const int64_t ri = row / p.N, rj = (int64_t)(row / (p.N * p.N)) * p.N + row % p.N;
            const int nt = p.ldy;   // message width H
            const float4 a4 = *reinterpret_cast<const float4*>(p.ab + ri * 2 * nt + col);
            const float4 b4 = *reinterpret_cast<const float4*>(p.ab + rj * 2 * nt + nt + col);
            o.x += a4.x + b4.x; o.y += a4.y + b4.y; o.z += a4.z + b4.z; o.w += a4.w + b4.w;
          }
          o.x = silu_node<STRICT>(o.x); o.y = silu_node<STRICT>(o.y);
          o.z = silu_node<STRICT>(o.z); o.w = silu_node<STRICT>(o.w);
        }
      }
      if (p.mode == 3) *reinterpret_cast<float4*>(p.Y + ((int64_t)(col >> 4) * p.rows + row) * 16 + (col & 15)) = o;
      else *reinterpret_cast<float4*>(p.Y + (int64_t)row * p.ldy + col) = o;
    }
  } else if (warp == NPROD) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t IDESC = ptx::idesc_bf16(TM, NT);
      for (int c = 0; c < nch; ++c) {
        const int s = c % NSTG;
        ptx::mbar_wait(bar_full(s), (c / NSTG) & 1);
        ptx::tc_fence_after();
        HD_STAMP(16 + c, true);
        const uint32_t a_hi = sbase + s * S::STAGE, a_lo = a_hi + A_PART;
        const uint32_t w_hi = a_hi + S::NP * A_PART, w_lo = w_hi + W_PART;
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks) {
          const uint32_t acc_on = (c | ks) ? 1u : 0u;
          const uint64_t da_hi = ptx::smem_desc(a_hi + ks * 2 * A_KG, A_KG, 128);
          const uint64_t db_hi = ptx::smem_desc(w_hi + ks * 2 * W_KG, W_KG, 128);
          ptx::mma_bf16<1>(tmem, da_hi, db_hi, IDESC, acc_on);
          if constexpr (STRICT) {
            const uint64_t da_lo = ptx::smem_desc(a_lo + ks * 2 * A_KG, A_KG, 128);
            const uint64_t db_lo = ptx::smem_desc(w_lo + ks * 2 * W_KG, W_KG, 128);
            ptx::mma_bf16<1>(tmem, da_hi, db_lo, IDESC, 1u);
            ptx::mma_bf16<1>(tmem, da_lo, db_hi, IDESC, 1u);
          }
        }
        ptx::mma_commit<1>(bar_empty(s));
      }
      ptx::mma_commit<1>(bar_acc);
    }
    __syncwarp();
  } else {
    // =========================== weight copies ===========================
    if (lane == 0) {
      const uint8_t* img_hi = reinterpret_cast<const uint8_t*>(p.w_hi) + (size_t)ct * (K / 8) * W_KG;
      const uint8_t* img_lo = reinterpret_cast<const uint8_t*>(p.w_lo) + (size_t)ct * (K / 8) * W_KG;
      for (int c = 0; c < nch; ++c) {
        const int s = c % NSTG;
        ptx::mbar_wait(bar_empty(s), ((c / NSTG) & 1) ^ 1);
        ptx::mbar_expect_tx(bar_full(s), S::NP * W_PART);
        const uint32_t dst = sbase + s * S::STAGE + S::NP * A_PART;
        ptx::bulk_g2s(dst, img_hi + (size_t)c * W_PART, W_PART, bar_full(s));
        if constexpr (STRICT) ptx::bulk_g2s(dst + W_PART, img_lo + (size_t)c * W_PART, W_PART, bar_full(s));
      }
    }
    __syncwarp();
  }

  HD_STAMP(9, tid == 0);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == NPROD) ptx::tmem_dealloc<1>(tmem, NT);
  HD_STAMP(10, tid == 32 * NPROD);
}

template <bool STRICT, int NT, bool RAGGED = false, bool EXT = false>
static int launch2(const Params& a, int n_out_a, const Params* b, int n_out_b, cudaStream_t st,
                   const Params* c3 = nullptr, int n_out_c = 0) {
  if constexpr (!RAGGED && !EXT) {
    if (a.ragged) return launch2<STRICT, NT, true>(a, n_out_a, b, n_out_b, st, c3, n_out_c);
  }
  using S = Smem<STRICT, NT>;
  static bool configured = false;
  auto kern = linear_tc_k<STRICT, NT, RAGGED, EXT>;
  if (!configured) {
    HD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  Params2 pp{};
  pp.a = a;
  pp.b = b ? *b : a;
  pp.c = c3 ? *c3 : a;
  pp.tiles_a = n_out_a / NT;
  pp.tiles_ab = pp.tiles_a + (b ? n_out_b / NT : 0);
  dim3 grid((a.grid_rows + TM - 1) / TM, pp.tiles_ab + (c3 ? n_out_c / NT : 0));
  HD_CHECK_CUDA(launch_pdl(kern, grid, dim3(NTHREADS), S::TOTAL, st, pp));
  count_launch();
  return HD_OK;
}
template <bool STRICT, int NT>
static int launch(const Params& p, int n_out, cudaStream_t st) {
  return launch2<STRICT, NT>(p, n_out, nullptr, 0, st);
}

}  // namespace lin

static lin::Params make_params(const FwdCtx& c, const float* X1, int ld1, int K1, const float* X2, int ld2, int K2,
                               const void* w_hi, const void* w_lo, const float* bias, float* Y, int ldy, int mode,
                               const float* resid) {
  lin::Params p{};
  p.X1 = X1;
  p.X2 = X2 ? X2 : X1;
  p.ld1 = ld1;
  p.K1 = K1;
  p.ld2 = ld2;
  p.K2 = K2;
  p.w_hi = w_hi;
  p.w_lo = w_lo;
  p.bias = bias;
  p.Y = Y;
  p.ldy = ldy;
  p.rows = c.B * c.N;
  p.grid_rows = c.node_off && c.rows_bound > 0 ? c.rows_bound : p.rows;
  p.mode = mode;
  p.resid = resid ? resid : Y;
  p.sizes = c.sizes;
  p.N = c.N;
  p.B = c.B;
  p.ragged = c.node_off != nullptr;
  return p;
}

// Y = epilogue([X1 | X2] W^T + bias) on the tensor cores; W given as its bf16 hi/lo images in `tile_n`-row output
// tiles (hd_layout.cu): tile_n = 128 for the 512-wide pre-projection, 64 for the 256-wide node_mlp layers, so that
// every launch is a single wave of CTAs at the sampling path's batch sizes
int linear_tc(const FwdCtx& c, const float* X1, int ld1, int K1, const float* X2, int ld2, int K2, const void* w_hi,
              const void* w_lo, int n_out, int tile_n, const float* bias, float* Y, int ldy, int mode,
              const float* resid, bool strict) {
  if (K1 % lin::KC || K2 % lin::KC || (tile_n != 64 && tile_n != 128) || n_out % tile_n || (ld1 & 3) || (ld2 & 3) ||
      (ldy & 3)) {
    set_error("linear_tc: unsupported shape K1=%d K2=%d n_out=%d tile_n=%d", K1, K2, n_out, tile_n);
    return HD_E_INVALID;
  }
#ifdef HD_EXP_SKIP_NODE   // timing experiment only (scripts/step_ablation.sh): what the node launches cost inside the chain
  return HD_OK;
#endif
  const lin::Params p = make_params(c, X1, ld1, K1, X2, ld2, K2, w_hi, w_lo, bias, Y, ldy, mode, resid);
  if (tile_n == 128) return strict ? lin::launch<true, 128>(p, n_out, c.stream) : lin::launch<false, 128>(p, n_out, c.stream);
  return strict ? lin::launch<true, 64>(p, n_out, c.stream) : lin::launch<false, 64>(p, n_out, c.stream);
}

// The same kernel for the stage-2 layer (hd_egcl.cu): `rows` of any count (edges or nodes), 128-column output tiles,
// the epilogue modes 0-2 and 4-6 (Params).
int linear_tc_rows(cudaStream_t st, int rows, const float* X1, int ld1, int K1, const float* X2, int ld2, int K2,
                   const void* w_hi, const void* w_lo, int n_out, const float* bias, float* Y, int ldy, int mode,
                   const float* resid, const float* s, const float* ws, const float* ab, const int32_t* sizes, int N,
                   bool strict) {
  if (K1 % lin::KC || K2 % lin::KC || n_out % 128 || (ld1 & 3) || (ld2 & 3) || (ldy & 3) || rows < 1) {
    set_error("linear_tc_rows: unsupported shape rows=%d K1=%d K2=%d n_out=%d", rows, K1, K2, n_out);
    return HD_E_INVALID;
  }
  lin::Params p{};
  p.X1 = X1; p.X2 = X2 ? X2 : X1; p.ld1 = ld1; p.K1 = K1; p.ld2 = ld2; p.K2 = K2;
  p.w_hi = w_hi; p.w_lo = w_lo; p.bias = bias; p.Y = Y; p.ldy = ldy; p.rows = rows; p.grid_rows = rows; p.mode = mode;
  p.resid = resid ? resid : Y; p.sizes = sizes; p.N = N; p.B = 0; p.ragged = 0;
  p.s = s; p.ws = ws; p.ab = ab;
  return strict ? lin::launch2<true, 128, false, true>(p, n_out, nullptr, 0, st)
                : lin::launch2<false, 128, false, true>(p, n_out, nullptr, 0, st);
}

// node_mlp.2 (+ residual, mask) -> h_out and, in the same launch, the next sub-layer's A|B pre-projection computed
// from [h | hid] with the pre-multiplied weight image (hd_layout.cu "fused pre-projection"): both in 128-column tiles
// m2_* / bm2 / ab2 (optional): a second pre-projection from the same [h | hid], for the first sub-layer of the next block
int linear_tc_v2_and_preproject(const FwdCtx& c, const float* hid, const float* h, float* h_out, const void* v2_hi,
                                const void* v2_lo, const float* c2, const void* m_hi, const void* m_lo,
                                const float* bm, float* ab, bool strict, const void* m2_hi, const void* m2_lo,
                                const float* bm2, float* ab2) {
#ifdef HD_EXP_SKIP_NODE
  return HD_OK;
#endif
  const lin::Params a = make_params(c, hid, H, H, nullptr, 0, 0, v2_hi, v2_lo, c2, h_out, H, 2, h);
  const lin::Params b = make_params(c, h, H, H, hid, H, H, m_hi, m_lo, bm, ab, 2 * H, 3, nullptr);
  if (m2_hi) {
    const lin::Params c3 = make_params(c, h, H, H, hid, H, H, m2_hi, m2_lo, bm2, ab2, 2 * H, 3, nullptr);
    return strict ? lin::launch2<true, 128>(a, H, &b, 2 * H, c.stream, &c3, 2 * H)
                  : lin::launch2<false, 128>(a, H, &b, 2 * H, c.stream, &c3, 2 * H);
  }
  return strict ? lin::launch2<true, 128>(a, H, &b, 2 * H, c.stream) : lin::launch2<false, 128>(a, H, &b, 2 * H, c.stream);
}

}  // namespace hd
