// hd_node.cu - tensor-core node GEMMs of the tensor-core engines: every per-node nn.Linear of a sub-layer
// (the A|B pre-projection of edge_mlp.0 / coord_mlp.0, node_mlp.0, node_mlp.2; egnn_new.py:52-62 and SURVEY.md 7
// "layer-1 split") as   Y[r, o] = epilogue(bias[o] + sum_k [X1 | X2][r, k] * W[o, k]).
//
//   tile     : 128 node rows x 64 output columns per CTA, K streamed in chunks of 64 through a 4-stage ring
//   producer : 4 warps read the fp32 activations (L2-resident, written by the previous kernel), split them into
//              bf16 hi (+ lo in strict mode) and store them as canonical K-major core matrices (the MMA A operand)
//   weights  : 1 thread bulk-copies (TMA 1-D) the pre-built bf16 hi/lo image chunk of W for this column tile
//   MMA      : 1 thread issues tcgen05.mma kind::f16 M=128 N=64 K=16; strict = hi*hi + hi*lo + lo*hi into one fp32
//              TMEM accumulator (64 columns)
//   epilogue : the 4 producer warps read the accumulator (thread = row): + bias, then store / SiLU /
//              (resid + v) * node_mask
//
// Algorithmic HBM bytes: rows*(K + n_out)*4 activations + n_out*K*2(*2) weights; everything is L2-resident at the
// sizes of the sampling path (h is 2.6 MB at B=64, N=40), so the kernel is bound by L2->SM latency/bandwidth.
#include <cuda_bf16.h>

#include "hd_common.cuh"
#include "hd_ptx.cuh"

namespace hd {
namespace lin {

constexpr int TM = 128;                  // rows per CTA
constexpr int NT = 64;                   // output columns per CTA
constexpr int KC = 64;                   // K per stage
constexpr int NSTG = 4;
constexpr int A_KG = TM * 16;            // bytes between K-adjacent core matrices of the A stage image
constexpr int A_PART = (KC / 8) * A_KG;  // 16 KB: hi (or lo) of one stage
constexpr int W_KG = NT * 16;            // same for the weight image
constexpr int W_PART = (KC / 8) * W_KG;  // 8 KB
constexpr int NTHREADS = 192;            // 4 producer/epilogue warps + MMA warp + weight-copy warp

template <bool STRICT>
struct Smem {
  static constexpr int NP = STRICT ? 2 : 1;
  static constexpr int STAGE = (A_PART + W_PART) * NP;   // [A_hi][A_lo][W_hi][W_lo]
  static constexpr int OFF_BAR = NSTG * STAGE;           // full[NSTG], empty[NSTG], acc_full
  static constexpr int NBAR = 2 * NSTG + 1;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
};

struct Params {
  const float* X1;
  const float* X2;
  int ld1, K1, ld2, K2;   // K1, K2 multiples of KC
  const void* w_hi;       // bf16 images [n_out/64][K/8][64][8]
  const void* w_lo;
  const float* bias;      // [n_out] or null
  float* Y;
  int ldy, rows, mode;    // mode 0: store, 1: SiLU, 2: (resid + v) * node_mask
  const float* resid;
  const int32_t* sizes;
  int N;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <bool STRICT>
__global__ void __launch_bounds__(NTHREADS, 1) linear_tc_k(const Params p) {
  using S = Smem<STRICT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t bar0 = sbase + S::OFF_BAR;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (NSTG + s); };
  const uint32_t bar_acc = bar0 + 8u * (2 * NSTG);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + S::OFF_TMEM);

  const int K = p.K1 + p.K2, nch = K / KC, nch1 = p.K1 / KC;
  const int row0 = blockIdx.x * TM, ct = blockIdx.y;

  if (tid == 0) {
    for (int s = 0; s < NSTG; ++s) {
      ptx::mbar_init(bar_full(s), 4 + 1);   // 4 producer warps + the weight copy's expect_tx arrive
      ptx::mbar_init(bar_empty(s), 1);      // tcgen05.commit
    }
    ptx::mbar_init(bar_acc, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) ptx::tmem_alloc<1>(sbase + S::OFF_TMEM, NT);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < 4) {
    // =========================== producers ===========================
    const int rs = lane & 7, kq = lane >> 3;
    auto load = [&](float4 (&v)[16], int c) {
      const bool first = c < nch1;
      const float* base = first ? p.X1 : p.X2;
      const int ld = first ? p.ld1 : p.ld2;
      const int col = (first ? c : c - nch1) * KC;
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) {
        const int row = row0 + 32 * warp + 8 * rg + rs;
        const float* q = base + (int64_t)row * ld + col + 8 * kq;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          if (row < p.rows) {
            v[(rg * 2 + kh) * 2] = __ldg(reinterpret_cast<const float4*>(q + 32 * kh));
            v[(rg * 2 + kh) * 2 + 1] = __ldg(reinterpret_cast<const float4*>(q + 32 * kh + 4));
          } else {
            v[(rg * 2 + kh) * 2] = v[(rg * 2 + kh) * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    };
    auto convert_store = [&](const float4 (&v)[16], int s) {
      uint8_t* stage = smem + s * S::STAGE;
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) {
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          const float4 a = v[(rg * 2 + kh) * 2], b = v[(rg * 2 + kh) * 2 + 1];
          uint4 hi;
          hi.x = pack_bf16(a.x, a.y);
          hi.y = pack_bf16(a.z, a.w);
          hi.z = pack_bf16(b.x, b.y);
          hi.w = pack_bf16(b.z, b.w);
          uint8_t* dst = stage + (4 * kh + kq) * A_KG + (32 * warp + 8 * rg + rs) * 16;
          *reinterpret_cast<uint4*>(dst) = hi;
          if constexpr (STRICT) {
            uint4 lo;
            lo.x = pack_bf16(a.x - __uint_as_float(hi.x << 16), a.y - __uint_as_float(hi.x & 0xffff0000u));
            lo.y = pack_bf16(a.z - __uint_as_float(hi.y << 16), a.w - __uint_as_float(hi.y & 0xffff0000u));
            lo.z = pack_bf16(b.x - __uint_as_float(hi.z << 16), b.y - __uint_as_float(hi.z & 0xffff0000u));
            lo.w = pack_bf16(b.z - __uint_as_float(hi.w << 16), b.w - __uint_as_float(hi.w & 0xffff0000u));
            *reinterpret_cast<uint4*>(dst + A_PART) = lo;
          }
        }
      }
    };
    auto publish = [&](int s) {
      ptx::fence_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_full(s));
    };
    float4 va[16], vb[16];
    load(va, 0);
    if (nch > 1) load(vb, 1);
#pragma unroll 1
    for (int c = 0; c < nch; c += 2) {   // nch is even (K1, K2 multiples of 128) or 1.. handled by guards
      {
        const int s = c % NSTG;
        ptx::mbar_wait(bar_empty(s), ((c / NSTG) & 1) ^ 1);
        convert_store(va, s);
        if (c + 2 < nch) load(va, c + 2);
        publish(s);
      }
      if (c + 1 < nch) {
        const int s = (c + 1) % NSTG;
        ptx::mbar_wait(bar_empty(s), (((c + 1) / NSTG) & 1) ^ 1);
        convert_store(vb, s);
        if (c + 3 < nch) load(vb, c + 3);
        publish(s);
      }
    }
    // =========================== epilogue ===========================
    ptx::mbar_wait(bar_acc, 0);
    ptx::tc_fence_after();
    const int row = row0 + 32 * warp + lane;
    const bool in_rows = row < p.rows;
    bool real = true;
    if (p.mode == 2 && in_rows) real = (row % p.N) < p.sizes[row / p.N];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
    for (int half = 0; half < NT / 32; ++half) {
      float v[32];
      ptx::tmem_ld32(taddr + 32 * half, v);
      ptx::tmem_wait_ld();
      const int col = ct * NT + 32 * half;
      float* yrow = p.Y + (int64_t)row * p.ldy + col;
      const float* rrow = p.resid + (int64_t)row * p.ldy + col;
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        float4 o = make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
        if (p.bias) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4 * k4));
          o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        }
        if (p.mode == 1) {
          o.x = silu_acc(o.x); o.y = silu_acc(o.y); o.z = silu_acc(o.z); o.w = silu_acc(o.w);
        } else if (p.mode == 2) {
          if (in_rows && real) {
            const float4 rr = *reinterpret_cast<const float4*>(rrow + 4 * k4);
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          } else {
            o = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (in_rows) *reinterpret_cast<float4*>(yrow + 4 * k4) = o;
      }
    }
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t IDESC = ptx::idesc_bf16(TM, NT);
      for (int c = 0; c < nch; ++c) {
        const int s = c % NSTG;
        ptx::mbar_wait(bar_full(s), (c / NSTG) & 1);
        ptx::tc_fence_after();
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

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc<1>(tmem, NT);
}

template <bool STRICT>
static int launch(const Params& p, int n_out, cudaStream_t st) {
  using S = Smem<STRICT>;
  static bool configured = false;
  auto kern = linear_tc_k<STRICT>;
  if (!configured) {
    HD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  dim3 grid((p.rows + TM - 1) / TM, n_out / NT);
  kern<<<grid, NTHREADS, S::TOTAL, st>>>(p);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

}  // namespace lin

// Y = epilogue([X1 | X2] W^T + bias) on the tensor cores; W given as its bf16 hi/lo tile images (hd_layout.cu)
int linear_tc(const FwdCtx& c, const float* X1, int ld1, int K1, const float* X2, int ld2, int K2, const void* w_hi,
              const void* w_lo, int n_out, const float* bias, float* Y, int ldy, int mode, const float* resid,
              bool strict) {
  if (K1 % lin::KC || K2 % lin::KC || n_out % lin::NT || (ld1 & 3) || (ld2 & 3) || (ldy & 3)) {
    set_error("linear_tc: unsupported shape K1=%d K2=%d n_out=%d", K1, K2, n_out);
    return HD_E_INVALID;
  }
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
  p.mode = mode;
  p.resid = resid ? resid : Y;
  p.sizes = c.sizes;
  p.N = c.N;
  return strict ? lin::launch<true>(p, n_out, c.stream) : lin::launch<false>(p, n_out, c.stream);
}

}  // namespace hd
