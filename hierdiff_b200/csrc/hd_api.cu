// hd_api.cu - the C ABI (include/hierdiff_b200.h): argument checking, orchestration of one
// EGNN forward, and the small per-node / per-molecule kernels around the edge kernels
// (embedding, output head, velocity + centre-of-gravity projection, diffusion update).
#include <cmath>
#include <cstring>

#include "hd_common.cuh"

namespace hd {
const char* last_error();
int64_t launch_count();
int pack_weights(const hd_config& c, const Layout& L, const float* w, char* P, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// en_dynamics.py:56-74: xh*node_mask, split, append time.  One thread per (row, channel).
// ---------------------------------------------------------------------------------------
__global__ void prep_k(const float* __restrict__ z, const float* __restrict__ t, const float* __restrict__ context,
                       int C, const int32_t* __restrict__ sizes, int B, int N, int F, float* __restrict__ hin,
                       float* __restrict__ x, float* __restrict__ x0, float* __restrict__ x2,
                       int32_t* __restrict__ nanflag) {
  pdl_wait();
  pdl_trigger();
  const int D = 3 + F, Fi = F + 1 + C, W = D + 1 + C;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) *nanflag = 0;
  if (idx >= (int64_t)B * N * W) return;
  const int64_t r = idx / W;
  const int ch = (int)(idx % W);
  const int b = (int)(r / N), i = (int)(r % N);
  const float mk = i < sizes[b] ? 1.f : 0.f;
  if (ch < 3) {
    const float v = z[r * D + ch] * mk;
    x[r * 3 + ch] = v;
    x0[r * 3 + ch] = v;
    x2[r * 3 + ch] = 0.f;   // ping-pong buffer of the coordinate updates: padded rows stay 0
  } else if (ch < D) {
    hin[r * Fi + (ch - 3)] = z[r * D + ch] * mk;
  } else if (ch == D) {
    hin[r * Fi + F] = t[b];  // time channel is not masked (en_dynamics.py:66-74)
  } else {
    hin[r * Fi + F + (ch - D)] = context[r * C + (ch - D - 1)];  // nor is the context (en_dynamics.py:76-79)
  }
}

// egnn_new.py:197 h = embedding(h); rows of padded nodes are written as 0 (they never reach a real node:
// every path out of them is cut by edge_mask and each layer ends in *node_mask)
__global__ void __launch_bounds__(256) embed_k(const float* __restrict__ hin, int Fi, const float* __restrict__ wT,
                                               const float* __restrict__ bias, const int32_t* __restrict__ sizes,
                                               int N, float* __restrict__ h) {
  pdl_wait();
  pdl_trigger();
  const int64_t r = blockIdx.x;
  const int c = threadIdx.x, b = (int)(r / N), i = (int)(r % N);
  float v = 0.f;
  if (i < sizes[b]) {
    v = bias[c];
    for (int f = 0; f < Fi; ++f) v = fmaf(hin[r * Fi + f], wT[f * H + c], v);
  }
  h[r * H + c] = v;
}

// egnn_new.py:202-204 h = embedding_out(h) * node_mask.  One warp per output feature.
__global__ void __launch_bounds__(256) out_k(const float* __restrict__ h, const float* __restrict__ w,
                                             const float* __restrict__ bias, int Fi,
                                             const int32_t* __restrict__ sizes, int N, float* __restrict__ hout) {
  pdl_wait();
  pdl_trigger();
  const int64_t r = blockIdx.x;
  const int b = (int)(r / N), i = (int)(r % N), warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool real = i < sizes[b];
  for (int f = warp; f < Fi; f += 8) {
    float s = 0.f;
    if (real)
      for (int c = lane; c < H; c += 32) s = fmaf(h[r * H + c], w[f * H + c], s);
    s = warp_sum(s);
    if (lane == 0) hout[r * Fi + f] = real ? s + bias[f] : 0.f;
  }
}

// prep_k + embed_k in one launch (the sampling path): one CTA per node row.  Block 0 also resets the NaN flag and,
// for the tensor-core engines, builds the edge-row prefix row_off[b+1] = sum_{b' <= b} edge_rows(n_b') that
// tc::plan_k would otherwise compute in a launch of its own (row_off may be null).
__global__ void __launch_bounds__(256) prep_embed_k(const float* __restrict__ z, const float* __restrict__ t,
                                                    const float* __restrict__ context, int C,
                                                    const int32_t* __restrict__ sizes, int B, int N, int F,
                                                    const float* __restrict__ wT, const float* __restrict__ bias,
                                                    float* __restrict__ x, float* __restrict__ x0,
                                                    float* __restrict__ x2, float* __restrict__ h,
                                                    int32_t* __restrict__ nanflag, int32_t* __restrict__ row_off,
                                                    int32_t* __restrict__ node_off, int rows_bound,
                                                    int32_t* __restrict__ flags) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_in[64];
  __shared__ int s_part[8];
  const int D = 3 + F, Fi = F + 1 + C;
  const int64_t r = blockIdx.x;
  const int tid = threadIdx.x, b = (int)(r / N), i = (int)(r % N);
  const bool real = i < sizes[b];
  const float mk = real ? 1.f : 0.f;
  // ragged node rows (node_off != null): node (b,i) lives in row sum_{b' < b} n_b' + i of x / x0 / x2 / h, padded
  // nodes have no row.  Every block sums the sizes ahead of its molecule itself; block 0 also publishes the prefix
  // table for the kernels that follow.
  int64_t o = r;
  if (node_off) {
    if (!real && r != 0) return;   // a padded node owns no row (block 0 stays: it publishes the prefix tables)
    int part = 0;
    for (int k = tid; k < b; k += 256) part += sizes[k];
    part = warp_sum_int(part);
    if ((tid & 31) == 0) s_part[tid >> 5] = part;
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) base += s_part[w];
    o = base + i;
  }
  if (r == 0 && tid < 32) {
    if (tid == 0) *nanflag = 0;
    if (row_off) {
      int carry = 0;
      if (tid == 0) row_off[0] = 0;
      for (int base = 0; base < B; base += 32) {
        const int bb = base + tid;
        const int n = bb < B ? sizes[bb] : 0;
        int v = edge_rows(n);
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, v, o2);
          if (tid >= o2) v += u;
        }
        if (bb < B) row_off[bb + 1] = carry + v;
        carry += __shfl_sync(0xffffffffu, v, 31);
      }
    }
    if (node_off) {
      int carry = 0;
      if (tid == 0) node_off[0] = 0;
      for (int base = 0; base < B; base += 32) {
        const int bb = base + tid;
        int v = bb < B ? sizes[bb] : 0;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, v, o2);
          if (tid >= o2) v += u;
        }
        if (bb < B) node_off[bb + 1] = carry + v;
        carry += __shfl_sync(0xffffffffu, v, 31);
      }
      // the caller's bound on sum(sizes) sized the node-GEMM grids: rows beyond it would silently go missing
      if (tid == 0 && carry > rows_bound && flags) atomicOr(flags, HD_FLAG_MASK);
    }
  }
  const bool wr = real || !node_off;   // ragged: padded nodes own no row
  if (tid < 3) {
    const float v = z[r * D + tid] * mk;
    if (wr) {
      x[o * 3 + tid] = v;
      x0[o * 3 + tid] = v;
      x2[o * 3 + tid] = 0.f;   // ping-pong buffer of the coordinate updates: padded rows stay 0
    }
  } else if (tid < D) {
    s_in[tid - 3] = z[r * D + tid] * mk;
  } else if (tid == D) {
    s_in[F] = t[b];          // time and context are not masked (en_dynamics.py:66-79)
  } else if (tid < D + 1 + C) {
    s_in[F + (tid - D)] = context[r * C + (tid - D - 1)];
  }
  __syncthreads();
  float v = 0.f;
  if (real) {
    v = bias[tid];
    for (int f = 0; f < Fi; ++f) v = fmaf(s_in[f], wT[f * H + tid], v);
  }
  if (wr) h[o * H + tid] = v;
}

// out_k + vel_k in one launch (the sampling path): one CTA per node row
__global__ void __launch_bounds__(256) out_vel_k(const float* __restrict__ h, const float* __restrict__ w,
                                                 const float* __restrict__ bias, int Fi, int F,
                                                 const int32_t* __restrict__ sizes, int N,
                                                 const float* __restrict__ xf, const float* __restrict__ x0,
                                                 float* __restrict__ eps_raw, int32_t* __restrict__ nanflag,
                                                 const int32_t* __restrict__ node_off, int32_t* __restrict__ state) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_out[64];
  const int64_t r = blockIdx.x;
  if (state) {
    // sampling loop: this forward belongs to step k = state[0] (no CTA of this kernel writes it); its NaN flag is
    // state[2 + (k & 1)].  Block 0 latches k for the tail kernel and clears the flag of step k+1 - last read by the
    // tail of step k-1, which has completed.
    const int k = state[0];
    nanflag = state + 2 + (k & 1);
    if (r == 0 && threadIdx.x == 0) {
      state[1] = k;
      state[2 + ((k + 1) & 1)] = 0;
    }
  }
  const int tid = threadIdx.x, b = (int)(r / N), i = (int)(r % N), warp = tid >> 5, lane = tid & 31;
  const bool real = i < sizes[b];
  const int64_t o = node_off ? (real ? node_off[b] + i : 0) : r;   // row of h / xf / x0 (ragged: real nodes only)
  for (int f = warp; f < Fi; f += 8) {
    float s = 0.f;
    if (real)
      for (int c = lane; c < H; c += 32) s = fmaf(h[o * H + c], w[f * H + c], s);
    s = warp_sum(s);
    if (lane == 0) s_out[f] = real ? s + bias[f] : 0.f;
  }
  __syncthreads();
  const int D = 3 + F;   // the first F of the Fi output channels: time and context are sliced off (en_dynamics.py:99-105)
  if (tid < D) {
    float v;
    if (tid < 3) {
      v = real ? xf[o * 3 + tid] - x0[o * 3 + tid] : 0.f;
      if (isnan(v)) atomicOr(nanflag, 1);
    } else {
      v = s_out[tid - 3];
    }
    eps_raw[r * D + tid] = v;
  }
}

// en_dynamics.py:89,103-111: vel = (x_final - x)*mask, drop the time channel, NaN detection.
__global__ void vel_k(const float* __restrict__ xf, const float* __restrict__ x0, const float* __restrict__ hout,
                      const int32_t* __restrict__ sizes, int B, int N, int F, int Fi, float* __restrict__ eps_raw,
                      int32_t* __restrict__ nanflag) {
  pdl_wait();
  pdl_trigger();
  const int D = 3 + F;   // the first F of the Fi output channels: time and context are sliced off (en_dynamics.py:99-105)
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * N * D) return;
  const int64_t r = idx / D;
  const int ch = (int)(idx % D), b = (int)(r / N), i = (int)(r % N);
  float v;
  if (ch < 3) {
    v = (xf[r * 3 + ch] - x0[r * 3 + ch]) * (i < sizes[b] ? 1.f : 0.f);
    if (isnan(v)) atomicOr(nanflag, 1);
  } else {
    v = hout[r * Fi + (ch - 3)];
  }
  eps_raw[idx] = v;
}

// masked mean over the nodes of molecule b for the first 3 channels of v [N, D] (models/utils.py:53-56):
// sum over ALL N rows / n_b.  Called by all threads of a block (>= 96 threads); result valid in all threads.
__device__ __forceinline__ void block_mean3(const float* v, int N, int D, int n, float* s_mean, float out[3]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 3) {
    float s = 0.f;
    for (int i = lane; i < N; i += 32) s += v[i * D + warp];
    s = warp_sum(s);
    if (lane == 0) s_mean[warp] = s / (float)n;
  }
  __syncthreads();
  out[0] = s_mean[0];
  out[1] = s_mean[1];
  out[2] = s_mean[2];
  __syncthreads();
}

// en_dynamics.py:109-116: NaN guard (whole batch) + remove_mean_with_mask; one block per molecule.
// raw != 0 (HD_ENGINE_RAW_VELOCITY): neither step is applied - the caller combines several node sets first (pocket
// conditioning) - but a NaN is still reported through `flags`.
__global__ void __launch_bounds__(128) cog_k(const float* __restrict__ eps_raw, const int32_t* __restrict__ sizes,
                                             int N, int F, const int32_t* __restrict__ nanflag,
                                             float* __restrict__ eps, int32_t* __restrict__ flags, int raw) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sv[];  // [N][D]
  __shared__ float s_mean[3];
  const int D = 3 + F, b = blockIdx.x, n = sizes[b];
  const bool nan = *nanflag != 0;
  const float* src = eps_raw + (int64_t)b * N * D;
  if (raw) {
    for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) eps[(int64_t)b * N * D + idx] = src[idx];
    if (nan && flags && b == 0 && threadIdx.x == 0) atomicOr(flags, HD_FLAG_NAN);
    return;
  }
  for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) {
    float v = src[idx];
    if (nan && idx % D < 3) v = 0.f;
    sv[idx] = v;
  }
  __syncthreads();
  float mean[3];
  block_mean3(sv, N, D, n, s_mean, mean);
  float* dst = eps + (int64_t)b * N * D;
  for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) {
    const int i = idx / D, ch = idx % D;
    float v = sv[idx];
    if (ch < 3 && i < n) v -= mean[ch];
    dst[idx] = v;
  }
  if (nan && flags && b == 0 && threadIdx.x == 0) atomicOr(flags, HD_FLAG_NAN);
}

// diffusion_qm9.py:445-456 with the two raw randn draws; one block per molecule.
__device__ __forceinline__ void load_noise(const float* rx, const float* rh, int N, int F, int n, float* nz,
                                           float* s_mean) {
  const int D = 3 + F;
  for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) {
    const int i = idx / D, ch = idx % D;
    const float mk = i < n ? 1.f : 0.f;
    nz[idx] = (ch < 3 ? rx[i * 3 + ch] : rh[i * F + (ch - 3)]) * mk;
  }
  __syncthreads();
  float mean[3];
  block_mean3(nz, N, D, n, s_mean, mean);
  for (int idx = threadIdx.x; idx < N * 3; idx += blockDim.x) {
    const int i = idx / 3, ch = idx % 3;
    if (i < n) nz[i * D + ch] -= mean[ch];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(128) combine_noise_k(const float* __restrict__ rx, const float* __restrict__ rh,
                                                       const int32_t* __restrict__ sizes, int N, int F,
                                                       float* __restrict__ z) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float nz[];
  __shared__ float s_mean[3];
  const int D = 3 + F, b = blockIdx.x;
  load_noise(rx + (int64_t)b * N * 3, rh + (int64_t)b * N * F, N, F, sizes[b], nz, s_mean);
  for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) z[(int64_t)b * N * D + idx] = nz[idx];
}

// diffusion_qm9.py:328-345; one block per molecule
__global__ void __launch_bounds__(128) reverse_step_k(const float* __restrict__ zt, const float* __restrict__ eps,
                                                      const float* __restrict__ rx, const float* __restrict__ rh,
                                                      const int32_t* __restrict__ sizes, int N, int F,
                                                      const float* __restrict__ sched, int sched_per_mol,
                                                      float* __restrict__ zs, int32_t* __restrict__ flags) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];  // nz [N*D], ev [N*D], zv [N*D]
  __shared__ float s_mean[3];
  __shared__ float s_chk[3];
  const int D = 3 + F, b = blockIdx.x, n = sizes[b];
  float* nz = sm;
  float* ev = sm + N * D;
  float* zv = ev + N * D;
  const float* sc = sched + (sched_per_mol ? 3 * b : 0);
  const float alpha = sc[0], ceps = sc[1], sigma = sc[2];
  const float* zsrc = zt + (int64_t)b * N * D;
  const float* esrc = eps + (int64_t)b * N * D;
  if (threadIdx.x < 3) s_chk[threadIdx.x] = 0.f;
  load_noise(rx + (int64_t)b * N * 3, rh + (int64_t)b * N * F, N, F, n, nz, s_mean);
  // assert_mean_zero_with_mask(zt_x) (models/utils.py:65-75), evaluated per molecule
  float pad_abs = 0.f, max_abs = 0.f;
  for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) {
    const float v = zsrc[idx];
    zv[idx] = v;
    ev[idx] = esrc[idx];
    if (idx % D < 3) {
      if (idx / D >= n) pad_abs = fmaxf(pad_abs, fabsf(v));
      max_abs = fmaxf(max_abs, fabsf(v));
    }
  }
  // float max over non-negative values == max over their int bit patterns
  atomicMax(reinterpret_cast<int*>(&s_chk[0]), __float_as_int(pad_abs));
  atomicMax(reinterpret_cast<int*>(&s_chk[1]), __float_as_int(max_abs));
  __syncthreads();
  float zsum[3];
  block_mean3(zv, N, D, 1, s_mean, zsum);  // n=1 -> plain sums
  if (flags && threadIdx.x == 0) {
    const float err = fmaxf(fabsf(zsum[0]), fmaxf(fabsf(zsum[1]), fabsf(zsum[2])));
    int f = 0;
    if (!(s_chk[0] < 1e-4f)) f |= HD_FLAG_MASK;
    if (!(err / (s_chk[1] + 1e-10f) < 1e-2f)) f |= HD_FLAG_COG;
    if (f) atomicOr(flags, f);
  }
  float mean[3];
  block_mean3(ev, N, D, n, s_mean, mean);  // :330 second CoG removal on eps_x
  for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) {
    const int i = idx / D, ch = idx % D;
    float e = ev[idx];
    if (ch < 3 && i < n) e -= mean[ch];
    const float mu = zv[idx] / alpha - ceps * e;  // :331
    zv[idx] = mu + sigma * nz[idx];               // :337, :442
  }
  __syncthreads();
  block_mean3(zv, N, D, n, s_mean, mean);  // :340-344
  float* dst = zs + (int64_t)b * N * D;
  for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) {
    const int i = idx / D, ch = idx % D;
    float v = zv[idx];
    if (ch < 3 && i < n) v -= mean[ch];
    dst[idx] = v;
  }
}

// diffusion_qm9.py:294-310, :174-179
__global__ void __launch_bounds__(128) final_decode_k(const float* __restrict__ z0, const float* __restrict__ eps0,
                                                      const float* __restrict__ rx, const float* __restrict__ rh,
                                                      const int32_t* __restrict__ sizes, int N, int F,
                                                      const float* __restrict__ sched, int sched_per_mol,
                                                      float norm_x, float norm_h, float bias_h,
                                                      float* __restrict__ x, float* __restrict__ h) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float nz[];
  __shared__ float s_mean[3];
  const int D = 3 + F, b = blockIdx.x, n = sizes[b];
  const float* sc = sched + (sched_per_mol ? 3 * b : 0);
  const float alpha0 = sc[0], sigma0 = sc[1], sigma_x = sc[2];
  load_noise(rx + (int64_t)b * N * 3, rh + (int64_t)b * N * F, N, F, n, nz, s_mean);
  const float* zsrc = z0 + (int64_t)b * N * D;
  const float* esrc = eps0 + (int64_t)b * N * D;
  for (int idx = threadIdx.x; idx < N * D; idx += blockDim.x) {
    const int i = idx / D, ch = idx % D;
    if (ch < 3) {
      const float mu = 1.0f / alpha0 * (zsrc[idx] - sigma0 * esrc[idx]);  // :244
      x[((int64_t)b * N + i) * 3 + ch] = (mu + sigma_x * nz[idx]) * norm_x;
    } else {
      h[((int64_t)b * N + i) * F + (ch - 3)] = (zsrc[idx] * norm_h + bias_h) * (i < n ? 1.f : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------
// The sampling loop's own kernels (hd_sampler_begin / _step / _final): everything between two EGNN stacks of
// DiffusionQM9.sample (diffusion_qm9.py:375-384) in ONE launch - the centre-of-gravity projection and NaN guard that end
// _forward (en_dynamics.py:109-116), the reverse update (diffusion_qm9.py:328-345), the step counter, and the input side
// of the next forward (en_dynamics.py:56-79 masking / time / context channels, egnn_new.py:197 embedding) together with
// the first sub-layer's A|B pre-projection folded through the embedding (hd_layout.cu fuse_embed_k).
// ---------------------------------------------------------------------------------------
struct SamplerArgs {
  // chain state
  float* z;                  // [B,N,3+F], updated in place
  const float* rx;           // [B,N,3]   raw randn draw of this step
  const float* rh;           // [B,N,F]
  const float* t_table;      // [T+1]
  const float* sched_table;  // [T+1][sched_rows][3]
  const float* context;      // [B,N,C] or null
  const int32_t* sizes;
  int32_t* state;            // Workspace::state
  int32_t* flags;
  const float* eps_raw;      // [B,N,3+F] written by out_vel_k
  int B, N, F, C, T, sched_rows, ragged;
  // input side of the next forward
  const float *emb_wT, *emb_b, *emb_abT, *emb_abb;   // emb_abT null: no A|B image (fp32 engine)
  float *x, *x0, *x2, *h, *ab;
  int32_t *row_off, *node_off;                       // prefix tables (written by the begin kernel, read afterwards)
  int rows_bound;
  // final decode
  float norm_x, norm_h, bias_h;
  float *x_out, *h_out;
};

// Input side of an EGNN forward for molecule b from its z rows in shared memory (zv [N][D]): x, x0 (masked), the
// embedding h and, for the tensor-core engines, the first sub-layer's A|B operands.  row0 = first workspace row of the
// molecule; inp = shared scratch [N][EMB_FMAX].  Called by all 256 threads of the block; thread = channel.  The thread
// keeps its columns of the (folded) embedding weights in registers and walks over the nodes without any barrier.
constexpr int EMB_FMAX = 12;               // input channels (F + time + context) held in registers; more: re-read
__device__ __forceinline__ void embed_molecule(const SamplerArgs& p, const float* zv, int b, int n, int64_t row0,
                                               float t, float* inp) {
  const int N = p.N, F = p.F, C = p.C, D = 3 + F, Fi = F + 1 + C, tid = threadIdx.x;
  const int64_t rows = (int64_t)p.B * N;
  const int count = p.ragged ? n : N;      // ragged node rows: padded nodes own no row
  const bool ab = p.emb_abT != nullptr;
  // per-node input vectors [features * mask | t | context], zero-padded to EMB_FMAX (time and context are not masked,
  // en_dynamics.py:66-79), and the masked coordinates
  for (int i = tid >> 4; i < count; i += 16) {
    const int f = tid & 15;
    if (f < EMB_FMAX) {
      float v = 0.f;
      if (i < n) {
        if (f < F) v = zv[i * D + 3 + f];
        else if (f == F) v = t;
        else if (f < Fi) v = p.context[((int64_t)b * N + i) * C + (f - F - 1)];
      }
      inp[i * EMB_FMAX + f] = v;
    } else if (f < EMB_FMAX + 3) {
      const int ch = f - EMB_FMAX;
      const float v = i < n ? zv[i * D + ch] : 0.f;
      p.x[(row0 + i) * 3 + ch] = v;
      p.x0[(row0 + i) * 3 + ch] = v;
    }
  }
  float w[EMB_FMAX], wa[EMB_FMAX], wb[EMB_FMAX];
#pragma unroll
  for (int f = 0; f < EMB_FMAX; ++f) {
    w[f] = f < Fi ? p.emb_wT[f * H + tid] : 0.f;
    wa[f] = ab && f < Fi ? p.emb_abT[f * 2 * H + tid] : 0.f;
    wb[f] = ab && f < Fi ? p.emb_abT[f * 2 * H + H + tid] : 0.f;
  }
  const float bias = p.emb_b[tid], bias_a = ab ? p.emb_abb[tid] : 0.f, bias_b = ab ? p.emb_abb[H + tid] : 0.f;
  __syncthreads();
  float* hp = p.h + row0 * H + tid;
  // K-chunk-major operand image of the edge kernel: [col / 16][row][col % 16], A (cols 0..H-1) then B
  float* ap = p.ab + ((int64_t)(tid >> 4) * rows + row0) * 16 + (tid & 15);
  float* bp = p.ab + ((int64_t)((H + tid) >> 4) * rows + row0) * 16 + (tid & 15);
  for (int i = 0; i < count; ++i, hp += H, ap += 16, bp += 16) {
    float v = 0.f, a0 = 0.f, a1 = 0.f;
    if (i < n) {
      v = bias;
      a0 = bias_a;
      a1 = bias_b;
      const float4* in4 = reinterpret_cast<const float4*>(inp + i * EMB_FMAX);   // broadcast reads
#pragma unroll
      for (int q = 0; q < EMB_FMAX / 4; ++q) {
        const float4 in = in4[q];
        v = fmaf(in.x, w[4 * q], v);          a0 = fmaf(in.x, wa[4 * q], a0);          a1 = fmaf(in.x, wb[4 * q], a1);
        v = fmaf(in.y, w[4 * q + 1], v);      a0 = fmaf(in.y, wa[4 * q + 1], a0);      a1 = fmaf(in.y, wb[4 * q + 1], a1);
        v = fmaf(in.z, w[4 * q + 2], v);      a0 = fmaf(in.z, wa[4 * q + 2], a0);      a1 = fmaf(in.z, wb[4 * q + 2], a1);
        v = fmaf(in.w, w[4 * q + 3], v);      a0 = fmaf(in.w, wa[4 * q + 3], a0);      a1 = fmaf(in.w, wb[4 * q + 3], a1);
      }
      for (int f = EMB_FMAX; f < Fi; ++f) {   // wider inputs than the register file holds: re-read per node
        const float in = f < F ? zv[i * D + 3 + f] : (f == F ? t : p.context[((int64_t)b * N + i) * C + (f - F - 1)]);
        v = fmaf(in, __ldg(p.emb_wT + f * H + tid), v);
        if (ab) {
          a0 = fmaf(in, __ldg(p.emb_abT + f * 2 * H + tid), a0);
          a1 = fmaf(in, __ldg(p.emb_abT + f * 2 * H + H + tid), a1);
        }
      }
    }
    *hp = v;
    if (ab) {
      *ap = a0;
      *bp = a1;
    }
  }
}

// element loop over a molecule's [N][D] block without integer division when D <= 16: 16 threads per node
#define HD_FOR_EACH_ELEM(i, ch, N, D, body)                                              \
  if ((D) <= 16) {                                                                       \
    const int ch = threadIdx.x & 15;                                                     \
    if (ch < (D))                                                                        \
      for (int i = threadIdx.x >> 4; i < (N); i += (int)(blockDim.x >> 4)) { body }      \
  } else {                                                                               \
    for (int _e = threadIdx.x; _e < (N) * (D); _e += blockDim.x) {                        \
      const int i = _e / (D), ch = _e - i * (D);                                         \
      body                                                                               \
    }                                                                                    \
  }

// First kernel of a chain: loop state, the prefix tables of the edge kernel / ragged layout, and the input side of the
// first forward.  One CTA per molecule, 256 threads.
__global__ void __launch_bounds__(256) sampler_begin_k(const SamplerArgs p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];   // inp [N*EMB_FMAX] (16-byte aligned rows), zv [N*D]
  __shared__ int s_part[8];
  const int D = 3 + p.F, b = blockIdx.x, tid = threadIdx.x, n = p.sizes[b];
  if (b == 0 && tid < 32) {
    if (tid < 4) p.state[tid] = 0;
    int carry_r = 0, carry_n = 0;
    if (tid == 0) {
      if (p.row_off) p.row_off[0] = 0;
      if (p.node_off) p.node_off[0] = 0;
    }
    for (int base = 0; base < p.B; base += 32) {
      const int bb = base + tid;
      const int nn = bb < p.B ? p.sizes[bb] : 0;
      int vr = edge_rows(nn), vn = nn;
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const int ur = __shfl_up_sync(0xffffffffu, vr, o2), un = __shfl_up_sync(0xffffffffu, vn, o2);
        if (tid >= o2) {
          vr += ur;
          vn += un;
        }
      }
      if (bb < p.B) {
        if (p.row_off) p.row_off[bb + 1] = carry_r + vr;
        if (p.node_off) p.node_off[bb + 1] = carry_n + vn;
      }
      carry_r += __shfl_sync(0xffffffffu, vr, 31);
      carry_n += __shfl_sync(0xffffffffu, vn, 31);
    }
    // the caller's bound on sum(sizes) sized the node-GEMM grids: rows beyond it would silently go missing
    if (tid == 0 && p.ragged && carry_n > p.rows_bound && p.flags) atomicOr(p.flags, HD_FLAG_MASK);
  }
  // first workspace row of this molecule: every CTA sums the sizes ahead of it itself (the table block 0 publishes is
  // for the kernels that follow)
  int64_t row0 = (int64_t)b * p.N;
  if (p.ragged) {
    int part = 0;
    for (int k = tid; k < b; k += 256) part += p.sizes[k];
    part = warp_sum_int(part);
    if ((tid & 31) == 0) s_part[tid >> 5] = part;
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) base += s_part[w];
    row0 = base;
  }
  const float* zsrc = p.z + (int64_t)b * p.N * D;
  float* zv = sm + p.N * EMB_FMAX;
  for (int idx = tid; idx < p.N * D; idx += 256) zv[idx] = zsrc[idx];
  // x2, the ping-pong partner of the coordinate updates: padded rows must be 0 and are never written afterwards
  for (int idx = tid; idx < (p.ragged ? n : p.N) * 3; idx += 256) p.x2[row0 * 3 + idx] = 0.f;
  __syncthreads();
  embed_molecule(p, zv, b, n, row0, p.t_table[0], sm);
}

// Between two forwards (FINAL: after the last one).  One CTA per molecule, 256 threads.
template <bool FINAL>
__global__ void __launch_bounds__(256) sampler_tail_k(const SamplerArgs p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];  // nz [N*D], ev [N*D], zv [N*D]; nz is reused as the embedding's input scratch
  __shared__ float s_mean[3];
  __shared__ float s_chk[3];
  const int N = p.N, F = p.F, D = 3 + F, b = blockIdx.x, tid = threadIdx.x, n = p.sizes[b];
  float* nz = sm;
  float* ev = sm + N * max(D, EMB_FMAX);
  float* zv = ev + N * D;
  const int k = min(p.state[1], p.T);          // the step whose forward has just run (latched by out_vel_k)
  const bool nan = p.state[2 + (k & 1)] != 0;  // NaN guard of that forward (whole batch, en_dynamics.py:109-111)
  const float* sc = p.sched_table + ((int64_t)k * p.sched_rows + (p.sched_rows > 1 ? b : 0)) * 3;
  const float c0 = sc[0], c1 = sc[1], c2 = sc[2];
  const float* zsrc = p.z + (int64_t)b * N * D;
  const float* esrc = p.eps_raw + (int64_t)b * N * D;
  const float* rxs = p.rx + (int64_t)b * N * 3;
  const float* rhs = p.rh + (int64_t)b * N * F;
  if (tid < 3) s_chk[tid] = 0.f;
  // the step's noise (diffusion_qm9.py:445-456: masked, positions centred below), z_t and eps into shared memory
  float pad_abs = 0.f, max_abs = 0.f;
  HD_FOR_EACH_ELEM(i, ch, N, D, {
    const int idx = i * D + ch;
    const float mk = i < n ? 1.f : 0.f;
    nz[idx] = (ch < 3 ? rxs[i * 3 + ch] : rhs[i * F + (ch - 3)]) * mk;
    const float v = zsrc[idx];
    zv[idx] = v;
    float e = esrc[idx];
    if (nan && ch < 3) e = 0.f;
    ev[idx] = e;
    if (ch < 3) {
      if (i >= n) pad_abs = fmaxf(pad_abs, fabsf(v));
      max_abs = fmaxf(max_abs, fabsf(v));
    }
  })
  if (nan && p.flags && b == 0 && tid == 0) atomicOr(p.flags, HD_FLAG_NAN);
  __syncthreads();
  float mean[3], emean[3];
  block_mean3(nz, N, D, n, s_mean, mean);      // sample_center_gravity_zero_gaussian_with_mask (models/utils.py:126-135)
  block_mean3(ev, N, D, n, s_mean, emean);     // remove_mean_with_mask that ends _forward (en_dynamics.py:116)
  for (int idx = tid; idx < n * 3; idx += 256) {
    const int i = idx / 3, ch = idx - 3 * i;
    nz[i * D + ch] -= mean[ch];
    ev[i * D + ch] -= emean[ch];
  }
  __syncthreads();
  if (FINAL) {
    // sample_p_xh_given_z0 (diffusion_qm9.py:294-310, :174-179): {alpha_0, sigma_0, sigma_x} = c0, c1, c2
    HD_FOR_EACH_ELEM(i, ch, N, D, {
      const int idx = i * D + ch;
      if (ch < 3) {
        const float mu = 1.0f / c0 * (zv[idx] - c1 * ev[idx]);   // :244
        p.x_out[((int64_t)b * N + i) * 3 + ch] = (mu + c2 * nz[idx]) * p.norm_x;
      } else {
        p.h_out[((int64_t)b * N + i) * F + (ch - 3)] = (zv[idx] * p.norm_h + p.bias_h) * (i < n ? 1.f : 0.f);
      }
    })
    return;
  }
  // sample_p_zs_given_zt after the network call (diffusion_qm9.py:328-345): {alpha_t|s, sigma2_t|s/alpha_t|s/sigma_t,
  // sigma_t|s sigma_s / sigma_t} = c0, c1, c2.  assert_mean_zero_with_mask(zt_x) (models/utils.py:65-75) per molecule:
  atomicMax(reinterpret_cast<int*>(&s_chk[0]), __float_as_int(pad_abs));   // float max of non-negatives == int max
  atomicMax(reinterpret_cast<int*>(&s_chk[1]), __float_as_int(max_abs));
  __syncthreads();
  float zsum[3];
  block_mean3(zv, N, D, 1, s_mean, zsum);      // n = 1 -> plain sums
  if (p.flags && tid == 0) {
    const float err = fmaxf(fabsf(zsum[0]), fmaxf(fabsf(zsum[1]), fabsf(zsum[2])));
    int f = 0;
    if (!(s_chk[0] < 1e-4f)) f |= HD_FLAG_MASK;
    if (!(err / (s_chk[1] + 1e-10f) < 1e-2f)) f |= HD_FLAG_COG;
    if (f) atomicOr(p.flags, f);
  }
  block_mean3(ev, N, D, n, s_mean, mean);      // :330 second CoG removal on eps_x
  HD_FOR_EACH_ELEM(i, ch, N, D, {
    const int idx = i * D + ch;
    float e = ev[idx];
    if (ch < 3 && i < n) e -= mean[ch];
    const float mu = zv[idx] / c0 - c1 * e;    // :331
    zv[idx] = mu + c2 * nz[idx];               // :337, :442
  })
  __syncthreads();
  block_mean3(zv, N, D, n, s_mean, mean);      // :340-344
  float* dst = p.z + (int64_t)b * N * D;
  HD_FOR_EACH_ELEM(i, ch, N, D, {
    const int idx = i * D + ch;
    float v = zv[idx];
    if (ch < 3 && i < n) v -= mean[ch];
    zv[idx] = v;
    dst[idx] = v;
  })
  if (b == 0 && tid == 0) p.state[0] = k + 1;  // read by the next forward's out_vel_k only
  __syncthreads();
  const int64_t row0 = p.ragged ? p.node_off[b] : (int64_t)b * N;
  embed_molecule(p, zv, b, n, row0, p.t_table[min(k + 1, p.T)], nz);
}

__device__ __forceinline__ float softplus_t(float v) { return v > 20.0f ? v : log1pf(expf(v)); }  // F.softplus
__device__ __forceinline__ float logsigmoid_t(float v) { return fminf(v, 0.0f) - log1pf(expf(-fabsf(v))); }
__device__ __forceinline__ float sigmoid_t(float v) { return 1.0f / (1.0f + expf(-v)); }

__global__ void step_scalars_k(const float* __restrict__ gs, const float* __restrict__ gt, int count,
                               float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const float s = gs[k], t = gt[k];
  const float sigma2_ts = -expm1f(softplus_t(s) - softplus_t(t));            // :189-191
  const float alpha_ts = expf(0.5f * (logsigmoid_t(-t) - logsigmoid_t(-s)));  // :194-200
  const float sigma_ts = sqrtf(sigma2_ts);
  const float sigma_s = sqrtf(sigmoid_t(s)), sigma_t = sqrtf(sigmoid_t(t));  // :148-150
  out[3 * k + 0] = alpha_ts;
  out[3 * k + 1] = sigma2_ts / alpha_ts / sigma_t;  // :331
  out[3 * k + 2] = sigma_ts * sigma_s / sigma_t;    // :334
}

__global__ void final_scalars_k(const float* __restrict__ g0, int count, float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const float g = g0[k];
  out[3 * k + 0] = sqrtf(sigmoid_t(-g));  // alpha_0
  out[3 * k + 1] = sqrtf(sigmoid_t(g));   // sigma_0
  out[3 * k + 2] = expf(-(-0.5f * g));    // SNR(-0.5*gamma_0)
}

// ---------------------------------------------------------------------------------------
// orchestration
// ---------------------------------------------------------------------------------------
static int check_common(const hd_config* cfg, const void* packed, const int32_t* sizes, int B, int N, int engine) {
  if (!cfg || !packed || !sizes) {
    set_error("null argument");
    return HD_E_INVALID;
  }
  if (B < 1 || N < 1 || N > 128) {
    set_error("unsupported batch shape B=%d N=%d (need B>=1, 1<=N<=128)", B, N);
    return HD_E_INVALID;
  }
  if (engine != HD_ENGINE_FP32 && engine != HD_ENGINE_TC_STRICT && engine != HD_ENGINE_TC_FAST) {
    set_error("unknown engine %d", engine);
    return HD_E_INVALID;
  }
  return HD_OK;
}

static int sub_gcl(const FwdCtx& c, int si, float* h, const float* x, const float* x0, int engine) {
  if (engine == HD_ENGINE_FP32) return fp32_gcl(c, si, h, x, x0);
  return tc_gcl(c, si, h, x, x0, engine);
}
static int sub_equiv(const FwdCtx& c, int si, const float* h, const float* x, const float* x0, float* xo,
                     int engine) {
  if (engine == HD_ENGINE_FP32) return fp32_equiv(c, si, h, x, x0, xo);
  return tc_equiv(c, si, h, x, x0, xo, engine);
}

// EGNN blocks (egnn_new.py:198-199, :139-152) on ws.h / ws.x; returns the buffers holding the final h and x
static int run_blocks(const FwdCtx& c, int engine, float** h_final, float** x_final, bool ab_ready = false) {
  float* h = reinterpret_cast<float*>(c.ws + c.W.h);
  float* h2 = reinterpret_cast<float*>(c.ws + c.W.h2);
  float* x = reinterpret_cast<float*>(c.ws + c.W.x);
  float* x2 = reinterpret_cast<float*>(c.ws + c.W.x2);
  const float* x0 = reinterpret_cast<const float*>(c.ws + c.W.x0);
  int si = 0, rc;
  c.ab_ready = ab_ready;   // the sampling loop's tail kernel has already produced block 0's A|B operands
  c.ab2_ready = false;
  for (int l = 0; l < c.cfg->n_layers; ++l) {
    for (int s = 0; s < c.cfg->inv_sublayers; ++s) {
      if (engine != HD_ENGINE_FP32 && c.L->subs[si].fuse_next) {
        // tensor-core engines: node_mlp.2 and the next sub-layer's pre-projection share a launch; h ping-pongs.  The
        // first sub-layer of a block finds its own operands in ws.ab2 (written by the previous block's last GCL)
        const bool use_ab2 = s == 0 && c.ab2_ready;
        if ((rc = tc_gcl_fused(c, si++, h, h2, x, x0, engine, use_ab2))) return rc;
        float* tmp = h;
        h = h2;
        h2 = tmp;
      } else if ((rc = sub_gcl(c, si++, h, x, x0, engine))) {
        return rc;
      }
    }
    if ((rc = sub_equiv(c, si++, h, x, x0, x2, engine))) return rc;
    float* tmp = x;
    x = x2;
    x2 = tmp;
  }
  *h_final = h;
  *x_final = x;
  return HD_OK;
}

}  // namespace hd

using namespace hd;

extern "C" {

HD_API int32_t hd_abi_version(void) { return HD_ABI_VERSION; }
HD_API const char* hd_last_error(void) { return hd::last_error(); }
HD_API int32_t hd_engine_available(int32_t engine) {
  if (engine == HD_ENGINE_FP32) return 1;
  if (engine == HD_ENGINE_TC_STRICT || engine == HD_ENGINE_TC_FAST) return tc_available() ? 1 : 0;
  return 0;
}

HD_API int64_t hd_launch_count(void) { return hd::launch_count(); }

HD_API int64_t hd_weight_count(const hd_config* cfg) {
  Layout L;
  if (!cfg || !make_layout(*cfg, &L)) return HD_E_INVALID;
  return L.flat_count;
}

HD_API int64_t hd_packed_bytes(const hd_config* cfg) {
  Layout L;
  if (!cfg || !make_layout(*cfg, &L)) return HD_E_INVALID;
  return L.total_bytes;
}

HD_API int32_t hd_pack_weights(const hd_config* cfg, const float* w_flat, void* packed, hd_stream_t stream) {
  Layout L;
  if (!cfg || !w_flat || !packed) {
    set_error("null argument");
    return HD_E_INVALID;
  }
  if (!make_layout(*cfg, &L)) return HD_E_INVALID;
  return pack_weights(*cfg, L, w_flat, static_cast<char*>(packed), static_cast<cudaStream_t>(stream));
}

HD_API int64_t hd_workspace_bytes(const hd_config* cfg, int32_t B, int32_t N) {
  if (!cfg || B < 1 || N < 1) return HD_E_INVALID;
  return make_workspace(*cfg, B, N).total_bytes;
}

HD_API int32_t hd_dynamics_forward(const hd_config* cfg, const void* packed, const float* z, const float* t,
                            const int32_t* sizes, int32_t B, int32_t N, float* eps, void* workspace,
                            int32_t* flags, int32_t engine, hd_stream_t stream) {
  return hd_dynamics_forward_ctx(cfg, packed, z, t, nullptr, 0, sizes, B, N, eps, workspace, flags, engine, stream);
}

HD_API int32_t hd_dynamics_forward_ctx(const hd_config* cfg, const void* packed, const float* z, const float* t,
                                const float* context, int32_t context_nf, const int32_t* sizes, int32_t B,
                                int32_t N, float* eps, void* workspace, int32_t* flags, int32_t engine,
                                hd_stream_t stream) {
  return hd_dynamics_forward_ragged(cfg, packed, z, t, context, context_nf, sizes, B, N, 0, eps, workspace, flags,
                                    engine, stream);
}

HD_API int32_t hd_dynamics_forward_ragged(const hd_config* cfg, const void* packed, const float* z, const float* t,
                                   const float* context, int32_t context_nf, const int32_t* sizes, int32_t B,
                                   int32_t N, int32_t live_rows, float* eps, void* workspace, int32_t* flags,
                                   int32_t engine, hd_stream_t stream) {
  if (live_rows < 0 || (int64_t)live_rows > (int64_t)B * N) {
    set_error("live_rows=%d outside [0, B*N]", live_rows);
    return HD_E_INVALID;
  }
  if (live_rows > 0) engine |= HD_ENGINE_RAGGED_ROWS;
  const bool ragged = (engine & HD_ENGINE_RAGGED_ROWS) != 0;
  const int raw = (engine & HD_ENGINE_RAW_VELOCITY) ? 1 : 0;
  engine &= ~(HD_ENGINE_RAGGED_ROWS | HD_ENGINE_RAW_VELOCITY);
  int rc = check_common(cfg, packed, sizes, B, N, engine);
  if (rc) return rc;
  if (!z || !t || !eps || !workspace) {
    set_error("null argument");
    return HD_E_INVALID;
  }
  if (context_nf < 0 || (context_nf > 0 && !context) || cfg->in_node_nf - 1 - context_nf < 0) {
    set_error("bad context (context_nf=%d, in_node_nf=%d)", context_nf, cfg->in_node_nf);
    return HD_E_INVALID;
  }
  Layout L;
  if (!make_layout(*cfg, &L)) return HD_E_INVALID;
  FwdCtx c{cfg, &L, static_cast<const char*>(packed), static_cast<char*>(workspace), make_workspace(*cfg, B, N),
           sizes, B, N, static_cast<cudaStream_t>(stream)};
  c.x_prezeroed = true;   // prep_k below writes x (masked) and zeroes x2
  const int Fi = cfg->in_node_nf, C = context_nf, F = Fi - 1 - C, D = 3 + F;
  const int64_t BN = (int64_t)B * N;
  auto WF = [&](int64_t off) { return reinterpret_cast<float*>(c.ws + off); };
  auto PF = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  int32_t* nanflag = reinterpret_cast<int32_t*>(c.ws + c.W.nanflag);
  if (Fi > 64 || D > 256) {
    set_error("in_node_nf=%d too large for the fused input kernel", Fi);
    return HD_E_INVALID;
  }
  int32_t* row_off = engine == HD_ENGINE_FP32 ? nullptr : reinterpret_cast<int32_t*>(c.ws + c.W.row_off);
  // HD_ENGINE_RAGGED_ROWS on a tensor-core engine: only real nodes own a row of h / x / agg, so the node GEMMs do not
  // pay for padding (B*N rows are still allocated, and the grids still cover them: see lin::Params::ragged)
  int32_t* node_off = row_off && ragged ? reinterpret_cast<int32_t*>(c.ws + c.W.node_off) : nullptr;
  if (row_off && B > 4096) {
    set_error("tensor-core engine supports at most 4096 molecules per call (got %d)", B);
    return HD_E_INVALID;
  }
  HD_CHECK_CUDA(launch_pdl(prep_embed_k, dim3((unsigned)BN), dim3(256), 0, c.stream, z, t, context, C, sizes, B, N, F,
                          PF(L.emb_wT), PF(L.emb_b), WF(c.W.x), WF(c.W.x0), WF(c.W.x2), WF(c.W.h), nanflag, row_off,
                          node_off, live_rows > 0 ? live_rows : (int)BN, flags));
  count_launch();
  c.planned = row_off != nullptr;
  c.node_off = node_off;
  c.rows_bound = node_off && live_rows > 0 ? live_rows : 0;
  float *hf = nullptr, *xf = nullptr;
  if ((rc = run_blocks(c, engine, &hf, &xf))) return rc;
  HD_CHECK_CUDA(launch_pdl(out_vel_k, dim3((unsigned)BN), dim3(256), 0, c.stream, (const float*)hf, PF(L.out_w),
                          PF(L.out_b), Fi, F, sizes, N, (const float*)xf, (const float*)WF(c.W.x0), WF(c.W.eps_raw),
                          nanflag, (const int32_t*)node_off, (int32_t*)nullptr));
  count_launch();
  HD_CHECK_CUDA(launch_pdl(cog_k, dim3(B), dim3(128), sizeof(float) * N * D, c.stream, (const float*)WF(c.W.eps_raw),
                          sizes, N, F, (const int32_t*)nanflag, eps, flags, raw));
  count_launch();
  return HD_OK;
}

HD_API int32_t hd_egnn_forward(const hd_config* cfg, const void* packed, const float* h_in, const float* x_in,
                        const int32_t* sizes, int32_t B, int32_t N, float* h_out, float* x_out, void* workspace,
                        int32_t engine, hd_stream_t stream) {
  int rc = check_common(cfg, packed, sizes, B, N, engine);
  if (rc) return rc;
  if (!h_in || !x_in || !h_out || !x_out || !workspace) {
    set_error("null argument");
    return HD_E_INVALID;
  }
  Layout L;
  if (!make_layout(*cfg, &L)) return HD_E_INVALID;
  FwdCtx c{cfg, &L, static_cast<const char*>(packed), static_cast<char*>(workspace), make_workspace(*cfg, B, N),
           sizes, B, N, static_cast<cudaStream_t>(stream)};
  const int Fi = cfg->in_node_nf;
  const int64_t BN = (int64_t)B * N;
  auto WF = [&](int64_t off) { return reinterpret_cast<float*>(c.ws + off); };
  auto PF = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  HD_CHECK_CUDA(cudaMemcpyAsync(WF(c.W.x), x_in, BN * 3 * 4, cudaMemcpyDeviceToDevice, c.stream));
  HD_CHECK_CUDA(cudaMemcpyAsync(WF(c.W.x0), x_in, BN * 3 * 4, cudaMemcpyDeviceToDevice, c.stream));
  embed_k<<<(unsigned)BN, 256, 0, c.stream>>>(h_in, Fi, PF(L.emb_wT), PF(L.emb_b), sizes, N, WF(c.W.h));
  HD_CHECK_LAUNCH();
  float *hf = nullptr, *xf = nullptr;
  if ((rc = run_blocks(c, engine, &hf, &xf))) return rc;
  out_k<<<(unsigned)BN, 256, 0, c.stream>>>(hf, PF(L.out_w), PF(L.out_b), Fi, sizes, N, h_out);
  HD_CHECK_LAUNCH();
  HD_CHECK_CUDA(cudaMemcpyAsync(x_out, xf, BN * 3 * 4, cudaMemcpyDeviceToDevice, c.stream));
  return HD_OK;
}

HD_API int32_t hd_gcl_forward(const hd_config* cfg, const void* packed, int32_t block, int32_t sub, float* h,
                       const float* x, const float* x0, const int32_t* sizes, int32_t B, int32_t N,
                       void* workspace, int32_t engine, hd_stream_t stream) {
  int rc = check_common(cfg, packed, sizes, B, N, engine);
  if (rc) return rc;
  if (!h || !x || !x0 || !workspace || block < 0 || block >= cfg->n_layers || sub < 0 ||
      sub >= cfg->inv_sublayers) {
    set_error("bad argument (block=%d sub=%d)", block, sub);
    return HD_E_INVALID;
  }
  Layout L;
  if (!make_layout(*cfg, &L)) return HD_E_INVALID;
  FwdCtx c{cfg, &L, static_cast<const char*>(packed), static_cast<char*>(workspace), make_workspace(*cfg, B, N),
           sizes, B, N, static_cast<cudaStream_t>(stream)};
  return sub_gcl(c, block * (cfg->inv_sublayers + 1) + sub, h, x, x0, engine);
}

HD_API int32_t hd_equiv_update(const hd_config* cfg, const void* packed, int32_t block, const float* h, const float* x,
                        const float* x0, const int32_t* sizes, int32_t B, int32_t N, float* x_out,
                        void* workspace, int32_t engine, hd_stream_t stream) {
  int rc = check_common(cfg, packed, sizes, B, N, engine);
  if (rc) return rc;
  if (!h || !x || !x0 || !x_out || x_out == x || !workspace || block < 0 || block >= cfg->n_layers) {
    set_error("bad argument (block=%d; x_out must not alias x)", block);
    return HD_E_INVALID;
  }
  Layout L;
  if (!make_layout(*cfg, &L)) return HD_E_INVALID;
  FwdCtx c{cfg, &L, static_cast<const char*>(packed), static_cast<char*>(workspace), make_workspace(*cfg, B, N),
           sizes, B, N, static_cast<cudaStream_t>(stream)};
  return sub_equiv(c, block * (cfg->inv_sublayers + 1) + cfg->inv_sublayers, h, x, x0, x_out, engine);
}

HD_API int32_t hd_edge_kernel_only(const hd_config* cfg, const void* packed, int32_t block, int32_t sub,
                                   const float* x, const float* x0, const int32_t* sizes, int32_t B, int32_t N,
                                   void* workspace, int32_t engine, hd_stream_t stream) {
  int rc = check_common(cfg, packed, sizes, B, N, engine);
  if (rc) return rc;
  if (!x || !x0 || !workspace || block < 0 || block >= cfg->n_layers || sub < 0 || sub > cfg->inv_sublayers) {
    set_error("bad argument (block=%d sub=%d)", block, sub);
    return HD_E_INVALID;
  }
  Layout L;
  if (!make_layout(*cfg, &L)) return HD_E_INVALID;
  FwdCtx c{cfg, &L, static_cast<const char*>(packed), static_cast<char*>(workspace), make_workspace(*cfg, B, N),
           sizes, B, N, static_cast<cudaStream_t>(stream)};
  const int si = block * (cfg->inv_sublayers + 1) + sub;
  if (engine == HD_ENGINE_FP32) return fp32_edge_only(c, si, x, x0);
  return tc_edge_only(c, si, x, x0, engine);
}

// ---- the sampling loop (hd_sampler_begin / _step / _final) -------------------------------------------------------
struct SamplerCall {
  Layout L;
  FwdCtx c;
  SamplerArgs a;
  int engine;
  size_t tail_smem;
};

static int sampler_setup(SamplerCall* sc, const hd_config* cfg, const void* packed, float* z, const float* t_table,
                         const float* sched_table, int sched_rows, int T, const float* context, int context_nf,
                         const int32_t* sizes, int B, int N, int live_rows, void* workspace, int32_t* flags, int engine,
                         hd_stream_t stream) {
  if (live_rows < 0 || (int64_t)live_rows > (int64_t)B * N) {
    set_error("live_rows=%d outside [0, B*N]", live_rows);
    return HD_E_INVALID;
  }
  if (live_rows > 0) engine |= HD_ENGINE_RAGGED_ROWS;
  const bool ragged_req = (engine & HD_ENGINE_RAGGED_ROWS) != 0;
  engine &= ~HD_ENGINE_RAGGED_ROWS;
  int rc = check_common(cfg, packed, sizes, B, N, engine);
  if (rc) return rc;
  if (!z || !t_table || !workspace || T < 1 || (sched_table && sched_rows != 1 && sched_rows != B)) {
    set_error("bad argument (null pointer, T=%d, sched_rows=%d)", T, sched_rows);
    return HD_E_INVALID;
  }
  if (context_nf < 0 || (context_nf > 0 && !context) || cfg->in_node_nf - 1 - context_nf < 0) {
    set_error("bad context (context_nf=%d, in_node_nf=%d)", context_nf, cfg->in_node_nf);
    return HD_E_INVALID;
  }
  if (!make_layout(*cfg, &sc->L)) return HD_E_INVALID;
  const int Fi = cfg->in_node_nf, C = context_nf, F = Fi - 1 - C, D = 3 + F;
  if (Fi > 64 || D > 256) {
    set_error("in_node_nf=%d too large for the fused input kernel", Fi);
    return HD_E_INVALID;
  }
  const bool tc = engine != HD_ENGINE_FP32;
  if (tc && B > 4096) {
    set_error("tensor-core engine supports at most 4096 molecules per call (got %d)", B);
    return HD_E_INVALID;
  }
  sc->engine = engine;
  sc->c = FwdCtx{cfg, &sc->L, static_cast<const char*>(packed), static_cast<char*>(workspace),
                 make_workspace(*cfg, B, N), sizes, B, N, static_cast<cudaStream_t>(stream)};
  FwdCtx& c = sc->c;
  auto WF = [&](int64_t off) { return reinterpret_cast<float*>(c.ws + off); };
  auto PF = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  const bool ragged = tc && ragged_req;
  c.x_prezeroed = true;
  c.planned = tc;
  c.node_off = ragged ? reinterpret_cast<int32_t*>(c.ws + c.W.node_off) : nullptr;
  c.rows_bound = ragged && live_rows > 0 ? live_rows : 0;
  SamplerArgs& a = sc->a;
  a = SamplerArgs{};
  a.z = z;
  a.t_table = t_table;
  a.sched_table = sched_table;
  a.context = context;
  a.sizes = sizes;
  a.state = reinterpret_cast<int32_t*>(c.ws + c.W.state);
  a.flags = flags;
  a.eps_raw = WF(c.W.eps_raw);
  a.B = B; a.N = N; a.F = F; a.C = C; a.T = T; a.sched_rows = sched_rows; a.ragged = ragged ? 1 : 0;
  a.emb_wT = PF(sc->L.emb_wT);
  a.emb_b = PF(sc->L.emb_b);
  a.emb_abT = tc ? PF(sc->L.emb_abT) : nullptr;
  a.emb_abb = tc ? PF(sc->L.emb_abb) : nullptr;
  a.x = WF(c.W.x); a.x0 = WF(c.W.x0); a.x2 = WF(c.W.x2); a.h = WF(c.W.h); a.ab = WF(c.W.ab);
  a.row_off = tc ? reinterpret_cast<int32_t*>(c.ws + c.W.row_off) : nullptr;
  a.node_off = const_cast<int32_t*>(c.node_off);
  a.rows_bound = live_rows > 0 ? live_rows : B * N;
  sc->tail_smem = sizeof(float) * N * (2 * D + (D > EMB_FMAX ? D : EMB_FMAX));
  return HD_OK;
}

// EGNN blocks + output head of the forward whose input side the previous begin / tail kernel prepared
static int sampler_forward(SamplerCall* sc) {
  FwdCtx& c = sc->c;
  const hd_config* cfg = c.cfg;
  const int Fi = cfg->in_node_nf;
  auto WF = [&](int64_t off) { return reinterpret_cast<float*>(c.ws + off); };
  auto PF = [&](int64_t off) { return reinterpret_cast<const float*>(c.packed + off); };
  float *hf = nullptr, *xf = nullptr;
  int rc = run_blocks(c, sc->engine, &hf, &xf, sc->engine != HD_ENGINE_FP32);
  if (rc) return rc;
  HD_CHECK_CUDA(launch_pdl(out_vel_k, dim3((unsigned)((int64_t)c.B * c.N)), dim3(256), 0, c.stream, (const float*)hf,
                          PF(sc->L.out_w), PF(sc->L.out_b), Fi, sc->a.F, c.sizes, c.N, (const float*)xf,
                          (const float*)WF(c.W.x0), WF(c.W.eps_raw), (int32_t*)nullptr, (const int32_t*)c.node_off,
                          sc->a.state));
  count_launch();
  return HD_OK;
}

static int check_mol(const int32_t* sizes, int B, int N, int F) {
  if (!sizes || B < 1 || N < 1 || N > 1024 || F < 0 || F > 64) {
    set_error("bad shape B=%d N=%d F=%d", B, N, F);
    return HD_E_INVALID;
  }
  return HD_OK;
}

HD_API int32_t hd_combine_noise(const float* randn_x, const float* randn_h, const int32_t* sizes, int32_t B, int32_t N,
                         int32_t F, float* z, hd_stream_t stream) {
  int rc = check_mol(sizes, B, N, F);
  if (rc) return rc;
  combine_noise_k<<<B, 128, sizeof(float) * N * (3 + F), static_cast<cudaStream_t>(stream)>>>(randn_x, randn_h,
                                                                                              sizes, N, F, z);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

HD_API int32_t hd_step_scalars(const float* gamma_s, const float* gamma_t, int32_t count, float* sched,
                        hd_stream_t stream) {
  if (!gamma_s || !gamma_t || !sched || count < 1) {
    set_error("bad argument");
    return HD_E_INVALID;
  }
  step_scalars_k<<<(count + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(gamma_s, gamma_t, count,
                                                                                     sched);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

HD_API int32_t hd_final_scalars(const float* gamma_0, int32_t count, float* sched, hd_stream_t stream) {
  if (!gamma_0 || !sched || count < 1) {
    set_error("bad argument");
    return HD_E_INVALID;
  }
  final_scalars_k<<<(count + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(gamma_0, count, sched);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

HD_API int32_t hd_reverse_step(const float* zt, const float* eps, const float* randn_x, const float* randn_h,
                        const int32_t* sizes, int32_t B, int32_t N, int32_t F, const float* sched,
                        int32_t sched_per_mol, float* zs, int32_t* flags, hd_stream_t stream) {
  int rc = check_mol(sizes, B, N, F);
  if (rc) return rc;
  if (!zt || !eps || !randn_x || !randn_h || !sched || !zs) {
    set_error("null argument");
    return HD_E_INVALID;
  }
  HD_CHECK_CUDA(launch_pdl(reverse_step_k, dim3(B), dim3(128), sizeof(float) * 3 * N * (3 + F),
                          static_cast<cudaStream_t>(stream), zt, eps, randn_x, randn_h, sizes, N, F, sched,
                          sched_per_mol, zs, flags));
  count_launch();
  return HD_OK;
}

HD_API int32_t hd_final_decode(const float* z0, const float* eps0, const float* randn_x, const float* randn_h,
                        const int32_t* sizes, int32_t B, int32_t N, int32_t F, const float* sched,
                        int32_t sched_per_mol, float norm_x, float norm_h, float bias_h, float* x, float* h,
                        hd_stream_t stream) {
  int rc = check_mol(sizes, B, N, F);
  if (rc) return rc;
  if (!z0 || !eps0 || !randn_x || !randn_h || !sched || !x || !h) {
    set_error("null argument");
    return HD_E_INVALID;
  }
  final_decode_k<<<B, 128, sizeof(float) * N * (3 + F), static_cast<cudaStream_t>(stream)>>>(
      z0, eps0, randn_x, randn_h, sizes, N, F, sched, sched_per_mol, norm_x, norm_h, bias_h, x, h);
  HD_CHECK_LAUNCH();
  return HD_OK;
}

HD_API int32_t hd_sampler_begin(const hd_config* cfg, const void* packed, const float* z, const float* t_table, int32_t T,
                                const float* context, int32_t context_nf, const int32_t* sizes, int32_t B, int32_t N,
                                int32_t live_rows, void* workspace, int32_t* flags, int32_t engine, hd_stream_t stream) {
  SamplerCall sc;
  int rc = sampler_setup(&sc, cfg, packed, const_cast<float*>(z), t_table, nullptr, 1, T, context, context_nf, sizes, B, N,
                         live_rows, workspace, flags, engine, stream);
  if (rc) return rc;
  HD_CHECK_CUDA(launch_pdl(sampler_begin_k, dim3(B), dim3(256), sizeof(float) * N * (3 + sc.a.F + EMB_FMAX), sc.c.stream,
                          sc.a));
  count_launch();
  return HD_OK;
}

HD_API int32_t hd_sampler_step(const hd_config* cfg, const void* packed, float* z, const float* randn_x,
                               const float* randn_h, const float* t_table, const float* sched_table, int32_t sched_rows,
                               int32_t T, const float* context, int32_t context_nf, const int32_t* sizes, int32_t B,
                               int32_t N, int32_t live_rows, void* workspace, int32_t* flags, int32_t engine,
                               hd_stream_t stream) {
  SamplerCall sc;
  if (!randn_x || !randn_h || !sched_table) {
    set_error("null argument");
    return HD_E_INVALID;
  }
  int rc = sampler_setup(&sc, cfg, packed, z, t_table, sched_table, sched_rows, T, context, context_nf, sizes, B, N,
                         live_rows, workspace, flags, engine, stream);
  if (rc) return rc;
  if ((rc = sampler_forward(&sc))) return rc;
  sc.a.rx = randn_x;
  sc.a.rh = randn_h;
  if (sc.tail_smem > 48 * 1024) {
    set_error("N=%d too large for the fused tail kernel", N);
    return HD_E_INVALID;
  }
  HD_CHECK_CUDA(launch_pdl(sampler_tail_k<false>, dim3(B), dim3(256), sc.tail_smem, sc.c.stream, sc.a));
  count_launch();
  return HD_OK;
}

HD_API int32_t hd_sampler_final(const hd_config* cfg, const void* packed, const float* z, const float* randn_x,
                                const float* randn_h, const float* t_table, const float* sched_table,
                                int32_t sched_rows, int32_t T, const float* context, int32_t context_nf,
                                const int32_t* sizes, int32_t B, int32_t N, int32_t live_rows, float norm_x,
                                float norm_h, float bias_h, float* x, float* h, void* workspace, int32_t* flags,
                                int32_t engine, hd_stream_t stream) {
  SamplerCall sc;
  if (!randn_x || !randn_h || !sched_table || !x || !h) {
    set_error("null argument");
    return HD_E_INVALID;
  }
  int rc = sampler_setup(&sc, cfg, packed, const_cast<float*>(z), t_table, sched_table, sched_rows, T, context, context_nf,
                         sizes, B, N, live_rows, workspace, flags, engine, stream);
  if (rc) return rc;
  if ((rc = sampler_forward(&sc))) return rc;
  sc.a.rx = randn_x;
  sc.a.rh = randn_h;
  sc.a.norm_x = norm_x;
  sc.a.norm_h = norm_h;
  sc.a.bias_h = bias_h;
  sc.a.x_out = x;
  sc.a.h_out = h;
  if (sc.tail_smem > 48 * 1024) {
    set_error("N=%d too large for the fused tail kernel", N);
    return HD_E_INVALID;
  }
  HD_CHECK_CUDA(launch_pdl(sampler_tail_k<true>, dim3(B), dim3(256), sc.tail_smem, sc.c.stream, sc.a));
  count_launch();
  return HD_OK;
}

}  // extern "C"
