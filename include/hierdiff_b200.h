/*
 * hierdiff_b200.h - C ABI of the B200-native HierDiff coarse-grained sampling path.
 *
 * The drop-in boundary of this repository.  The reference (qiangbo1222/HierDiff) has no
 * FFI of its own for this path - the path is PyTorch library calls behind a Python class
 * surface - so each entry point below names the reference Python function whose
 * arithmetic it replaces (paths relative to /root/reference/endiffusion).  The Python
 * mirror in the hierdiff_b200 package binds these through ctypes (INTEGRATION.md shows the stub
 * a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name ends in _host;
 *   - the library BORROWS all buffers, allocates nothing, never synchronises and never
 *     throws: every call only enqueues kernels on `stream` (a cudaStream_t passed as
 *     void*), so any sequence of calls is CUDA-graph capturable;
 *   - int return value: 0 = ok, negative = HD_E_* (message via hd_last_error());
 *   - tensors are fp32, row-major, in the reference's padded layout: z/eps [B,N,3+F],
 *     h [B*N,*], x [B*N,3]; masks are given as `sizes[b]` = number of real nodes of
 *     molecule b (node_mask[b,i] = i < sizes[b]; edge_mask[b,i,j] = i,j < sizes[b], i != j;
 *     diffusion_qm9.py:350-359), 1 <= sizes[b] <= N (the reference's size histogram has no empty molecule).
 */
#ifndef HIERDIFF_B200_H
#define HIERDIFF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HD_ABI_VERSION 10

#if defined(__GNUC__)
#define HD_API __attribute__((visibility("default")))
#else
#define HD_API
#endif

/* error codes */
#define HD_OK 0
#define HD_E_INVALID (-1)     /* bad argument / unsupported configuration */
#define HD_E_CUDA (-2)        /* a CUDA runtime call failed */
#define HD_E_UNSUPPORTED (-3) /* configuration valid for the reference but not built here */

/* which kernels run the EGNN sub-layers */
#define HD_ENGINE_FP32 0      /* CUDA-core fp32 FFMA kernels (reference-order arithmetic)          */
#define HD_ENGINE_TC_STRICT 1 /* tcgen05 tensor cores, bf16x3 split operands, fp32 accumulate      */
#define HD_ENGINE_TC_FAST 2   /* tcgen05 tensor cores, single bf16 operands, fast SiLU             */
/* May be OR-ed into the `engine` argument of hd_dynamics_forward[_ctx]: a performance hint that sum(sizes) is well
 * below B*N.  The tensor-core engines then keep one workspace row per REAL node (instead of B*N rows), so the
 * per-node GEMMs skip the padding.  Results are bit-identical with and without the hint. */
#define HD_ENGINE_RAGGED_ROWS 0x100
/* May be OR-ed into the `engine` argument of hd_dynamics_forward[_ctx|_ragged]: return eps = [velocity | h] BEFORE the
 * NaN guard and the centre-of-gravity projection of en_dynamics.py:109-116 (a NaN still sets HD_FLAG_NAN).  For callers
 * that combine several node sets into one projection: pocket-conditioned dynamics (en_dynamics.py:83-88 with
 * mol_shape < n_nodes), where the ligand and the frozen pocket share one mean. */
#define HD_ENGINE_RAW_VELOCITY 0x200

/* bits of the device-side status word (`flags`) that kernels OR into */
#define HD_FLAG_NAN 1       /* en_dynamics.py:109-111 NaN guard fired (velocity zeroed)            */
#define HD_FLAG_COG 2       /* models/utils.py:65-70 assert_mean_zero_with_mask(z_t) would fail    */
#define HD_FLAG_MASK 4      /* models/utils.py:47-50,72-75 masked entries not ~0                   */

typedef void* hd_stream_t; /* cudaStream_t */

/* EGNN hyper-parameters: models/layers/egnn_new.py:156-190 (EGNN.__init__) as called by
 * models/module/en_dynamics.py:17-24.  in_node_nf counts the time channel (+context). */
typedef struct {
  int32_t n_layers;      /* EquivariantBlocks                                   */
  int32_t inv_sublayers; /* GCLs per block                                      */
  int32_t hidden_nf;     /* H; this build supports H == 256                     */
  int32_t in_node_nf;    /* features entering `embedding` (F + time + context)  */
  int32_t attention;     /* 0/1                                                 */
  int32_t tanh;          /* 0/1                                                 */
  float coords_range;    /* EGNN ctor value (30); per block = / n_layers        */
  float norm_constant;
  float normalization_factor;
  int32_t aggregation_mean; /* 0 = 'sum' (divide by normalization_factor), 1 = 'mean' */
} hd_config;

HD_API int32_t hd_abi_version(void);
/* 1 when `engine` (HD_ENGINE_*) is compiled into this library, else 0 */
HD_API int32_t hd_engine_available(int32_t engine);
/* Kernels this library has enqueued so far in this process (host-side counter; a captured launch counts
 * once, at capture time).  bench.py multiplies the per-step delta by the steps it replays. */
HD_API int64_t hd_launch_count(void);
HD_API const char* hd_last_error(void);

/* Number of floats of the flat parameter buffer: the `dynamics.egnn.*` tensors of the
 * reference state_dict concatenated in state_dict order (egnn_new.py:173-189):
 * embedding.{weight,bias}, embedding_out.{weight,bias}, then per block b:
 * gcl_k.{edge_mlp.0,edge_mlp.2,node_mlp.0,node_mlp.2,att_mlp.0}.{weight,bias} (k < inv_sublayers),
 * gcl_equiv.coord_mlp.{0,2}.{weight,bias}, coord_mlp.4.weight.  nn.Linear layout [out,in]. */
HD_API int64_t hd_weight_count(const hd_config* cfg);

/* Bytes of the packed (kernel-ready) weight image and the kernel that builds it from the flat
 * buffer: transposed fp32 copies for the FFMA engine, bf16 hi/lo operand tiles in tcgen05
 * canonical shared-memory order for the tensor-core engines.  Replaces nothing in the
 * reference (it keeps nn.Parameter tensors); call again after every load_state_dict. */
HD_API int64_t hd_packed_bytes(const hd_config* cfg);
HD_API int32_t hd_pack_weights(const hd_config* cfg, const float* w_flat, void* packed, hd_stream_t stream);

/* Scratch bytes one forward needs for a padded batch of B molecules x N nodes.  Zero-fill the buffer once after
 * allocating it: rows that belong to padded nodes are read as (masked) GEMM operand rows but never written. */
HD_API int64_t hd_workspace_bytes(const hd_config* cfg, int32_t B, int32_t N);

/* EGNN_dynamics_QM9._forward (models/module/en_dynamics.py:49-122) for mode 'egnn_dynamics',
 * condition_time=True, context=None:  eps = [velocity | h] [B,N,3+F], F = in_node_nf-1.
 * t is [B] (one time per molecule).  ORs HD_FLAG_NAN into *flags when the NaN guard fires.
 * Shape limits of the tensor-core engines: 1 <= N <= 128, 1 <= B <= 4096 (HD_E_INVALID otherwise). */
HD_API int32_t hd_dynamics_forward(const hd_config* cfg, const void* packed, const float* z, const float* t,
                            const int32_t* sizes, int32_t B, int32_t N, float* eps, void* workspace,
                            int32_t* flags, int32_t engine, hd_stream_t stream);

/* The same with conditioning (en_dynamics.py:76-79, :99-101): `context` [B,N,context_nf] is appended, unmasked, after
 * the time channel, so in_node_nf = F + 1 + context_nf; the context channels are sliced off the output again.
 * z / eps stay [B,N,3+F].  context may be NULL when context_nf == 0. */
HD_API int32_t hd_dynamics_forward_ctx(const hd_config* cfg, const void* packed, const float* z, const float* t,
                                const float* context, int32_t context_nf, const int32_t* sizes, int32_t B,
                                int32_t N, float* eps, void* workspace, int32_t* flags, int32_t engine,
                                hd_stream_t stream);

/* hd_dynamics_forward_ctx with ragged node rows (HD_ENGINE_RAGGED_ROWS) and a host-side bound: live_rows >=
 * sum(sizes) (1..B*N; 0 = no bound, as _ctx).  The per-node GEMM grids are sized for live_rows instead of B*N, so a
 * batch with much padding launches no dead CTAs at all.  A bound below sum(sizes) ORs HD_FLAG_MASK into *flags. */
HD_API int32_t hd_dynamics_forward_ragged(const hd_config* cfg, const void* packed, const float* z, const float* t,
                                   const float* context, int32_t context_nf, const int32_t* sizes, int32_t B,
                                   int32_t N, int32_t live_rows, float* eps, void* workspace, int32_t* flags,
                                   int32_t engine, hd_stream_t stream);

/* EGNN.forward (models/layers/egnn_new.py:192-205) on the canonical dense edge list of
 * en_dynamics.py:124-143.  h_in [B*N,in_node_nf], x_in [B*N,3] -> h_out [B*N,in_node_nf], x_out. */
HD_API int32_t hd_egnn_forward(const hd_config* cfg, const void* packed, const float* h_in, const float* x_in,
                        const int32_t* sizes, int32_t B, int32_t N, float* h_out, float* x_out,
                        void* workspace, int32_t engine, hd_stream_t stream);

/* One GCL (egnn_new.py:35-70) = sub-layer `sub` of block `block`, in place on h [B*N,H].
 * x [B*N,3] are the block-entry coordinates, x0 [B*N,3] the EGNN-entry coordinates
 * (edge_attr = [|x_i-x_j|^2, |x0_i-x0_j|^2], egnn_new.py:141-144,194). */
HD_API int32_t hd_gcl_forward(const hd_config* cfg, const void* packed, int32_t block, int32_t sub, float* h,
                       const float* x, const float* x0, const int32_t* sizes, int32_t B, int32_t N,
                       void* workspace, int32_t engine, hd_stream_t stream);

/* EquivariantUpdate (egnn_new.py:91-110) of block `block`: x_out = (x + sum_j trans_ij/norm)*mask. */
HD_API int32_t hd_equiv_update(const hd_config* cfg, const void* packed, int32_t block, const float* h,
                        const float* x, const float* x0, const int32_t* sizes, int32_t B, int32_t N,
                        float* x_out, void* workspace, int32_t engine, hd_stream_t stream);

/* sample_combined_position_feature_noise (train_module/diffusion_qm9.py:445-456) given the two
 * raw torch.randn draws: z = [CoG-free(randn_x*mask) | randn_h*mask]. */
HD_API int32_t hd_combine_noise(const float* randn_x, const float* randn_h, const int32_t* sizes, int32_t B,
                         int32_t N, int32_t F, float* z, hd_stream_t stream);

/* Schedule scalars of one reverse step from gamma_s, gamma_t (diffusion_qm9.py:181-204,:320-334):
 * sched[k] = {alpha_t|s, sigma2_t|s/alpha_t|s/sigma_t, sigma_t|s*sigma_s/sigma_t} for k < count. */
HD_API int32_t hd_step_scalars(const float* gamma_s, const float* gamma_t, int32_t count, float* sched,
                        hd_stream_t stream);
/* Final-decode scalars from gamma_0 (diffusion_qm9.py:294-304): {alpha_0, sigma_0, exp(0.5*gamma_0)}. */
HD_API int32_t hd_final_scalars(const float* gamma_0, int32_t count, float* sched, hd_stream_t stream);

/* sample_p_zs_given_zt after the network call (diffusion_qm9.py:328-345): re-centre eps_x,
 * mu = zt/alpha - c*eps, zs = mu + sigma*noise, re-centre zs_x.  sched is [B,3] when
 * sched_per_mol != 0, else [3].  ORs HD_FLAG_COG / HD_FLAG_MASK when the reference's
 * assert_mean_zero_with_mask(zt_x) would have raised. */
HD_API int32_t hd_reverse_step(const float* zt, const float* eps, const float* randn_x, const float* randn_h,
                        const int32_t* sizes, int32_t B, int32_t N, int32_t F, const float* sched,
                        int32_t sched_per_mol, float* zs, int32_t* flags, hd_stream_t stream);

/* sample_p_xh_given_z0 after the network call (diffusion_qm9.py:294-310, :174-179):
 * x = ((z0 - sigma_0*eps)/alpha_0 + sigma_x*noise)[..., :3]*norm_x, h = (z0_h*norm_h + bias_h)*mask. */
HD_API int32_t hd_final_decode(const float* z0, const float* eps0, const float* randn_x, const float* randn_h,
                        const int32_t* sizes, int32_t B, int32_t N, int32_t F, const float* sched,
                        int32_t sched_per_mol, float norm_x, float norm_h, float bias_h, float* x,
                        float* h, hd_stream_t stream);

/* The T-step loop of DiffusionQM9.sample (diffusion_qm9.py:361-386) with everything BETWEEN two EGNN stacks in one
 * kernel.  A chain is: hd_sampler_begin, T x hd_sampler_step, hd_sampler_final; the step index lives in `workspace`
 * (reset by _begin, advanced on the device by every _step), so one captured CUDA graph of a step serves all T.
 *   t_table [T+1]: time fed to the dynamics by the k-th executed forward (t/T for the steps, 0 for the final decode);
 *   sched_table [T+1][sched_rows][3], sched_rows = B or 1: row k < T = hd_step_scalars of the k-th executed step,
 *   row T = hd_final_scalars;  z [B,N,3+F] is updated in place;  randn_x / randn_h: this step's two raw draws.
 * _begin:  the input side of the first forward (en_dynamics.py:56-79 masking, time / context channels; egnn_new.py:197
 *          embedding, with the first sub-layer's pre-projection folded through it) from z = z_T.
 * _step:   the EGNN blocks and output head of the forward prepared by the previous call, then ONE kernel for:
 *          NaN guard + centre-of-gravity projection (en_dynamics.py:109-116), sample_p_zs_given_zt
 *          (diffusion_qm9.py:328-345, as hd_reverse_step, same status bits), step counter, the next forward's input side.
 * _final:  the last forward (t = 0) and sample_p_xh_given_z0 (diffusion_qm9.py:294-310, as hd_final_decode).
 * live_rows / HD_ENGINE_RAGGED_ROWS as hd_dynamics_forward_ragged.  N is limited by the tail kernel's shared memory
 * ((2*(3+F) + max(3+F, 12)) * N floats <= 48 KB). */
HD_API int32_t hd_sampler_begin(const hd_config* cfg, const void* packed, const float* z, const float* t_table, int32_t T,
                                const float* context, int32_t context_nf, const int32_t* sizes, int32_t B, int32_t N,
                                int32_t live_rows, void* workspace, int32_t* flags, int32_t engine, hd_stream_t stream);
HD_API int32_t hd_sampler_step(const hd_config* cfg, const void* packed, float* z, const float* randn_x,
                               const float* randn_h, const float* t_table, const float* sched_table, int32_t sched_rows,
                               int32_t T, const float* context, int32_t context_nf, const int32_t* sizes, int32_t B,
                               int32_t N, int32_t live_rows, void* workspace, int32_t* flags, int32_t engine,
                               hd_stream_t stream);
HD_API int32_t hd_sampler_final(const hd_config* cfg, const void* packed, const float* z, const float* randn_x,
                                const float* randn_h, const float* t_table, const float* sched_table,
                                int32_t sched_rows, int32_t T, const float* context, int32_t context_nf,
                                const int32_t* sizes, int32_t B, int32_t N, int32_t live_rows, float norm_x,
                                float norm_h, float bias_h, float* x, float* h, void* workspace, int32_t* flags,
                                int32_t engine, hd_stream_t stream);

/* Profiling hook: enqueue ONLY the fused edge kernel of sub-layer (block, sub) - sub == inv_sublayers selects
 * the EquivariantUpdate - on the operands a previous hd_gcl_forward / hd_equiv_update call left in `workspace`
 * (the A|B pre-projection).  Results go to workspace scratch.  Used by bench.py to time the dominant kernel
 * alone with CUDA events; not part of the sampling path. */
HD_API int32_t hd_edge_kernel_only(const hd_config* cfg, const void* packed, int32_t block, int32_t sub,
                                   const float* x, const float* x0, const int32_t* sizes, int32_t B, int32_t N,
                                   void* workspace, int32_t engine, hd_stream_t stream);

/* ---- stage-2 fine-grained decoder layer: one E_GCL of the reference ROOT models/egnn/gcl.py:9-209 as
 * models/edge_denoise.py:35-43 builds it (recurrent, agg = 'sum', coord_update, context_nf = 0, geo = False).  Replaces
 * E_GCL.forward (gcl.py:167-199).  `w` = the layer's parameters, flat fp32 in state_dict order:
 *   mes_mlp.0.{weight [H, 2H+1+De], bias}, mes_mlp.2.{weight [H,H], bias}, [edge_mlp.0.{weight [H, H+1+De], bias},
 *   edge_mlp.2.{weight, bias} if edge_update], node_mlp.0.{weight [H,2H], bias}, node_mlp.2.{weight, bias},
 *   coord_mlp.0.{weight, bias}, coord_mlp.2.weight [1,H], [att_mlp.0.{weight [1,H], bias [1]} if attention].
 * Edges: row == NULL selects the DENSE list of edge_denoise.py:506-524 (edge e = (b, i, j), row = b*N+i, col = b*N+j,
 * n_nodes = B*N, n_edges = B*N*N) with node_mask / edge_mask derived from `sizes` (prefix node masks, off-diagonal
 * edge masks; edge_mask / node_mask arguments ignored); otherwise row/col [n_edges] index the n_nodes rows of h
 * (int32 or int64 elements: index_bits = 32 / 64, the latter being torch's edge_index as it is),
 * edge_mask [n_edges] / node_mask [n_nodes] are optional float multipliers (NULL = the reference's None) and
 * sizes (NULL) / B / N are ignored; an empty list (n_edges = 0, all edge pointers NULL) is valid.  Messages aggregate over `col` (gcl.py:121).  h, h_out [n_nodes, H]; x, x_out
 * [n_nodes, 3] (x_out must not alias x); edge_attr [n_edges, De]; edge_out [n_edges, H] when edge_update (may alias
 * nothing).  edge_attr == NULL with edges_in_d = 1 and no edge update: the edge feature is |x_row - x_col|^2, computed in the
 * kernel (what edge_denoise.py:345-347, :396-398 pass to gcl_edge / gcl_denoise).  The dense reduction is deterministic;
 * the explicit list is reduced with fp32 atomics. */
typedef struct {
  int32_t hidden_nf;    /* H: input_nf = output_nf = hidden_nf */
  int32_t edges_in_d;   /* De */
  int32_t attention;
  int32_t tanh;
  float coords_range;
  int32_t edge_update;
} hd_egcl_config;
/* y [rows, out_nf] = act(x [rows, in_nf] . weight^T + bias): one nn.Linear (weight [out_nf, in_nf] row-major, bias may be
 * NULL), act 0 = none, 1 = SiLU.  The embeddings and prediction heads of Edge_denoise (edge_denoise.py:27-31, :54-56). */
HD_API int32_t hd_linear_forward(const float* x, int64_t rows, int32_t in_nf, const float* weight, const float* bias,
                                 int32_t out_nf, int32_t act, float* y, hd_stream_t stream);
HD_API int64_t hd_egcl_weight_count(const hd_egcl_config* cfg);
HD_API int64_t hd_egcl_workspace_bytes(const hd_egcl_config* cfg, int64_t n_nodes, int64_t n_edges);
/* Tensor-core engines (HD_ENGINE_TC_STRICT / _FAST) of the layer: the DENSE list with hidden_nf = edges_in_d = 256 (the
 * gcl_full_* stack, edge_denoise.py:35) runs every Linear on tcgen05 from a packed bf16 hi/lo weight image built once
 * by hd_egcl_pack_weights (hd_egcl_packed_bytes = 0: this configuration has no tensor-core path).  `packed` may be NULL
 * with HD_ENGINE_FP32; explicit edge lists always run the fp32 path (they are a few hundred edges). */
HD_API int64_t hd_egcl_packed_bytes(const hd_egcl_config* cfg);
HD_API int32_t hd_egcl_pack_weights(const hd_egcl_config* cfg, const float* w, void* packed, hd_stream_t stream);
HD_API int32_t hd_egcl_forward(const hd_egcl_config* cfg, const float* w, const void* packed, const float* h,
                               const float* x, const float* edge_attr, const void* row, const void* col,
                               int32_t index_bits, const float* edge_mask, const float* node_mask, const int32_t* sizes,
                               int32_t B, int32_t N, int64_t n_nodes, int64_t n_edges, float* h_out, float* x_out,
                               float* edge_out, void* workspace, int32_t engine, hd_stream_t stream);

/* ---- forward value of the diffusion loss / NLL: the glue of DiffusionQM9.forward -> nll ->
 * compute_loss (train_module/diffusion_qm9.py:701-751, :675-699, :530-673) around the network calls (hd_dynamics_forward
 * with one t per molecule).  No gradients - the validation / test value and the training objective's forward value.
 * hd_loss_prepare:   xh [B,N,3+F] = [ (x - CoG) / norm_x | (h - bias_h) / norm_h ] on the real nodes, 0 elsewhere
 *                    (forward :726, normalize :165-172); center = 0 skips the CoG removal; HD_FLAG_MASK as the reference's
 *                    assert (models/utils.py:47-50).
 * hd_loss_noise_mix: z = alpha(gamma[b]) * xh + sigma(gamma[b]) * eps (:573, :631); HD_FLAG_COG as :574's assert.
 * hd_loss_terms:     per molecule: error = sum (eps - net)^2 (:250-258), SNR weight (:588-593), kl_prior (:206-234), the
 *                    constants (:260-289), -log p(x,h|z0) (:460-528, from the t = 0 call when t0_always, else from the same
 *                    call for the molecules that drew t = 0) and their combination (:617-668); nll = loss - delta_log_px
 *                    (:694-697).  t_int / gamma_* are [B]; z_0 / eps_0 / net_0 may be NULL unless t0_always;
 *                    terms (optional) [B][4] = {kl_prior, estimator term, constants, L0}. */
typedef struct {
  int32_t T;            /* timesteps */
  int32_t t0_always;    /* eval: separate t = 0 network call (:617-640) */
  int32_t l2_training;  /* self.training and loss_type == 'l2' (:253, :588, :603, :657, :683) */
  int32_t int_nf;       /* integer-valued feature columns (5 'prop' / 3 'elem', :462-467) */
  int32_t cont_nf;      /* continuous feature columns (3 / 0) */
  float norm_x;         /* norm_values[0] */
  float norm_int;       /* norm_values[2] */
  float bias_int;       /* norm_biases[2] */
} hd_loss_config;
HD_API int32_t hd_loss_prepare(const float* x, const float* h, const int32_t* sizes, int32_t B, int32_t N, int32_t F,
                               float norm_x, float norm_h, float bias_h, int32_t center, float* xh, int32_t* flags,
                               hd_stream_t stream);
HD_API int32_t hd_loss_noise_mix(const float* xh, const float* eps, const float* gamma, const int32_t* sizes, int32_t B,
                                 int32_t N, int32_t F, float* z, int32_t* flags, hd_stream_t stream);
HD_API int32_t hd_loss_terms(const hd_loss_config* cfg, const float* xh, const float* z_t, const float* eps_t,
                             const float* net_t, const float* z_0, const float* eps_0, const float* net_0,
                             const float* t_int, const float* gamma_s, const float* gamma_t, const float* gamma_0,
                             const float* gamma_T, const int32_t* sizes, int32_t B, int32_t N, int32_t F, float* nll,
                             float* loss, float* error, float* terms, hd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HIERDIFF_B200_H */
