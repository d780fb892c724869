/*
 * hd_oracle.h - CPU restatement of HierDiff's coarse-grained sampling hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hierdiff_b200/ may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do (as the checker / the CPU arm, never as the product).
 *
 * Parity status: PINNED.  Every function here is checked in
 * tests/test_oracle_golden.py against fixtures that tests/golden/make_golden.py
 * recorded from the unmodified reference modules (torch fp32, CPU).
 *
 * The restatement follows the reference's AS-WRITTEN operation order
 * (concatenate [h_i, h_j, r, d0] then one 514->256 Linear, etc.), not the
 * restructured order the CUDA kernels use, so that the algebraic restructure is
 * itself under test.  Each Linear is an exactly rounded fp32 op: products are
 * accumulated in double and rounded to float once.
 *
 * Reference files restated (paths relative to /root/reference/endiffusion):
 *   models/layers/egnn_new.py:35-70 (GCL), :91-110 (EquivariantUpdate),
 *   :139-152 (EquivariantBlock), :192-205 (EGNN), :260-266 (coord2diff),
 *   :269-289 (unsorted_segment_sum)
 *   models/module/en_dynamics.py:49-122 (_forward)
 *   train_module/diffusion_qm9.py:148-204 (schedule algebra), :236-248,
 *   :294-345 (final decode, reverse step), :438-456 (noise)
 *   models/utils.py:43-57 (remove_mean_with_mask), :126-135, :156-159
 *   models/noise_model.py:75-105, :163-200 (GammaNetwork)
 */
#ifndef HD_ORACLE_H
#define HD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t n_layers;        /* EquivariantBlocks                         */
  int32_t inv_sublayers;   /* GCLs per block                            */
  int32_t hidden_nf;       /* H                                         */
  int32_t in_node_nf;      /* features entering the embedding (F + time) */
  int32_t attention;       /* 0/1                                       */
  int32_t tanh;            /* 0/1                                       */
  float coords_range;      /* EGNN ctor value (30); per block = /n_layers */
  float norm_constant;
  float normalization_factor;
} hdo_config;

/* Number of floats of the flat parameter buffer: the dynamics.egnn.* tensors
 * of the reference state_dict concatenated in state_dict order
 * (embedding.{weight,bias}, embedding_out.{weight,bias}, then per block:
 * gcl_k.{edge_mlp.0,edge_mlp.2,node_mlp.0,node_mlp.2,att_mlp.0}.{weight,bias},
 * gcl_equiv.coord_mlp.{0,2}.{weight,bias}, coord_mlp.4.weight). */
int64_t hdo_weight_count(const hdo_config* c);

/* optional intermediates of hdo_egnn_forward, each [B*N, .] (may be NULL) */
typedef struct {
  float* h_embed;   /* [BN,H] after embedding (padded rows included)          */
  float* h_gcl0;    /* [BN,H] after block 0, gcl_0                             */
  float* h_gcl1;    /* [BN,H] after block 0, gcl_1                             */
  float* x_block0;  /* [BN,3] after block 0                                    */
  float* h_final;   /* [BN,in_node_nf] egnn output h                          */
  float* x_final;   /* [BN,3] egnn output x                                    */
} hdo_trace;

/* en_dynamics.py:49-122.  z [B,N,3+F], t [B], sizes [B], eps [B,N,3+F].
 * F = in_node_nf-1.  Returns 0, or 1 if the NaN guard fired. */
int hdo_dynamics_forward(const hdo_config* c, const float* w, const float* z, const float* t,
                         const int32_t* sizes, int32_t B, int32_t N, float* eps, hdo_trace* tr);
/* same with per-node context channels [B*N, C] appended after time (en_dynamics.py:76-79, :99-101) */
int hdo_dynamics_forward_ctx(const hdo_config* c, const float* w, const float* z, const float* t,
                             const float* context, int32_t C, const int32_t* sizes, int32_t B, int32_t N,
                             float* eps, hdo_trace* tr);

/* diffusion_qm9.py:181-204, :320-334 from gamma_s, gamma_t (fp32 libm):
 * out[0]=alpha_t_given_s out[1]=sigma2_t_given_s/alpha_t_given_s/sigma_t
 * out[2]=sigma_t_given_s*sigma_s/sigma_t */
void hdo_step_scalars(float gamma_s, float gamma_t, float out[3]);
/* :294-304 from gamma_0: out[0]=alpha_0 out[1]=sigma_0 out[2]=sigma_x */
void hdo_final_scalars(float gamma_0, float out[3]);

/* diffusion_qm9.py:328-345 given eps = phi(z_t,t).  randn_x [B,N,3] and
 * randn_h [B,N,F] are the two raw torch.randn draws of :449-454.  sc is [B,3]:
 * the reference's schedule scalars are [B,1,1] tensors whose rows are NOT
 * bit-identical (GammaNetwork's fp32 GEMM rounds rows differently). */
void hdo_reverse_step(const float* zt, const float* eps, const float* randn_x, const float* randn_h,
                      const int32_t* sizes, int32_t B, int32_t N, int32_t F, const float* sc, float* zs);

/* diffusion_qm9.py:294-310, :174-179.  x [B,N,3], h [B,N,F]. */
void hdo_final_decode(const float* z0, const float* eps0, const float* randn_x, const float* randn_h,
                      const int32_t* sizes, int32_t B, int32_t N, int32_t F, const float* sc,
                      float norm_x, float norm_h, float bias_h, float* x, float* h);

/* noise_model.py:163-200.  p = [gamma_0, gamma_1, l1.w, l1.b, l2.w[1024],
 * l2.b[1024], l3.w[1024], l3.b] (2+2+2048+1025 floats). */
float hdo_gamma(const float* p, float t);

#ifdef __cplusplus
}
#endif
#endif
