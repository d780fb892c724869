/* Minimal C host of the C ABI (include/hierdiff_b200.h): plans the buffers of one EGNN_dynamics_QM9._forward call.
 * Build:  gcc -std=c99 -Iinclude examples/c_host.c -Lhierdiff_b200/_lib -lhierdiff_b200 -Wl,-rpath,hierdiff_b200/_lib
 * The device work itself (cudaMalloc, hd_pack_weights, hd_dynamics_forward on a stream) needs the CUDA runtime; this
 * file only shows that the header is plain C and which sizes a host has to provide. */
#include <stdio.h>

#include "hierdiff_b200.h"

int main(void) {
  hd_config cfg;
  cfg.n_layers = 4;
  cfg.inv_sublayers = 2;
  cfg.hidden_nf = 256;
  cfg.in_node_nf = 9; /* 8 features + time */
  cfg.attention = 1;
  cfg.tanh = 1;
  cfg.coords_range = 30.0f;
  cfg.norm_constant = 0.0f;
  cfg.normalization_factor = 10.0f;
  cfg.aggregation_mean = 0;
  printf("abi %d, parameters %lld floats, packed image %lld bytes, workspace(B=64,N=40) %lld bytes, engines: fp32=%d "
         "strict=%d fast=%d\n",
         (int)hd_abi_version(), (long long)hd_weight_count(&cfg), (long long)hd_packed_bytes(&cfg),
         (long long)hd_workspace_bytes(&cfg, 64, 40), (int)hd_engine_available(HD_ENGINE_FP32),
         (int)hd_engine_available(HD_ENGINE_TC_STRICT), (int)hd_engine_available(HD_ENGINE_TC_FAST));
  cfg.hidden_nf = 128; /* unsupported: reported through the status / error string, never thrown */
  if (hd_packed_bytes(&cfg) < 0) printf("hidden_nf=128 -> %s\n", hd_last_error());
  return 0;
}
