// Phase timing of the node GEMM kernel (one CTA's globaltimer stamps) at the C2 shape.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DHD_PHASE_TIMING -o gpurun_out/node_timing scripts/node_timing.cu
#include <cstdio>
#include <vector>
#include "../hierdiff_b200/csrc/hd_node.cu"

int main() {
  using namespace hd;
  const int rows = 2560;
  for (int variant = 0; variant < 3; ++variant) {
    const int K1 = 256, K2 = variant == 1 ? 256 : 0, n_out = variant == 0 ? 512 : 256, mode = variant;
    float *X, *Y, *bias;
    void *whi, *wlo;
    int32_t* sizes;
    cudaMalloc(&X, rows * 512 * 4);
    cudaMalloc(&Y, rows * 512 * 4);
    cudaMalloc(&bias, 512 * 4);
    cudaMalloc(&whi, 512 * 512 * 2);
    cudaMalloc(&wlo, 512 * 512 * 2);
    cudaMalloc(&sizes, 64 * 4);
    cudaMemset(X, 0x3c, rows * 512 * 4);
    cudaMemset(Y, 0, rows * 512 * 4);
    cudaMemset(bias, 0, 512 * 4);
    cudaMemset(whi, 0, 512 * 512 * 2);
    cudaMemset(wlo, 0, 512 * 512 * 2);
    std::vector<int32_t> hs(64, 40);
    cudaMemcpy(sizes, hs.data(), 64 * 4, cudaMemcpyHostToDevice);
    lin::Params p{};
    p.X1 = X; p.X2 = X + 256; p.ld1 = 512; p.ld2 = 512; p.K1 = K1; p.K2 = K2;
    p.w_hi = whi; p.w_lo = wlo; p.bias = bias; p.Y = Y; p.ldy = n_out; p.rows = rows; p.mode = mode;
    p.resid = Y; p.sizes = sizes; p.N = 40; p.grid_rows = rows; p.B = 64;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      for (int k = 0; k < 10; ++k) { if (variant == 0) lin::launch<true, 128>(p, n_out, 0); else lin::launch<true, 64>(p, n_out, 0); }
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long ph[64];
    cudaMemcpyFromSymbol(ph, lin::g_phase, sizeof(ph));
    printf("variant %d (K=%d n_out=%d mode=%d): %.2f us/launch back-to-back; err=%s\n", variant, K1 + K2, n_out, mode,
           ms * 100.0f, cudaGetErrorString(cudaGetLastError()));
    const char* names[] = {"entry", "setup done", "-", "-", "-",
                           "-", "all published", "acc ready", "tile staged", "stores issued", "end"};
    for (int i = 1; i <= 10; ++i) if (names[i][0] != '-') printf("  %-22s +%6lld ns\n", names[i], (long long)(ph[i] - ph[0]));
    for (int c = 0; c < (K1 + K2) / 64; ++c) printf("  mma full[%d]            +%6lld ns\n", c, (long long)(ph[16 + c] - ph[0]));
  }
  return 0;
}
namespace hd {  // stubs for symbols hd_node.cu references
void set_error(const char*, ...) {}
void count_launch() {}
bool pdl_enabled() { return false; }
}
