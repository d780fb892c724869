// Per-role cycle accounting of the fused edge kernel (one CTA) at the C2 shape, synthetic operands.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DHD_PHASE_TIMING -o build/edge_timing scripts/edge_timing.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hierdiff_b200/csrc/hd_tc.cu"

namespace hd {
void set_error(const char*, ...) {}
void count_launch() {}
bool pdl_enabled() { return false; }
int linear_tc(const FwdCtx&, const float*, int, int, const float*, int, int, const void*, const void*, int, int,
              const float*, float*, int, int, const float*, bool) { return 0; }
int linear_tc_v2_and_preproject(const FwdCtx&, const float*, const float*, float*, const void*, const void*, const float*,
                                const void*, const void*, const float*, float*, bool, const void*, const void*,
                                const float*, float*) { return 0; }
}

template <bool GCL, bool STRICT>
static void run(const char* name, hd::tc::Params p) {
  using namespace hd;
  long long zero[2][3][16] = {};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int it = 0; it < 3; ++it) {
#ifdef HD_PHASE_TIMING
    cudaMemcpyToSymbol(tc::g_acc, zero, sizeof(zero));
#endif
    cudaEventRecord(e0);
    for (int k = 0; k < 10; ++k) tc::launch_edge<GCL, STRICT, 2>(p, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
  }
  long long acc2[2][3][16] = {};
#ifdef HD_PHASE_TIMING
  cudaMemcpyFromSymbol(acc2, tc::g_acc, sizeof(acc2));
#else
  (void)zero;
  printf("%s: %.2f us/launch; err=%s\n", name, ms * 100.f, cudaGetErrorString(cudaGetLastError()));
#ifdef HD_TIMELINE
  {
    long long tl[2][128];
    cudaMemcpyFromSymbol(tl, tc::g_tl, sizeof(tl));
    unsigned long long sp[160][2];
    cudaMemcpyFromSymbol(sp, tc::g_span, sizeof(sp));
    unsigned long long e_min = ~0ull, e_max = 0, x_min = ~0ull, x_max = 0;
    for (int c = 0; c < 148; ++c) {
      if (sp[c][0] < e_min) e_min = sp[c][0];
      if (sp[c][0] > e_max) e_max = sp[c][0];
      if (sp[c][1] < x_min) x_min = sp[c][1];
      if (sp[c][1] > x_max) x_max = sp[c][1];
    }
    printf("  all CTAs (last launch, globaltimer ns): entry spread %llu, first exit %llu, last exit %llu after first entry; CTA 10: entry +%llu exit +%llu\n",
           e_max - e_min, x_min - e_min, x_max - e_min, sp[10][0] - e_min, sp[10][1] - e_min);
    for (int c = 0; c < 148; c += 2) printf("%s%llu-%llu", c ? " " : "   ", (sp[c][0] - e_min) / 100, (sp[c][1] - e_min) / 100);
    printf("\n");
    for (int cta = 0; cta < 2; ++cta) {
      const long long t0 = tl[cta][0];
      auto us = [&](int s) { return (tl[cta][s] - t0) / 1965.0; };
      printf("  rank %d: setup done %.2f | ranges %.2f | W ready %.2f | exit %.2f (us after entry)\n", cta, us(1), us(2), us(3), us(4));
      for (int t = 0; t < 8; ++t)
        printf("    tile %d: meta %.2f | produced %.2f | mma issued %.2f | acc seen %.2f | epilogue done %.2f\n", t, us(64 + t),
               us(16 + t), us(80 + t), us(32 + t), us(48 + t));
    }
  }
#endif
  return;
#endif
  for (int cta = 0; cta < 2; ++cta) {
  long long (*acc)[16] = acc2[cta];
  printf(" -- CTA %d (cluster rank %d)\n", 10 + cta, cta);
  printf("%s: %.2f us/launch; err=%s  (cycles per launch)\n", name, ms * 100.f, cudaGetErrorString(cudaGetLastError()));
    const char* pn[] = {"tile prologue+meta", "wait empty stage", "half steps", "publish", "tile barrier", "issue loads"};
  const char* en[] = {"wait accumulator", "pass 1", "dot exchange", "pass 2", "scratch barrier", "combine+release"};
  const char* mn[] = {"wait free acc", "wait operands", "issue"};
  long long s = 0;
  for (int i = 0; i < 6; ++i) { printf("  producer  %-20s %8lld\n", pn[i], acc[0][i] / 10); s += acc[0][i] / 10; }
  printf("  producer  total %lld\n", s); s = 0;
  for (int i = 0; i < 6; ++i) { printf("  epilogue  %-20s %8lld\n", en[i], acc[1][i] / 10); s += acc[1][i] / 10; }
  printf("  epilogue  total %lld\n", s); s = 0;
  for (int i = 0; i < 3; ++i) { printf("  mma       %-20s %8lld\n", mn[i], acc[2][i] / 10); s += acc[2][i] / 10; }
  printf("  mma       total %lld\n", s);
  }
}

int main() {
  using namespace hd;
  const int B = getenv("HD_B") ? atoi(getenv("HD_B")) : 64, N = 40, BN = B * N;
  std::vector<float> h(BN * 512);
  for (auto& v : h) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  float *ab, *x, *vec, *out;
  void* w;
  int32_t *sizes, *row_off;
  cudaMalloc(&ab, BN * 512 * 4); cudaMemcpy(ab, h.data(), BN * 512 * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&x, BN * 3 * 4); cudaMemcpy(x, h.data(), BN * 3 * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&vec, 8 * 256 * 4); cudaMemcpy(vec, h.data(), 8 * 256 * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, BN * 256 * 4);
  cudaMalloc(&w, 2 * 256 * 256 * 2); cudaMemset(w, 0x3c, 2 * 256 * 256 * 2);
  std::vector<int32_t> hs(B, N);
  cudaMalloc(&sizes, B * 4); cudaMemcpy(sizes, hs.data(), B * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&row_off, (B + 1) * 4);
  tc::plan_k<<<1, 32>>>(sizes, B, row_off);
  tc::Params p{};
  p.a_img = ab; p.kc_stride = (int64_t)BN * 16; p.b_img = ab + 16 * p.kc_stride;
  p.x = x; p.x0 = x; p.sizes = sizes; p.row_off = row_off;
  p.w_hi = w; p.w_lo = (char*)w + 256 * 256 * 2;
  p.b2 = vec; p.wa = vec + 256; p.ba = vec + 512; p.wr = vec + 768; p.wd = vec + 1024;
  p.out = out; p.B = B; p.N = N; p.attention = 1; p.use_tanh = 1; p.range = 7.5f; p.norm_constant = 0.f; p.norm_div = 10.f;
  run<true, true>("gcl strict", p);
  run<true, false>("gcl fast", p);
  run<false, true>("equiv strict", p);
  return 0;
}
