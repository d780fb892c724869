// Launch-period microbenchmark: how long does an (almost) empty kernel with the edge kernel's launch shape take back to back?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/exp/launch_gap scripts/launch_gap.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: nothing, 1: cluster sync, 2: + TMEM alloc/dealloc (cta_group::2)
__global__ void __launch_bounds__(576, 1) k(int* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (MODE >= 2 && threadIdx.x / 32 == 16) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (MODE >= 1) {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (threadIdx.x == 0 && out) out[blockIdx.x] = 1;
  if (MODE >= 2) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    if (threadIdx.x / 32 == 16) {
      uint32_t t = *reinterpret_cast<volatile uint32_t*>(smem);
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(t) : "memory");
    }
  }
}

template <int MODE>
static void run(const char* name, int smem, int cluster, bool pdl) {
  auto kern = k<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster > 1) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; ++na; }
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0; ++na;
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    for (int i = 0; i < 200; ++i) cudaLaunchKernelEx(&cfg, kern, (int*)nullptr);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
  }
  // the same inside a CUDA graph
  cudaStream_t st; cudaStreamCreate(&st); cfg.stream = st;
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < 200; ++i) cudaLaunchKernelEx(&cfg, kern, (int*)nullptr);
  cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
  float msg = 0;
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
    cudaEventElapsedTime(&msg, e0, e1);
  }
  printf("%-44s smem %6d cluster %d pdl %d: %.2f us/launch stream, %.2f us/launch graph  (%s)\n", name, smem, cluster, pdl, ms * 5.f, msg * 5.f,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run<0>("empty", 0, 1, false);
  run<0>("empty", 0, 1, true);
  run<0>("empty + 227 KB", 232448, 1, false);
  run<0>("empty + 227 KB", 232448, 1, true);
  run<1>("cluster sync", 0, 2, false);
  run<1>("cluster sync + 227 KB", 232448, 2, false);
  run<1>("cluster sync + 227 KB", 232448, 2, true);
  run<2>("cluster + TMEM + 227 KB", 232448, 2, false);
  run<2>("cluster + TMEM + 227 KB", 232448, 2, true);
  return 0;
}
