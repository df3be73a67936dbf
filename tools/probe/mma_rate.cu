// mma_rate.cu -- issue rate of the warp-level (legacy) tensor-core MMAs on sm_100a, per SM sub-partition.
// Decides which instruction the 64-pixel training chunk should use (lbdrn_train_fp32.cuh): 3xTF32 m16n8k8 or an
// fp16 hi+lo split on m16n8k16.  Prints cycles per MMA per sub-partition for `chains` independent accumulators per warp and
// `warps` warps per sub-partition.   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if (KIND == 0)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else if (KIND == 1)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else if (KIND == 2)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else  // m16n8k8 f16 (half the k of KIND 1)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b[0]));
}

template <int KIND, int CHAINS>
__global__ void rate_kernel(long long* cycles, float* sink, int iters) {
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = 0x3c003c00u + threadIdx.x * 0;   // 1.0h pairs (or a tf32 pattern; value is irrelevant)
  b[0] = b[1] = 0x00000000u;
  float d[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c)
    for (int i = 0; i < 4; ++i) d[c][i] = (float)(threadIdx.x + c);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma<KIND>(d[c], a, b);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c)
    for (int i = 0; i < 4; ++i) s += d[c][i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND, int CHAINS>
void run(const char* name, int warps_per_smsp, long long* dc, float* sink) {
  const int iters = 4096;
  const int threads = 128 * warps_per_smsp;   // warp w sits on sub-partition w % 4
  rate_kernel<KIND, CHAINS><<<148, threads>>>(dc, sink, iters);
  rate_kernel<KIND, CHAINS><<<148, threads>>>(dc, sink, iters);
  cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
  const double per = (double)c / ((double)iters * CHAINS * warps_per_smsp);
  printf("%-18s chains=%d warps/smsp=%d : %7.2f cycles per MMA per sub-partition\n", name, CHAINS, warps_per_smsp, per);
}

int main() {
  long long* dc;
  float* sink;
  cudaMalloc(&dc, sizeof(long long));
  cudaMalloc(&sink, 148 * 1024 * sizeof(float));
  for (int w : {1, 4}) {
    run<0, 1>("tf32 m16n8k8", w, dc, sink);
    run<0, 2>("tf32 m16n8k8", w, dc, sink);
    run<0, 4>("tf32 m16n8k8", w, dc, sink);
    run<0, 8>("tf32 m16n8k8", w, dc, sink);
    run<1, 1>("f16 m16n8k16", w, dc, sink);
    run<1, 2>("f16 m16n8k16", w, dc, sink);
    run<1, 4>("f16 m16n8k16", w, dc, sink);
    run<1, 8>("f16 m16n8k16", w, dc, sink);
    run<2, 4>("bf16 m16n8k16", w, dc, sink);
    run<2, 8>("bf16 m16n8k16", w, dc, sink);
    run<3, 4>("f16 m16n8k8", w, dc, sink);
    run<3, 8>("f16 m16n8k8", w, dc, sink);
  }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
