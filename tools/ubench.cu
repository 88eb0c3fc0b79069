// tools/ubench.cu — latency microbenchmarks that shaped k_solve (not part of the product):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k_lat(double* out, long long* cyc, int n, double a, double b) {
  __shared__ int chase[1024];
  __shared__ unsigned long long bar[2];
  __shared__ double sval[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 1024; i += blockDim.x) chase[i] = (i * 17 + 1) & 1023;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0, t1;
  double x = a + tid * 1e-9;
  // 0: dependent DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) x = fma(x, b, a);
  t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  // 1: dependent DMUL
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) x = x * b;
  t1 = clock64();
  if (tid == 0) cyc[1] = t1 - t0;
  // 2: dependent rcp.approx.ftz.f64
  double y = 1.0 + x * 1e-30;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(y));
  t1 = clock64();
  if (tid == 0) cyc[2] = t1 - t0;
  // 3: shared-memory pointer chase
  int p = tid & 1023;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) p = chase[p];
  t1 = clock64();
  if (tid == 0) cyc[3] = t1 - t0;
  // 4: dependent shuffles
  double z = x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) z = __shfl_sync(0xffffffffu, z, (lane + 1) & 31);
  t1 = clock64();
  if (tid == 0) cyc[4] = t1 - t0;
  // 5: __syncthreads
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < 64; ++i) __syncthreads();
  t1 = clock64();
  if (tid == 0) cyc[5] = (t1 - t0) * (n / 64);
  // 6: mbarrier ping-pong between warp 0 and warp 1 (store value, arrive; peer waits, loads)
  __syncthreads();
  t0 = clock64();
  if (warp < 2 && blockDim.x >= 64) {
    for (int i = 0; i < 64; ++i) {
      const int ph = i & 1;
      if (warp == 0) {
        if (lane == 0) { sval[0] = x + i; asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory"); }
        asm volatile("{\n.reg .pred p;\nW0_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D0_%=;\nbra W0_%=;\nD0_%=:\n}\n" ::"r"(smem_u32(&bar[1])), "r"(ph) : "memory");
        x += sval[1];
      } else {
        asm volatile("{\n.reg .pred p;\nW1_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1_%=;\nbra W1_%=;\nD1_%=:\n}\n" ::"r"(smem_u32(&bar[0])), "r"(ph) : "memory");
        if (lane == 0) { sval[1] = sval[0] * 0.5; asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory"); }
      }
    }
  }
  t1 = clock64();
  if (tid == 0) cyc[6] = (t1 - t0) * (n / 64) / 2;   // per one-way hop
  // 7: DFMA issue: 8 independent chains
  double c0 = x, c1 = x + 1, c2 = x + 2, c3 = x + 3, c4 = x + 4, c5 = x + 5, c6 = x + 6, c7 = x + 7;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    c0 = fma(c0, b, a); c1 = fma(c1, b, a); c2 = fma(c2, b, a); c3 = fma(c3, b, a);
    c4 = fma(c4, b, a); c5 = fma(c5, b, a); c6 = fma(c6, b, a); c7 = fma(c7, b, a);
  }
  t1 = clock64();
  if (tid == 0) cyc[7] = (t1 - t0) / 8;
  out[tid] = x + y + p + z + c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 64);
  const int n = 1024;
  const char* names[8] = {"dependent DFMA", "dependent DMUL", "dependent rcp.approx.f64", "LDS pointer chase",
                          "dependent SHFL", "__syncthreads", "mbarrier hop (store+arrive -> wait+load)", "DFMA issue (8 chains)"};
  for (int threads : {64, 128, 352, 512}) {
    k_lat<<<1, threads>>>(out, cyc, n, 1.0000001, 0.9999999);
    k_lat<<<1, threads>>>(out, cyc, n, 1.0000001, 0.9999999);
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("threads %d:", threads);
    for (int i = 0; i < 8; ++i) printf("  %s %.1f", names[i], (double)h[i] / n);
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
