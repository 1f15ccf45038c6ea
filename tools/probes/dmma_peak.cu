// FP64 tensor (mma.sync m8n8k4 / m16n8k8) vs FP64 FMA issue-rate probe for sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dmma_k4(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}
__global__ void dmma_m16k8(double* out, int iters) {
  double a[4] = {threadIdx.x * 1e-3, 0.5, 0.25, 0.125}, b[2] = {1.0 + threadIdx.x * 1e-4, 0.3};
  double c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  }
  double s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 12345.678) out[0] = s;
}
__global__ void dfma(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-9;
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 12345.678) out[0] = s;
}
int main() {
  double* d; cudaMalloc(&d, 8);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = sms * 8, block = 256;
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); dmma_k4<<<grid, block>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    double fl = (double)grid * (block / 32) * iters * 8.0 * (2.0 * 8 * 8 * 4);
    printf("dmma m8n8k4  : %.2f TFLOP/s (%s)\n", fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaEventRecord(e0); dmma_m16k8<<<grid, block>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    fl = (double)grid * (block / 32) * iters * 8.0 * (2.0 * 16 * 8 * 8);
    printf("dmma m16n8k8 : %.2f TFLOP/s (%s)\n", fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaEventRecord(e0); dfma<<<grid, block>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    fl = (double)grid * block * iters * 16.0 * 2.0;
    printf("dfma         : %.2f TFLOP/s\n", fl / ms / 1e9);
  }
  return 0;
}
