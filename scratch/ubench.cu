#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* t, double a, double b, int n) {
  double x = a, y = b;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { x = fma(x, y, a); }            // dependent DFMA
  long long t1 = clock64();
  for (int i = 0; i < n; i++) { x = x + y; }                   // dependent DADD
  long long t2 = clock64();
  for (int i = 0; i < n; i++) { x = fmin(fmax(x, -y), y) + a; }  // minmax + add
  long long t3 = clock64();
  for (int i = 0; i < n; i++) { x = __shfl_sync(0xffffffffu, x, (i + 1) & 31) + a; }  // shfl double + add
  long long t4 = clock64();
  for (int i = 0; i < n; i++) { x = a / (x + 3.0); }           // division
  long long t5 = clock64();
  float xf = (float)x, yf = (float)y;
  for (int i = 0; i < n; i++) { xf = fmaf(xf, yf, 1.0f); }     // dependent FFMA
  long long t6 = clock64();
  double z0 = a, z1 = b, z2 = a + b, z3 = a - b;
  for (int i = 0; i < n; i++) { z0 = fma(z0, y, a); z1 = fma(z1, y, a); z2 = fma(z2, y, a); z3 = fma(z3, y, a); }  // 4 independent DFMA chains
  long long t7 = clock64();
  for (int i = 0; i < n; i++) { x = sqrt(x * x + 1.0); }
  long long t8 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { t[0]=t1-t0; t[1]=t2-t1; t[2]=t3-t2; t[3]=t4-t3; t[4]=t5-t4; t[5]=t6-t5; t[6]=t7-t6; t[7]=t8-t7; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + xf + z0 + z1 + z2 + z3;
}
int main() {
  double* out; long long* t; cudaMalloc(&out, 1 << 22); cudaMalloc(&t, 64);
  const char* names[] = {"dep DFMA", "dep DADD", "dep minmax+add", "dep shfl64+add", "dep ddiv+add", "dep FFMA", "4x indep DFMA (per iter)", "dep dsqrt(fma)"};
  for (int cfg = 0; cfg < 3; cfg++) {
    int blocks = cfg == 0 ? 1 : 148 * (cfg == 1 ? 1 : 8), threads = cfg == 0 ? 32 : (cfg == 1 ? 32 : 32);
    int n = 4096;
    k<<<blocks, threads>>>(out, t, 1.0000001, 0.9999999, n);
    cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, t, 64, cudaMemcpyDeviceToHost);
    printf("config: %d blocks x %d threads\n", blocks, threads);
    for (int i = 0; i < 8; i++) printf("  %-28s %.1f cycles/iter\n", names[i], (double)h[i] / n);
  }
  // throughput: many warps of independent DFMA
  return 0;
}
