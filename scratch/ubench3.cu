// Throughput of the instruction classes the step kernel lives on, per SM, as a function of resident warps.
// Each warp runs 8 independent chains of one op; time by CUDA events over the whole grid.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, const double* in, int n) {
  __shared__ double sh[8][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double x[8];
  for (int i = 0; i < 8; i++) { x[i] = in[(lane + i) & 63]; sh[i][threadIdx.x & 63] = x[i]; }
  const double y = in[lane + 3], z = in[lane + 5];
  __syncthreads();
  for (int it = 0; it < n; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0) x[i] = fma(x[i], y, z);
      if (OP == 1) x[i] = x[i] + y;
      if (OP == 2) x[i] = x[i] < y ? z : x[i] + 1e-9;            // DSETP + select (+ DADD)
      if (OP == 3) x[i] = __shfl_sync(0xffffffffu, x[i], (lane + 1) & 31);
      if (OP == 4) { const int j = (int)(__double_as_longlong(x[i]) & 31); x[i] = sh[i][(j + w * 32) & 63] + 1.0; }  // LDS.64 + DADD
      if (OP == 5) x[i] = __longlong_as_double(__double_as_longlong(x[i]) + 3);  // integer pipe
      if (OP == 6) x[i] = (float)x[i] * 1.0001f;                                    // conversions (XU?)
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, double* out, double* in) {
  const int n = 20000;
  printf("%-28s", name);
  for (int wps = 2; wps <= 16; wps += (wps < 4 ? 2 : 4)) {  // warps per SM: 2, 4, 8, 12, 16
    const int blocks = 148 * wps / 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 64>>>(out, in, 100);
    cudaEventRecord(e0);
    k<OP><<<blocks, 64>>>(out, in, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 2 * 8 * n;  // warp-instructions of the measured class
    printf("  w%-2d %.3f/clk/SM", wps, ops / 148 / (ms * 1e-3 * 1.965e9));
  }
  printf("\n");
}
int main() {
  double *out, *in; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 1024);
  double h[128]; for (int i = 0; i < 128; i++) h[i] = 1.0 + 1e-3 * i; cudaMemcpy(in, h, 1024, cudaMemcpyHostToDevice);
  printf("warp-instructions per clock per SM (1.965 GHz assumed), 8 independent chains per warp\n");
  run<0>("DFMA", out, in); run<1>("DADD", out, in); run<2>("DSETP+SEL+DADD (per group)", out, in);
  run<3>("SHFL 64-bit (2 SHFL)", out, in); run<4>("LDS.64 + DADD", out, in); run<5>("IADD64 (int pipe)", out, in); run<6>("F2F/FMUL/F2F", out, in);
  return 0;
}
