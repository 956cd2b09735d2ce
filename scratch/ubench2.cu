#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
struct S { double Minv[2][26][32]; double pad[64]; };
__global__ void k(double* out, long long* t, const double* in, int n) {
  extern __shared__ unsigned char raw[];
  S& sm = *reinterpret_cast<S*>(raw);
  int lane = threadIdx.x;
  for (int i = lane; i < 2 * 26 * 32; i += 32) (&sm.Minv[0][0][0])[i] = in[i % 97] * 1e-3;
  __syncwarp();
  double x = in[lane], y = in[lane + 32];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { x = shfl(x, (i + lane) & 31) + y; }   // dependent shfl64 + DADD
  long long t1 = clock64();
  for (int i = 0; i < n; i++) { int j = ((int)__double_as_longlong(x)) & 15; x = sm.Minv[0][j][lane] + y; }  // dependent LDS + DADD
  long long t2 = clock64();
  // friction-like loop: 26 rows, two hands
  double va[2] = {x, y}, f[2] = {0.0, 0.0}, base[2] = {in[lane + 3], in[lane + 5]}, b[2] = {in[lane + 7], in[lane + 9]};
  double R[2] = {0.3, 0.4}, ARi[2] = {0.01, 0.02}, AR[2] = {100, 50}, loss[2] = {0.01, 0.01}, impr = 0;
  for (int it = 0; it < n / 16; it++) {
    for (int i = 0; i < 26; i++) {
      const bool mine = lane == i;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        double res = base[h] + va[h];
        double xx = f[h] - res * ARi[h];
        double fn = xx < -loss[h] ? -loss[h] : (xx > loss[h] ? loss[h] : xx);
        double dl = fn - f[h];
        double dbc = shfl(dl, i);
        impr -= mine ? dl * (0.5 * dl * AR[h] + res) : 0.0;
        f[h] = mine ? fn : f[h];
        base[h] = mine ? b[h] + R[h] * fn : base[h];
        va[h] += sm.Minv[h][i][lane] * dbc;
      }
    }
  }
  long long t3 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; }
  out[blockIdx.x * 32 + lane] = x + va[0] + va[1] + impr + f[0] + f[1];
}
int main() {
  double *out, *in; long long* t; cudaMalloc(&out, 1 << 22); cudaMalloc(&t, 64); cudaMalloc(&in, 1024);
  double h_in[128]; for (int i = 0; i < 128; i++) h_in[i] = 0.001 * (i + 1); cudaMemcpy(in, h_in, 1024, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 36000);
  for (int cfg = 0; cfg < 2; cfg++) {
    int blocks = cfg == 0 ? 1 : 148 * 6, n = 4096;
    k<<<blocks, 32, 36000>>>(out, t, in, n);
    cudaDeviceSynchronize();
    long long h[3]; cudaMemcpy(h, t, 24, cudaMemcpyDeviceToHost);
    printf("blocks=%d: dep shfl64+dadd %.1f cyc/iter; dep LDS+dadd %.1f cyc/iter; friction sweep %.1f cycles per row-step (2 hands)\n", blocks,
           (double)h[0] / n, (double)h[1] / n, (double)h[2] / (n / 16 * 26));
  }
  return 0;
}
