// rp_pgs_kernel.cu — EXPERIMENTAL (not on the product path): the structure-of-arrays Gauss-Seidel solve of
// rp_pgs.cuh as a kernel, lane == environment, and a debug entry point that runs it on a host-built problem
// buffer.  Exists to measure what the split pipeline of DESIGN.md §5 would gain before the step kernel is cut
// into setup / solve / finish; rp_step never calls it.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "rp_pgs.cuh"

namespace {
// Variant with the scalar rows staged: the friction-loss and joint-limit rows, their forces and va (elements
// PG_FR .. PG_CK-1, 676 doubles per environment) live in shared memory for the whole solve, so that a sweep
// over the 104 scalar rows touches global memory only when a row commits (one column of M^-1) and for the
// contact records.  One warp per block (32 environments x 676 x 8 B = 173 KB).
struct StagedView {
  double* P;
  size_t stride;
  int env;
  double* sm;  // [PG_CK - PG_FR][32], this lane's column
  __device__ __forceinline__ double& at(int e) const {
    return (e >= rp::PG_FR && e < rp::PG_CK) ? sm[(e - rp::PG_FR) * 32] : P[(size_t)e * stride + env];
  }
  __device__ __forceinline__ double ld(int e) const {
    return (e >= rp::PG_FR && e < rp::PG_CK) ? sm[(e - rp::PG_FR) * 32] : __ldg(P + (size_t)e * stride + env);
  }
};
// global-memory view whose read-only elements go through the non-coherent path (independent of the stores)
struct GlobalView {
  double* P;
  size_t stride;
  int env;
  __device__ __forceinline__ double& at(int e) const { return P[(size_t)e * stride + env]; }
  __device__ __forceinline__ double ld(int e) const { return __ldg(P + (size_t)e * stride + env); }
};
__global__ void __launch_bounds__(64) rp_pgs_kernel(double* P, int nenv, int maxiter, double tol, double scale) {
  extern __shared__ double va_s[];  // [2*NL][blockDim.x]: lane-private columns, bank == lane
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= nenv) return;
  GlobalView v{P, (size_t)nenv, env};
  rp::pgs_solve_env(v, maxiter, tol, scale, va_s + threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(32) rp_pgs_kernel_staged(double* P, int nenv, int maxiter, double tol, double scale) {
  extern __shared__ double sh[];  // [PG_CK - PG_FR][32] staged rows, then [2*NL][32] va scratch
  constexpr int NS = rp::PG_CK - rp::PG_FR;
  const int lane = threadIdx.x, env = blockIdx.x * 32 + lane;
  if (env >= nenv) return;
  double* mine = sh + lane;
  for (int e = 0; e < NS; e++) mine[e * 32] = P[(size_t)(rp::PG_FR + e) * nenv + env];
  StagedView v{P, (size_t)nenv, env, mine};
  rp::pgs_solve_env(v, maxiter, tol, scale, sh + NS * 32 + lane, 32);
  for (int e = 0; e < NS; e++) P[(size_t)(rp::PG_FR + e) * nenv + env] = mine[e * 32];
}
}  // namespace

// P_host: [PG_SIZE][nenv] doubles (in: problem, out: solution of the last run).  The kernel is timed `repeat`
// times, each on a fresh device copy of the problem; ms_out receives the mean kernel time.
extern "C" int rp_debug_pgs_solve(double* P_host, int nenv, int maxiter, double tol, double scale, int repeat, float* ms_out, int staged) {
  const size_t bytes = sizeof(double) * (size_t)rp::PG_SIZE * nenv;
  double *P0 = nullptr, *P = nullptr;
  if (cudaMalloc((void**)&P0, bytes) != cudaSuccess || cudaMalloc((void**)&P, bytes) != cudaSuccess) { cudaFree(P0); return 1; }
  cudaMemcpy(P0, P_host, bytes, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = staged ? 32 : 64, blocks = (nenv + threads - 1) / threads;
  const size_t smem = staged ? sizeof(double) * 32 * (rp::PG_CK - rp::PG_FR + 2 * RP_NL) : sizeof(double) * 2 * RP_NL * threads;
  if (staged && cudaFuncSetAttribute(rp_pgs_kernel_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaFree(P0); cudaFree(P); return 4;
  }
  float total = 0;
  for (int r = 0; r < (repeat > 0 ? repeat : 1) + 1; r++) {  // first pass: warm-up
    cudaMemcpy(P, P0, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    if (staged) rp_pgs_kernel_staged<<<blocks, threads, smem>>>(P, nenv, maxiter, tol, scale);
    else rp_pgs_kernel<<<blocks, threads, smem>>>(P, nenv, maxiter, tol, scale);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(P0); cudaFree(P); return 2; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0) total += ms;
  }
  if (ms_out) *ms_out = total / (repeat > 0 ? repeat : 1);
  cudaMemcpy(P_host, P, bytes, cudaMemcpyDeviceToHost);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(P0); cudaFree(P);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
extern "C" int rp_debug_pgs_size(void) { return rp::PG_SIZE; }
