// fp64_issue.cu -- does a warp-wide DFMA block the SMSP issue port for 2 cycles on B200, or only the fp64 pipe?
// A: 8 independent DFMA chains.  B: A + one independent integer op per DFMA.  C: A + two integer ops per DFMA.
// If B takes as long as A, non-fp64 instructions are free up to a 1:1 mix; if B ~ 1.5x A, every instruction
// costs an issue slot on top of the 2 cycles a DFMA holds the port.
#include <cstdio>
#include <cuda_runtime.h>

template <int NI>
__global__ void k(double *out, int *iout, int iters, double a, double b, int ia) {
   double x[8];
   int y[8];
#pragma unroll
   for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x * 1e-3 + j; y[j] = threadIdx.x + j; }
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
         for (int j = 0; j < 8; ++j) {
            x[j] = fma(x[j], a, b);
            if (NI >= 1) y[j] = (y[j] ^ ia) + it;
            if (NI >= 2) y[j] = (y[j] | 5) - ia;
         }
      }
   }
   double s = 0; int t = 0;
#pragma unroll
   for (int j = 0; j < 8; ++j) { s += x[j]; t += y[j]; }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int NI>
float run(int blocks, int threads, int iters, double *out, int *iout) {
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   k<NI><<<blocks, threads>>>(out, iout, iters, 1.0000001, 1e-9, 3);
   cudaEventRecord(e0);
   k<NI><<<blocks, threads>>>(out, iout, iters, 1.0000001, 1e-9, 3);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   return ms;
}

int main() {
   double *out; int *iout;
   cudaMalloc(&out, 148 * 8 * 1024 * 8); cudaMalloc(&iout, 148 * 8 * 1024 * 4);
   const int iters = 4000;
   for (int wps : {4, 8, 16}) { // warps per SM
      int threads = 128, blocks = 148 * wps * 32 / threads;
      float a = run<0>(blocks, threads, iters, out, iout), b = run<1>(blocks, threads, iters, out, iout), c = run<2>(blocks, threads, iters, out, iout);
      double dfma = (double)blocks * threads * iters * 32.0;
      printf("warps/SM %2d: DFMA only %.3f ms (%.1f DFMA/clk/SM @1.9GHz) | +1 int/DFMA %.3f ms (x%.2f) | +2 int/DFMA %.3f ms (x%.2f)\n", wps, a,
             dfma / (a * 1e-3) / 148 / 1.9e9, b, b / a, c, c / a);
   }
   return 0;
}
