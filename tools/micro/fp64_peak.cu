// FP64 throughput microbenchmark for B200: DFMA (register operands), DFMA with uniform-register
// coefficient feed, DMMA m8n8k4 and (if it compiles) m16n8k8.  Prints TFLOP/s of each.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_dfma(double *out, double a, double b)
{
   double x[8];
#pragma unroll
   for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-3 + i; }
   for (int it = 0; it < ITERS; it++)
   {
#pragma unroll
      for (int i = 0; i < 8; i++) { x[i] = fma(x[i], a, b); }
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 8; i++) { s += x[i]; }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma884(double *out, double a, double b)
{
   double c[8][2];
#pragma unroll
   for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
   for (int it = 0; it < ITERS; it++)
   {
#pragma unroll
      for (int i = 0; i < 8; i++) { dmma884(c[i][0], c[i][1], a, b); }
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 8; i++) { s += c[i][0] + c[i][1]; }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef BIG_DMMA
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2])
{
   asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__global__ void k_dmma1688(double *out, double a, double b)
{
   double c[4][4];
   double av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
#pragma unroll
   for (int i = 0; i < 4; i++) { for (int j = 0; j < 4; j++) { c[i][j] = threadIdx.x + i + j; } }
   for (int it = 0; it < ITERS; it++)
   {
#pragma unroll
      for (int i = 0; i < 4; i++) { dmma1688(c[i], av, bv); }
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 4; i++) { for (int j = 0; j < 4; j++) { s += c[i][j]; } }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

template <typename F>
float timeit(F f)
{
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   f(); cudaDeviceSynchronize();
   cudaEventRecord(e0);
   for (int i = 0; i < 5; i++) { f(); }
   cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   return ms / 5;
}

int main()
{
   int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
   const int blocks = nsm * 8, threads = 256;
   double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
   for (int warps = 1; warps <= 8; warps *= 2)
   {
      const int th = warps * 32 * 4 > 1024 ? 1024 : warps * 32 * 4;
      float ms = timeit([&] { k_dfma<<<nsm, th>>>(out, 0.999, 1e-3); });
      printf("DFMA   %2d warps/SMSP(1 block/SM x %4d thr): %7.2f TFLOP/s\n", th / 128, th,
             2.0 * 8 * ITERS * (double)nsm * th / ms * 1e-9);
      ms = timeit([&] { k_dmma884<<<nsm, th>>>(out, 0.999, 1e-3); });
      printf("DMMA884 %2d warps/SMSP                      : %7.2f TFLOP/s\n", th / 128,
             2.0 * 256 * 8 * ITERS * (double)nsm * (th / 32) / ms * 1e-9);
#ifdef BIG_DMMA
      ms = timeit([&] { k_dmma1688<<<nsm, th>>>(out, 0.999, 1e-3); });
      printf("DMMA1688 %2d warps/SMSP                     : %7.2f TFLOP/s\n", th / 128,
             2.0 * 1024 * 4 * ITERS * (double)nsm * (th / 32) / ms * 1e-9);
#endif
   }
   float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, 0.999, 1e-3); });
   printf("DFMA   full occupancy: %7.2f TFLOP/s\n", 2.0 * 8 * ITERS * (double)blocks * threads / ms * 1e-9);
   ms = timeit([&] { k_dmma884<<<blocks, threads>>>(out, 0.999, 1e-3); });
   printf("DMMA884 full occupancy: %7.2f TFLOP/s\n", 2.0 * 256 * 8 * ITERS * (double)blocks * (threads / 32) / ms * 1e-9);
   return 0;
}
