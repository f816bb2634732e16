// Dependent-issue latencies and single-warp / multi-warp throughput of the fp64 instructions the NID
// kernels are made of (B200, sm_100a). One warp per SM sub-partition unless stated.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void lat(double* out, long long* cyc, double a, double b, int iters) {
  double x = a + threadIdx.x * 1e-9, y = b;
  double x2 = x + 1, x3 = x + 2, x4 = x + 3;
  int k = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (OP == 0) x = fma(x, y, a);                         // dependent DFMA
      if (OP == 1) x = x + y;                                // dependent DADD
      if (OP == 2) { x = fma(x, y, a); x2 = fma(x2, y, a); x3 = fma(x3, y, a); x4 = fma(x4, y, a); }  // 4 chains
      if (OP == 3) { k = __double2int_rz(x); x = __hiloint2double(0x43300000, k) - 4503599627370496.0 + 0.5; }  // F2I + magic
      if (OP == 4) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.5; }
      if (OP == 5) x = x * y;                                // dependent DMUL
      if (OP == 6) { x = (double)(int)x + 0.25; }            // F2I + I2F
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + x2 + x3 + x4 + k;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void smem_lat(double* out, long long* cyc, int iters) {
  __shared__ double h[16 * 256];
  for (int i = threadIdx.x; i < 16 * 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  double* row = h + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 16; j++) row[(j & 3) * 256] += 1.0;  // dependent RMW through shared memory
  }
  long long t1 = clock64();
  out[threadIdx.x] = row[0];
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void tex_lat(cudaTextureObject_t tex, int* out, long long* cyc, int iters) {
  int ix = threadIdx.x * 7 % 600, iy = threadIdx.x * 13 % 400;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    uchar4 g = tex2Dgather<uchar4>(tex, (float)ix + 1.f, (float)iy + 1.f, 0);
    ix = (ix + g.x + 1) % 600; iy = (iy + g.y + 1) % 400;  // dependent
  }
  long long t1 = clock64();
  out[threadIdx.x] = ix + iy;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMallocManaged(&cyc, 8);
  const char* names[] = {"DFMA dep", "DADD dep", "DFMA 4 chains (per 4)", "F2I.F64 + magic I2D + DADD", "MUFU.RCP64H + DADD", "DMUL dep", "F2I + I2F.F64 + DADD"};
  const int iters = 256;
  for (int warps = 1; warps <= 16; warps *= 2)
    for (int op = 0; op < 7; op++) {
      auto run = [&](auto kern) { kern<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, iters); cudaDeviceSynchronize(); };
      switch (op) { case 0: run(lat<0>); break; case 1: run(lat<1>); break; case 2: run(lat<2>); break; case 3: run(lat<3>); break; case 4: run(lat<4>); break; case 5: run(lat<5>); break; case 6: run(lat<6>); break; }
      printf("warps/CTA %2d  %-28s %7.2f cycles per op (per warp)\n", warps, names[op], (double)*cyc / (iters * 16));
    }
  smem_lat<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  printf("LDS.64 + DADD + STS.64 dependent RMW: %.2f cycles\n", (double)*cyc / (iters * 16));
  cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned char>();
  cudaMallocArray(&arr, &cd, 640, 480, cudaArrayTextureGather);
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  int* io; cudaMalloc(&io, 4096);
  tex_lat<<<1, 32>>>(tex, io, cyc, 1024); cudaDeviceSynchronize();
  printf("TLD4 (L1/L2 hit mix) dependent gather + int ops: %.1f cycles, err=%s\n", (double)*cyc / 1024, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
