// Issue throughput of the conversion / fp64 instructions the NID pixel kernels are made of (B200, sm_100a):
// 16 warps per SM sub-partition-quad (512 threads/CTA, 1 CTA/SM), 4 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void thr(double* out, long long* cyc, float fa, double da, int ia, int iters) {
  float f0 = fa + threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
  double d0 = da + threadIdx.x, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3;
  int i0 = ia + threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (OP == 0) { d0 = fma(d0, da, da); d1 = fma(d1, da, da); d2 = fma(d2, da, da); d3 = fma(d3, da, da); }
      if (OP == 1) {  // F2F.F64.F32 (+ F2F.F32.F64 back to keep the chain): 2 conversions per step
        d0 = (double)f0; d1 = (double)f1; d2 = (double)f2; d3 = (double)f3;
        f0 = (float)d0 + fa; f1 = (float)d1 + fa; f2 = (float)d2 + fa; f3 = (float)d3 + fa;
      }
      if (OP == 2) {  // I2F.F64.S32 + F2I.F64
        d0 = (double)i0; d1 = (double)i1; d2 = (double)i2; d3 = (double)i3;
        i0 = (int)d0 + ia; i1 = (int)d1 + ia; i2 = (int)d2 + ia; i3 = (int)d3 + ia;
      }
      if (OP == 3) {  // magic u32 -> f64 (LOP/MOV + DADD) + F2I.F64
        d0 = __hiloint2double(0x43300000, i0) - 4503599627370496.0; d1 = __hiloint2double(0x43300000, i1) - 4503599627370496.0;
        d2 = __hiloint2double(0x43300000, i2) - 4503599627370496.0; d3 = __hiloint2double(0x43300000, i3) - 4503599627370496.0;
        i0 = __double2int_rz(d0) + ia; i1 = __double2int_rz(d1) + ia; i2 = __double2int_rz(d2) + ia; i3 = __double2int_rz(d3) + ia;
      }
      if (OP == 4) {  // F2I.F64 alone is not chainable; DADD only (reference for OP 3)
        d0 = d0 + da; d1 = d1 + da; d2 = d2 + da; d3 = d3 + da;
      }
      if (OP == 5) {  // F2F.F64.F32 only: f32 chain advanced by an FADD
        d0 += (double)f0; d1 += (double)f1; d2 += (double)f2; d3 += (double)f3;
        f0 += fa; f1 += fa; f2 += fa; f3 += fa;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3 + f0 + f1 + f2 + f3 + i0 + i1 + i2 + i3;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// gather throughput: every lane fetches a 2x2 footprint at a pseudo-random position inside a 160x120 window
__global__ void gat_tex(cudaTextureObject_t tex, int* out, long long* cyc, int iters) {
  unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
  int acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      s = s * 1664525u + 1013904223u;
      int ix = (s >> 8) % 160, iy = (s >> 20) % 120;
      uchar4 g = tex2Dgather<uchar4>(tex, (float)ix + 1.f, (float)iy + 1.f, 0);
      acc += g.x + g.y + g.z + g.w;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <typename T>
__global__ void gat_ldg(const T* __restrict__ img, int* out, long long* cyc, int iters) {
  unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
  int acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      s = s * 1664525u + 1013904223u;
      int ix = (s >> 8) % 160, iy = (s >> 20) % 120;
      T g = __ldg(img + iy * 640 + ix);
      acc += *reinterpret_cast<int*>(&g);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMallocManaged(&cyc, 8);
  const char* names[] = {"DFMA", "F2F.F64.F32 + F2F.F32.F64 + FADD", "I2F.F64.S32 + F2I.F64 + IADD", "magic u32->f64 (DADD) + F2I.F64 + IADD", "DADD", "F2F.F64.F32 + DADD + FADD"};
  const int iters = 64;
  for (int op = 0; op < 6; op++) {
    auto run = [&](auto kern) { kern<<<148, 512>>>(out, cyc, 1.5f, 1.0000001, 3, iters); cudaDeviceSynchronize(); };
    switch (op) { case 0: run(thr<0>); break; case 1: run(thr<1>); break; case 2: run(thr<2>); break; case 3: run(thr<3>); break; case 4: run(thr<4>); break; case 5: run(thr<5>); break; }
    printf("%-42s %8.2f cycles per step per warp-quad-chain (16 warps/SM, 4 chains) => %.3f steps/clk/SMSP\n", names[op],
           (double)*cyc / (iters * 8 * 4), 4.0 * (iters * 8 * 4) / (double)*cyc);
  }
  cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned char>();
  cudaMallocArray(&arr, &cd, 640, 480, cudaArrayTextureGather);
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  int* io; cudaMalloc(&io, 1 << 22);
  void* img; cudaMalloc(&img, 640 * 480 * 16); cudaMemset(img, 1, 640 * 480 * 16);
  for (int warps = 4; warps <= 32; warps *= 2) {
    gat_tex<<<148, 32 * warps>>>(tex, io, cyc, 256); cudaDeviceSynchronize();
    double a = (double)*cyc / 1024;
    gat_ldg<unsigned><<<148, 32 * warps>>>((const unsigned*)img, io, cyc, 256); cudaDeviceSynchronize();
    double b = (double)*cyc / 1024;
    gat_ldg<uint4><<<148, 32 * warps>>>((const uint4*)img, io, cyc, 256); cudaDeviceSynchronize();
    double c = (double)*cyc / 1024;
    printf("warps/SM %2d: cycles per gather per warp: TLD4 u8 %.1f | LDG.32 %.1f | LDG.128 %.1f   (err=%s)\n", warps, a, b, c, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
