// "Sorted" evaluation path (the production path for cells of a few thousand pixels and more).
//
// Idea. The reference image never changes between the evaluations of one pair, and a pixel's four
// reference spline weights depend only on its 8-bit reference intensity v. nid_prepare therefore
// regroups the valid pixels of every cell by v (257 classes: 0..255, plus 256 = "no reference sample":
// valid depth but out of bounds at the prepare pose, which the CPU edge still counts in the target
// marginal, types_six_dof_expmap.cpp:593-602 with zero bs_value_ref_ rows). For a class the joint
// histogram update  P_j[k_r+m][k_t+n] += w_ref[m] * w_t[n]  (types_six_dof_expmap.cpp:598-601)
// factors into  w_ref[m] * ( sum_i w_t,i[n] ): one *un-weighted* soft histogram h_v[B] per class, 4
// accumulations per pixel instead of 20, and
//     P_j[r][t] = sum_v w_ref,v[r - k_r(v)] * h_v[t],      P_t[t] = sum_v h_v[t].
// The same factorisation turns the Jacobian's 16-term table lookup into a 4-term one against a
// per-class row  Wv[t] = V[t] + sum_k w_ref,v[k] W[k_r(v)+k][t].
//
// Work unit = "task": up to 256 consecutive pixels of one (cell, class) segment, processed by ONE WARP
// with lane-private accumulators in shared memory (no atomics: 64-bit shared atomics are CAS loops on
// sm_100, profiles/r01_atom_bench_microbenchmark.txt), merged by the warp in a fixed lane order and
// written as a partial. Partials are combined per cell in task order => results are bit-reproducible.
#include <math.h>

#include <algorithm>

#include "nid_ctx.h"
#include "nid_device.cuh"

namespace nid {

#define NID_LAUNCH_CHECK(c, what)                         \
  do {                                                    \
    (c)->launches++;                                      \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) return check_cuda(e__, what); \
  } while (0)

__device__ __forceinline__ void unpack_task(int2 t, int& start, int& count, int& cls, int& cell) {
  start = t.x;
  count = t.y & 0x1ff;
  cls = (t.y >> 9) & 0x1ff;
  cell = (t.y >> 18) & 0x3fff;
}

// ------------------------------------------------------------------------------------------------
// prepare: stable counting-sort scatter of a cell's valid pixels into (class, row-major) order.
// One CTA per cell. key: 0..255 reference intensity (in bounds at the prepare pose), 256 valid but out
// of bounds, -1 invalid depth.
__global__ void __launch_bounds__(256) k_scatter(EvalParams p, int pair, const int* __restrict__ seg_start,
                                                 double* __restrict__ sx, double* __restrict__ sy,
                                                 double* __restrict__ sz) {
  __shared__ int run[NID_NCLS];
  const int c = blockIdx.x;
  const size_t base = (size_t)pair * p.N;
  for (int k = threadIdx.x; k < NID_NCLS; k += blockDim.x) run[k] = seg_start[c * NID_NCLS + k];
  __syncthreads();
  const int r0 = (c / p.cell) * p.rb, c0 = (c % p.cell) * p.cb;
  const int npx = p.rb * p.cb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = 0; t0 < npx; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    int key = -1;
    size_t i = 0;
    if (t < npx) {
      i = base + (size_t)(r0 + t / p.cb) * p.cols + (c0 + t % p.cb);
      if (!isnan(p.pwx[i])) key = p.inb0[i] ? (int)p.im0[i] : 256;
    }
    const unsigned mask = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(mask) - 1;
    const int rank = __popc(mask & ((1u << lane) - 1u));
    const int cnt = __popc(mask);
    int pos = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
      if (warp == w && lane == leader && key >= 0) {
        pos = run[key];
        run[key] = pos + cnt;
      }
      __syncthreads();
    }
    pos = __shfl_sync(0xffffffffu, pos, leader) + rank;
    if (key >= 0) {
      sx[base + pos] = p.pwx[i];
      sy[base + pos] = p.pwy[i];
      sz[base + pos] = p.pwz[i];
    }
  }
}

// counts per (cell, class) from existing in-bounds flags (nid_import_prepare path)
__global__ void k_count_classes(EvalParams p, int pair, unsigned int* __restrict__ cnt) {
  const size_t base = (size_t)pair * p.N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.N; i += gridDim.x * blockDim.x) {
    int row = i / p.cols, col = i % p.cols;
    if (isnan(p.pwx[base + i]) || row >= p.rb * p.cell || col >= p.cb * p.cell) continue;
    int c = (row / p.rb) * p.cell + (col / p.cb);
    int key = p.inb0[base + i] ? (int)p.im0[base + i] : 256;
    atomicAdd(&cnt[((size_t)pair * p.ncell + c) * NID_NCLS + key], 1u);
  }
}

// ------------------------------------------------------------------------------------------------
// Target-image taps. TEX = true fetches 2x2 footprints with tex2Dgather from a CUDA array (one
// instruction per four taps, through the texture pipe, which leaves the L1 load/store pipe to the
// shared-memory accumulators); TEX = false uses byte loads.
// gather component order for the footprint with top-left texel (ix, iy):
//   .w = (ix, iy)  .z = (ix+1, iy)  .x = (ix, iy+1)  .y = (ix+1, iy+1)
__device__ __forceinline__ uchar4 gather2x2(cudaTextureObject_t tex, int ix, int iy) {
  return tex2Dgather<uchar4>(tex, (float)ix + 1.0f, (float)iy + 1.0f, 0);
}

template <bool TEX>
__device__ __forceinline__ double sample_center(cudaTextureObject_t tex, const uint8_t* __restrict__ im, int cols,
                                                double u, double v) {
  const int ix = (int)u, iy = (int)v;  // u, v >= 0
  const double dx = u - u2d((unsigned)ix), dy = v - u2d((unsigned)iy);
  const double dxdy = dx * dy;
  unsigned p00, p01, p10, p11;
#ifdef NID_ABL_NOTAP
  p00 = ix & 255; p01 = iy & 255; p10 = (ix + iy) & 255; p11 = (ix ^ iy) & 255;
#else
  if (TEX) {
    const uchar4 g = gather2x2(tex, ix, iy);
    p00 = g.w; p01 = g.z; p10 = g.x; p11 = g.y;
  } else {
    const uint8_t* r0 = im + (size_t)iy * cols + ix;
    const uint8_t* r1 = r0 + cols;
    p00 = __ldg(r0); p01 = __ldg(r0 + 1); p10 = __ldg(r1); p11 = __ldg(r1 + 1);
  }
#endif
  // types_six_dof_expmap.h:321-326, same term order
  return dxdy * u2d(p11) + (dy - dxdy) * u2d(p10) + (dx - dxdy) * u2d(p01) + (1.0 - dx - dy + dxdy) * u2d(p00);
}

// centre sample + central-difference gradient (types_six_dof_expmap.cpp:434-435): for u,v >= 1 the five
// bilinear samples share their fractional weights, so 12 taps (four 2x2 footprints) suffice; the first
// image row/column, where (int)(u-1) truncates towards zero, takes the literal formula.
template <bool TEX>
__device__ __forceinline__ void sample_grad(cudaTextureObject_t tex, const uint8_t* __restrict__ im, int cols, double u,
                                            double v, double& ic, double& gx, double& gy) {
  const int ix = (int)u, iy = (int)v;
  if (ix >= 1 && iy >= 1) {
    const double dx = u - u2d((unsigned)ix), dy = v - u2d((unsigned)iy);
    const double w11 = dx * dy, w10 = dy - w11, w01 = dx - w11, w00 = 1.0 - dx - dy + w11;
    int a01, a02, a10, a11, a12, a13, a20, a21, a22, a23, a31, a32;
#ifdef NID_ABL_NOTAP
    a01 = ix & 255; a02 = iy & 255; a10 = 3; a11 = (ix + iy) & 255; a12 = 7; a13 = 9; a20 = 1; a21 = 100; a22 = ix & 127; a23 = 5; a31 = 8; a32 = 77;
#else
    if (TEX) {
      const uchar4 A = gather2x2(tex, ix - 1, iy), Bq = gather2x2(tex, ix + 1, iy);
      const uchar4 C = gather2x2(tex, ix, iy - 1), D = gather2x2(tex, ix, iy + 1);
      a10 = A.w; a11 = A.z; a20 = A.x; a21 = A.y;
      a12 = Bq.w; a13 = Bq.z; a22 = Bq.x; a23 = Bq.y;
      a01 = C.w; a02 = C.z;
      a31 = D.x; a32 = D.y;
    } else {
      const uint8_t* r0 = im + (size_t)(iy - 1) * cols + (ix - 1);
      const uint8_t* r1 = r0 + cols;
      const uint8_t* r2 = r1 + cols;
      const uint8_t* r3 = r2 + cols;
      a01 = __ldg(r0 + 1); a02 = __ldg(r0 + 2);
      a10 = __ldg(r1); a11 = __ldg(r1 + 1); a12 = __ldg(r1 + 2); a13 = __ldg(r1 + 3);
      a20 = __ldg(r2); a21 = __ldg(r2 + 1); a22 = __ldg(r2 + 2); a23 = __ldg(r2 + 3);
      a31 = __ldg(r3 + 1); a32 = __ldg(r3 + 2);
    }
#endif
    ic = w11 * u2d(a22) + w10 * u2d(a21) + w01 * u2d(a12) + w00 * u2d(a11);
    gx = (w11 * i2d_small(a23 - a21) + w10 * i2d_small(a22 - a20) + w01 * i2d_small(a13 - a11) + w00 * i2d_small(a12 - a10)) * 0.5;
    gy = (w11 * i2d_small(a32 - a12) + w10 * i2d_small(a31 - a11) + w01 * i2d_small(a22 - a02) + w00 * i2d_small(a21 - a01)) * 0.5;
  } else {
    ic = interp_u8(im, cols, u, v);
    gx = (interp_u8(im, cols, u + 1.0, v) - interp_u8(im, cols, u - 1.0, v)) / 2;
    gy = (interp_u8(im, cols, u, v + 1.0) - interp_u8(im, cols, u, v - 1.0)) / 2;
  }
}

// ------------------------------------------------------------------------------------------------
// Shared front end of both passes. A warp owns `pp` consecutive tasks (contiguous in the sorted arrays):
// lane l holds the descriptor of task l, the whole range is pushed towards L2 up front, and the pixel
// loop keeps the next pixel's point in registers while the current one is processed.
struct WarpTasks {
  int first, n;  // first task index, number of tasks (<= 32)
  int2 mine;     // descriptor held by this lane
};

__device__ __forceinline__ WarpTasks warp_tasks_begin(const EvalParams& p, int pair, int lane, int wg,
                                                      const double* sx, const double* sy, const double* sz) {
  WarpTasks w;
  w.first = wg * p.pp;
  w.n = min(p.pp, p.ntasks[pair] - w.first);
  w.mine = make_int2(0, 0);
  if (w.n <= 0) return w;
  if (lane < w.n) w.mine = p.tasks[(size_t)pair * p.max_tasks + w.first + lane];
  const int s0 = __shfl_sync(0xffffffffu, w.mine.x, 0);
  const int sl = __shfl_sync(0xffffffffu, w.mine.x, w.n - 1);
  const int cl = __shfl_sync(0xffffffffu, w.mine.y, w.n - 1) & 0x1ff;
  const int s1 = sl + cl;
  for (int o = s0 + lane * 16; o < s1; o += 32 * 16) {  // one 128-byte line per lane and array
    asm volatile("prefetch.global.L2 [%0];" ::"l"(sx + o));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(sy + o));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(sz + o));
  }
  return w;
}

// ------------------------------------------------------------------------------------------------
// Pass 1: per task the un-weighted target soft histogram h[B] of its pixels, accumulated in lane-private
// shared memory (no atomics) and merged over the 32 lanes in a fixed order.
// grid (ceil(ceil(max_tasks/pp)/8), jobs), 256 threads; shared: 8 warps x B x 32 doubles + spline table.
// (A power-sum variant -- sum f^j per span, spline applied once per task -- was measured slower: its
// per-task warp reductions outweigh the saved Horner evaluations at ~3 pixels per lane and task.)
template <bool TEX>
__global__ void __launch_bounds__(256) k_hist_sorted(EvalParams p) {
  extern __shared__ double sm[];
  const int B = p.bins;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int job = blockIdx.y + p.job0;
  const int pair = p.job_pair[job];
  double* coef = sm + (size_t)W * B * 32;  // [(B-3)*16] spline polynomial table
  for (int i = threadIdx.x; i < (B - 3) * 16; i += blockDim.x) coef[i] = p.bs_coef[i];
  __syncthreads();
  const size_t base = (size_t)pair * p.N;
  const double* sxp = p.sx + base;
  const double* syp = p.sy + base;
  const double* szp = p.sz + base;
  const WarpTasks wt_ = warp_tasks_begin(p, pair, lane, blockIdx.x * W + warp, sxp, syp, szp);
  if (wt_.n <= 0) return;
  double* h = sm + (size_t)warp * B * 32;  // h[tt*32 + lane]
  const double* cp = p.cam + 4 * pair;
  const Cam cam{cp[0], cp[1], cp[2], cp[3]};
  const Pose P = load_pose(p.poses + 16 * job);
  const uint8_t* im1 = p.im1 + base;
  const cudaTextureObject_t tex = TEX ? p.tex[pair] : 0;
  const double s = (double)(B - 3) / 255.0;
  for (int j = 0; j < wt_.n; j++) {
    const int start = __shfl_sync(0xffffffffu, wt_.mine.x, j);
    const int count = __shfl_sync(0xffffffffu, wt_.mine.y, j) & 0x1ff;
    for (int tt = 0; tt < B; tt++) h[tt * 32 + lane] = 0.0;
    const double* sx = sxp + start;
    const double* sy = syp + start;
    const double* sz = szp + start;
    int i = lane;
    bool have = i < count;
    double nx = 0, ny = 0, nz = 0;
    if (have) { nx = sx[i]; ny = sy[i]; nz = sz[i]; }
    while (have) {
      const double x0 = nx, y0 = ny, z0 = nz;
      i += 32;
      have = i < count;
      if (have) { nx = sx[i]; ny = sy[i]; nz = sz[i]; }
      double x1, y1, z1, u, v;
      warp_project(P, cam, x0, y0, z0, x1, y1, z1, u, v);
      if (inb_cost(u, v, p.rows, p.cols)) {
        const double ic = clamp_intensity(sample_center<TEX>(tex, im1, p.cols, u, v));
        const double ub = ic * s;
        const int kt = (int)ub;  // ub >= 0
        double wt[4], dw[4];
        bspline4_tab<false>(coef, ub, kt, wt, dw);
#pragma unroll
        for (int n = 0; n < 4; n++) h[(kt + n) * 32 + lane] += wt[n];
      }
    }
    __syncwarp();
    // fixed-order merge of the 32 lane-private copies: lane tt sums column tt (rotated start => no bank
    // conflicts), four independent chains to shorten the dependency
    double* out = p.G + ((size_t)job * p.g_stride + wt_.first + j) * B;
    for (int tt = lane; tt < B; tt += 32) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        a0 += h[tt * 32 + ((q + tt) & 31)];
        a1 += h[tt * 32 + ((q + 1 + tt) & 31)];
        a2 += h[tt * 32 + ((q + 2 + tt) & 31)];
        a3 += h[tt * 32 + ((q + 3 + tt) & 31)];
      }
      out[tt] = (a0 + a1) + (a2 + a3);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Assembly (a7 + table half of a8): one CTA per (cell, job). Combines the task partials of the cell in
// task order into P_t and P_j,
//     P_j[r][t] = sum_v w_ref,v[r - k_r(v)] * h_v[t]   over the classes v with k_r(v) in [r-3, r]
//     P_t[t]    = sum_v h_v[t]
// normalises by n_c, computes H_t, H_j, err (computeH.cu:261-300; types_six_dof_expmap.cpp:609-635,
// .h:227) and, when want_jac, stores the scaled tables for k_qtable:
//   W[r][t] = coefJ * (1 + log2 P_j[r][t]),  V[t] = coefT * (1 + log2 P_t[t])   (0 where P < 1e-30),
//   coefJ = -(s/(n_c Hj^2)) (Ht + Href), coefT = (s/(n_c Hj^2)) Hj  (types_six_dof_expmap.cpp:486-528).
// Tasks are ordered by class, and k_r(v) is monotone in v, so every histogram row only walks the
// contiguous task range of its classes (cls_task_start).
#define NID_ASM_THREADS 512
#define NID_ASM_MAXE 8  // ceil(64*64/512)
__global__ void __launch_bounds__(NID_ASM_THREADS) k_assemble(EvalParams p, int want_jac) {
  extern __shared__ double sm[];
  __shared__ double scratch[NID_ASM_THREADS / 32];
  __shared__ int s_cts[NID_NCLS + 1];
  const int B = p.bins, BB = B * B;
  double* Pall = sm;                 // [BB + B]
  double* red = sm + BB + B;         // [NID_ASM_THREADS] partial sums of P_t
  double* hvs = red + NID_ASM_THREADS;  // [NID_NCLS][B] per-class soft histograms
  const int c = blockIdx.x, job = blockIdx.y + p.job0;
  const int pair = p.job_pair[job];
  const int nc = p.n_c[pair * p.ncell + c];
  const size_t o = (size_t)job * p.ncell + c;
  if (nc < NID_MIN_CELL_POINTS) {
    if (threadIdx.x == 0) { p.ht[o] = nan(""); p.hj[o] = nan(""); p.err[o] = nan(""); }
    return;
  }
  const int* cts = p.cls_task_start + ((size_t)pair * p.ncell + c) * (NID_NCLS + 1);
  for (int i = threadIdx.x; i <= NID_NCLS; i += blockDim.x) s_cts[i] = cts[i];
  __syncthreads();
  const double* G = p.G + (size_t)job * p.g_stride * B;
  // ---- per-class sums over the class's tasks (task order), all classes in parallel
  for (int i = threadIdx.x; i < NID_NCLS * B; i += blockDim.x) {
    const int v = i / B, tt = i % B;
    double hv = 0.0;
    for (int t = s_cts[v]; t < s_cts[v + 1]; t++) hv += G[(size_t)t * B + tt];
    hvs[i] = hv;
  }
  __syncthreads();
  // ---- P_j rows from the classes with k_r in [r-3, r] (class order)
  double ej = 0.0;
#pragma unroll
  for (int e = 0; e < NID_ASM_MAXE; e++) {
    const int idx = threadIdx.x + e * NID_ASM_THREADS;
    if (idx < BB) {
      const int r = idx / B, tt = idx % B;
      double a = 0.0;
      const int vlo = p.row_cls[2 * r], vhi = p.row_cls[2 * r + 1];
      for (int v = vlo; v < vhi; v++) a += p.lut_w[4 * v + (r - p.lut_k[v])] * hvs[v * B + tt];
      const double q = a / (double)nc;
      Pall[idx] = q;
      ej -= (q < kSigma) ? 0.0 : q * log2(q);
    }
  }
  // ---- P_t: thread (g, tt) sums classes g, g+ng, ... ; groups are then added in order
  {
    const int ng = NID_ASM_THREADS / B;
    const int g = threadIdx.x / B, tt = threadIdx.x % B;
    double a = 0.0;
    if (g < ng)
      for (int v = g; v < NID_NCLS; v += ng) a += hvs[v * B + tt];
    red[threadIdx.x] = a;
  }
  __syncthreads();
  double et = 0.0;
  if ((int)threadIdx.x < B) {
    const int ng = NID_ASM_THREADS / B;
    double a = 0.0;
    for (int g = 0; g < ng; g++) a += red[g * B + threadIdx.x];
    const double q = a / (double)nc;
    Pall[BB + threadIdx.x] = q;
    et -= (q < kSigma) ? 0.0 : q * log2(q);
  }
  const double Hj = block_sum(ej, scratch);
  const double Ht = block_sum(et, scratch);
  const double Href = p.href[pair * p.ncell + c];
  if (threadIdx.x == 0) {
    p.ht[o] = Ht; p.hj[o] = Hj;
    p.err[o] = (2 * Hj - Href - Ht) / Hj;
  }
  __syncthreads();
  if (p.hist) for (int i = threadIdx.x; i < BB + B; i += blockDim.x) p.hist[o * (BB + B) + i] = Pall[i];
  if (want_jac) {
    const double s_over = ((double)(B - 3) / 255.0) / ((double)nc * Hj * Hj);
    const double coefJ = -s_over * (Ht + Href);
    const double coefT = s_over * Hj;
    double* wv = p.wv + o * (size_t)(BB + B);
    for (int i = threadIdx.x; i < BB + B; i += blockDim.x) {
      const double q = Pall[i];
      const double L = (q < kSigma) ? 0.0 : (1.0 + log2(q));
      wv[i] = L * (i < BB ? coefJ : coefT);
    }
  }
}

// Per-class / per-span quadratic of pass 2, one thread per (class v, span k):
//   c(f) = q0 + q1 f + q2 f^2 = sum_m N'_{k+m}(k+f) * Wv[v][k+m],
//   Wv[v][t] = V[t] + sum_kk w_ref,v[kk] W[k_r(v)+kk][t]     (class 256: V only)
// grid (ceil(257*NS/256), ncell, jobs)
__global__ void __launch_bounds__(256) k_qtable(EvalParams p) {
  extern __shared__ double sm[];  // W | V | spline table
  const int B = p.bins, BB = B * B, NS = B - 3;
  const int c = blockIdx.y, job = blockIdx.z + p.job0;
  const int pair = p.job_pair[job];
  if (p.n_c[pair * p.ncell + c] < NID_MIN_CELL_POINTS) return;
  const size_t o = (size_t)job * p.ncell + c;
  const double* wvg = p.wv + o * (size_t)(BB + B);
  double* coef = sm + BB + B;
  for (int i = threadIdx.x; i < BB + B; i += blockDim.x) sm[i] = wvg[i];
  for (int i = threadIdx.x; i < NS * 16; i += blockDim.x) coef[i] = p.bs_coef[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NID_NCLS * NS) return;
  const int v = i / NS, k = i % NS;
  double wv[4];
#pragma unroll
  for (int m = 0; m < 4; m++) wv[m] = sm[BB + k + m];
  if (v < 256) {
    const int kr = p.lut_k[v];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      const double wr = p.lut_w[4 * v + kk];
#pragma unroll
      for (int m = 0; m < 4; m++) wv[m] += wr * sm[(kr + kk) * B + k + m];
    }
  }
  double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const double* cf = coef + (k * 4 + m) * 4;
    q0 += cf[1] * wv[m];
    q1 += 2.0 * cf[2] * wv[m];
    q2 += 3.0 * cf[3] * wv[m];
  }
  double* qt = p.qt + o * (size_t)(NID_NCLS * NS * 3) + 3 * (size_t)i;
  qt[0] = q0; qt[1] = q1; qt[2] = q2;
}

// ------------------------------------------------------------------------------------------------
// Pass 2: per task the partial of  J[a] = sum_i g_i[a] * c_i,  c_i = q0 + f_i (q1 + f_i q2) with the
// quadratic of the pixel's (class, span); c_i = 0 at ub == 0 exactly (the reference's BsplineDer quirk).
template <bool TEX>
__global__ void __launch_bounds__(256) k_jac_sorted(EvalParams p) {
  extern __shared__ double sm[];
  const int B = p.bins, NS = B - 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int job = blockIdx.y + p.job0;
  const int pair = p.job_pair[job];
  const size_t base = (size_t)pair * p.N;
  const double* sxp = p.sx + base;
  const double* syp = p.sy + base;
  const double* szp = p.sz + base;
  const WarpTasks wt_ = warp_tasks_begin(p, pair, lane, blockIdx.x * W + warp, sxp, syp, szp);
  if (wt_.n <= 0) return;
  double* wq = sm + warp * (NS * 3);
  const double* cp = p.cam + 4 * pair;
  const Cam cam{cp[0], cp[1], cp[2], cp[3]};
  const Pose P = load_pose(p.poses + 16 * job);
  const uint8_t* im1 = p.im1 + base;
  const cudaTextureObject_t tex = TEX ? p.tex[pair] : 0;
  const double s = (double)(B - 3) / 255.0;
  for (int j = 0; j < wt_.n; j++) {
    const int start = __shfl_sync(0xffffffffu, wt_.mine.x, j);
    const int desc = __shfl_sync(0xffffffffu, wt_.mine.y, j);
    const int count = desc & 0x1ff, cls = (desc >> 9) & 0x1ff, cell = (desc >> 18) & 0x3fff;
    {
      const double* row = p.qt + (((size_t)job * p.ncell + cell) * NID_NCLS + cls) * (NS * 3);
      for (int tt = lane; tt < NS * 3; tt += 32) wq[tt] = row[tt];
    }
    __syncwarp();
    const double* sx = sxp + start;
    const double* sy = syp + start;
    const double* sz = szp + start;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    int i = lane;
    bool have = i < count;
    double nx = 0, ny = 0, nz = 0;
    if (have) { nx = sx[i]; ny = sy[i]; nz = sz[i]; }
    while (have) {
      const double x0 = nx, y0 = ny, z0 = nz;
      i += 32;
      have = i < count;
      if (have) { nx = sx[i]; ny = sy[i]; nz = sz[i]; }
      double x, y, z, u, v;
      warp_project(P, cam, x0, y0, z0, x, y, z, u, v);
      if (inb_jac(u, v, p.rows, p.cols)) {
        double ic, gx, gy;
        sample_grad<TEX>(tex, im1, p.cols, u, v, ic, gx, gy);
        ic = clamp_intensity(ic);
        const double ub = ic * s;
        const int k = (int)ub;
        const double f = ub - u2d((unsigned)k);
        const double* q = wq + 3 * k;
        double ci = fma(f, fma(f, q[2], q[1]), q[0]);
        if (ub == 0.0) ci = 0.0;
        // d(u,v)/d(xi), types_six_dof_expmap.cpp:438-450
        const double iz = 1.0 / z, iz2 = iz * iz;
        const double a = ci * gx * cam.fx, b = ci * gy * cam.fy;
        acc[0] += a * (-x * y * iz2) + b * (-(1.0 + y * y * iz2));
        acc[1] += a * (1.0 + x * x * iz2) + b * (x * y * iz2);
        acc[2] += a * (-y * iz) + b * (x * iz);
        acc[3] += a * iz;
        acc[4] += b * iz;
        acc[5] += a * (-x * iz2) + b * (-y * iz2);
      }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double vv = acc[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) vv += __shfl_xor_sync(0xffffffffu, vv, off);
      acc[k] = vv;
    }
    if (lane < 6) {
      double vv = acc[0];
#pragma unroll
      for (int k = 1; k < 6; k++) if (lane == k) vv = acc[k];
      p.jpart[((size_t)job * p.g_stride + wt_.first + j) * 6 + lane] = vv;
    }
    __syncwarp();
  }
}

// a8 tail: one warp per (job, cell): task partials summed in a fixed order -> der[6]
__global__ void __launch_bounds__(256) k_jac_final_sorted(EvalParams p, int n_jobs) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= n_jobs * p.ncell) return;
  const int job = p.job0 + wid / p.ncell, c = wid % p.ncell;
  const int pair = p.job_pair[job];
  double* der = p.der + ((size_t)job * p.ncell + c) * 6;
  if (p.n_c[pair * p.ncell + c] < NID_MIN_CELL_POINTS) {
    if (lane < 6) der[lane] = nan("");
    return;
  }
  const int t0 = p.cell_task_start[pair * (p.ncell + 1) + c];
  const int t1 = p.cell_task_start[pair * (p.ncell + 1) + c + 1];
  const double* jp = p.jpart + (size_t)job * p.g_stride * 6;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int t = t0 + lane; t < t1; t += 32) {
#pragma unroll
    for (int k = 0; k < 6; k++) acc[k] += jp[(size_t)t * 6 + k];
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double vv = acc[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) vv += __shfl_xor_sync(0xffffffffu, vv, off);
    if (lane == 0) der[k] = vv;
  }
}

// ================================================================================================ launchers
int launch_count_classes(nid_ctx* c, int pair) {
  EvalParams p = make_params(c, 1);
  int g = (c->N + 255) / 256;
  if (g > c->sm_count * 8) g = c->sm_count * 8;
  k_count_classes<<<g, 256, 0, c->stream>>>(p, pair, c->cnt);
  NID_LAUNCH_CHECK(c, "k_count_classes");
  return NID_OK;
}

int launch_scatter(nid_ctx* c, int pair) {
  EvalParams p = make_params(c, 1);
  k_scatter<<<c->ncell, 256, 0, c->stream>>>(p, pair, c->seg_start + (size_t)pair * (c->ncell * NID_NCLS + 1), c->sx, c->sy, c->sz);
  NID_LAUNCH_CHECK(c, "k_scatter");
  return NID_OK;
}

// warps per CTA of pass 1: as many as fit the lane-private histograms in shared memory (max 8)
static int hist_warps(const nid_ctx* c) {
  const size_t per_warp = sizeof(double) * (size_t)c->bins * 32;
  const size_t fixed = sizeof(double) * (size_t)(c->bins - 3) * 16;
  int w = (int)((200 * 1024 - fixed) / per_warp);
  return std::max(1, std::min(w, 8));
}
size_t hist_sorted_smem(const nid_ctx* c) {
  return sizeof(double) * ((size_t)(c->bins - 3) * 16 + (size_t)hist_warps(c) * (size_t)c->bins * 32);
}
size_t jac_sorted_smem(const nid_ctx* c) { return sizeof(double) * 8 * (size_t)(c->bins - 3) * 3; }
size_t assemble_smem(const nid_ctx* c) { return sizeof(double) * ((size_t)c->bins * c->bins + c->bins + NID_ASM_THREADS + (size_t)NID_NCLS * c->bins); }
size_t qtable_smem(const nid_ctx* c) { return sizeof(double) * ((size_t)c->bins * c->bins + c->bins + (size_t)(c->bins - 3) * 16); }

int launch_eval_sorted(nid_ctx* c, int job0, int n_jobs, int n_jobs_total, int want_jac) {
  EvalParams p = make_params(c, n_jobs_total);
  p.job0 = job0;
  // tasks per warp: enough warps to fill the machine a few times over, long-lived warps otherwise
  {
    const long long pieces = (long long)c->max_ntasks_prepared * n_jobs;
    const long long target = (long long)c->sm_count * 16 * 4;
    long long pp = c->opt_tasks_per_warp > 0 ? c->opt_tasks_per_warp : (pieces + target / 2) / target;
    p.pp = (int)std::max(1LL, std::min(pp, 16LL));
  }
  const int nwarps = (c->max_ntasks_prepared + p.pp - 1) / p.pp;
  const int hw = hist_warps(c);
  const bool tex = c->use_tex;
  ktime_mark(c, 0);
  if (tex) k_hist_sorted<true><<<dim3((nwarps + hw - 1) / hw, n_jobs), hw * 32, hist_sorted_smem(c), c->stream>>>(p);
  else k_hist_sorted<false><<<dim3((nwarps + hw - 1) / hw, n_jobs), hw * 32, hist_sorted_smem(c), c->stream>>>(p);
  NID_LAUNCH_CHECK(c, "k_hist_sorted");
  ktime_mark(c, 1);
  k_assemble<<<dim3(c->ncell, n_jobs), NID_ASM_THREADS, assemble_smem(c), c->stream>>>(p, want_jac);
  NID_LAUNCH_CHECK(c, "k_assemble");
  if (want_jac) {
    const int items = NID_NCLS * (c->bins - 3);
    k_qtable<<<dim3((items + 255) / 256, c->ncell, n_jobs), 256, qtable_smem(c), c->stream>>>(p);
    NID_LAUNCH_CHECK(c, "k_qtable");
  }
  ktime_mark(c, 2);
  if (want_jac) {
    if (tex) k_jac_sorted<true><<<dim3((nwarps + 7) / 8, n_jobs), 256, jac_sorted_smem(c), c->stream>>>(p);
    else k_jac_sorted<false><<<dim3((nwarps + 7) / 8, n_jobs), 256, jac_sorted_smem(c), c->stream>>>(p);
    NID_LAUNCH_CHECK(c, "k_jac_sorted");
    ktime_mark(c, 3);
    const int warps = n_jobs * c->ncell;
    k_jac_final_sorted<<<(warps * 32 + 255) / 256, 256, 0, c->stream>>>(p, n_jobs);
    NID_LAUNCH_CHECK(c, "k_jac_final_sorted");
    ktime_mark(c, 4);
    const int slot[4] = {0, 3, 1, 2};
    ktime_collect(c, 4, slot);
  } else {
    const int slot[2] = {0, 3};
    ktime_collect(c, 2, slot);
  }
  return NID_OK;
}

int sorted_init(nid_ctx* c) {
  cudaError_t e = cudaFuncSetAttribute(k_hist_sorted<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_sorted_smem(c));
  if (e != cudaSuccess) return check_cuda(e, "smem attr k_hist_sorted");
  e = cudaFuncSetAttribute(k_hist_sorted<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_sorted_smem(c));
  if (e != cudaSuccess) return check_cuda(e, "smem attr k_hist_sorted");
  cudaFuncSetAttribute(k_hist_sorted<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_hist_sorted<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  e = cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)assemble_smem(c));
  if (e != cudaSuccess) return check_cuda(e, "smem attr k_assemble");
  e = cudaFuncSetAttribute(k_qtable, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qtable_smem(c));
  if (e != cudaSuccess) return check_cuda(e, "smem attr k_qtable");
  return NID_OK;
}

}  // namespace nid
