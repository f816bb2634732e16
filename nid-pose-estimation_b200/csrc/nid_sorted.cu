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
  if (TEX) {
    const uchar4 g = gather2x2(tex, ix, iy);
    p00 = g.w; p01 = g.z; p10 = g.x; p11 = g.y;
  } else {
    const uint8_t* r0 = im + (size_t)iy * cols + ix;
    const uint8_t* r1 = r0 + cols;
    p00 = __ldg(r0); p01 = __ldg(r0 + 1); p10 = __ldg(r1); p11 = __ldg(r1 + 1);
  }
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

// Sum four per-lane values over the warp with a fixed butterfly (10 double shuffles instead of 20):
// after it, every lane holds S_0..S_3 = the warp totals of v[0..3]. Deterministic.
__device__ __forceinline__ void warp_sum4(double v[4], int lane) {
  const unsigned F = 0xffffffffu;
  const bool b0 = lane & 1, b1 = lane & 2;
  // xor 1: even lanes keep (v0,v1), odd lanes keep (v2,v3)
  double k0 = b0 ? v[2] : v[0], k1 = b0 ? v[3] : v[1];
  double s0 = b0 ? v[0] : v[2], s1 = b0 ? v[1] : v[3];
  k0 += __shfl_xor_sync(F, s0, 1);
  k1 += __shfl_xor_sync(F, s1, 1);
  // xor 2: keep one of the two
  double k = b1 ? k1 : k0, s = b1 ? k0 : k1;
  k += __shfl_xor_sync(F, s, 2);
  k += __shfl_xor_sync(F, k, 4);
  k += __shfl_xor_sync(F, k, 8);
  k += __shfl_xor_sync(F, k, 16);
  // lane l now holds the total of v[2*b0 + b1]: v0 in lane 0, v1 in lane 2, v2 in lane 1, v3 in lane 3
  v[0] = __shfl_sync(F, k, 0);
  v[1] = __shfl_sync(F, k, 2);
  v[2] = __shfl_sync(F, k, 1);
  v[3] = __shfl_sync(F, k, 3);
}

// ------------------------------------------------------------------------------------------------
// Pass 1. For the pixels of a task that fall into spline span k (ub = k + f), the four basis functions
// are cubics in f, so their sums over pixels need only the power sums S_j = sum f^j (j = 0..3):
//     h[k+m] += sum_j coef[k][m][j] * S_j[k].
// Each lane keeps the power sums of its current span in registers and spills them into lane-private
// shared memory only when its span changes; at the end of a task the touched spans are reduced over the
// warp (fixed butterfly), turned into h[B] with the polynomial table, and stored as the task's partial.
// grid (ceil(ceil(max_tasks/pp)/W), jobs), W warps per CTA.
#define NID_MROW 33  // row stride (doubles) of the lane-private moment store
template <bool TEX>
__global__ void __launch_bounds__(256) k_hist_sorted(EvalParams p) {
  extern __shared__ double sm[];
  const int B = p.bins, NS = B - 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int job = blockIdx.y + p.job0;
  const int pair = p.job_pair[job];
  double* coef = sm;                                    // [NS*16]
  double* M = sm + NS * 16 + (size_t)warp * (NS * 4 * NID_MROW + B);  // [NS*4][NID_MROW] lane-private power sums
  double* hout = M + NS * 4 * NID_MROW;                 // [B]
  for (int i = threadIdx.x; i < NS * 16; i += blockDim.x) coef[i] = p.bs_coef[i];
  for (int r = 0; r < NS * 4; r++) M[r * NID_MROW + lane] = 0.0;
  __syncthreads();
  const size_t base = (size_t)pair * p.N;
  const double* sxp = p.sx + base;
  const double* syp = p.sy + base;
  const double* szp = p.sz + base;
  const WarpTasks wt_ = warp_tasks_begin(p, pair, lane, blockIdx.x * W + warp, sxp, syp, szp);
  if (wt_.n <= 0) return;
  const double* cp = p.cam + 4 * pair;
  const Cam cam{cp[0], cp[1], cp[2], cp[3]};
  const Pose P = load_pose(p.poses + 16 * job);
  const uint8_t* im1 = p.im1 + base;
  const cudaTextureObject_t tex = TEX ? p.tex[pair] : 0;
  const double s = (double)(B - 3) / 255.0;
  for (int j = 0; j < wt_.n; j++) {
    const int start = __shfl_sync(0xffffffffu, wt_.mine.x, j);
    const int count = __shfl_sync(0xffffffffu, wt_.mine.y, j) & 0x1ff;
    for (int tt = lane; tt < B; tt += 32) hout[tt] = 0.0;
    const double* sx = sxp + start;
    const double* sy = syp + start;
    const double* sz = szp + start;
    int cur_k = -1, kmin = 1 << 20, kmax = -1;
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
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
        const int k = (int)ub;  // ub >= 0
        const double f = ub - u2d((unsigned)k);
        if (k != cur_k) {
          if (cur_k >= 0) {
            double* m = M + (cur_k * 4) * NID_MROW + lane;
            m[0] += r0; m[NID_MROW] += r1; m[2 * NID_MROW] += r2; m[3 * NID_MROW] += r3;
          }
          cur_k = k; r0 = 0.0; r1 = 0.0; r2 = 0.0; r3 = 0.0;
          kmin = min(kmin, k); kmax = max(kmax, k);
        }
        const double f2 = f * f;
        r0 += 1.0; r1 += f; r2 += f2; r3 += f2 * f;
      }
    }
    if (cur_k >= 0) {
      double* m = M + (cur_k * 4) * NID_MROW + lane;
      m[0] += r0; m[NID_MROW] += r1; m[2 * NID_MROW] += r2; m[3 * NID_MROW] += r3;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, off));
      kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, off));
    }
    __syncwarp();
    for (int k = kmin; k <= kmax; k++) {
      double* m = M + (k * 4) * NID_MROW + lane;
      double S[4] = {m[0], m[NID_MROW], m[2 * NID_MROW], m[3 * NID_MROW]};
      m[0] = 0.0; m[NID_MROW] = 0.0; m[2 * NID_MROW] = 0.0; m[3 * NID_MROW] = 0.0;
      warp_sum4(S, lane);
      if (lane < 4) {
        const double* c = coef + (k * 4 + lane) * 4;
        hout[k + lane] += c[0] * S[0] + c[1] * S[1] + c[2] * S[2] + c[3] * S[3];
      }
      __syncwarp();
    }
    double* out = p.G + ((size_t)job * p.g_stride + wt_.first + j) * B;
    for (int tt = lane; tt < B; tt += 32) out[tt] = hout[tt];
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Assembly (a7 + table half of a8): one CTA per (cell, job). Combines the task partials of the cell in
// task order into P_t and P_j, normalises by n_c, computes H_t, H_j, err (computeH.cu:261-300;
// types_six_dof_expmap.cpp:609-635, .h:227) and, when want_jac, the per-class / per-span quadratic
//   c(f) = q0 + q1 f + q2 f^2 = sum_m N'_{k+m}(k+f) * Wv[v][k+m],
//   Wv[v][t] = V[t] + sum_kk w_ref,v[kk] W[k_r(v)+kk][t],
//   W[r][t] = coefJ * (1 + log2 P_j[r][t]),  V[t] = coefT * (1 + log2 P_t[t])   (0 where P < 1e-30),
//   coefJ = -(s/(n_c Hj^2)) (Ht + Href), coefT = (s/(n_c Hj^2)) Hj  (types_six_dof_expmap.cpp:486-528),
// that pass 2 evaluates per pixel.
#define NID_ASM_BATCH 64
#define NID_ASM_MAXE 16  // ceil(64*64/256)
__global__ void __launch_bounds__(256) k_assemble(EvalParams p, int want_jac) {
  extern __shared__ double sm[];
  __shared__ double scratch[8];
  __shared__ int s_cls[NID_ASM_BATCH];
  const int B = p.bins, BB = B * B, NS = B - 3;
  double* Gs = sm;                        // [NID_ASM_BATCH][B]
  double* Pall = sm + NID_ASM_BATCH * B;  // [BB + B]
  const int c = blockIdx.x, job = blockIdx.y + p.job0;
  const int pair = p.job_pair[job];
  const int nc = p.n_c[pair * p.ncell + c];
  const size_t o = (size_t)job * p.ncell + c;
  if (nc < NID_MIN_CELL_POINTS) {
    if (threadIdx.x == 0) { p.ht[o] = nan(""); p.hj[o] = nan(""); p.err[o] = nan(""); }
    return;
  }
  const int t0 = p.cell_task_start[pair * (p.ncell + 1) + c];
  const int t1 = p.cell_task_start[pair * (p.ncell + 1) + c + 1];
  const int2* tasks = p.tasks + (size_t)pair * p.max_tasks;
  const double* G = p.G + (size_t)job * p.g_stride * B;
  double acc[NID_ASM_MAXE];
#pragma unroll
  for (int e = 0; e < NID_ASM_MAXE; e++) acc[e] = 0.0;
  double acc_t = 0.0;
  for (int b0 = t0; b0 < t1; b0 += NID_ASM_BATCH) {
    const int nb = min(NID_ASM_BATCH, t1 - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * B; i += blockDim.x) Gs[i] = G[(size_t)b0 * B + i];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_cls[i] = (tasks[b0 + i].y >> 9) & 0x1ff;
    __syncthreads();
    if ((int)threadIdx.x < B)
      for (int j = 0; j < nb; j++) acc_t += Gs[j * B + threadIdx.x];
#pragma unroll
    for (int e = 0; e < NID_ASM_MAXE; e++) {
      const int idx = threadIdx.x + e * 256;
      if (idx < BB) {
        const int r = idx / B, tt = idx % B;
        double a = acc[e];
        for (int j = 0; j < nb; j++) {
          const int v = s_cls[j];
          if (v < 256) {
            const int m = r - p.lut_k[v];
            if (m >= 0 && m < 4) a += p.lut_w[4 * v + m] * Gs[j * B + tt];
          }
        }
        acc[e] = a;
      }
    }
  }
  __syncthreads();
  double ej = 0.0, et = 0.0;
#pragma unroll
  for (int e = 0; e < NID_ASM_MAXE; e++) {
    const int idx = threadIdx.x + e * 256;
    if (idx < BB) {
      double q = acc[e] / (double)nc;
      Pall[idx] = q;
      ej -= (q < kSigma) ? 0.0 : q * log2(q);
    }
  }
  if ((int)threadIdx.x < B) {
    double q = acc_t / (double)nc;
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
    __syncthreads();
    for (int i = threadIdx.x; i < BB + B; i += blockDim.x) {
      const double q = Pall[i];
      const double L = (q < kSigma) ? 0.0 : (1.0 + log2(q));
      Pall[i] = L * (i < BB ? coefJ : coefT);
    }
    __syncthreads();
    double* qt = p.qt + o * (size_t)(NID_NCLS * NS * 3);
    for (int i = threadIdx.x; i < NID_NCLS * NS; i += blockDim.x) {
      const int v = i / NS, k = i % NS;
      double wv[4];
#pragma unroll
      for (int m = 0; m < 4; m++) wv[m] = Pall[BB + k + m];
      if (v < 256) {
        const int kr = p.lut_k[v];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const double wr = p.lut_w[4 * v + kk];
#pragma unroll
          for (int m = 0; m < 4; m++) wv[m] += wr * Pall[(kr + kk) * B + k + m];
        }
      }
      double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const double* cf = p.bs_coef + (k * 4 + m) * 4;
        q0 += cf[1] * wv[m];
        q1 += 2.0 * cf[2] * wv[m];
        q2 += 3.0 * cf[3] * wv[m];
      }
      qt[3 * i] = q0; qt[3 * i + 1] = q1; qt[3 * i + 2] = q2;
    }
  }
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

// warps per CTA of pass 1: as many as fit the lane-private moment store in shared memory (max 8)
static int hist_warps(const nid_ctx* c) {
  const size_t per_warp = sizeof(double) * ((size_t)(c->bins - 3) * 4 * NID_MROW + c->bins);
  const size_t fixed = sizeof(double) * (size_t)(c->bins - 3) * 16;
  int w = (int)((200 * 1024 - fixed) / per_warp);
  return std::max(1, std::min(w, 8));
}
size_t hist_sorted_smem(const nid_ctx* c) {
  return sizeof(double) * ((size_t)(c->bins - 3) * 16 + (size_t)hist_warps(c) * ((size_t)(c->bins - 3) * 4 * NID_MROW + c->bins));
}
size_t jac_sorted_smem(const nid_ctx* c) { return sizeof(double) * 8 * (size_t)(c->bins - 3) * 3; }
size_t assemble_smem(const nid_ctx* c) { return sizeof(double) * (NID_ASM_BATCH * c->bins + c->bins * c->bins + c->bins); }

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
  k_assemble<<<dim3(c->ncell, n_jobs), 256, assemble_smem(c), c->stream>>>(p, want_jac);
  NID_LAUNCH_CHECK(c, "k_assemble");
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
  return NID_OK;
}

}  // namespace nid
