// "Sorted" evaluation path (the production path for cells of a few thousand pixels and more).
//
// Idea 1 -- class factorisation. The reference image never changes between the evaluations of one pair,
// and a pixel's four reference spline weights depend only on its 8-bit reference intensity v.
// nid_prepare therefore regroups the valid pixels of every cell by v (257 classes: 0..255, plus 256 =
// "no reference sample": valid depth but out of bounds at the prepare pose, which the CPU edge still
// counts in the target marginal, types_six_dof_expmap.cpp:593-602 with zero bs_value_ref_ rows). For a
// class the joint histogram update  P_j[k_r+m][k_t+n] += w_ref[m] * w_t[n]
// (types_six_dof_expmap.cpp:598-601) factors into  w_ref[m] * ( sum_i w_t,i[n] ): one *un-weighted* soft
// histogram h_v[B] per class, 4 accumulations per pixel instead of 20, and
//     P_j[r][t] = sum_v w_ref,v[r - k_r(v)] * h_v[t],      P_t[t] = sum_v h_v[t].
// The same factorisation turns the Jacobian's 16-term table lookup into one quadratic per (class, span).
//
// Idea 2 -- one lane per task, sliced-ELL storage. A "task" is up to L consecutive pixels of one
// (cell, class) segment. 64-bit shared-memory atomics are CAS loops on sm_100
// (profiles/r01_atom_bench_microbenchmark.txt), so nothing is shared: ONE THREAD owns a task, walks its
// pixels sequentially and accumulates into a private shared-memory row (pass 1) or registers (pass 2).
// Tasks are ordered by length and packed 32 to a "slice" (one warp); the pixels of a slice are stored
// interleaved -- group g of lane l at slice_off + (g*32 + l)*4 -- so every warp load is a fully coalesced
// 128-bit access and the lanes of a warp run out of work together (SELL-C-sigma, as in sparse mat-vec).
// Each task writes one partial; partials are combined per cell in task order => no floating-point
// atomics anywhere and bit-reproducible results.
//
// Idea 3 -- a fast path that cannot change a decision. Per pixel only the depth z and the pixel index are
// stored (12 B); the camera-frame point is z * (M0 cxn + M1 cyn + M2) + M3 with M = T_cw1 * T_wc0 composed
// once per job, and one reciprocal serves the projection and the Jacobian. That differs from the
// reference's operation order (CudaPoints3d.cu:20-28, computeH.cu:152-158) by a few ulp, which is harmless
// everywhere except at the reference's discontinuities: the in-bounds tests, the (int) truncation of
// (u, v) and the `>= 255 -> 254.999` clamp on saturated plateaus. A pixel whose (u, v) lands within 2^-24
// of an integer, or whose four taps are all 255, is therefore re-evaluated with the reference's exact
// sequence (exact_uv), so every decision is bit-identical to the reference's and everything else agrees
// to ~1e-13.
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

#define NID_PAD_ID 0xFFFFFFFFu

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
// prepare: stable counting-sort scatter of a cell's valid pixels into the sliced-ELL layout. One CTA per
// cell. key: 0..255 reference intensity (in bounds at the prepare pose), 256 valid but out of bounds,
// -1 invalid depth. The r-th pixel (row-major order) of class `key` goes to pixel o = r % L of task
// cls_task_start[key] + r / L, i.e. to  task_pos[task] + (o/4)*128 + o%4.
// PTS: store the world point (pairs set from caller-supplied points); else the depth z only.
template <bool PTS>
__global__ void __launch_bounds__(256) k_scatter_sell(EvalParams p, int pair, int L, const double* __restrict__ depth,
                                                      const int* __restrict__ task_pos, double* __restrict__ sd0,
                                                      double* __restrict__ sd1, double* __restrict__ sd2,
                                                      unsigned* __restrict__ sid) {
  __shared__ int run[NID_NCLS];
  __shared__ int s_cts[NID_NCLS + 1];
  const int c = blockIdx.x;
  if (p.n_c[pair * p.ncell + c] < NID_MIN_CELL_POINTS) return;  // inactive cells have no tasks
  const size_t base = (size_t)pair * p.N;
  const int* cts = p.cls_task_start + ((size_t)pair * p.ncell + c) * (NID_NCLS + 1);
  for (int k = threadIdx.x; k < NID_NCLS; k += blockDim.x) run[k] = 0;
  for (int k = threadIdx.x; k <= NID_NCLS; k += blockDim.x) s_cts[k] = cts[k];
  __syncthreads();
  const int r0 = (c / p.cell) * p.rb, c0 = (c % p.cell) * p.cb;
  const int npx = p.rb * p.cb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t sbase = (size_t)pair * p.sell_cap;
  for (int t0 = 0; t0 < npx; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    int key = -1, row = 0, col = 0;
    size_t i = 0;
    if (t < npx) {
      row = r0 + t / p.cb; col = c0 + t % p.cb;
      i = base + (size_t)row * p.cols + col;
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
      const int task = s_cts[key] + pos / L, o = pos % L;
      const size_t dst = sbase + (size_t)task_pos[task] + (size_t)(o >> 2) * 128 + (o & 3);
      if (PTS) { sd0[dst] = p.pwx[i]; sd1[dst] = p.pwy[i]; sd2[dst] = p.pwz[i]; }
      else sd0[dst] = depth[(size_t)row * p.cols + col];
      sid[dst] = ((unsigned)row << 16) | (unsigned)col;
    }
  }
}

// Packed target texture: texel = I | (Gx+256) << 8 | (Gy+256) << 17 with the central differences
// Gx = I(x+1,y) - I(x-1,y), Gy = I(x,y+1) - I(x,y-1) (types_six_dof_expmap.cpp:434-435 before the /2);
// one 2x2 gather then holds everything the bilinear samples of I, dI/du and dI/dv need. Border texels
// use clamped neighbours and are never consumed (the first row/column takes the literal formula).
__global__ void k_pack_tex(int rows, int cols, const uint8_t* __restrict__ im, unsigned* __restrict__ out) {
  const int N = rows * cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int y = i / cols, x = i % cols;
    const int xm = max(x - 1, 0), xp = min(x + 1, cols - 1), ym = max(y - 1, 0), yp = min(y + 1, rows - 1);
    const int I = im[i];
    const int gx = (int)im[y * cols + xp] - (int)im[y * cols + xm];
    const int gy = (int)im[yp * cols + x] - (int)im[ym * cols + x];
    out[i] = (unsigned)I | ((unsigned)(gx + 256) << 8) | ((unsigned)(gy + 256) << 17);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-job geometry, built once per CTA in shared memory.
struct Geo {
  double M[12];   // fast path: T_cw1 * T_wc0 (depth form) or T_cw1 (point form), 3x4, M[3*c + r]
  double T1[12];  // T_cw1, 3x4 (exact path)
  double T0[16];  // T_wc0 column-major (exact path, depth form)
  double fx, fy, cx, cy, ifx, ify, ncx, ncy;  // ncx = -cx/fx
};

template <bool PTS>
__device__ __forceinline__ void build_geo(const EvalParams& p, int job, int pair, Geo* g) {
  if (threadIdx.x < 12) {
    const int c = threadIdx.x / 3, r = threadIdx.x % 3;
    const double* T1 = p.poses + 16 * job;
    const double* T0 = p.Twc0 + 16 * pair;
    g->T1[threadIdx.x] = T1[4 * c + r];
    double m;
    if (PTS) m = T1[4 * c + r];
    else {
      m = T1[r] * T0[4 * c] + T1[4 + r] * T0[4 * c + 1] + T1[8 + r] * T0[4 * c + 2];
      if (c == 3) m += T1[12 + r];
    }
    g->M[threadIdx.x] = m;
  } else if (threadIdx.x < 28) {
    g->T0[threadIdx.x - 12] = p.Twc0[16 * pair + threadIdx.x - 12];
  } else if (threadIdx.x == 28) {
    const double* cp = p.cam + 4 * pair;
    g->fx = cp[0]; g->fy = cp[1]; g->cx = cp[2]; g->cy = cp[3];
    g->ifx = 1.0 / cp[0]; g->ify = 1.0 / cp[1];
    g->ncx = -cp[2] / cp[0]; g->ncy = -cp[3] / cp[1];
  }
}

// The reference's exact sequence for one pixel: CudaPoints3d.cu:20-28 then computeH.cu:152-158.
template <bool PTS>
__device__ __noinline__ void exact_uv(const Geo* g, double a0, double a1, double a2, unsigned id, double* out5) {
  const Cam cam{g->fx, g->fy, g->cx, g->cy};
  double xw = a0, yw = a1, zw = a2;
  if (!PTS) backproject(g->T0, cam, a0, (int)(id >> 16), (int)(id & 0xffffu), xw, yw, zw);
  Pose P;
#pragma unroll
  for (int i = 0; i < 12; i++) P.m[i] = g->T1[i];
  double x1, y1, z1, u, v;
  warp_project(P, cam, xw, yw, zw, x1, y1, z1, u, v);
  out5[0] = x1; out5[1] = y1; out5[2] = z1; out5[3] = u; out5[4] = v;
}

// Result of the shared front end of both passes.
struct Px {
  double x, y, z, rz;  // camera-frame point and 1/z
  double dx, dy;       // fractional parts of (u, v)
  int ix, iy;
  bool cost, jac;      // in-bounds for the cost (u+3<=cols) / for the Jacobian (u+3<=cols-1)
  bool exact;          // (u, v) came from the reference's exact sequence
  bool fix;            // fast path could not decide: (u, v) within 2^-24 of an integer
};

// not near an integer: fraction in [2^-24, 1 - 2^-21)
__device__ __forceinline__ bool frac_is_safe(double d) {
  return (unsigned)(__double2hiint(d) - 0x3E700000) < (unsigned)(0x3FEFFFFF - 0x3E700000);
}

__device__ __forceinline__ void px_from_exact(const double* e, int rows, int cols, Px& r) {
  r.x = e[0]; r.y = e[1]; r.z = e[2]; r.rz = 1.0 / e[2];  // types_six_dof_expmap.cpp:437
  const double u = e[3], v = e[4];
  r.cost = inb_cost(u, v, rows, cols);
  r.jac = inb_jac(u, v, rows, cols);
  r.ix = 0; r.iy = 0; r.dx = 0.0; r.dy = 0.0;
  if (r.cost) {
    r.ix = (int)u; r.iy = (int)v;
    r.dx = u - (double)r.ix; r.dy = v - (double)r.iy;
  }
  r.exact = true;
  r.fix = false;
}

// Branch-free fast path (so that the pixels of a group interleave). `valid` = not a padding slot.
template <bool PTS>
__device__ __forceinline__ void front(const Geo* g, int rows, int cols, double a0, double a1, double a2, unsigned id,
                                      bool valid, Px& r) {
  double x1, y1, z1;
  if (PTS) {
    x1 = fma(g->M[0], a0, fma(g->M[3], a1, fma(g->M[6], a2, g->M[9])));
    y1 = fma(g->M[1], a0, fma(g->M[4], a1, fma(g->M[7], a2, g->M[10])));
    z1 = fma(g->M[2], a0, fma(g->M[5], a1, fma(g->M[8], a2, g->M[11])));
  } else {
    const double cxn = fma(u2d(id & 0xffffu), g->ifx, g->ncx);
    const double cyn = fma(u2d(id >> 16), g->ify, g->ncy);
    const double d0 = fma(g->M[0], cxn, fma(g->M[3], cyn, g->M[6]));
    const double d1 = fma(g->M[1], cxn, fma(g->M[4], cyn, g->M[7]));
    const double d2 = fma(g->M[2], cxn, fma(g->M[5], cyn, g->M[8]));
    x1 = fma(a0, d0, g->M[9]);
    y1 = fma(a0, d1, g->M[10]);
    z1 = fma(a0, d2, g->M[11]);
  }
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(z1));
  double e = fma(-z1, y, 1.0);
  y = fma(y, e, y);
  e = fma(-z1, y, 1.0);
  y = fma(y, e, y);
  const double u = fma(g->fx * x1, y, g->cx);
  const double v = fma(g->fy * y1, y, g->cy);
  r.x = x1; r.y = y1; r.z = z1; r.rz = y;
  r.exact = false;
  const int ix = __double2int_rz(u), iy = __double2int_rz(v);  // NaN -> 0, saturating
  // u <= -1 or u >= cols+1 (same for v) is out of bounds whatever the last bits are
  const bool inr = valid && (unsigned)ix <= (unsigned)cols && (unsigned)iy <= (unsigned)rows;
  const double dx = u - u2d((unsigned)ix), dy = v - u2d((unsigned)iy);
  const bool safe = frac_is_safe(dx) && frac_is_safe(dy);
  // with u, v not within 2^-24 of an integer the integer comparisons decide exactly like
  // `u>=0 && u+3<=cols && v>=0 && v+3<=rows` (types_six_dof_expmap.cpp:565; :433 with cols-1)
  r.cost = inr && safe && ix <= cols - 4 && iy <= rows - 4;
  r.jac = inr && safe && ix <= cols - 5 && iy <= rows - 4;
  r.fix = inr && !safe;
  r.ix = r.cost ? ix : 0; r.iy = r.cost ? iy : 0;
  r.dx = dx; r.dy = dy;
}

// saturated plateau: the clamp `>= 255 -> 254.999` (types_six_dof_expmap.cpp:572) depends on the last bit
// of the bilinear weights, so the fractions must be the reference's own
template <bool PTS>
__device__ __forceinline__ void make_exact(const Geo* g, int rows, int cols, double a0, double a1, double a2, unsigned id,
                                           Px& r) {
  double ex[5];
  exact_uv<PTS>(g, a0, a1, a2, id, ex);
  px_from_exact(ex, rows, cols, r);
}

// types_six_dof_expmap.h:321-326 in the reference's term order, without contraction
__device__ __forceinline__ double bilinear_ref(double dx, double dy, unsigned p00, unsigned p01, unsigned p10, unsigned p11) {
  const double dxdy = __dmul_rn(dx, dy);
  const double w00 = __dadd_rn(__dsub_rn(__dsub_rn(1.0, dx), dy), dxdy);
  double a = __dmul_rn(dxdy, u2d(p11));
  a = __dadd_rn(a, __dmul_rn(__dsub_rn(dy, dxdy), u2d(p10)));
  a = __dadd_rn(a, __dmul_rn(__dsub_rn(dx, dxdy), u2d(p01)));
  return __dadd_rn(a, __dmul_rn(w00, u2d(p00)));
}

// gather component order for the footprint with top-left texel (ix, iy):
//   .w = (ix, iy)  .z = (ix+1, iy)  .x = (ix, iy+1)  .y = (ix+1, iy+1)
__device__ __forceinline__ uchar4 gather_u8(cudaTextureObject_t tex, int ix, int iy) {
  return tex2Dgather<uchar4>(tex, (float)ix + 1.0f, (float)iy + 1.0f, 0);
}
__device__ __forceinline__ uint4 gather_u32(cudaTextureObject_t tex, int ix, int iy) {
  return tex2Dgather<uint4>(tex, (float)ix + 1.0f, (float)iy + 1.0f, 0);
}

// Cubic B-spline basis on span k, f = ub - k. Interior spans of the clamped uniform knot vector
// (2 <= k <= NS-3) share the uniform cubic basis; the two spans at either end take the per-span
// polynomial table (nid_api.cu: build_bspline_table).
__device__ __forceinline__ void bspline4_uniform(double f, double w[4]) {
  const double s6 = 1.0 / 6.0;
  w[0] = fma(f, fma(f, fma(f, -s6, 0.5), -0.5), s6);
  w[1] = fma(f * f, fma(f, 0.5, -1.0), 2.0 / 3.0);
  w[2] = fma(f, fma(f, fma(f, -0.5, 0.5), 0.5), s6);
  w[3] = f * f * f * s6;
}
__device__ __forceinline__ bool span_is_uniform(int k, int NS) { return (unsigned)(k - 2) <= (unsigned)(NS - 5); }
__device__ __forceinline__ void bspline4_edge(const double* __restrict__ coef, int k, double f, double w[4]) {
  const double2* c2 = reinterpret_cast<const double2*>(coef + (size_t)k * 16);
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const double2 a = c2[2 * m], b = c2[2 * m + 1];
    w[m] = fma(f, fma(f, fma(f, b.y, b.x), a.y), a.x);
  }
}

// ------------------------------------------------------------------------------------------------
// One group = four consecutive pixels of a lane's task (32 B of depths + 16 B of pixel ids per lane).
template <bool PTS>
struct Group {
  unsigned id[4];
  double a0[4], a1[4], a2[4];
  __device__ __forceinline__ void load(const double* q0, const double* q1, const double* q2, const unsigned* qi, size_t o) {
    const uint4 i4 = *reinterpret_cast<const uint4*>(qi + o);
    id[0] = i4.x; id[1] = i4.y; id[2] = i4.z; id[3] = i4.w;
    const double2 xa = *reinterpret_cast<const double2*>(q0 + o), xb = *reinterpret_cast<const double2*>(q0 + o + 2);
    a0[0] = xa.x; a0[1] = xa.y; a0[2] = xb.x; a0[3] = xb.y;
    if (PTS) {
      const double2 ya = *reinterpret_cast<const double2*>(q1 + o), yb = *reinterpret_cast<const double2*>(q1 + o + 2);
      const double2 za = *reinterpret_cast<const double2*>(q2 + o), zb = *reinterpret_cast<const double2*>(q2 + o + 2);
      a1[0] = ya.x; a1[1] = ya.y; a1[2] = yb.x; a1[3] = yb.y;
      a2[0] = za.x; a2[1] = za.y; a2[2] = zb.x; a2[3] = zb.y;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) { a1[j] = 0.0; a2[j] = 0.0; }
    }
  }
};

// Pass 1 on W pixels of a group at once: projection (branch-free), W gathers in flight, spline weights,
// then the accumulations in pixel order into the lane's private row h[b * 256].
template <bool PTS, int W>
__device__ __forceinline__ void hist_pixels(const EvalParams& p, const Geo* g, const Group<PTS>& G, int j0,
                                            cudaTextureObject_t tex, const double* __restrict__ coef, double s, int NS,
                                            double* __restrict__ h) {
  Px r[W];
  bool anyfix = false;
#pragma unroll
  for (int j = 0; j < W; j++) {
    front<PTS>(g, p.rows, p.cols, G.a0[j0 + j], G.a1[j0 + j], G.a2[j0 + j], G.id[j0 + j], G.id[j0 + j] != NID_PAD_ID, r[j]);
    anyfix |= r[j].fix;
  }
  if (anyfix) {
#pragma unroll
    for (int j = 0; j < W; j++)
      if (r[j].fix) make_exact<PTS>(g, p.rows, p.cols, G.a0[j0 + j], G.a1[j0 + j], G.a2[j0 + j], G.id[j0 + j], r[j]);
  }
  uchar4 t[W];
#pragma unroll
  for (int j = 0; j < W; j++) t[j] = gather_u8(tex, r[j].ix, r[j].iy);
  bool sat = false;
#pragma unroll
  for (int j = 0; j < W; j++) sat |= r[j].cost && !r[j].exact && (t[j].x & t[j].y & t[j].z & t[j].w) == 255;
  if (sat) {
#pragma unroll
    for (int j = 0; j < W; j++)
      if (r[j].cost && !r[j].exact && (t[j].x & t[j].y & t[j].z & t[j].w) == 255) {
        make_exact<PTS>(g, p.rows, p.cols, G.a0[j0 + j], G.a1[j0 + j], G.a2[j0 + j], G.id[j0 + j], r[j]);
        t[j] = gather_u8(tex, r[j].ix, r[j].iy);
      }
  }
  double wt[W][4], fr[W];
  int kt[W];
  bool edge = false;
#pragma unroll
  for (int j = 0; j < W; j++) {
    const double ic = clamp_intensity(bilinear_ref(r[j].dx, r[j].dy, t[j].w, t[j].z, t[j].x, t[j].y));
    const double ub = ic * s;
    kt[j] = r[j].cost ? (int)ub : 0;  // 0 <= ub < NS
    fr[j] = ub - u2d((unsigned)kt[j]);
    bspline4_uniform(fr[j], wt[j]);
    edge |= r[j].cost && !span_is_uniform(kt[j], NS);
  }
  if (edge) {
#pragma unroll
    for (int j = 0; j < W; j++)
      if (r[j].cost && !span_is_uniform(kt[j], NS)) bspline4_edge(coef, kt[j], fr[j], wt[j]);
  }
#pragma unroll
  for (int j = 0; j < W; j++) {
    if (!r[j].cost) continue;
    double* hk = h + kt[j] * 256;
#pragma unroll
    for (int n = 0; n < 4; n++) hk[n * 256] += wt[j][n];
  }
}

// Pass 1: per task the un-weighted target soft histogram h[B] of its pixels.
// grid (jobs, ceil(max_slices/8)), 256 threads; shared: rows [B][256] + spline table + geometry. The job
// index is the fast grid dimension and slices are ordered longest first, so the long CTAs of every job
// start first and the short ones fill the tail.
template <bool PTS, int W>
__global__ void __launch_bounds__(256, W == 1 ? 4 : (W == 2 ? 3 : 2)) k_hist_sell(EvalParams p) {
  extern __shared__ double sm[];
  __shared__ Geo geo;
  const int B = p.bins, NS = B - 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int job = blockIdx.x + p.job0;
  const int pair = p.job_pair[job];
  double* coef = sm + (size_t)B * 256;
  for (int i = threadIdx.x; i < NS * 16; i += blockDim.x) coef[i] = p.bs_coef[i];
  build_geo<PTS>(p, job, pair, &geo);
  __syncthreads();
  const int slice = blockIdx.y * 8 + warp;
  if (slice >= p.nslices[pair]) return;
  const int* so = p.sl_off + (size_t)pair * (p.max_slices + 1) + slice;
  const int off0 = so[0], ngroups = (so[1] - off0) >> 7;
  const int task = p.sl_task[((size_t)pair * p.max_slices + slice) * 32 + lane];
  double* h = sm + threadIdx.x;  // h[b * 256]
  for (int b = 0; b < B; b++) h[b * 256] = 0.0;
  const size_t sbase = (size_t)pair * p.sell_cap + off0 + lane * 4;
  const double* q0 = p.sd0 + sbase;
  const double* q1 = PTS ? p.sd1 + sbase : nullptr;
  const double* q2 = PTS ? p.sd2 + sbase : nullptr;
  const unsigned* qi = p.sid + sbase;
  const cudaTextureObject_t tex = p.tex[pair];
  const double s = (double)NS / 255.0;
  for (int gi = 0; gi < ngroups; gi++) {
    Group<PTS> G;
    G.load(q0, q1, q2, qi, (size_t)gi * 128);
#pragma unroll
    for (int j0 = 0; j0 < 4; j0 += W) hist_pixels<PTS, W>(p, &geo, G, j0, tex, coef, s, NS, h);
  }
  if (task >= 0) {
    double* out = p.G + ((size_t)job * p.g_stride + task) * B;
    for (int b = 0; b < B; b++) out[b] = h[b * 256];
  }
}

// ------------------------------------------------------------------------------------------------
// Per-class soft histograms: hvs[job][cell][v][t] = sum over the tasks of class v (task order) of G[task][t].
// One thread per output, grid (ceil(257*B/256), ncell, jobs): ~1.6 M independent short sums, full occupancy.
__global__ void __launch_bounds__(256) k_class_sum(EvalParams p) {
  const int B = p.bins;
  const int c = blockIdx.y, job = blockIdx.z + p.job0;
  const int pair = p.job_pair[job];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NID_NCLS * B) return;
  if (p.n_c[pair * p.ncell + c] < NID_MIN_CELL_POINTS) return;
  const int v = i / B, tt = i % B;
  const int* cts = p.cls_task_start + ((size_t)pair * p.ncell + c) * (NID_NCLS + 1);
  const int t0 = cts[v], t1 = cts[v + 1];
  const double* G = p.G + (size_t)job * p.g_stride * B + tt;
  double hv = 0.0;
  for (int t = t0; t < t1; t++) hv += G[(size_t)t * B];
  p.hvs[((size_t)job * p.ncell + c) * (NID_NCLS * B) + i] = hv;
}

// Assembly (a7 + table half of a8): one CTA per (cell, job). From the per-class soft histograms h_v,
//     P_j[r][t] = sum_kk sum_{v: k_r(v) = r-kk} w_ref,v[kk] * h_v[t]     (kk = 0..3, classes in order)
//     P_t[t]    = sum_v h_v[t]
// normalises by n_c, computes H_t, H_j, err (computeH.cu:261-300; types_six_dof_expmap.cpp:609-635,
// .h:227) and, when want_jac, stores the scaled tables for pass 2:
//   W[r][t] = coefJ * (1 + log2 P_j[r][t]),  V[t] = coefT * (1 + log2 P_t[t])   (0 where P < 1e-30),
//   coefJ = -(s/(n_c Hj^2)) (Ht + Href), coefT = (s/(n_c Hj^2)) Hj  (types_six_dof_expmap.cpp:486-528).
// k_r(v) is monotone in v, so the classes of a span are a contiguous range (span_start).
#define NID_ASM_THREADS 512
__global__ void __launch_bounds__(NID_ASM_THREADS, 2) k_assemble(EvalParams p, int want_jac) {
  extern __shared__ double sm[];
  __shared__ double scratch[NID_ASM_THREADS / 32];
  const int B = p.bins, BB = B * B, NS = B - 3;
  double* Pall = sm;                    // [BB + B]
  double* red = sm + BB + B;            // [NID_ASM_THREADS] partial sums of P_t
  double* hvs = red + NID_ASM_THREADS;  // [NID_NCLS][B] per-class soft histograms
  double* part = hvs + NID_NCLS * B;    // [4][BB] per-span partial sums of P_j
  double* wl = part + 4 * BB;           // [256][4] reference weights
  const int c = blockIdx.x, job = blockIdx.y + p.job0;
  const int pair = p.job_pair[job];
  const int nc = p.n_c[pair * p.ncell + c];
  const size_t o = (size_t)job * p.ncell + c;
  if (nc < NID_MIN_CELL_POINTS) {
    if (threadIdx.x == 0) { p.ht[o] = nan(""); p.hj[o] = nan(""); p.err[o] = nan(""); }
    return;
  }
  {
    const double* src = p.hvs + o * (size_t)(NID_NCLS * B);
    for (int i = threadIdx.x; i < NID_NCLS * B; i += blockDim.x) hvs[i] = src[i];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) wl[i] = p.lut_w[i];
  }
  __syncthreads();
  // ---- P_j: item (kk, r, t) sums the classes of span r-kk
  for (int it = threadIdx.x; it < 4 * BB; it += blockDim.x) {
    const int kk = it / BB, idx = it % BB;
    const int r = idx / B, tt = idx % B;
    const int k = r - kk;
    double a = 0.0;
    if (k >= 0 && k < NS) {
      const int vlo = p.span_start[k], vhi = p.span_start[k + 1];
      for (int v = vlo; v < vhi; v++) a += wl[4 * v + kk] * hvs[v * B + tt];
    }
    part[it] = a;
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
  double ej = 0.0, et = 0.0;
  for (int idx = threadIdx.x; idx < BB; idx += blockDim.x) {
    const double a = ((part[idx] + part[BB + idx]) + part[2 * BB + idx]) + part[3 * BB + idx];
    const double q = a / (double)nc;
    Pall[idx] = q;
    ej -= (q < kSigma) ? 0.0 : q * log2(q);
  }
  if ((int)threadIdx.x >= NID_ASM_THREADS - B) {  // the last B threads (idle in the loop above for B <= 22)
    const int tt = threadIdx.x - (NID_ASM_THREADS - B);
    const int ng = NID_ASM_THREADS / B;
    double a = 0.0;
    for (int g = 0; g < ng; g++) a += red[g * B + tt];
    const double q = a / (double)nc;
    Pall[BB + tt] = q;
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

// ------------------------------------------------------------------------------------------------
// Pass 2: per task the partial of  J[a] = sum_i g_i[a] * c_i,  c_i = q0 + f_i (q1 + f_i q2) with the
// quadratic of the pixel's (class, span); c_i = 0 at ub == 0 exactly (the reference's BsplineDer quirk).
// grid (ceil(max_slices/4), jobs), 128 threads; shared: the lanes' quadratic rows [3*NS][128].
__device__ __forceinline__ double biased9_to_double(unsigned v) {  // v in [0, 511] -> (double)(v - 256), exact
  return __hiloint2double(0x43300000, (int)v) - (4503599627370496.0 + 256.0);
}

// Pass-2 pixel by the reference's literal sequence: exact (u, v) for the cached intensity
// (types_six_dof_expmap.cpp:562-575), the second projection fx*(x/z)+cx for the Jacobian bounds test and the
// four gradient samples (:407-435). Taken when (u, v) is within 2^-24 of an integer, on saturated plateaus,
// and in the first image row / column, where (int)(u-1) truncates towards zero so that the five bilinear
// samples do not share their fractions. out: {x, y, z, 1/z, ic, 2 gx, 2 gy}; returns the Jacobian validity.
template <bool PTS>
__device__ __noinline__ bool jac_pixel_literal(const Geo* g, int rows, int cols, const uint8_t* __restrict__ im,
                                               double a0, double a1, double a2, unsigned id, double* out7) {
  double e[5];
  exact_uv<PTS>(g, a0, a1, a2, id, e);
  const Cam cam{g->fx, g->fy, g->cx, g->cy};
  const double u = e[3], v = e[4];
  double u2, v2;
  project_jac(cam, e[0], e[1], e[2], u2, v2);
  if (!inb_cost(u, v, rows, cols) || !inb_jac(u2, v2, rows, cols)) return false;
  out7[0] = e[0]; out7[1] = e[1]; out7[2] = e[2]; out7[3] = 1.0 / e[2];
  out7[4] = interp_u8(im, cols, u, v);
  out7[5] = interp_u8(im, cols, u2 + 1.0, v2) - interp_u8(im, cols, u2 - 1.0, v2);
  out7[6] = interp_u8(im, cols, u2, v2 + 1.0) - interp_u8(im, cols, u2, v2 - 1.0);
  return true;
}

// Pass 2 on W pixels of a group at once; acc[6] are the lane's Jacobian partial sums.
template <bool PTS, int W>
__device__ __forceinline__ void jac_pixels(const EvalParams& p, const Geo* g, const Group<PTS>& G, int j0,
                                           cudaTextureObject_t tex2, const uint8_t* __restrict__ im1, double s, double hfx,
                                           double hfy, const double* __restrict__ wq, double acc[6]) {
  Px r[W];
#pragma unroll
  for (int j = 0; j < W; j++)
    front<PTS>(g, p.rows, p.cols, G.a0[j0 + j], G.a1[j0 + j], G.a2[j0 + j], G.id[j0 + j], G.id[j0 + j] != NID_PAD_ID, r[j]);
  uint4 t[W];
#pragma unroll
  for (int j = 0; j < W; j++) t[j] = gather_u32(tex2, r[j].jac ? r[j].ix : 0, r[j].jac ? r[j].iy : 0);
#pragma unroll
  for (int j = 0; j < W; j++) {
    double ic = bilinear_ref(r[j].dx, r[j].dy, t[j].w & 0xffu, t[j].z & 0xffu, t[j].x & 0xffu, t[j].y & 0xffu);
    const double w11 = r[j].dx * r[j].dy, w10 = r[j].dy - w11, w01 = r[j].dx - w11, w00 = 1.0 - r[j].dx - r[j].dy + w11;
    double gx2 = fma(w11, biased9_to_double((t[j].y >> 8) & 0x1ffu),
                     fma(w10, biased9_to_double((t[j].x >> 8) & 0x1ffu),
                         fma(w01, biased9_to_double((t[j].z >> 8) & 0x1ffu), w00 * biased9_to_double((t[j].w >> 8) & 0x1ffu))));
    double gy2 = fma(w11, biased9_to_double(t[j].y >> 17),
                     fma(w10, biased9_to_double(t[j].x >> 17),
                         fma(w01, biased9_to_double(t[j].z >> 17), w00 * biased9_to_double(t[j].w >> 17))));
    // rare: undecided by the fast path, saturated plateau, or first image row / column
    const bool sat = (t[j].x & t[j].y & t[j].z & t[j].w & 0xffu) == 0xffu;
    if (r[j].fix || (r[j].jac && (sat || r[j].ix < 1 || r[j].iy < 1))) {
      double o7[7];
      r[j].jac = jac_pixel_literal<PTS>(g, p.rows, p.cols, im1, G.a0[j0 + j], G.a1[j0 + j], G.a2[j0 + j], G.id[j0 + j], o7);
      if (r[j].jac) {
        r[j].x = o7[0]; r[j].y = o7[1]; r[j].z = o7[2]; r[j].rz = o7[3];
        ic = o7[4]; gx2 = o7[5]; gy2 = o7[6];
      }
    }
    ic = clamp_intensity(ic);
    const double ub = r[j].jac ? ic * s : 0.0;
    const int k = (int)ub;
    const double f = ub - u2d((unsigned)k);
    const double* q = wq + 3 * k * 128;
    double ci = fma(f, fma(f, q[256], q[128]), q[0]);
    if (ub == 0.0) ci = 0.0;  // the reference's BsplineDer quirk
    if (!r[j].jac) continue;  // (a padding slot may carry z = 0 and non-finite coordinates)
    // d(u,v)/d(xi), types_six_dof_expmap.cpp:438-450, in normalised coordinates xn = x/z, yn = y/z
    const double iz = r[j].rz, xn = r[j].x * iz, yn = r[j].y * iz;
    const double a = ci * gx2 * hfx, b = ci * gy2 * hfy;
    const double xy = xn * yn;
    acc[0] = fma(-b, fma(yn, yn, 1.0), fma(-a, xy, acc[0]));
    acc[1] = fma(b, xy, fma(a, fma(xn, xn, 1.0), acc[1]));
    acc[2] = fma(b, xn, fma(-a, yn, acc[2]));
    const double aiz = a * iz, biz = b * iz;
    acc[3] += aiz;
    acc[4] += biz;
    acc[5] = fma(-yn, biz, fma(-xn, aiz, acc[5]));
  }
}

// grid (jobs, ceil(max_slices/4)), 128 threads; shared: the lanes' quadratic rows [3*NS][128].
template <bool PTS, int W>
__global__ void __launch_bounds__(128, W == 1 ? 5 : 4) k_jac_sell(EvalParams p) {
  extern __shared__ double sm[];
  __shared__ Geo geo;
  const int B = p.bins, NS = B - 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int job = blockIdx.x + p.job0;
  const int pair = p.job_pair[job];
  build_geo<PTS>(p, job, pair, &geo);
  {
    // derivative of the per-span basis polynomials: dco[k][j][m] = (j+1) * coef[k][m][j+1]
    double* dco = sm + 3 * NS * 128 + 4 * (B * (B + 1) + B);
    for (int i = threadIdx.x; i < NS * 12; i += blockDim.x) {
      const int k = i / 12, j = (i % 12) / 4, m = i % 4;
      dco[i] = (double)(j + 1) * p.bs_coef[(k * 4 + m) * 4 + j + 1];
    }
  }
  __syncthreads();
  const int slice = blockIdx.y * 4 + warp;
  if (slice >= p.nslices[pair]) return;
  const int* so = p.sl_off + (size_t)pair * (p.max_slices + 1) + slice;
  const int off0 = so[0], ngroups = (so[1] - off0) >> 7;
  const int task = p.sl_task[((size_t)pair * p.max_slices + slice) * 32 + lane];
  double* wq = sm + threadIdx.x;  // wq[i * 128]
  // ---- per-class / per-span quadratic of the lane's task (class v, cell of the slice):
  //   c(f) = q0 + q1 f + q2 f^2 = sum_m N'_{k+m}(k+f) * Wv[k+m],
  //   Wv[t] = V[t] + sum_kk w_ref,v[kk] W[k_r(v)+kk][t]     (class 256: V only)
  // from the cell's scaled log tables W|V (k_assemble), staged per warp with rows padded to B+1.
  {
    const int BP = B + 1;
    double* Ww = sm + 3 * NS * 128 + warp * (B * BP + B);
    double* Vw = Ww + B * BP;
    const double* dco = sm + 3 * NS * 128 + 4 * (B * BP + B);  // [NS][3][4] derivative coefficients
    const int desc = task >= 0 ? p.tasks[(size_t)pair * p.max_tasks + task].y : 0;
    const int cell = __shfl_sync(0xffffffffu, (desc >> 18) & 0x3fff, 0);  // lane 0 always owns a task
    const double* wvg = p.wv + ((size_t)job * p.ncell + cell) * (size_t)(B * B + B);
    for (int i = lane; i < B * B; i += 32) Ww[(i / B) * BP + i % B] = wvg[i];
    for (int i = lane; i < B; i += 32) Vw[i] = wvg[B * B + i];
    __syncwarp();
    if (task >= 0) {
      const int cls = (desc >> 9) & 0x1ff;
      double wr[4] = {0.0, 0.0, 0.0, 0.0};
      int kr = 0;
      if (cls < 256) {
        kr = p.lut_k[cls];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) wr[kk] = p.lut_w[4 * cls + kk];
      }
      const double* Wr = Ww + kr * BP;
      auto wv_at = [&](int t) {
        double a = Vw[t];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) a += wr[kk] * Wr[kk * BP + t];
        return a;
      };
      double w0 = wv_at(0), w1 = wv_at(1), w2 = wv_at(2), w3;
      for (int k = 0; k < NS; k++) {
        w3 = wv_at(k + 3);
        const double* cf = dco + k * 12;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        a0 += cf[0] * w0; a1 += cf[4] * w0; a2 += cf[8] * w0;
        a0 += cf[1] * w1; a1 += cf[5] * w1; a2 += cf[9] * w1;
        a0 += cf[2] * w2; a1 += cf[6] * w2; a2 += cf[10] * w2;
        a0 += cf[3] * w3; a1 += cf[7] * w3; a2 += cf[11] * w3;
        wq[(3 * k) * 128] = a0; wq[(3 * k + 1) * 128] = a1; wq[(3 * k + 2) * 128] = a2;
        w0 = w1; w1 = w2; w2 = w3;
      }
    }
  }
  const size_t sbase = (size_t)pair * p.sell_cap + off0 + lane * 4;
  const double* q0 = p.sd0 + sbase;
  const double* q1 = PTS ? p.sd1 + sbase : nullptr;
  const double* q2 = PTS ? p.sd2 + sbase : nullptr;
  const unsigned* qi = p.sid + sbase;
  const cudaTextureObject_t tex2 = p.tex2[pair];
  const uint8_t* im1 = p.im1 + (size_t)pair * p.N;
  const double s = (double)NS / 255.0;
  const double hfx = 0.5 * geo.fx, hfy = 0.5 * geo.fy;  // the /2 of the central differences folded in
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int gi = 0; gi < ngroups; gi++) {
    Group<PTS> G;
    G.load(q0, q1, q2, qi, (size_t)gi * 128);
#pragma unroll
    for (int j0 = 0; j0 < 4; j0 += W) jac_pixels<PTS, W>(p, &geo, G, j0, tex2, im1, s, hfx, hfy, wq, acc);
  }
  // one partial per slice: fixed-order butterfly over the 32 lanes (lanes without a task hold zeros)
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
    for (int k = 1; k < 6; k++) vv = lane == k ? acc[k] : vv;
    p.jpart[((size_t)job * p.max_slices + slice) * 6 + lane] = vv;
  }
}

// a8 tail: one warp per (job, cell): the slice partials of the cell summed in a fixed order -> der[6]
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
  const int s0 = p.cell_slice_start[pair * (p.ncell + 1) + c];
  const int s1 = p.cell_slice_start[pair * (p.ncell + 1) + c + 1];
  const double* jp = p.jpart + (size_t)job * p.max_slices * 6;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int t = s0 + lane; t < s1; t += 32) {
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
  const size_t sb = (size_t)pair * c->sell_cap;
  cudaError_t e = cudaMemsetAsync(c->sid + sb, 0xFF, sizeof(unsigned) * c->sell_cap, c->stream);
  if (e != cudaSuccess) return check_cuda(e, "memset sid");
  e = cudaMemsetAsync(c->sd0 + sb, 0, sizeof(double) * c->sell_cap, c->stream);
  if (e != cudaSuccess) return check_cuda(e, "memset sd0");
  const int* tp = c->task_pos + (size_t)pair * c->max_tasks;
  const double* depth = c->depth + (size_t)pair * c->N;
  if (c->sell_points) {
    cudaMemsetAsync(c->sd1 + sb, 0, sizeof(double) * c->sell_cap, c->stream);
    cudaMemsetAsync(c->sd2 + sb, 0, sizeof(double) * c->sell_cap, c->stream);
    k_scatter_sell<true><<<c->ncell, 256, 0, c->stream>>>(p, pair, c->task_px, depth, tp, c->sd0, c->sd1, c->sd2, c->sid);
  } else {
    k_scatter_sell<false><<<c->ncell, 256, 0, c->stream>>>(p, pair, c->task_px, depth, tp, c->sd0, nullptr, nullptr, c->sid);
  }
  NID_LAUNCH_CHECK(c, "k_scatter_sell");
  return NID_OK;
}

int launch_pack_tex(nid_ctx* c, int pair, unsigned* d_out) {
  int g = (c->N + 255) / 256;
  if (g > c->sm_count * 8) g = c->sm_count * 8;
  k_pack_tex<<<g, 256, 0, c->stream>>>(c->rows, c->cols, c->im1 + (size_t)pair * c->N, d_out);
  NID_LAUNCH_CHECK(c, "k_pack_tex");
  return NID_OK;
}

size_t hist_sell_smem(const nid_ctx* c) { return sizeof(double) * ((size_t)c->bins * 256 + (size_t)(c->bins - 3) * 16); }
size_t jac_sell_smem(const nid_ctx* c) {
  const size_t B = c->bins, NS = B - 3;
  return sizeof(double) * (NS * 3 * 128 + 4 * (B * (B + 1) + B) + NS * 12);
}
size_t assemble_smem(const nid_ctx* c) { return sizeof(double) * ((size_t)5 * c->bins * c->bins + c->bins + NID_ASM_THREADS + (size_t)NID_NCLS * c->bins + 1024); }

template <bool PTS>
static void launch_hist_w(nid_ctx* c, const EvalParams& p, int ns, int n_jobs) {
  const dim3 grid(n_jobs, (ns + 7) / 8);
  const size_t sm = hist_sell_smem(c);
  switch (c->opt_ilp_hist) {
    case 1: k_hist_sell<PTS, 1><<<grid, 256, sm, c->stream>>>(p); break;
    case 2: k_hist_sell<PTS, 2><<<grid, 256, sm, c->stream>>>(p); break;
    default: k_hist_sell<PTS, 4><<<grid, 256, sm, c->stream>>>(p); break;
  }
}
template <bool PTS>
static void launch_jac_w(nid_ctx* c, const EvalParams& p, int ns, int n_jobs) {
  const dim3 grid(n_jobs, (ns + 3) / 4);
  const size_t sm = jac_sell_smem(c);
  switch (c->opt_ilp_jac) {
    case 1: k_jac_sell<PTS, 1><<<grid, 128, sm, c->stream>>>(p); break;
    default: k_jac_sell<PTS, 2><<<grid, 128, sm, c->stream>>>(p); break;
  }
}

int launch_eval_sorted(nid_ctx* c, int job0, int n_jobs, int n_jobs_total, int want_jac) {
  EvalParams p = make_params(c, n_jobs_total);
  p.job0 = job0;
  const int ns = c->max_nslices_prepared;
  const bool pts = c->sell_points;
  ktime_mark(c, 0);
  if (pts) launch_hist_w<true>(c, p, ns, n_jobs); else launch_hist_w<false>(c, p, ns, n_jobs);
  NID_LAUNCH_CHECK(c, "k_hist_sell");
  ktime_mark(c, 1);
  k_class_sum<<<dim3((NID_NCLS * c->bins + 255) / 256, c->ncell, n_jobs), 256, 0, c->stream>>>(p);
  NID_LAUNCH_CHECK(c, "k_class_sum");
  k_assemble<<<dim3(c->ncell, n_jobs), NID_ASM_THREADS, assemble_smem(c), c->stream>>>(p, want_jac);
  NID_LAUNCH_CHECK(c, "k_assemble");
  ktime_mark(c, 2);
  if (want_jac) {
    if (pts) launch_jac_w<true>(c, p, ns, n_jobs); else launch_jac_w<false>(c, p, ns, n_jobs);
    NID_LAUNCH_CHECK(c, "k_jac_sell");
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
  cudaError_t e;
  if (c->bins > NID_SORTED_MAX_BINS) return NID_OK;  // natural-order kernels only
#define NID_SMEM_ATTR(k, bytes)                                                                     \
  e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));            \
  if (e != cudaSuccess) return check_cuda(e, "smem attr " #k);
  NID_SMEM_ATTR((k_hist_sell<true, 1>), hist_sell_smem(c));
  NID_SMEM_ATTR((k_hist_sell<false, 1>), hist_sell_smem(c));
  NID_SMEM_ATTR((k_hist_sell<true, 2>), hist_sell_smem(c));
  NID_SMEM_ATTR((k_hist_sell<false, 2>), hist_sell_smem(c));
  NID_SMEM_ATTR((k_hist_sell<true, 4>), hist_sell_smem(c));
  NID_SMEM_ATTR((k_hist_sell<false, 4>), hist_sell_smem(c));
  NID_SMEM_ATTR((k_jac_sell<true, 1>), jac_sell_smem(c));
  NID_SMEM_ATTR((k_jac_sell<false, 1>), jac_sell_smem(c));
  NID_SMEM_ATTR((k_jac_sell<true, 2>), jac_sell_smem(c));
  NID_SMEM_ATTR((k_jac_sell<false, 2>), jac_sell_smem(c));
  NID_SMEM_ATTR(k_assemble, assemble_smem(c));
#undef NID_SMEM_ATTR
  return NID_OK;
}

}  // namespace nid
