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
// Pass 1: one warp per task -> un-weighted target soft histogram h[B] of the task's pixels.
// grid (ceil(max_tasks/8), jobs), 256 threads; shared: 8 warps x B x 32 doubles.
__global__ void __launch_bounds__(256) k_hist_sorted(EvalParams p) {
  extern __shared__ double sm[];
  const int B = p.bins;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int job = blockIdx.y + p.job0;
  const int pair = p.job_pair[job];
  const int t = blockIdx.x * 8 + warp;
  double* coef = sm + (size_t)8 * B * 32;  // [(B-3)*16] spline polynomial table
  for (int i = threadIdx.x; i < (B - 3) * 16; i += blockDim.x) coef[i] = p.bs_coef[i];
  __syncthreads();
  if (t >= p.ntasks[pair]) return;
  int start, count, cls, cell;
  unpack_task(p.tasks[(size_t)pair * p.max_tasks + t], start, count, cls, cell);
  if (p.n_c[pair * p.ncell + cell] < NID_MIN_CELL_POINTS) return;
  double* h = sm + (size_t)warp * B * 32;  // h[tt*32 + lane]
  for (int tt = 0; tt < B; tt++) h[tt * 32 + lane] = 0.0;

  const size_t base = (size_t)pair * p.N;
  const double* cp = p.cam + 4 * pair;
  const Cam cam{cp[0], cp[1], cp[2], cp[3]};
  const Pose P = load_pose(p.poses + 16 * job);
  const uint8_t* im1 = p.im1 + base;
  const double s = (double)(B - 3) / 255.0;
  const double* sx = p.sx + base + start;
  const double* sy = p.sy + base + start;
  const double* sz = p.sz + base + start;
  for (int i = lane; i < count; i += 32) {
    double x1, y1, z1, u, v;
    warp_project(P, cam, sx[i], sy[i], sz[i], x1, y1, z1, u, v);
    if (!inb_cost(u, v, p.rows, p.cols)) continue;
    const double ic = clamp_intensity(interp_u8_fast(im1, p.cols, u, v));
    const double ub = ic * s;
    const int kt = (int)ub;  // ub >= 0
    double wt[4], dw[4];
    bspline4_tab<false>(coef, ub, kt, wt, dw);
#pragma unroll
    for (int n = 0; n < 4; n++) h[(kt + n) * 32 + lane] += wt[n];
  }
  __syncwarp();
  // fixed-order merge of the 32 lane-private copies: lane tt sums column tt (rotated start => no bank conflicts)
  double* out = p.G + ((size_t)job * p.g_stride + t) * B;
  for (int tt = lane; tt < B; tt += 32) {
    double acc = 0.0;
#pragma unroll 8
    for (int j = 0; j < 32; j++) acc += h[tt * 32 + ((j + tt) & 31)];
    out[tt] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Assembly (a7 + table half of a8): one CTA per (cell, job). Combines the task partials of the cell in
// task order into P_t and P_j, normalises by n_c, computes H_t, H_j, err (computeH.cu:261-300;
// types_six_dof_expmap.cpp:609-635, .h:227) and, when want_jac, the scaled tables
//   W[r][t] = coefJ * (1 + log2 P_j[r][t]),  V[t] = coefT * (1 + log2 P_t[t])   (0 where P < 1e-30)
// with coefJ = -(s/(n_c Hj^2)) (Ht + Href), coefT = (s/(n_c Hj^2)) Hj  (types_six_dof_expmap.cpp:486-528).
#define NID_ASM_BATCH 64
#define NID_ASM_MAXE 16  // ceil(64*64/256)
__global__ void __launch_bounds__(256) k_assemble(EvalParams p, int want_jac) {
  extern __shared__ double sm[];
  __shared__ double scratch[8];
  __shared__ int s_cls[NID_ASM_BATCH];
  const int B = p.bins, BB = B * B;
  double* Gs = sm;                      // [NID_ASM_BATCH][B]
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
    double* wv = p.wv + o * (BB + B);
    for (int i = threadIdx.x; i < BB + B; i += blockDim.x) {
      const double q = Pall[i];
      const double L = (q < kSigma) ? 0.0 : (1.0 + log2(q));
      wv[i] = L * (i < BB ? coefJ : coefT);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pass 2: one warp per task -> partial of  J[a] = sum_i g_i[a] * sum_m dw_i[m] * Wv[kt_i + m].
__global__ void __launch_bounds__(256) k_jac_sorted(EvalParams p) {
  extern __shared__ double sm[];
  const int B = p.bins, BB = B * B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int job = blockIdx.y + p.job0;
  const int pair = p.job_pair[job];
  const int t = blockIdx.x * 8 + warp;
  double* coef = sm + 8 * B;  // [(B-3)*16] spline polynomial table
  for (int i = threadIdx.x; i < (B - 3) * 16; i += blockDim.x) coef[i] = p.bs_coef[i];
  __syncthreads();
  if (t >= p.ntasks[pair]) return;
  int start, count, cls, cell;
  unpack_task(p.tasks[(size_t)pair * p.max_tasks + t], start, count, cls, cell);
  if (p.n_c[pair * p.ncell + cell] < NID_MIN_CELL_POINTS) return;
  double* wrow = sm + warp * B;
  {
    const double* wv = p.wv + ((size_t)job * p.ncell + cell) * (BB + B);
    for (int tt = lane; tt < B; tt += 32) {
      double a = wv[BB + tt];
      if (cls < 256) {
        const int kr = p.lut_k[cls];
#pragma unroll
        for (int k = 0; k < 4; k++) a += p.lut_w[4 * cls + k] * wv[(kr + k) * B + tt];
      }
      wrow[tt] = a;
    }
  }
  __syncwarp();
  const size_t base = (size_t)pair * p.N;
  const double* cp = p.cam + 4 * pair;
  const Cam cam{cp[0], cp[1], cp[2], cp[3]};
  const Pose P = load_pose(p.poses + 16 * job);
  const uint8_t* im1 = p.im1 + base;
  const double s = (double)(B - 3) / 255.0;
  const double* sx = p.sx + base + start;
  const double* sy = p.sy + base + start;
  const double* sz = p.sz + base + start;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = lane; i < count; i += 32) {
    double x, y, z, u, v;
    warp_project(P, cam, sx[i], sy[i], sz[i], x, y, z, u, v);
    if (!inb_jac(u, v, p.rows, p.cols)) continue;
    double ic, gx, gy;
    sample_grad_u8(im1, p.cols, u, v, ic, gx, gy);
    ic = clamp_intensity(ic);
    const double ub = ic * s;
    const int kt = (int)ub;
    double wt[4], dw[4];
    bspline4_tab<true>(coef, ub, kt, wt, dw);
    double ci = 0.0;
#pragma unroll
    for (int m = 0; m < 4; m++) ci += dw[m] * wrow[kt + m];
    const double iz = 1.0 / z, iz2 = iz * iz;
    const double a = ci * gx * cam.fx, b = ci * gy * cam.fy;
    acc[0] += a * (-x * y * iz2) + b * (-(1.0 + y * y * iz2));
    acc[1] += a * (1.0 + x * x * iz2) + b * (x * y * iz2);
    acc[2] += a * (-y * iz) + b * (x * iz);
    acc[3] += a * iz;
    acc[4] += b * iz;
    acc[5] += a * (-x * iz2) + b * (-y * iz2);
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
    p.jpart[((size_t)job * p.g_stride + t) * 6 + lane] = vv;
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

size_t hist_sorted_smem(const nid_ctx* c) { return sizeof(double) * (8 * c->bins * 32 + (c->bins - 3) * 16); }
size_t jac_sorted_smem(const nid_ctx* c) { return sizeof(double) * (8 * c->bins + (c->bins - 3) * 16); }
size_t assemble_smem(const nid_ctx* c) { return sizeof(double) * (NID_ASM_BATCH * c->bins + c->bins * c->bins + c->bins); }

int launch_eval_sorted(nid_ctx* c, int job0, int n_jobs, int n_jobs_total, int want_jac) {
  EvalParams p = make_params(c, n_jobs_total);
  p.job0 = job0;
  const int tblocks = (c->max_ntasks_prepared + 7) / 8;
  ktime_mark(c, 0);
  k_hist_sorted<<<dim3(tblocks, n_jobs), 256, hist_sorted_smem(c), c->stream>>>(p);
  NID_LAUNCH_CHECK(c, "k_hist_sorted");
  ktime_mark(c, 1);
  k_assemble<<<dim3(c->ncell, n_jobs), 256, assemble_smem(c), c->stream>>>(p, want_jac);
  NID_LAUNCH_CHECK(c, "k_assemble");
  ktime_mark(c, 2);
  if (want_jac) {
    k_jac_sorted<<<dim3(tblocks, n_jobs), 256, jac_sorted_smem(c), c->stream>>>(p);
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
  cudaError_t e = cudaFuncSetAttribute(k_hist_sorted, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_sorted_smem(c));
  if (e != cudaSuccess) return check_cuda(e, "smem attr k_hist_sorted");
  e = cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)assemble_smem(c));
  if (e != cudaSuccess) return check_cuda(e, "smem attr k_assemble");
  return NID_OK;
}

}  // namespace nid
