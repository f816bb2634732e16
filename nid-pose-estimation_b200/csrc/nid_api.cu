// C-ABI of the B200-native NID path (include/nid_b200.h): context, staging, and the host-side
// Levenberg-Marquardt driver (the only CPU arithmetic is the 6x6 solve and the pose update, as in the
// reference: optimization_algorithm_levenberg.cpp:61-225).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <string>
#include <chrono>
#include <vector>

#include "../host/nid_host_math.hpp"
#include "nid_ctx.h"

namespace nid {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NID_OK;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return NID_ERR_CUDA;
}

#define CU(call, what)                                   \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return check_cuda(e_, what);  \
  } while (0)
#define OKR(call)               \
  do {                          \
    int r_ = (call);            \
    if (r_ != NID_OK) return r_; \
  } while (0)

static int strips_for(const nid_ctx* c, int n_jobs) {
  int S;
  if (c->opt_force_strips > 0) S = std::min(c->opt_force_strips, c->rb);
  else {
    long long want = 2LL * c->sm_count;
    long long per = (long long)c->ncell * n_jobs;
    S = (int)((want + per - 1) / per);
    S = std::max(1, std::min(S, std::min(c->rb, 32)));
  }
  // the natural-order kernels write n_jobs * S (job, strip) partials: never more than the buffers hold
  if (c->part_slots > 0 && n_jobs > 0) S = (int)std::max<size_t>(1, std::min<size_t>((size_t)S, c->part_slots / (size_t)n_jobs));
  return S;
}

EvalParams make_params(nid_ctx* c, int n_jobs) {
  EvalParams p;
  memset(&p, 0, sizeof(p));
  p.rows = c->rows; p.cols = c->cols; p.cell = c->cell; p.bins = c->bins;
  p.rb = c->rb; p.cb = c->cb; p.N = c->N; p.ncell = c->ncell;
  p.S = strips_for(c, n_jobs);
  p.strip_rows = (c->rb + p.S - 1) / p.S;
  p.hist_stride = c->bins * c->bins + c->bins;
  p.pwx = c->pwx; p.pwy = c->pwy; p.pwz = c->pwz;
  p.im0 = c->im0; p.im1 = c->im1; p.inb0 = c->inb0;
  p.n_c = c->n_c; p.href = c->href; p.cam = c->cam;
  p.lut_w = c->lut_w; p.lut_k = c->lut_k;
  p.poses = c->poses; p.job_pair = c->job_pair;
  const bool sorted = use_sorted(c);
  p.part = c->part; p.jpart = sorted ? c->jpart_s : c->jpart; p.hist = c->opt_keep_hist ? c->hist : nullptr;
  p.ht = c->ht; p.hj = c->hj; p.err = c->err; p.der = c->der; p.gn = c->gn;
  p.sd0 = c->sd0; p.sd1 = c->sd1; p.sd2 = c->sd2; p.sid = c->sid;
  p.sv = c->sv; p.span_mode = c->span_mode ? 1 : 0;
  p.stage_bulk = (c->opt_stage_bulk < 0 ? (long long)c->rb * c->cb < NID_SMALL_CELL_PX : c->opt_stage_bulk) ? 1 : 0;
  p.sl_off = c->sl_off; p.sl_task = c->sl_task; p.sl_desc = c->sl_desc; p.task_cls = c->task_cls; p.sl_cell = c->sl_cell; p.nslices = c->nslices;
  p.sell_cap = c->sell_cap; p.max_slices = c->max_slices; p.Twc0 = c->Twc0;
  p.tasks = c->tasks; p.ntasks = c->ntasks; p.cell_task_start = c->cell_task_start; p.cell_slice_start = c->cell_slice_start;
  p.cls_task_start = c->cls_task_start; p.span_start = c->span_start; p.wv = c->wv;
  p.max_tasks = c->max_tasks; p.g_stride = (int)c->g_stride;
  p.G = c->G; p.fp1 = c->fp1; p.tex2 = c->d_tex2; p.k1tex = c->d_k1tex;
  return p;
}

// Path selection. The sorted path (no floating-point atomics) wins wherever it is available: measured on B200 at
// 640x480, cost+Jacobian: 4x4 cells/16 bins 77k vs 9k evals/s, the reference's default 16x16 cells/10 bins 29k vs 9k
// (tools/time_config.py). The natural-order kernels remain for bin counts outside [6, 40] and as a second,
// independently written implementation the parity tests hold to the same bar.
// Task kind of the sorted path, fixed per geometry: cells of fewer than NID_SPAN_PX pixels take span tasks;
// caller-supplied world points and more than 20 bins keep class tasks. MEASURED (B200, 640x480, 16x16 cells, 10 bins):
// span tasks halve the assembly (2.1 -> 1.1 us per evaluation) but their 16 shared-memory read-modify-writes per pixel
// saturate the shared-memory pipe in pass 1 (5.0 -> 7.8 us) and pass 2 gains nothing (7.5 us either way): 60.9k
// against 68.5k evaluations/s. The automatic threshold is therefore 0 (class tasks everywhere); span tasks stay
// available and tested behind nid_set_option("sorted_mode", 2).
#ifndef NID_SPAN_PX
#define NID_SPAN_PX 0
#endif
static bool want_span(const nid_ctx* c) {
  if (!use_sorted(c) || c->sell_points || c->bins > 20 || c->opt_sorted_mode == 1) return false;
  return c->opt_sorted_mode == 2 || (long long)c->rb * c->cb < NID_SPAN_PX;
}
static void update_task_kind(nid_ctx* c) {
  const bool span = want_span(c);
  if (span != c->span_mode) {
    c->span_mode = span;
    std::fill(c->pair_prepared.begin(), c->pair_prepared.end(), 0);  // the pixel store is laid out per task kind
    std::fill(c->pair_sorted.begin(), c->pair_sorted.end(), 0);
  }
}

bool use_sorted(const nid_ctx* c) {
  // the assembly tables must fit in shared memory; the end-block fold needs the two 3x3 blocks disjoint (B >= 6, the
  // reference's smallest knot table, computeH.cu:100)
  if (c->bins > NID_SORTED_MAX_BINS || c->bins < NID_SORTED_MIN_BINS) return false;
  if (c->opt_path == 1) return false;
  return true;
}

template <typename T>
static int dalloc(T** p, size_t n, const char* what) {
  cudaError_t e = cudaMalloc((void**)p, sizeof(T) * (n ? n : 1));
  if (e != cudaSuccess) return check_cuda(e, what);
  return NID_OK;
}

// per-job scratch of the active path, allocated on first use / grown when a pair with more tasks is prepared
int ensure_job_buffers(nid_ctx* c) {
  const size_t J = c->job_cap, NC = c->ncell;
  const size_t hs = (size_t)c->bins * c->bins + c->bins;
  if (c->opt_keep_hist && !c->hist) OKR(dalloc(&c->hist, J * NC * hs, "hist"));
  if (use_sorted(c)) {
    const size_t rows_per_task = c->span_mode ? 4 : 1;
    const size_t need = (size_t)std::max(c->max_ntasks_prepared, 1);
    if (c->g_stride < need || c->g_rows < rows_per_task) {
      CU(cudaStreamSynchronize(c->stream), "sync before growing job buffers");
      if (c->G) cudaFree(c->G);
      c->G = nullptr;
      OKR(dalloc(&c->G, J * std::max(need, c->g_stride) * c->bins * rows_per_task, "G"));
      c->g_stride = std::max(need, c->g_stride);
      c->g_rows = rows_per_task;
    }
    if (!c->jpart_s) OKR(dalloc(&c->jpart_s, J * (size_t)c->max_slices * 6, "jpart_s"));  // one partial per slice
    if (!c->wv) {
      OKR(dalloc(&c->wv, J * NC * (size_t)nid::wv_stride(c->bins), "wv"));
      CU(cudaMemset(c->wv, 0, sizeof(double) * J * NC * (size_t)nid::wv_stride(c->bins)), "memset wv");  // (the padding is copied, never used)
    }
  } else if (!c->part) {
    c->part_slots = J + 2 * (size_t)c->sm_count + 64;
    OKR(dalloc(&c->part, c->part_slots * NC * hs, "part"));
    OKR(dalloc(&c->jpart, c->part_slots * NC * 6, "jpart"));
  }
  return NID_OK;
}

// footprint-packed plane and gather texture of the targets of pairs [pair0, pair0 + n)
static int update_textures(nid_ctx* c, int pair0, int n) {
  for (int i = 0; i < n; i++) c->k1_ready[pair0 + i] = 0;  // kernel 1's planes follow the target image (rebuilt on next use)
  for (int p0 = pair0; p0 < pair0 + n; p0 += c->setup_batch) {
    const int nb = std::min(c->setup_batch, pair0 + n - p0);
    OKR(launch_pack(c, p0, nb));
    for (int i = 0; i < nb; i++)
      CU(cudaMemcpy2DToArrayAsync(c->tex2_arrays[p0 + i], 0, 0, c->d_pack + (size_t)i * c->N, sizeof(unsigned) * c->cols,
                                  sizeof(unsigned) * c->cols, c->rows, cudaMemcpyDeviceToDevice, c->stream), "packed im1 -> texture array");
  }
  return NID_OK;
}
static int update_texture(nid_ctx* c, int pair) { return update_textures(c, pair, 1); }

// Pinned bump arena for the small per-pair arrays (poses, intrinsics) that the asynchronous set-up calls copy to the
// device: a region is only handed out again after the stream was synchronised.
static int arena_take(nid_ctx* c, size_t bytes, void** out) {
  bytes = (bytes + 63) & ~(size_t)63;
  if (bytes > c->h_arena_cap) {
    CU(cudaStreamSynchronize(c->stream), "sync before growing the staging arena");
    if (c->h_arena) cudaFreeHost(c->h_arena);
    c->h_arena = nullptr; c->h_arena_cap = 0; c->h_arena_used = 0;
    const size_t cap = std::max(bytes * 4, (size_t)1 << 20);
    CU(cudaMallocHost((void**)&c->h_arena, cap), "pinned staging arena");
    c->h_arena_cap = cap;
  }
  if (c->h_arena_used + bytes > c->h_arena_cap) {
    CU(cudaStreamSynchronize(c->stream), "sync to recycle the staging arena");
    c->h_arena_used = 0;
  }
  *out = c->h_arena + c->h_arena_used;
  c->h_arena_used += bytes;
  return NID_OK;
}

// Staging of n_jobs (pair, pose) jobs: the next slot of the pinned ring is taken (waiting, if need be, for the copies
// that last read it), filled, and copied to the device asynchronously on the context stream.
// mode 0: pairs must be prepared (evaluation); 1: pairs must be set (hard-binned NID, kernel 1).
static int stage_jobs(nid_ctx* c, int n_jobs, const int* job_pair, const double* poses, int mode = 0) {
  if (n_jobs < 1 || n_jobs > c->max_jobs) { set_error("n_jobs out of range"); return NID_ERR_ARG; }
  for (int j = 0; j < n_jobs; j++) {
    int pr = job_pair ? job_pair[j] : 0;
    if (pr < 0 || pr >= c->n_pairs) { set_error("job_pair out of range"); return NID_ERR_ARG; }
    if (mode == 1) {
      if (!c->pair_set[pr]) { set_error("job_pair invalid or pair not set"); return NID_ERR_ARG; }
      continue;
    }
    if (!c->pair_prepared[pr]) { set_error("pair not prepared (call nid_prepare)"); return NID_ERR_STATE; }
    if (use_sorted(c) && !c->pair_sorted[pr]) {
      set_error("pair not prepared for the sorted path (call nid_prepare after selecting the path)");
      return NID_ERR_STATE;
    }
  }
  const int slot = (c->stage_slot + 1) % NID_STAGE_RING;
  CU(cudaEventSynchronize(c->stage_ev[slot]), "wait for the staging slot");  // (returns at once for an unrecorded event)
  c->stage_slot = slot;
  c->h_poses = c->h_poses_ring + (size_t)slot * 16 * c->job_cap;
  c->h_job_pair = c->h_job_pair_ring + (size_t)slot * c->job_cap;
  for (int j = 0; j < n_jobs; j++) c->h_job_pair[j] = job_pair ? job_pair[j] : 0;
  memcpy(c->h_poses, poses, sizeof(double) * 16 * n_jobs);
  CU(cudaMemcpyAsync(c->poses, c->h_poses, sizeof(double) * 16 * n_jobs, cudaMemcpyHostToDevice, c->stream), "H2D poses");
  CU(cudaMemcpyAsync(c->job_pair, c->h_job_pair, sizeof(int) * n_jobs, cudaMemcpyHostToDevice, c->stream), "H2D job_pair");
  CU(cudaEventRecord(c->stage_ev[slot], c->stream), "record staging event");
  c->staged_jobs = n_jobs;
  return NID_OK;
}

static int fetch(nid_ctx* c, int n_jobs, int want_jac, double* Ht, double* Hj, double* der) {
  const size_t nc = (size_t)n_jobs * c->ncell;
  double* h = c->h_out;
  if (Ht) CU(cudaMemcpyAsync(h, c->ht, sizeof(double) * nc, cudaMemcpyDeviceToHost, c->stream), "D2H ht");
  if (Hj) CU(cudaMemcpyAsync(h + nc, c->hj, sizeof(double) * nc, cudaMemcpyDeviceToHost, c->stream), "D2H hj");
  if (want_jac && der) CU(cudaMemcpyAsync(h + 2 * nc, c->der, sizeof(double) * 6 * nc, cudaMemcpyDeviceToHost, c->stream), "D2H der");
  CU(cudaStreamSynchronize(c->stream), "sync");
  if (Ht) memcpy(Ht, h, sizeof(double) * nc);
  if (Hj) memcpy(Hj, h + nc, sizeof(double) * nc);
  if (want_jac && der) memcpy(der, h + 2 * nc, sizeof(double) * 6 * nc);
  return NID_OK;
}

}  // namespace nid

using namespace nid;

extern "C" {

const char* nid_last_error(void) { return g_err.c_str(); }
int nid_version(void) { return 100; }

int nid_create(nid_ctx** out, int device, int rows, int cols, int cell, int bins, int degree, int n_pairs, int max_jobs) {
  cudaGetLastError();  // a stale error of another library in this process must not fail our first launch check
  if (!out) { set_error("ctx out pointer is NULL"); return NID_ERR_ARG; }
  *out = nullptr;
  if (degree != 3) { set_error("only bs_degree == 3 (order-4 B-splines) is supported, as in the reference"); return NID_ERR_UNSUPPORTED; }
  if (rows > 65535 || cols > 65535) { set_error("rows, cols must be < 65536"); return NID_ERR_ARG; }
  if (cell > 90) { set_error("cell must be <= 90 (the task descriptors keep the cell index in 13 bits)"); return NID_ERR_ARG; }
  if (rows < 8 || cols < 8 || cell < 1 || bins < 6 || bins > 64 || n_pairs < 1 || max_jobs < 1 || cell > rows || cell > cols) {
    set_error("bad geometry (need rows,cols>=8, 1<=cell<=min(rows,cols), 6<=bins<=64, n_pairs,max_jobs>=1)");
    return NID_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
    return NID_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("device index out of range"); return NID_ERR_ARG; }
  CU(cudaSetDevice(device), "cudaSetDevice");
  nid_ctx* c = new nid_ctx();
  c->device = device; c->rows = rows; c->cols = cols; c->cell = cell; c->bins = bins; c->degree = degree;
  c->N = rows * cols; c->ncell = cell * cell; c->rb = rows / cell; c->cb = cols / cell;
  c->n_pairs = n_pairs; c->max_jobs = max_jobs;
  c->job_cap = std::max(max_jobs, NID_MIN_JOB_SLOTS);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  c->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  const size_t N = c->N, P = n_pairs, J = c->job_cap, NC = c->ncell;
  const size_t hs = (size_t)bins * bins + bins;
  OKR(dalloc(&c->pwx, P * N, "pwx")); OKR(dalloc(&c->pwy, P * N, "pwy")); OKR(dalloc(&c->pwz, P * N, "pwz"));
  OKR(dalloc(&c->im0, P * N, "im0")); OKR(dalloc(&c->im1, P * N, "im1")); OKR(dalloc(&c->inb0, P * N, "inb0"));
  OKR(dalloc(&c->n_c, P * NC, "n_c")); OKR(dalloc(&c->href, P * NC, "href"));
  OKR(dalloc(&c->cam, P * 4, "cam")); OKR(dalloc(&c->Twc0, P * 16, "Twc0"));
  OKR(dalloc(&c->cnt, P * NC * NID_NCLS, "cnt"));
  // pixels per task, by geometry only (never by the capacity of the context: a shard of a job must compute the same
  // bits as the whole job). Measured at 640x480, 96 evaluations in flight: 4x4 cells 144k evaluations/s with 32-pixel
  // tasks, 137k with 24, 126k with 16 (twice the task rows to assemble); 8x8 cells 112k / 110k / 103k; the reference's
  // default 16x16 cells (1200 pixels, ~5 per reference intensity) 72.5k with 32, 78.7k with 16 or 20, 77k with 12 (before
  // the other small-cell choices of NID_SMALL_CELL_PX): their
  // tasks are short anyway and a slice is as long as its longest task. Shorter tasks also cut the latency of a lone
  // solve (more, shorter slices: 1.68 -> 1.37 ms at 4x4 cells with 16): option "task_px" for latency-bound callers.
  c->task_px = (long long)c->rb * c->cb < NID_SMALL_CELL_PX ? 16 : 32;
  const int min_task_px = 8;
  c->max_tasks = (int)(N / min_task_px + NC * NID_NCLS + 1);
  c->max_slices = c->max_tasks / 32 + (int)NC + 1;
  // pixel slots per pair: every task is padded to a multiple of 4 and every cell's slices to their longest task
  c->sell_cap = (N + 3 * (size_t)c->max_tasks + NC * 32 * (size_t)(NID_TASK_PX_MAX + 4) + 255) / 128 * 128;
  OKR(dalloc(&c->depth, P * N, "depth"));
  OKR(dalloc(&c->sl_off, P * ((size_t)c->max_slices + 1), "sl_off"));
  OKR(dalloc(&c->sl_task, P * (size_t)c->max_slices * 32, "sl_task"));
  OKR(dalloc(&c->sl_desc, P * (size_t)c->max_slices * 32, "sl_desc"));
  OKR(dalloc(&c->task_cls, P * (size_t)c->max_tasks, "task_cls"));
  OKR(dalloc(&c->sl_cell, P * (size_t)c->max_slices, "sl_cell"));
  OKR(dalloc(&c->task_pos, P * (size_t)c->max_tasks, "task_pos"));
  OKR(dalloc(&c->nslices, P, "nslices"));
  CU(cudaMemset(c->nslices, 0, sizeof(int) * P), "memset nslices");
  c->h_nslices.assign(P, 0);
  OKR(dalloc(&c->tasks, P * (size_t)c->max_tasks, "tasks"));
  OKR(dalloc(&c->ntasks, P, "ntasks"));
  CU(cudaMemset(c->ntasks, 0, sizeof(int) * P), "memset ntasks");
  OKR(dalloc(&c->cell_task_start, P * (NC + 1), "cell_task_start"));
  OKR(dalloc(&c->cell_slice_start, P * (NC + 1), "cell_slice_start"));
  OKR(dalloc(&c->cls_task_start, P * NC * (NID_NCLS + 1), "cls_task_start"));
  {
    // first class of every span: k_r(v) = floor(v' (B-3)/255), v' = 254.999 for v = 255, is monotone in v
    const int NS = bins - 3;
    std::vector<int> ss(NS + 1, 256);
    auto kr = [bins](int v) { double o = v >= 255 ? 254.999 : (double)v; return (int)std::floor(o * (bins - 3) / 255.0); };
    for (int v = 255; v >= 0; v--)
      for (int k = 0; k <= kr(v) && k <= NS; k++) ss[k] = v;
    ss[NS] = 256;
    OKR(dalloc(&c->span_start, ss.size(), "span_start"));
    CU(cudaMemcpy(c->span_start, ss.data(), sizeof(int) * ss.size(), cudaMemcpyHostToDevice), "H2D span_start");
  }
  c->h_ntasks.assign(P, 0);
  // packed target images as gather-able textures (tex2Dgather needs a CUDA array created with cudaArrayTextureGather)
  {
    bool ok = true;
    // packed I | Gx | Gy texels, 32 bit
    c->tex2_arrays.assign(P, nullptr);
    c->h_tex2.assign(P, 0);
    cudaChannelFormatDesc cd2 = cudaCreateChannelDesc<unsigned int>();
    for (size_t i = 0; i < P && ok; i++) {
      if (cudaMallocArray(&c->tex2_arrays[i], &cd2, cols, rows, cudaArrayTextureGather) != cudaSuccess) { ok = false; break; }
      cudaResourceDesc rd;
      memset(&rd, 0, sizeof(rd));
      rd.resType = cudaResourceTypeArray;
      rd.res.array.array = c->tex2_arrays[i];
      cudaTextureDesc td;
      memset(&td, 0, sizeof(td));
      td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
      td.filterMode = cudaFilterModePoint;
      td.readMode = cudaReadModeElementType;
      td.normalizedCoords = 0;
      if (cudaCreateTextureObject(&c->h_tex2[i], &rd, &td, nullptr) != cudaSuccess) ok = false;
    }
    if (!ok) {
      cudaGetLastError();
      set_error("could not create the packed gather textures for the target images");
      return NID_ERR_CUDA;
    }
    OKR(dalloc(&c->d_tex2, P, "d_tex2"));
    OKR(dalloc(&c->d_k1tex, P, "d_k1tex"));
    CU(cudaMemset(c->d_k1tex, 0, sizeof(cudaTextureObject_t) * P), "memset k1tex");
    c->k1_arrays.assign(P, nullptr);
    c->h_k1tex.assign(P, 0);
    c->k1_ready.assign(P, 0);
    CU(cudaMemcpy(c->d_tex2, c->h_tex2.data(), sizeof(cudaTextureObject_t) * P, cudaMemcpyHostToDevice), "H2D tex2 handles");
    c->setup_batch = (int)std::min<size_t>(P, 32);
    OKR(dalloc(&c->d_pack, N * (size_t)c->setup_batch, "d_pack"));
    OKR(dalloc(&c->lay_tot, (size_t)c->setup_batch * NC * 3, "lay_tot"));
    OKR(dalloc(&c->lay_base, (size_t)c->setup_batch * NC * 3, "lay_base"));
    OKR(dalloc(&c->prep_poses, (size_t)c->setup_batch * 16, "prep_poses"));
    OKR(dalloc(&c->depth_factor, P, "depth_factor"));
    c->pair_u16.assign(P, 0);
    OKR(dalloc(&c->fp1, P * N, "fp1"));
    c->h_Twc0.assign(P * 16, 0.0);
    c->h_cam.assign(P * 4, 1.0);
  }
  OKR(dalloc(&c->d_depth, N, "d_depth")); OKR(dalloc(&c->d_img64, N, "d_img64")); OKR(dalloc(&c->d_flag, 2, "d_flag"));
  CU(cudaMemset(c->d_flag, 0, sizeof(int) * 2), "memset flag");
  OKR(dalloc(&c->lut_w, 256 * 4, "lut_w")); OKR(dalloc(&c->lut_k, 256, "lut_k"));
  OKR(dalloc(&c->poses, J * 16, "poses")); OKR(dalloc(&c->job_pair, J, "job_pair"));
  CU(cudaMemset(c->job_pair, 0, sizeof(int) * J), "memset job_pair");
  CU(cudaMemset(c->poses, 0, sizeof(double) * 16 * J), "memset poses");
  OKR(dalloc(&c->aux_pose, 16, "aux_pose"));
  (void)hs;
  OKR(dalloc(&c->ht, J * NC, "ht")); OKR(dalloc(&c->hj, J * NC, "hj")); OKR(dalloc(&c->err, J * NC, "err"));
  OKR(dalloc(&c->der, J * NC * 6, "der")); OKR(dalloc(&c->gn, J * 44, "gn"));
  OKR(dalloc(&c->hard, J * (NC + 1), "hard"));
  CU(cudaMallocHost((void**)&c->h_poses_ring, sizeof(double) * 16 * J * NID_STAGE_RING), "pinned poses");
  CU(cudaMallocHost((void**)&c->h_job_pair_ring, sizeof(int) * J * NID_STAGE_RING), "pinned job_pair");
  memset(c->h_poses_ring, 0, sizeof(double) * 16 * J * NID_STAGE_RING);
  memset(c->h_job_pair_ring, 0, sizeof(int) * J * NID_STAGE_RING);
  c->h_poses = c->h_poses_ring;
  c->h_job_pair = c->h_job_pair_ring;
  for (int i = 0; i < NID_STAGE_RING; i++) CU(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming), "cudaEventCreate (staging)");
  CU(cudaMallocHost((void**)&c->h_aux_pose, sizeof(double) * 16), "pinned aux pose");
  CU(cudaMallocHost((void**)&c->h_out, sizeof(double) * J * (NC * 8 + 44)), "pinned out");
  c->pair_set.assign(P, 0);
  c->pair_prepared.assign(P, 0);
  c->pair_sorted.assign(P, 0);
  OKR(launch_build_lut(c));
  OKR(sorted_init(c));
  c->span_mode = want_span(c);
  CU(cudaStreamSynchronize(c->stream), "sync after lut");
  *out = c;
  return NID_OK;
}

int nid_destroy(nid_ctx* c) {
  if (!c) return NID_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  void* ptrs[] = {c->pwx, c->pwy, c->pwz, c->im0, c->im1, c->inb0, c->n_c, c->href, c->cam, c->Twc0, c->cnt, c->d_depth,
                  c->d_img64, c->d_flag, c->d_pix, c->d_pix4, c->d_pix4_jobs, c->chunk_cnt, c->d_bsv, c->d_bsi, c->lut_w, c->lut_k, c->poses,
                  c->job_pair, c->aux_pose, c->part, c->jpart, c->hist, c->ht, c->hj, c->err, c->der, c->gn, c->hard,
                  c->depth, c->sd0, c->sd1, c->sd2, c->sid, c->sv, c->sl_off, c->sl_task, c->sl_desc, c->task_cls, c->sl_cell, c->nslices, c->task_pos,
                  c->lay_tot, c->lay_base, c->prep_poses, c->depth16, c->depth_factor, c->d_tex2, c->d_pack, c->tasks, c->ntasks, c->cell_task_start, c->cell_slice_start, c->G, c->jpart_s, c->fp1, c->cls_task_start, c->span_start, c->wv};
  for (auto t : c->h_k1tex) if (t) cudaDestroyTextureObject(t);
  for (auto arr : c->k1_arrays) if (arr) cudaFreeArray(arr);
  if (c->d_k1tex) cudaFree(c->d_k1tex);
  if (c->d_k1pack) cudaFree(c->d_k1pack);
  for (auto t : c->h_tex2) if (t) cudaDestroyTextureObject(t);
  for (auto arr : c->tex2_arrays) if (arr) cudaFreeArray(arr);
  for (void* p : ptrs) if (p) cudaFree(p);
  if (c->h_poses_ring) cudaFreeHost(c->h_poses_ring);
  if (c->h_job_pair_ring) cudaFreeHost(c->h_job_pair_ring);
  if (c->h_aux_pose) cudaFreeHost(c->h_aux_pose);
  if (c->h_lm_lists) cudaFreeHost(c->h_lm_lists);
  destroy_latency_graph(c);
  if (c->h_lm_stage) cudaFreeHost(c->h_lm_stage);
  if (c->d_lm_stage) cudaFree(c->d_lm_stage);
  if (c->d_lm_lists) cudaFree(c->d_lm_lists);
  for (int i = 0; i < NID_STAGE_RING; i++) if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->h_arena) cudaFreeHost(c->h_arena);
  if (c->h_res) cudaFreeHost(c->h_res);
  for (int i = 0; i < 2; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 5; i++) if (c->kev[i]) cudaEventDestroy(c->kev[i]);
  cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  delete c;
  return NID_OK;
}

int nid_sync(nid_ctx* c) {
  if (!c) { set_error("NULL ctx"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  CU(cudaStreamSynchronize(c->stream), "nid_sync");
  c->h_arena_used = 0;
  return NID_OK;
}

// Geometry of pairs [pair0, pair0 + n): T_wc0 [n][16] and intr [n][5] go through the pinned arena (the device keeps
// fx fy cx cy and the depth factor in separate tables), then the world points. depth64 / depth16: [n][N], one of them.
static int set_pairs_geometry(nid_ctx* c, int pair0, int n, const double* depth64, const uint16_t* depth16, const double* T_wc0,
                              const double* intr) {
  if (pair0 < 0 || n < 1 || pair0 + n > c->n_pairs) { set_error("pair range out of bounds"); return NID_ERR_ARG; }
  if ((!depth64 && !depth16) || !T_wc0 || !intr) { set_error("NULL argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  const size_t N = c->N;
  if (depth16) {
    if (!c->depth16) OKR(dalloc(&c->depth16, (size_t)c->n_pairs * N, "depth16"));
    CU(cudaMemcpyAsync(c->depth16 + (size_t)pair0 * N, depth16, sizeof(uint16_t) * N * n, cudaMemcpyDefault, c->stream), "H2D depth (u16)");
  } else {
    CU(cudaMemcpyAsync(c->depth + (size_t)pair0 * N, depth64, sizeof(double) * N * n, cudaMemcpyDefault, c->stream), "H2D depth");
  }
  double* a = nullptr;
  OKR(arena_take(c, sizeof(double) * 21 * n, (void**)&a));
  double *aT = a, *acam = a + 16 * (size_t)n, *afac = a + 20 * (size_t)n;
  memcpy(aT, T_wc0, sizeof(double) * 16 * n);
  for (int i = 0; i < n; i++) {
    memcpy(acam + 4 * i, intr + 5 * (size_t)i, sizeof(double) * 4);
    afac[i] = intr[5 * (size_t)i + 4];
  }
  CU(cudaMemcpyAsync(c->Twc0 + 16 * (size_t)pair0, aT, sizeof(double) * 16 * n, cudaMemcpyHostToDevice, c->stream), "H2D Twc0");
  CU(cudaMemcpyAsync(c->cam + 4 * (size_t)pair0, acam, sizeof(double) * 4 * n, cudaMemcpyHostToDevice, c->stream), "H2D intr");
  CU(cudaMemcpyAsync(c->depth_factor + pair0, afac, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream), "H2D depth factor");
  memcpy(c->h_Twc0.data() + 16 * (size_t)pair0, aT, sizeof(double) * 16 * n);
  memcpy(c->h_cam.data() + 4 * (size_t)pair0, acam, sizeof(double) * 4 * n);
  OKR(launch_points(c, pair0, n, depth16 != nullptr));
  for (int i = 0; i < n; i++) { c->pair_prepared[pair0 + i] = 0; c->pair_u16[pair0 + i] = depth16 ? 1 : 0; }
  return NID_OK;
}
static int set_pair_common(nid_ctx* c, int pair, const double* depth, const double T_wc0[16], const double intr[5]) {
  if (pair < 0 || pair >= c->n_pairs) { set_error("pair index out of range"); return NID_ERR_ARG; }
  return set_pairs_geometry(c, pair, 1, depth, nullptr, T_wc0, intr);
}

int nid_set_pair(nid_ctx* c, int pair, const double* depth, const uint8_t* im0, const uint8_t* im1, const double T_wc0[16],
                 const double intr[5]) {
  if (!c) { set_error("NULL ctx"); return NID_ERR_ARG; }
  if (!im0 || !im1) { set_error("NULL image"); return NID_ERR_ARG; }
  OKR(set_pair_common(c, pair, depth, T_wc0, intr));
  CU(cudaMemcpyAsync(c->im0 + (size_t)pair * c->N, im0, c->N, cudaMemcpyDefault, c->stream), "H2D im0");
  CU(cudaMemcpyAsync(c->im1 + (size_t)pair * c->N, im1, c->N, cudaMemcpyDefault, c->stream), "H2D im1");
  OKR(update_texture(c, pair));
  CU(cudaStreamSynchronize(c->stream), "sync set_pair");
  c->pair_set[pair] = 1;
  return NID_OK;
}

int nid_set_pairs_u16(nid_ctx* c, int pair0, int n, const uint16_t* depth_raw, const uint8_t* im0, const uint8_t* im1,
                      const double* T_wc0, const double* intr) {
  if (!c) { set_error("NULL ctx"); return NID_ERR_ARG; }
  if (!im0 || !im1 || !depth_raw) { set_error("NULL image"); return NID_ERR_ARG; }
  if (c->sell_points) { set_error("this context holds caller-supplied world points (nid_set_pair_points)"); return NID_ERR_STATE; }
  OKR(set_pairs_geometry(c, pair0, n, nullptr, depth_raw, T_wc0, intr));
  const size_t N = c->N;
  CU(cudaMemcpyAsync(c->im0 + (size_t)pair0 * N, im0, N * n, cudaMemcpyDefault, c->stream), "H2D im0");
  CU(cudaMemcpyAsync(c->im1 + (size_t)pair0 * N, im1, N * n, cudaMemcpyDefault, c->stream), "H2D im1");
  OKR(update_textures(c, pair0, n));
  for (int i = 0; i < n; i++) c->pair_set[pair0 + i] = 1;
  return NID_OK;  // (asynchronous: see the header)
}

int nid_set_target(nid_ctx* c, int pair, const uint8_t* im1) {
  if (!c || !im1) { set_error("bad argument"); return NID_ERR_ARG; }
  if (pair < 0 || pair >= c->n_pairs) { set_error("pair index out of range"); return NID_ERR_ARG; }
  if (!c->pair_set[pair]) { set_error("pair not set (nid_set_target replaces the target of an existing pair)"); return NID_ERR_STATE; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  CU(cudaMemcpyAsync(c->im1 + (size_t)pair * c->N, im1, c->N, cudaMemcpyDefault, c->stream), "H2D im1");
  OKR(update_texture(c, pair));
  CU(cudaStreamSynchronize(c->stream), "sync set_target");
  c->pair_prepared[pair] = 0;
  return NID_OK;
}

static int upload_images_f64(nid_ctx* c, int pair, const double* im0, const double* im1);
static int build_sorted_layout(nid_ctx* c, int pair0, int n);
static int finish_sorted_layout(nid_ctx* c, int pair0, int n, int* bs_counter, double* Href);

int nid_set_pair_f64(nid_ctx* c, int pair, const double* depth, const double* im0, const double* im1,
                     const double T_wc0[16], const double intr[5]) {
  if (!c) { set_error("NULL ctx"); return NID_ERR_ARG; }
  OKR(set_pair_common(c, pair, depth, T_wc0, intr));
  OKR(upload_images_f64(c, pair, im0, im1));
  c->pair_set[pair] = 1;
  return NID_OK;
}

static int upload_images_f64(nid_ctx* c, int pair, const double* im0, const double* im1) {
  CU(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream), "memset flag");
  if (im0) {
    CU(cudaMemcpyAsync(c->d_img64, im0, sizeof(double) * c->N, cudaMemcpyDefault, c->stream), "H2D im0 f64");
    OKR(launch_check_integral(c, c->d_img64, c->im0 + (size_t)pair * c->N, 1));
  }
  if (im1) {
    CU(cudaMemcpyAsync(c->d_img64, im1, sizeof(double) * c->N, cudaMemcpyDefault, c->stream), "H2D im1 f64");
    OKR(launch_check_integral(c, c->d_img64, c->im1 + (size_t)pair * c->N, 0));
    OKR(update_texture(c, pair));
  }
  int flag = 0;
  CU(cudaMemcpyAsync(&flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream), "D2H flag");
  CU(cudaStreamSynchronize(c->stream), "sync upload images");
  if (flag) {
    set_error("image holds non-8-bit values (non-integral or outside [0,255]); only 8-bit gray is supported");
    return NID_ERR_UNSUPPORTED;
  }
  return NID_OK;
}

int nid_set_pair_points(nid_ctx* c, int pair, const double* points_3d, const double* im0, const double* im1,
                        const double intr[5]) {
  if (!c || pair < 0 || pair >= c->n_pairs) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  if (points_3d) {
    if (!intr) { set_error("intr required with points"); return NID_ERR_ARG; }
    if (!c->d_pix) OKR(dalloc(&c->d_pix, (size_t)8 * c->N, "d_pix"));
    CU(cudaMemcpyAsync(c->d_pix, points_3d, sizeof(double) * 3 * c->N, cudaMemcpyDefault, c->stream), "H2D points3d");
    CU(cudaMemcpyAsync(c->cam + 4 * pair, intr, sizeof(double) * 4, cudaMemcpyDefault, c->stream), "H2D intr");
    memcpy(c->h_cam.data() + 4 * (size_t)pair, intr, sizeof(double) * 4);
    OKR(launch_points_soa(c, pair, c->d_pix));
    c->pair_prepared[pair] = 0;
    if (!c->sell_points) {
      // caller-supplied world points cannot be rebuilt from (depth, pixel): the sorted store keeps the points
      CU(cudaStreamSynchronize(c->stream), "sync before switching to the point form");
      c->sell_points = true;
      std::fill(c->pair_prepared.begin(), c->pair_prepared.end(), 0);
      update_task_kind(c);
    }
  }
  OKR(upload_images_f64(c, pair, im0, im1));
  c->pair_set[pair] = 1;
  return NID_OK;
}

int nid_import_prepare(nid_ctx* c, int pair, const double* bs_value, const int* bs_counter, const double* Href) {
  if (!c || pair < 0 || pair >= c->n_pairs || !bs_value || !bs_counter || !Href) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!c->pair_set[pair]) { set_error("pair not set"); return NID_ERR_STATE; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  if (!c->d_bsv) { OKR(dalloc(&c->d_bsv, (size_t)4 * c->N, "d_bsv")); OKR(dalloc(&c->d_bsi, (size_t)c->N, "d_bsi")); }
  CU(cudaMemcpyAsync(c->d_bsv, bs_value, sizeof(double) * 4 * c->N, cudaMemcpyDefault, c->stream), "H2D bs_value");
  OKR(launch_import_flags(c, pair, c->d_bsv));
  CU(cudaMemcpyAsync(c->n_c + pair * c->ncell, bs_counter, sizeof(int) * c->ncell, cudaMemcpyDefault, c->stream), "H2D n_c");
  CU(cudaMemcpyAsync(c->href + pair * c->ncell, Href, sizeof(double) * c->ncell, cudaMemcpyDefault, c->stream), "H2D href");
  CU(cudaMemsetAsync(c->cnt + (size_t)pair * c->ncell * NID_NCLS, 0, sizeof(unsigned int) * c->ncell * NID_NCLS, c->stream), "memset cnt");
  OKR(launch_count_classes(c, pair));
  OKR(build_sorted_layout(c, pair, 1));
  OKR(finish_sorted_layout(c, pair, 1, nullptr, nullptr));
  c->pair_prepared[pair] = 1;
  return NID_OK;
}

int nid_get_inbounds(nid_ctx* c, int pair, uint8_t* flags) {
  if (!c || pair < 0 || pair >= c->n_pairs || !flags) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!c->pair_prepared[pair]) { set_error("pair not prepared"); return NID_ERR_STATE; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  CU(cudaMemcpyAsync(flags, c->inb0 + (size_t)pair * c->N, c->N, cudaMemcpyDefault, c->stream), "D2H inb0");
  CU(cudaStreamSynchronize(c->stream), "sync inb0");
  return NID_OK;
}

int nid_get_points3d(nid_ctx* c, int pair, double* points_3d) {
  if (!c || pair < 0 || pair >= c->n_pairs || !points_3d) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!c->pair_set[pair]) { set_error("pair not set"); return NID_ERR_STATE; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  if (!c->d_pix) OKR(dalloc(&c->d_pix, (size_t)8 * c->N, "d_pix"));
  OKR(launch_points_aos(c, pair, c->d_pix));
  CU(cudaMemcpyAsync(points_3d, c->d_pix, sizeof(double) * 3 * c->N, cudaMemcpyDefault, c->stream), "copy points3d");
  CU(cudaStreamSynchronize(c->stream), "sync points3d");
  return NID_OK;
}

// Regroup the valid pixels of pairs [pair0, pair0 + n) by (cell, reference class), cut the segments into tasks of at most
// task_px pixels, order the tasks by length and pack them 32 to a slice (nid_sorted.cu). Everything happens on the
// device, from the class counts k_prepare left there: table kernels, then the regrouping scatter. Asynchronous.
static int build_sorted_layout(nid_ctx* c, int pair0, int n) {
  for (int i = 0; i < n; i++) c->pair_sorted[pair0 + i] = 0;
  if (!use_sorted(c)) return NID_OK;  // the natural-order kernels need none of this
  if (!c->sd0) {
    OKR(dalloc(&c->sd0, (size_t)c->n_pairs * c->sell_cap, "sd0"));
    OKR(dalloc(&c->sid, (size_t)c->n_pairs * c->sell_cap, "sid"));
  }
  if (c->sell_points && !c->sd1) {
    OKR(dalloc(&c->sd1, (size_t)c->n_pairs * c->sell_cap, "sd1"));
    OKR(dalloc(&c->sd2, (size_t)c->n_pairs * c->sell_cap, "sd2"));
  }
  if (c->span_mode && !c->sv) OKR(dalloc(&c->sv, (size_t)c->n_pairs * c->sell_cap, "sv"));
  if (!c->chunk_cnt) {
    const size_t nchunks = ((size_t)c->rb * c->cb + 255) / 256;
    OKR(dalloc(&c->chunk_cnt, (size_t)c->setup_batch * c->ncell * nchunks * NID_NCLS, "chunk_cnt"));
  }
  for (int p0 = pair0; p0 < pair0 + n; p0 += c->setup_batch)
    OKR(launch_layout_and_scatter(c, p0, std::min(c->setup_batch, pair0 + n - p0)));
  return NID_OK;
}

// One synchronisation for the whole range: n_c and H_ref for the caller, task and slice counts for the launch geometry.
static int finish_sorted_layout(nid_ctx* c, int pair0, int n, int* bs_counter, double* Href) {
  const size_t NC = c->ncell;
  const size_t bytes = (size_t)n * (NC * (sizeof(double) + sizeof(int)) + 2 * sizeof(int)) + sizeof(int) + 64;
  if (c->h_res_cap < bytes) {
    CU(cudaStreamSynchronize(c->stream), "sync before growing the result buffer");
    if (c->h_res) cudaFreeHost(c->h_res);
    c->h_res = nullptr; c->h_res_cap = 0;
    CU(cudaMallocHost((void**)&c->h_res, bytes * 2), "pinned prepare results");
    c->h_res_cap = bytes * 2;
  }
  double* h_href = (double*)c->h_res;
  int* h_nc = (int*)(h_href + (size_t)n * NC);
  int* h_nt = h_nc + (size_t)n * NC;
  int* h_ns = h_nt + n;
  int* h_ovf = h_ns + n;
  CU(cudaMemcpyAsync(h_nc, c->n_c + (size_t)pair0 * NC, sizeof(int) * NC * n, cudaMemcpyDeviceToHost, c->stream), "D2H n_c");
  CU(cudaMemcpyAsync(h_href, c->href + (size_t)pair0 * NC, sizeof(double) * NC * n, cudaMemcpyDeviceToHost, c->stream), "D2H href");
  const bool sorted = use_sorted(c);
  if (sorted) {
    CU(cudaMemcpyAsync(h_nt, c->ntasks + pair0, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream), "D2H ntasks");
    CU(cudaMemcpyAsync(h_ns, c->nslices + pair0, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream), "D2H nslices");
    CU(cudaMemcpyAsync(h_ovf, c->d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream), "D2H overflow flag");
  }
  CU(cudaStreamSynchronize(c->stream), "sync prepare");
  c->h_arena_used = 0;
  if (sorted) {
    if (*h_ovf) { set_error("sliced pixel store overflow"); return NID_ERR_STATE; }
    for (int i = 0; i < n; i++) {
      c->h_ntasks[pair0 + i] = h_nt[i];
      c->h_nslices[pair0 + i] = h_ns[i];
      c->pair_sorted[pair0 + i] = 1;
    }
    c->max_ntasks_prepared = 0;
    for (int v : c->h_ntasks) c->max_ntasks_prepared = std::max(c->max_ntasks_prepared, v);
    c->max_nslices_prepared = 0;
    for (int v : c->h_nslices) c->max_nslices_prepared = std::max(c->max_nslices_prepared, v);
  }
  if (bs_counter) memcpy(bs_counter, h_nc, sizeof(int) * NC * n);
  if (Href) memcpy(Href, h_href, sizeof(double) * NC * n);
  return NID_OK;
}

int nid_prepare_pairs(nid_ctx* c, int pair0, int n, const double* T_cw1, int* bs_counter, double* Href) {
  if (!c || !T_cw1 || n < 1 || pair0 < 0 || pair0 + n > c->n_pairs) { set_error("bad argument"); return NID_ERR_ARG; }
  for (int i = 0; i < n; i++)
    if (!c->pair_set[pair0 + i]) { set_error("pair not set"); return NID_ERR_STATE; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  for (int p0 = pair0; p0 < pair0 + n; p0 += c->setup_batch) {
    const int nb = std::min(c->setup_batch, pair0 + n - p0);
    double* a = nullptr;
    OKR(arena_take(c, sizeof(double) * 16 * nb, (void**)&a));
    memcpy(a, T_cw1 + 16 * (size_t)(p0 - pair0), sizeof(double) * 16 * nb);
    CU(cudaMemcpyAsync(c->prep_poses, a, sizeof(double) * 16 * nb, cudaMemcpyHostToDevice, c->stream), "H2D initial poses");
    OKR(launch_prepare(c, p0, nb, c->prep_poses));
    OKR(launch_href(c, p0, nb));
    OKR(build_sorted_layout(c, p0, nb));
  }
  OKR(finish_sorted_layout(c, pair0, n, bs_counter, Href));
  for (int i = 0; i < n; i++) c->pair_prepared[pair0 + i] = 1;
  return NID_OK;
}

int nid_prepare(nid_ctx* c, int pair, const double T_cw1[16], int* bs_counter, double* Href) {
  if (!c || pair < 0 || pair >= c->n_pairs || !T_cw1) { set_error("bad argument"); return NID_ERR_ARG; }
  return nid_prepare_pairs(c, pair, 1, T_cw1, bs_counter, Href);
}

int nid_get_ref_weights(nid_ctx* c, int pair, double* bs_value, int* bs_index) {
  if (!c || pair < 0 || pair >= c->n_pairs) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!c->pair_prepared[pair]) { set_error("pair not prepared"); return NID_ERR_STATE; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  if (!c->d_bsv) { OKR(dalloc(&c->d_bsv, (size_t)4 * c->N, "d_bsv")); OKR(dalloc(&c->d_bsi, (size_t)c->N, "d_bsi")); }
  OKR(launch_ref_weights(c, pair));
  if (bs_value) CU(cudaMemcpyAsync(bs_value, c->d_bsv, sizeof(double) * 4 * c->N, cudaMemcpyDefault, c->stream), "D2H bs_value");
  if (bs_index) CU(cudaMemcpyAsync(bs_index, c->d_bsi, sizeof(int) * c->N, cudaMemcpyDefault, c->stream), "D2H bs_index");
  CU(cudaStreamSynchronize(c->stream), "sync ref_weights");
  return NID_OK;
}

int nid_stage_jobs(nid_ctx* c, int n_jobs, const int* job_pair, const double* poses) {
  if (!c || !poses) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  return stage_jobs(c, n_jobs, job_pair, poses);
}

int nid_eval_staged(nid_ctx* c, int n_jobs, int want_jac) {
  if (!c || n_jobs < 1 || n_jobs > c->max_jobs) { set_error("bad argument"); return NID_ERR_ARG; }
  if (n_jobs > c->staged_jobs) { set_error("nid_eval_staged: more jobs than the last nid_stage_jobs staged"); return NID_ERR_STATE; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  return launch_eval(c, n_jobs, want_jac);
}

int nid_fetch_results(nid_ctx* c, int n_jobs, int want_jac, double* Ht, double* Hj, double* der) {
  if (!c || n_jobs < 1 || n_jobs > c->max_jobs) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  return fetch(c, n_jobs, want_jac, Ht, Hj, der);
}

int nid_eval_jobs(nid_ctx* c, int n_jobs, const int* job_pair, const double* poses, int want_jac, double* Ht, double* Hj,
                  double* der) {
  if (!c || !poses) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  OKR(stage_jobs(c, n_jobs, job_pair, poses));
  OKR(launch_eval(c, n_jobs, want_jac));
  return fetch(c, n_jobs, want_jac, Ht, Hj, der);
}

int nid_eval(nid_ctx* c, int pair, const double T_cw1[16], int want_jac, double* Ht, double* Hj, double* der) {
  return nid_eval_jobs(c, 1, &pair, T_cw1, want_jac, Ht, Hj, der);
}

int nid_eval_gn(nid_ctx* c, int pair, const double T_cw1[16], double delta, double* chi2, double* H36, double* b6,
                double* err, double* J) {
  if (!c || !T_cw1) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  OKR(stage_jobs(c, 1, &pair, T_cw1));
  OKR(launch_eval(c, 1, 1));
  OKR(launch_gn(c, 1, delta));
  double* h = c->h_out;
  const size_t nc = c->ncell;
  CU(cudaMemcpyAsync(h, c->gn, sizeof(double) * 44, cudaMemcpyDeviceToHost, c->stream), "D2H gn");
  CU(cudaMemcpyAsync(h + 44, c->err, sizeof(double) * nc, cudaMemcpyDeviceToHost, c->stream), "D2H err");
  CU(cudaMemcpyAsync(h + 44 + nc, c->der, sizeof(double) * 6 * nc, cudaMemcpyDeviceToHost, c->stream), "D2H der");
  CU(cudaStreamSynchronize(c->stream), "sync gn");
  if (chi2) *chi2 = h[0];
  if (H36) memcpy(H36, h + 1, sizeof(double) * 36);
  if (b6) memcpy(b6, h + 37, sizeof(double) * 6);
  if (err) memcpy(err, h + 44, sizeof(double) * nc);
  if (J) memcpy(J, h + 44 + nc, sizeof(double) * 6 * nc);
  return NID_OK;
}

// ---------------------------------------------------------------------------------------------- LM
namespace {
struct LM {
  nidhost::Pose7 est, backup;
  double lambda = -1., ni = 2.;
  int nBad = 0, it = 0, qmax = 0;
  double currentChi = 0, iniChi = 0, rho = 0;
  double H[36], b[6], x[6] = {0, 0, 0, 0, 0, 0};
  bool ok2 = true;
  int phase = 0;  // 0 need jac, 1 need trial, 2 done
  int jac_evals = 0, cost_evals = 0;
  int pair = 0;
  bool slot_holds_est = false;  // the problem's job slot holds the evaluation (histograms, err, tables) of `est`
};

void lm_start_trial(LM& s) {
  s.backup = s.est;  // push
  double Hl[36];
  memcpy(Hl, s.H, sizeof(Hl));
  for (int j = 0; j < 6; j++) Hl[7 * j] += s.lambda;  // setLambda (block_solver.hpp:573-599)
  s.ok2 = nidhost::ldlt6_solve(Hl, s.b, s.x);
  s.est = nidhost::pose_mul(nidhost::pose_exp(s.x), s.est);  // oplusImpl
  s.phase = 1;
}
}  // namespace

// One LM step of problem s from the 44 doubles {chi2, H[36], b[6], -} of its job
// (optimization_algorithm_levenberg.cpp:98-141 after a cost+Jacobian job, :173-202 after a trial-pose cost job).
static void lm_absorb(nid_ctx* c, LM& s, const double* g, int n, int max_iters) {
  const int maxTrials = 10;
  const double tau = 1e-5, goodUp = 2. / 3., goodLo = 1. / 3.;
  if (s.phase == 0) {
    s.jac_evals++;
    s.currentChi = g[0];
    s.iniChi = s.currentChi;
    memcpy(s.H, g + 1, sizeof(double) * 36);
    memcpy(s.b, g + 37, sizeof(double) * 6);
    if (s.it == 0) {
      double md = 0.;
      for (int j = 0; j < 6; j++) md = std::max(std::fabs(s.H[7 * j]), md);
      s.lambda = tau * md;
      s.ni = 2;
      s.nBad = 0;
    }
    s.rho = 0;
    s.qmax = 0;
    lm_start_trial(s);
    return;
  }
  s.cost_evals++;
  double tempChi = g[0];
  if (!s.ok2) tempChi = std::numeric_limits<double>::max();
  double rho = s.currentChi - tempChi;
  double scale = 0.;
  for (int j = 0; j < 6; j++) scale += s.x[j] * (s.lambda * s.x[j] + s.b[j]);
  scale += 1e-3;
  rho /= scale;
  if (rho > 0 && std::isfinite(tempChi)) {
    double alpha = 1. - std::pow((2 * rho - 1), 3);
    alpha = std::min(alpha, goodUp);
    double sf = std::max(goodLo, alpha);
    s.lambda *= sf;
    s.ni = 2;
    s.currentChi = tempChi;
  } else {
    s.lambda *= s.ni;
    s.ni *= 2;
    s.est = s.backup;  // pop
    s.slot_holds_est = false;
  }
  s.rho = rho;
  s.qmax++;
  if (rho < 0 && s.qmax < maxTrials) {
    lm_start_trial(s);
    return;
  }
  bool terminate = (s.qmax == maxTrials || rho == 0);
  if (!terminate) {
    if ((s.iniChi - s.currentChi) * 1e3 < s.iniChi) s.nBad++;
    else s.nBad = 0;
    if (s.nBad >= 3) terminate = true;
  }
  s.it++;
  s.phase = (!terminate && s.it < max_iters) ? 0 : 2;
  if (c->lm_trace && n == 1 && s.it <= c->lm_trace_cap) {
    double* t = c->lm_trace + 10 * (s.it - 1);
    t[0] = s.currentChi; t[1] = s.lambda; t[2] = s.qmax;
    memcpy(t + 3, s.est.t, sizeof(double) * 3);
    memcpy(t + 6, s.est.q, sizeof(double) * 4);
  }
}

// Latency mode of the LM driver (a handful of problems, the reference's own use: one pair per process,
// NID_pose_estimation.cpp:350). With one job in flight the device is nearly idle and a solve is a chain of ~34
// dependent rounds (10 linearisations + ~24 trial poses), each a few small kernels and a host round trip. The trial
// poses of an outer iteration are known in advance: after a rejection the schedule retries with lambda * ni, ni * 2
// (optimization_algorithm_levenberg.cpp:186-199). So every round evaluates, per problem, the next K trial poses of that
// sequence at once and completely (histograms, err, chi2 AND Jacobian / Gauss-Newton block); the host then walks the
// unchanged state machine over the cached results: the first accepted trial ends the iteration and its Gauss-Newton
// block is already there for the next one. Same decisions, same numbers (the kernels are deterministic), about one
// round per outer iteration instead of three and a half.
static int solve_speculative(nid_ctx* c, std::vector<LM>& st, int n, int max_iters, double delta, int K) {
  struct Cand { nidhost::Pose7 pose; double gn[44]; bool valid; };
  std::vector<Cand> cache((size_t)n * K);
  for (auto& q : cache) q.valid = false;
  // one staging block per round: the poses of all n * K slots, then the list of the slots in use (one H2D copy); the
  // kernels read both from there for the duration of the solve
  const int nslots = n * K;
  const size_t stage_bytes = sizeof(double) * 16 * (size_t)c->job_cap + sizeof(int) * (size_t)c->job_cap;
  if (!c->h_lm_stage) {
    CU(cudaMallocHost((void**)&c->h_lm_stage, stage_bytes), "pinned LM staging");
    CU(cudaMalloc((void**)&c->d_lm_stage, stage_bytes), "LM staging");
  }
  struct Swap {  // restores the context's pose buffers on every way out
    nid_ctx* c; double* poses; double* h_poses;
    ~Swap() { c->poses = poses; c->h_poses = h_poses; }
  } swap{c, c->poses, c->h_poses};
  c->poses = reinterpret_cast<double*>(c->d_lm_stage);
  c->h_poses = reinterpret_cast<double*>(c->h_lm_stage);
  int* const list = reinterpret_cast<int*>(c->h_poses + 16 * (size_t)nslots);
  int* const d_list = reinterpret_cast<int*>(c->poses + 16 * (size_t)nslots);
  for (int j = 0; j < n; j++)
    for (int k = 0; k < K; k++) c->h_job_pair[j * K + k] = st[j].pair;
  CU(cudaMemcpyAsync(c->job_pair, c->h_job_pair, sizeof(int) * nslots, cudaMemcpyHostToDevice, c->stream), "H2D job_pair");
  CU(cudaStreamSynchronize(c->stream), "sync before LM");
  auto same_pose = [](const nidhost::Pose7& a, const nidhost::Pose7& b) {
    return memcmp(a.t, b.t, sizeof(a.t)) == 0 && memcmp(a.q, b.q, sizeof(a.q)) == 0;
  };
  const bool tt_on = getenv("NID_LM_TIMES") != nullptr;
  double tt[4] = {0, 0, 0, 0};
  int tt_rounds = 0;
  auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (;;) {
    const double t0 = now();
    // 1. let every problem consume what the cache already holds
    bool progress = true;
    while (progress) {
      progress = false;
      for (int j = 0; j < n; j++) {
        LM& s = st[j];
        if (s.phase == 2) continue;
        for (int k = 0; k < K; k++) {
          Cand& q = cache[(size_t)j * K + k];
          if (q.valid && same_pose(q.pose, s.est)) {
            lm_absorb(c, s, q.gn, n, max_iters);  // phase 0 reads chi2/H/b, phase 1 reads chi2: both are in the block
            progress = true;
            break;
          }
        }
      }
    }
    // 2. the next round: per unfinished problem the pose it waits for and the trials that would follow rejections
    int nj = 0;
    for (int j = 0; j < n; j++) {
      LM& s = st[j];
      for (int k = 0; k < K; k++) cache[(size_t)j * K + k].valid = false;
      if (s.phase == 2) continue;
      LM sim = s;
      for (int k = 0; k < K; k++) {
        const int slot = j * K + k;
        Cand& q = cache[slot];
        q.pose = sim.est;
        q.valid = true;
        nidhost::pose_to_mat16(sim.est, c->h_poses + 16 * (size_t)slot);
        nj++;
        if (sim.phase != 1 || sim.qmax + 1 >= 10) break;  // only trial poses have successors; maxTrials = 10
        // the trial after a rejection of this one (lm_absorb's else branch, then lm_start_trial)
        sim.lambda *= sim.ni;
        sim.ni *= 2;
        sim.est = sim.backup;
        sim.qmax++;
        lm_start_trial(sim);
      }
    }
    if (nj == 0) break;
    const double t1 = now();
    // every slot is evaluated every round (one fixed chain: see launch_latency_round); the slots this round has no use
    // for repeat their problem's first pose and their results are ignored
    for (int j = 0; j < n; j++) {
      for (int k = 0; k < K; k++) {
        const int slot = j * K + k;
        if (cache[slot].valid) continue;
        if (k > 0) memcpy(c->h_poses + 16 * (size_t)slot, c->h_poses + 16 * (size_t)(j * K), sizeof(double) * 16);
        else nidhost::pose_to_mat16(st[j].est, c->h_poses + 16 * (size_t)slot);
      }
    }
    for (int i = 0; i < nslots; i++) list[i] = i;
    OKR(launch_latency_round(c, nslots, d_list, list, sizeof(double) * 16 * nslots + sizeof(int) * nslots, delta, c->h_out));
    const double t2 = now();
    CU(cudaStreamSynchronize(c->stream), "sync lm");
    const double t3 = now();
    for (int i = 0; i < nslots; i++)
      if (cache[i].valid) memcpy(cache[i].gn, c->h_out + 44 * (size_t)i, sizeof(double) * 44);
    tt[0] += t1 - t0; tt[1] += t2 - t1; tt[2] += t3 - t2; tt[3] += now() - t3; tt_rounds++;
  }
  if (tt_on) fprintf(stderr, "lm rounds %d: host %.1f issue %.1f wait %.1f copy %.1f us per round\n", tt_rounds, tt[0] / tt_rounds, tt[1] / tt_rounds, tt[2] / tt_rounds, tt[3] / tt_rounds);
  return NID_OK;
}

// Lock-step LM over n problems. The problems are cut into two halves that ping-pong: while the device evaluates the
// jobs of one half (its own stream and its own range of job slots), the host absorbs the results of the other half,
// solves the 6x6 systems and stages the next poses. Each problem sees exactly the schedule of a solo solve.
//
// Every problem keeps ONE job slot for the whole solve, and every round hands the kernels two lists of slots: list 1
// gets pass 1 + assembly (histograms, entropies, err, chi2 and the scaled log tables), list 2 gets pass 2 + the
// Gauss-Newton block. A trial pose is on list 1 only. A cost+Jacobian job is on list 2 and -- this is the point --
// on list 1 only when its slot does not already hold the evaluation of that very pose: after an accepted trial the
// next outer iteration linearises at the pose the trial was evaluated at (optimization_algorithm_levenberg.cpp:173-199
// then :98), and the sorted kernels are deterministic, so the histograms, err and tables in the slot are bit for bit
// what a recomputation would produce. The reference evaluates them again (CudaComputeH(true) repeats the cost half);
// skipping that repeats nothing observable and saves a third of the device work of a solve (option "lm_reuse").
int nid_solve_jobs(nid_ctx* c, int n, const int* job_pair, double* poses7, int max_iters, double delta, int* stats) {
  if (!c || n < 1 || n > c->max_jobs || !poses7) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  std::vector<LM> st(n);
  for (int j = 0; j < n; j++) {
    LM& s = st[j];
    s.pair = job_pair ? job_pair[j] : j;
    if (s.pair < 0 || s.pair >= c->n_pairs) { set_error("job_pair out of range"); return NID_ERR_ARG; }
    if (!c->pair_prepared[s.pair] || (use_sorted(c) && !c->pair_sorted[s.pair])) { set_error("pair not prepared"); return NID_ERR_STATE; }
    memcpy(s.est.t, poses7 + 7 * j, sizeof(double) * 3);
    memcpy(s.est.q, poses7 + 7 * j + 3, sizeof(double) * 4);
    s.phase = max_iters > 0 ? 0 : 2;
  }
  OKR(ensure_job_buffers(c));
  c->staged_jobs = 0;  // the solver rewrites the staged poses and job table
  const bool sorted = use_sorted(c);
  const bool reuse = sorted && c->opt_lm_reuse;
  if (sorted && n <= 4 && c->opt_lm_spec >= 2) {
    const int K = std::min(c->opt_lm_spec, c->job_cap / n);
    if (K >= 2) {
      for (int j = 0; j < n; j++) c->h_job_pair[j] = st[j].pair;
      int r = solve_speculative(c, st, n, max_iters, delta, K);
      if (r != NID_OK) return r;
      for (int j = 0; j < n; j++) {
        memcpy(poses7 + 7 * j, st[j].est.t, sizeof(double) * 3);
        memcpy(poses7 + 7 * j + 3, st[j].est.q, sizeof(double) * 4);
        if (stats) { stats[3 * j] = st[j].it; stats[3 * j + 1] = st[j].jac_evals; stats[3 * j + 2] = st[j].cost_evals; }
      }
      return NID_OK;
    }
  }
  // (the natural-order kernels size their per-job partial buffers by the job count of a launch: one range there)
  const int nh = (n >= 8 && sorted) ? 2 : 1;
  if (nh == 2 && !c->stream2) {
    CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking), "cudaStreamCreate (LM)");
  }
  if (!c->h_lm_lists) {
    CU(cudaMallocHost((void**)&c->h_lm_lists, sizeof(int) * 2 * (size_t)c->job_cap), "pinned LM lists");
    OKR(dalloc(&c->d_lm_lists, 2 * (size_t)c->job_cap, "LM lists"));
  }
  CU(cudaStreamSynchronize(c->stream), "sync before LM");
  // slot of problem j = j; half h owns the slots [lo, hi) and the list storage [2 lo, 2 hi): list 1 then list 2
  struct Half { int lo, hi, n1, n2, na; cudaStream_t stream; };
  Half H[2];
  H[0] = {0, nh == 2 ? n / 2 : n, 0, 0, 0, c->stream};
  H[1] = {n / 2, n, 0, 0, 0, c->stream2};
  cudaStream_t const main_stream = c->stream;
  for (int j = 0; j < n; j++) c->h_job_pair[j] = st[j].pair;
  CU(cudaMemcpyAsync(c->job_pair, c->h_job_pair, sizeof(int) * n, cudaMemcpyHostToDevice, main_stream), "H2D job_pair");
  CU(cudaStreamSynchronize(main_stream), "sync job_pair");
  int rc = NID_OK;
  auto issue = [&](Half& h) -> int {
    int* l1 = c->h_lm_lists + 2 * (size_t)h.lo;
    int* l2 = l1 + (h.hi - h.lo);
    h.n1 = h.n2 = h.na = 0;
    for (int j = h.lo; j < h.hi; j++) {
      LM& s = st[j];
      if (s.phase == 2) continue;
      h.na++;
      nidhost::pose_to_mat16(s.est, c->h_poses + 16 * (size_t)j);
      if (s.phase == 0) {
        l2[h.n2++] = j;
        if (!(reuse && s.slot_holds_est)) l1[h.n1++] = j;
      } else {
        l1[h.n1++] = j;
      }
      s.slot_holds_est = true;  // after this round the slot holds the evaluation of s.est (a rejection resets it)
    }
    if (h.na == 0) return NID_OK;
    const int cnt = h.hi - h.lo;
    CU(cudaMemcpyAsync(c->poses + 16 * (size_t)h.lo, c->h_poses + 16 * (size_t)h.lo, sizeof(double) * 16 * cnt, cudaMemcpyHostToDevice, h.stream),
       "H2D poses");
    int* d1 = c->d_lm_lists + 2 * (size_t)h.lo;
    int* d2 = d1 + cnt;
    CU(cudaMemcpyAsync(d1, l1, sizeof(int) * 2 * cnt, cudaMemcpyHostToDevice, h.stream), "H2D job lists");
    c->stream = h.stream;  // the launchers issue on the context's current stream
    int r = NID_OK;
    if (sorted) {
      if (h.n1 > 0) r = launch_sorted_pass1(c, d1, l1, 0, h.n1, 1);
      if (r == NID_OK && h.n2 > 0) r = launch_sorted_pass2(c, d2, l2, 0, h.n2);
      if (r == NID_OK && h.n2 > 0) r = launch_gn_list(c, d2, 0, h.n2, delta, 1);
      if (r == NID_OK && h.n1 > 0) r = launch_gn_list(c, d1, 0, h.n1, delta, 0);  // chi2 of the trial poses (and, harmlessly, again of full jobs)
    } else {
      // natural-order kernels: contiguous ranges, everything recomputed -- jobs are compacted into the slots [lo, lo + na)
      r = NID_ERR_STATE;
    }
    c->stream = main_stream;
    if (r != NID_OK) return r;
    CU(cudaMemcpyAsync(c->h_out + 44 * (size_t)h.lo, c->gn + 44 * (size_t)h.lo, sizeof(double) * 44 * cnt, cudaMemcpyDeviceToHost, h.stream),
       "D2H gn");
    return NID_OK;
  };
  // natural-order path: the round-1 scheme (compacted contiguous ranges, full evaluations)
  std::vector<int> order;
  auto issue_natural = [&](Half& h) -> int {
    order.clear();
    int nj = 0;
    for (int j = h.lo; j < h.hi; j++) if (st[j].phase == 0) { order.push_back(j); nj++; }
    for (int j = h.lo; j < h.hi; j++) if (st[j].phase == 1) order.push_back(j);
    h.na = (int)order.size();
    if (h.na == 0) return NID_OK;
    for (int k = 0; k < h.na; k++) {
      nidhost::pose_to_mat16(st[order[k]].est, c->h_poses + 16 * (size_t)k);
      c->h_job_pair[k] = st[order[k]].pair;
    }
    CU(cudaMemcpyAsync(c->poses, c->h_poses, sizeof(double) * 16 * h.na, cudaMemcpyHostToDevice, h.stream), "H2D poses");
    CU(cudaMemcpyAsync(c->job_pair, c->h_job_pair, sizeof(int) * h.na, cudaMemcpyHostToDevice, h.stream), "H2D job_pair");
    const int r = launch_eval_mixed(c, 0, nj, h.na - nj, delta);
    if (r != NID_OK) return r;
    CU(cudaMemcpyAsync(c->h_out, c->gn, sizeof(double) * 44 * h.na, cudaMemcpyDeviceToHost, h.stream), "D2H gn");
    return NID_OK;
  };
  if (!sorted) {
    Half& h = H[0];
    rc = issue_natural(h);
    while (rc == NID_OK && h.na > 0) {
      if (cudaStreamSynchronize(h.stream) != cudaSuccess) { rc = check_cuda(cudaGetLastError(), "sync lm"); break; }
      const std::vector<int> done = order;
      for (int k = 0; k < (int)done.size(); k++) lm_absorb(c, st[done[k]], c->h_out + 44 * (size_t)k, n, max_iters);
      rc = issue_natural(h);
    }
  } else {
    for (int i = 0; i < nh && rc == NID_OK; i++) rc = issue(H[i]);
    while (rc == NID_OK && (H[0].na > 0 || (nh == 2 && H[1].na > 0))) {
      for (int i = 0; i < nh && rc == NID_OK; i++) {
        Half& h = H[i];
        if (h.na == 0) continue;
        if (cudaStreamSynchronize(h.stream) != cudaSuccess) { rc = check_cuda(cudaGetLastError(), "sync lm"); break; }
        for (int j = h.lo; j < h.hi; j++)
          if (st[j].phase != 2) lm_absorb(c, st[j], c->h_out + 44 * (size_t)j, n, max_iters);
        rc = issue(h);
      }
    }
  }
  c->stream = main_stream;
  if (nh == 2) cudaStreamSynchronize(c->stream2);
  cudaStreamSynchronize(main_stream);
  if (rc != NID_OK) return rc;
  for (int j = 0; j < n; j++) {
    memcpy(poses7 + 7 * j, st[j].est.t, sizeof(double) * 3);
    memcpy(poses7 + 7 * j + 3, st[j].est.q, sizeof(double) * 4);
    if (stats) { stats[3 * j] = st[j].it; stats[3 * j + 1] = st[j].jac_evals; stats[3 * j + 2] = st[j].cost_evals; }
  }
  return NID_OK;
}

int nid_solve(nid_ctx* c, int pair, double pose7[7], int max_iters, double delta, double* trace, int* stats) {
  if (!c) { set_error("NULL ctx"); return NID_ERR_ARG; }
  c->lm_trace = trace;
  c->lm_trace_cap = max_iters;
  int r = nid_solve_jobs(c, 1, &pair, pose7, max_iters, delta, stats);
  c->lm_trace = nullptr;
  return r;
}

int nid_hard_eval_jobs(nid_ctx* c, int n_jobs, const int* job_pair, const double* poses, double* total, double* nid_cells) {
  if (!c || !poses) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  OKR(stage_jobs(c, n_jobs, job_pair, poses, 1));
  c->staged_jobs = 0;  // (not evaluation jobs: the pairs need not be prepared)
  OKR(launch_hard(c, n_jobs));
  const size_t w = c->ncell + 1;
  CU(cudaMemcpyAsync(c->h_out, c->hard, sizeof(double) * w * n_jobs, cudaMemcpyDeviceToHost, c->stream), "D2H hard");
  CU(cudaStreamSynchronize(c->stream), "sync hard");
  for (int j = 0; j < n_jobs; j++) {
    if (total) total[j] = c->h_out[j * w + c->ncell];
    if (nid_cells) memcpy(nid_cells + (size_t)j * c->ncell, c->h_out + j * w, sizeof(double) * c->ncell);
  }
  return NID_OK;
}

static int warp_sample_common(nid_ctx* c, int pair, const double T_cw1[16], int f64) {
  if (!c || pair < 0 || pair >= c->n_pairs || !T_cw1) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!c->pair_set[pair]) { set_error("pair not set"); return NID_ERR_STATE; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  if (f64 && !c->d_pix) OKR(dalloc(&c->d_pix, (size_t)8 * c->N, "d_pix"));
  if (!f64 && !c->d_pix4) OKR(dalloc(&c->d_pix4, (size_t)4 * c->N, "d_pix4"));
  memcpy(c->h_aux_pose, T_cw1, sizeof(double) * 16);  // (both callers synchronise before returning)
  CU(cudaMemcpyAsync(c->aux_pose, c->h_aux_pose, sizeof(double) * 16, cudaMemcpyHostToDevice, c->stream), "H2D pose");
  return launch_warp_sample(c, pair, c->aux_pose, f64);
}

int nid_warp_sample(nid_ctx* c, int pair, const double T_cw1[16], float* out) {
  OKR(warp_sample_common(c, pair, T_cw1, 0));
  if (out) CU(cudaMemcpyAsync(out, c->d_pix4, sizeof(float) * 4 * c->N, cudaMemcpyDefault, c->stream), "D2H pix4");
  CU(cudaStreamSynchronize(c->stream), "sync warp_sample");
  return NID_OK;
}

// Kernel 1's target texture of one pair: three stacked fp16 planes (I, Gx/2, Gy/2) in a gather-able CUDA array, created
// and filled the first time the pair is used by nid_warp_sample_jobs after its target image changed.
static int ensure_k1_texture(nid_ctx* c, int pair) {
  if (c->k1_ready[pair]) return NID_OK;
  if (!c->k1_arrays[pair]) {
    cudaChannelFormatDesc cd = cudaCreateChannelDescHalf();
    CU(cudaMallocArray(&c->k1_arrays[pair], &cd, c->cols, 3 * c->rows, cudaArrayTextureGather), "cudaMallocArray (kernel-1 planes)");
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->k1_arrays[pair];
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    CU(cudaCreateTextureObject(&c->h_k1tex[pair], &rd, &td, nullptr), "cudaCreateTextureObject (kernel-1 planes)");
    CU(cudaMemcpyAsync(c->d_k1tex + pair, &c->h_k1tex[pair], sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice, c->stream), "H2D k1 texture handle");
  }
  if (!c->d_k1pack) OKR(dalloc(&c->d_k1pack, (size_t)3 * c->N, "d_k1pack"));
  OKR(launch_pack_k1(c, pair, c->d_k1pack));
  CU(cudaMemcpy2DToArrayAsync(c->k1_arrays[pair], 0, 0, c->d_k1pack, sizeof(unsigned short) * c->cols, sizeof(unsigned short) * c->cols,
                              3 * c->rows, cudaMemcpyDeviceToDevice, c->stream), "fp16 planes -> texture array");
  c->k1_ready[pair] = 1;
  return NID_OK;
}

int nid_warp_sample_jobs(nid_ctx* c, int n_jobs, const int* job_pair, const double* poses, float* out) {
  if (!c || !poses) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  if (n_jobs < 1 || n_jobs > c->max_jobs) { set_error("n_jobs out of range"); return NID_ERR_ARG; }
  if (c->sell_points) { set_error("nid_warp_sample_jobs needs depth pairs (nid_set_pair), not caller-supplied points"); return NID_ERR_UNSUPPORTED; }
  if (!c->d_pix4_jobs) OKR(dalloc(&c->d_pix4_jobs, (size_t)4 * c->N * c->max_jobs, "d_pix4_jobs"));
  OKR(stage_jobs(c, n_jobs, job_pair, poses, 1));
  c->staged_jobs = 0;  // (not evaluation jobs: the pairs need not be prepared)
  // pairs uploaded as raw 16-bit depth are read as such (2 B/px); a mixed batch falls back to the fp64 planes, which
  // every pair has
  bool u16 = true;
  for (int j = 0; j < n_jobs; j++) {
    u16 = u16 && c->pair_u16[c->h_job_pair[j]];
    OKR(ensure_k1_texture(c, c->h_job_pair[j]));
  }
  OKR(launch_warp_sample_jobs(c, n_jobs, (float4*)c->d_pix4_jobs, u16));
  if (out) CU(cudaMemcpyAsync(out, c->d_pix4_jobs, sizeof(float) * 4 * c->N * (size_t)n_jobs, cudaMemcpyDefault, c->stream), "D2H pix4 jobs");
  CU(cudaStreamSynchronize(c->stream), "sync warp_sample_jobs");
  return NID_OK;
}

int nid_warp_sample_f64(nid_ctx* c, int pair, const double T_cw1[16], double* out) {
  OKR(warp_sample_common(c, pair, T_cw1, 1));
  if (out) CU(cudaMemcpyAsync(out, c->d_pix, sizeof(double) * 8 * c->N, cudaMemcpyDefault, c->stream), "D2H pix8");
  CU(cudaStreamSynchronize(c->stream), "sync warp_sample_f64");
  return NID_OK;
}

int nid_debug_hist(nid_ctx* c, int job, int cell_index, double* P_t, double* P_j) {
  if (!c || job < 0 || job >= c->max_jobs || cell_index < 0 || cell_index >= c->ncell) { set_error("bad argument"); return NID_ERR_ARG; }
  const size_t hs = (size_t)c->bins * c->bins + c->bins;
  if (!c->hist) { set_error("set option keep_hist=1 before the evaluation"); return NID_ERR_STATE; }
  const double* src = c->hist + ((size_t)job * c->ncell + cell_index) * hs;
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  CU(cudaStreamSynchronize(c->stream), "sync");
  if (P_j) CU(cudaMemcpy(P_j, src, sizeof(double) * c->bins * c->bins, cudaMemcpyDeviceToHost), "D2H P_j");
  if (P_t) CU(cudaMemcpy(P_t, src + c->bins * c->bins, sizeof(double) * c->bins, cudaMemcpyDeviceToHost), "D2H P_t");
  return NID_OK;
}

long long nid_launch_count(nid_ctx* c) { return c ? c->launches : 0; }

int nid_kernel_times(nid_ctx* c, double ms[4], long long calls[4]) {
  if (!c || !ms || !calls) { set_error("bad argument"); return NID_ERR_ARG; }
  for (int i = 0; i < 4; i++) { ms[i] = c->kernel_ms[i]; calls[i] = c->kernel_calls[i]; }
  return NID_OK;
}

void* nid_stream(nid_ctx* c) { return c ? (void*)c->stream : nullptr; }

int nid_event_record(nid_ctx* c, int slot) {
  if (!c || slot < 0 || slot > 1) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  if (!c->ev[slot]) CU(cudaEventCreate(&c->ev[slot]), "cudaEventCreate");
  CU(cudaEventRecord(c->ev[slot], c->stream), "cudaEventRecord");
  return NID_OK;
}

int nid_event_elapsed_ms(nid_ctx* c, float* ms) {
  if (!c || !ms || !c->ev[0] || !c->ev[1]) { set_error("bad argument or events not recorded"); return NID_ERR_ARG; }
  CU(cudaEventSynchronize(c->ev[1]), "cudaEventSynchronize");
  CU(cudaEventElapsedTime(ms, c->ev[0], c->ev[1]), "cudaEventElapsedTime");
  return NID_OK;
}

int nid_set_option(nid_ctx* c, const char* key, int value) {
  if (!c || !key) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!strcmp(key, "force_strips")) {
    if (value < 0) { set_error("force_strips must be >= 0 (0 = automatic)"); return NID_ERR_ARG; }
    c->opt_force_strips = value;
    return NID_OK;
  }
  if (!strcmp(key, "path")) {
    if (value < 0 || value > 2) { set_error("path must be 0 (auto), 1 (natural) or 2 (sorted)"); return NID_ERR_ARG; }
    if (value == 2 && (c->bins > NID_SORTED_MAX_BINS || c->bins < NID_SORTED_MIN_BINS)) { set_error("the sorted path supports 6 to 40 bins"); return NID_ERR_UNSUPPORTED; }
    c->opt_path = value;
    update_task_kind(c);
    return NID_OK;
  }
  if (!strcmp(key, "sorted_mode")) {
    if (value < 0 || value > 2) { set_error("sorted_mode must be 0 (automatic), 1 (class tasks) or 2 (span tasks)"); return NID_ERR_ARG; }
    if (value == 2 && (c->bins > 20 || c->sell_points || !use_sorted(c))) { set_error("span tasks need the sorted path, depth pairs and at most 20 bins"); return NID_ERR_UNSUPPORTED; }
    c->opt_sorted_mode = value;
    update_task_kind(c);
    return NID_OK;
  }
  if (!strcmp(key, "keep_hist")) { c->opt_keep_hist = value; return NID_OK; }
  if (!strcmp(key, "lm_reuse")) { c->opt_lm_reuse = value ? 1 : 0; return NID_OK; }
  if (!strcmp(key, "lm_graph")) { c->opt_lm_graph = value ? 1 : 0; return NID_OK; }
  if (!strcmp(key, "asm_wide")) { c->opt_asm_wide = value < 0 ? 0 : std::min(value, 2); return NID_OK; }
  if (!strcmp(key, "stage_bulk")) { c->opt_stage_bulk = value < 0 ? -1 : (value ? 1 : 0); return NID_OK; }
  if (!strcmp(key, "lm_speculate")) {
    if (value < 0 || value > 8) { set_error("lm_speculate must be 0 (off) .. 8 trial poses per round"); return NID_ERR_ARG; }
    c->opt_lm_spec = value;
    return NID_OK;
  }
  if (!strcmp(key, "task_px")) {
    if (value < 8 || value > NID_TASK_PX_MAX || (value & 3)) { set_error("task_px must be a multiple of 4 in [8, 256]"); return NID_ERR_ARG; }
    if (value != c->task_px) {
      c->task_px = value;
      std::fill(c->pair_prepared.begin(), c->pair_prepared.end(), 0);  // the pixel store is laid out per task
    }
    return NID_OK;
  }
  if (!strcmp(key, "time_kernels")) {
    c->opt_time_kernels = value;
    for (int i = 0; i < 4; i++) { c->kernel_ms[i] = 0; c->kernel_calls[i] = 0; }
    return NID_OK;
  }
  set_error(std::string("unknown option ") + key);
  return NID_ERR_ARG;
}

}  // extern "C"
