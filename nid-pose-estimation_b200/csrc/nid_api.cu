// C-ABI of the B200-native NID path (include/nid_b200.h): context, staging, and the host-side
// Levenberg-Marquardt driver (the only CPU arithmetic is the 6x6 solve and the pose update, as in the
// reference: optimization_algorithm_levenberg.cpp:61-225).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <string>
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
  if (c->opt_force_strips > 0) return std::min(c->opt_force_strips, c->rb);
  long long want = 2LL * c->sm_count;
  long long per = (long long)c->ncell * n_jobs;
  int S = (int)((want + per - 1) / per);
  S = std::max(1, std::min(S, std::min(c->rb, 32)));
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
  p.sl_off = c->sl_off; p.sl_task = c->sl_task; p.sl_cell = c->sl_cell; p.nslices = c->nslices;
  p.sell_cap = c->sell_cap; p.max_slices = c->max_slices; p.Twc0 = c->Twc0;
  p.tasks = c->tasks; p.ntasks = c->ntasks; p.cell_task_start = c->cell_task_start; p.cell_slice_start = c->cell_slice_start;
  p.cls_task_start = c->cls_task_start; p.span_start = c->span_start; p.wv = c->wv;
  p.max_tasks = c->max_tasks; p.g_stride = (int)c->g_stride;
  p.G = c->G; p.fp1 = c->fp1; p.tex2 = c->d_tex2;
  return p;
}

// Path selection. The sorted path (no floating-point atomics) wins wherever it is available: measured on B200 at
// 640x480, cost+Jacobian: 4x4 cells/16 bins 77k vs 9k evals/s, the reference's default 16x16 cells/10 bins 29k vs 9k
// (tools/time_config.py). The natural-order kernels remain for bin counts outside [8, 40] and as a second,
// independently written implementation the parity tests hold to the same bar.
bool use_sorted(const nid_ctx* c) {
  // the assembly tables must fit in shared memory; below 8 bins every span is an end span
  if (c->bins > NID_SORTED_MAX_BINS || c->bins < 8) return false;
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
  const size_t J = c->max_jobs, NC = c->ncell;
  const size_t hs = (size_t)c->bins * c->bins + c->bins;
  if (c->opt_keep_hist && !c->hist) OKR(dalloc(&c->hist, J * NC * hs, "hist"));
  if (use_sorted(c)) {
    const size_t need = (size_t)std::max(c->max_ntasks_prepared, 1);
    if (c->g_stride < need) {
      CU(cudaStreamSynchronize(c->stream), "sync before growing job buffers");
      if (c->G) cudaFree(c->G);
      c->G = nullptr;
      OKR(dalloc(&c->G, J * need * c->bins, "G"));
      c->g_stride = need;
    }
    if (!c->jpart_s) OKR(dalloc(&c->jpart_s, J * (size_t)c->max_slices * 6, "jpart_s"));  // one partial per slice
    if (!c->wv) OKR(dalloc(&c->wv, J * NC * hs, "wv"));
  } else if (!c->part) {
    c->part_slots = J + 2 * (size_t)c->sm_count + 64;
    OKR(dalloc(&c->part, c->part_slots * NC * hs, "part"));
    OKR(dalloc(&c->jpart, c->part_slots * NC * 6, "jpart"));
  }
  return NID_OK;
}

static int update_texture(nid_ctx* c, int pair) {
  OKR(launch_pack_fp(c, pair));
  OKR(launch_pack_tex(c, pair, c->d_pack));
  CU(cudaMemcpy2DToArrayAsync(c->tex2_arrays[pair], 0, 0, c->d_pack, sizeof(unsigned) * c->cols, sizeof(unsigned) * c->cols,
                              c->rows, cudaMemcpyDeviceToDevice, c->stream), "packed im1 -> texture array");
  return NID_OK;
}

static int stage_jobs(nid_ctx* c, int n_jobs, const int* job_pair, const double* poses) {
  if (n_jobs < 1 || n_jobs > c->max_jobs) { set_error("n_jobs out of range"); return NID_ERR_ARG; }
  for (int j = 0; j < n_jobs; j++) {
    int pr = job_pair ? job_pair[j] : 0;
    if (pr < 0 || pr >= c->n_pairs) { set_error("job_pair out of range"); return NID_ERR_ARG; }
    if (!c->pair_prepared[pr]) { set_error("pair not prepared (call nid_prepare)"); return NID_ERR_STATE; }
    if (use_sorted(c) && !c->pair_sorted[pr]) {
      set_error("pair not prepared for the sorted path (call nid_prepare after selecting the path)");
      return NID_ERR_STATE;
    }
    c->h_job_pair[j] = pr;
  }
  memcpy(c->h_poses, poses, sizeof(double) * 16 * n_jobs);
  CU(cudaMemcpyAsync(c->poses, c->h_poses, sizeof(double) * 16 * n_jobs, cudaMemcpyHostToDevice, c->stream), "H2D poses");
  CU(cudaMemcpyAsync(c->job_pair, c->h_job_pair, sizeof(int) * n_jobs, cudaMemcpyHostToDevice, c->stream), "H2D job_pair");
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
  if (rows < 8 || cols < 8 || cell < 1 || bins < 7 || bins > 64 || n_pairs < 1 || max_jobs < 1 || cell > rows || cell > cols) {
    set_error("bad geometry (need rows,cols>=8, 1<=cell<=min(rows,cols), 7<=bins<=64, n_pairs,max_jobs>=1)");
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
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  c->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  const size_t N = c->N, P = n_pairs, J = max_jobs, NC = c->ncell;
  const size_t hs = (size_t)bins * bins + bins;
  OKR(dalloc(&c->pwx, P * N, "pwx")); OKR(dalloc(&c->pwy, P * N, "pwy")); OKR(dalloc(&c->pwz, P * N, "pwz"));
  OKR(dalloc(&c->im0, P * N, "im0")); OKR(dalloc(&c->im1, P * N, "im1")); OKR(dalloc(&c->inb0, P * N, "inb0"));
  OKR(dalloc(&c->n_c, P * NC, "n_c")); OKR(dalloc(&c->href, P * NC, "href"));
  OKR(dalloc(&c->cam, P * 4, "cam")); OKR(dalloc(&c->Twc0, P * 16, "Twc0"));
  OKR(dalloc(&c->cnt, P * NC * NID_NCLS, "cnt"));
  // pixels per task: short tasks when few evaluations are in flight (more threads), longer ones for batches
  c->task_px = 32;
  const int min_task_px = 16;
  c->max_tasks = (int)(N / min_task_px + NC * NID_NCLS + 1);
  c->max_slices = c->max_tasks / 32 + (int)NC + 1;
  // pixel slots per pair: every task is padded to a multiple of 4 and every cell's slices to their longest task
  c->sell_cap = (N + 3 * (size_t)c->max_tasks + NC * 32 * (size_t)(NID_TASK_PX_MAX + 4) + 255) / 128 * 128;
  OKR(dalloc(&c->depth, P * N, "depth"));
  OKR(dalloc(&c->sl_off, P * ((size_t)c->max_slices + 1), "sl_off"));
  OKR(dalloc(&c->sl_task, P * (size_t)c->max_slices * 32, "sl_task"));
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
    CU(cudaMemcpy(c->d_tex2, c->h_tex2.data(), sizeof(cudaTextureObject_t) * P, cudaMemcpyHostToDevice), "H2D tex2 handles");
    OKR(dalloc(&c->d_pack, N, "d_pack"));
    OKR(dalloc(&c->fp1, P * N, "fp1"));
    c->h_Twc0.assign(P * 16, 0.0);
    c->h_cam.assign(P * 4, 1.0);
  }
  OKR(dalloc(&c->d_depth, N, "d_depth")); OKR(dalloc(&c->d_img64, N, "d_img64")); OKR(dalloc(&c->d_flag, 1, "d_flag"));
  OKR(dalloc(&c->lut_w, 256 * 4, "lut_w")); OKR(dalloc(&c->lut_k, 256, "lut_k"));
  OKR(dalloc(&c->poses, J * 16, "poses")); OKR(dalloc(&c->job_pair, J, "job_pair"));
  (void)hs;
  OKR(dalloc(&c->ht, J * NC, "ht")); OKR(dalloc(&c->hj, J * NC, "hj")); OKR(dalloc(&c->err, J * NC, "err"));
  OKR(dalloc(&c->der, J * NC * 6, "der")); OKR(dalloc(&c->gn, J * 44, "gn"));
  OKR(dalloc(&c->hard, J * (NC + 1), "hard"));
  CU(cudaMallocHost((void**)&c->h_poses, sizeof(double) * 16 * J), "pinned poses");
  CU(cudaMallocHost((void**)&c->h_job_pair, sizeof(int) * J), "pinned job_pair");
  CU(cudaMallocHost((void**)&c->h_out, sizeof(double) * J * (NC * 8 + 44)), "pinned out");
  c->pair_set.assign(P, 0);
  c->pair_prepared.assign(P, 0);
  c->pair_sorted.assign(P, 0);
  OKR(launch_build_lut(c));
  OKR(sorted_init(c));
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
                  c->job_pair, c->part, c->jpart, c->hist, c->ht, c->hj, c->err, c->der, c->gn, c->hard,
                  c->depth, c->sd0, c->sd1, c->sd2, c->sid, c->sl_off, c->sl_task, c->sl_cell, c->nslices, c->task_pos,
                  c->d_tex2, c->d_pack, c->tasks, c->ntasks, c->cell_task_start, c->cell_slice_start, c->G, c->jpart_s, c->fp1, c->cls_task_start, c->span_start, c->wv};
  for (auto t : c->h_tex2) if (t) cudaDestroyTextureObject(t);
  for (auto arr : c->tex2_arrays) if (arr) cudaFreeArray(arr);
  for (void* p : ptrs) if (p) cudaFree(p);
  if (c->h_poses) cudaFreeHost(c->h_poses);
  if (c->h_job_pair) cudaFreeHost(c->h_job_pair);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->h_cnt) cudaFreeHost(c->h_cnt);
  for (int i = 0; i < 2; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 5; i++) if (c->kev[i]) cudaEventDestroy(c->kev[i]);
  cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  delete c;
  return NID_OK;
}

int nid_sync(nid_ctx* c) {
  CU(cudaStreamSynchronize(c->stream), "nid_sync");
  return NID_OK;
}

static int set_pair_common(nid_ctx* c, int pair, const double* depth, const double T_wc0[16], const double intr[5]) {
  if (pair < 0 || pair >= c->n_pairs) { set_error("pair index out of range"); return NID_ERR_ARG; }
  if (!depth || !T_wc0 || !intr) { set_error("NULL argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  CU(cudaMemcpyAsync(c->depth + (size_t)pair * c->N, depth, sizeof(double) * c->N, cudaMemcpyDefault, c->stream), "H2D depth");
  CU(cudaMemcpyAsync(c->Twc0 + 16 * pair, T_wc0, sizeof(double) * 16, cudaMemcpyDefault, c->stream), "H2D Twc0");
  CU(cudaMemcpyAsync(c->cam + 4 * pair, intr, sizeof(double) * 4, cudaMemcpyDefault, c->stream), "H2D intr");
  memcpy(c->h_Twc0.data() + 16 * (size_t)pair, T_wc0, sizeof(double) * 16);
  memcpy(c->h_cam.data() + 4 * (size_t)pair, intr, sizeof(double) * 4);
  OKR(launch_points(c, pair));
  c->pair_prepared[pair] = 0;
  return NID_OK;
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
static int build_sorted_layout(nid_ctx* c, int pair);

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
  OKR(build_sorted_layout(c, pair));
  CU(cudaStreamSynchronize(c->stream), "sync import");
  c->pair_prepared[pair] = 1;
  return NID_OK;
}

int nid_get_inbounds(nid_ctx* c, int pair, uint8_t* flags) {
  if (!c || pair < 0 || pair >= c->n_pairs || !flags) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!c->pair_prepared[pair]) { set_error("pair not prepared"); return NID_ERR_STATE; }
  CU(cudaMemcpyAsync(flags, c->inb0 + (size_t)pair * c->N, c->N, cudaMemcpyDefault, c->stream), "D2H inb0");
  CU(cudaStreamSynchronize(c->stream), "sync inb0");
  return NID_OK;
}

int nid_get_points3d(nid_ctx* c, int pair, double* points_3d) {
  if (!c || pair < 0 || pair >= c->n_pairs || !points_3d) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!c->pair_set[pair]) { set_error("pair not set"); return NID_ERR_STATE; }
  if (!c->d_pix) OKR(dalloc(&c->d_pix, (size_t)8 * c->N, "d_pix"));
  OKR(launch_points_aos(c, pair, c->d_pix));
  CU(cudaMemcpyAsync(points_3d, c->d_pix, sizeof(double) * 3 * c->N, cudaMemcpyDefault, c->stream), "copy points3d");
  CU(cudaStreamSynchronize(c->stream), "sync points3d");
  return NID_OK;
}

// Regroup the pair's valid pixels by (cell, reference class), cut the segments into tasks of at most
// task_px pixels, order the tasks by length and pack them 32 to a slice (nid_sorted.cu). The class counts
// come from the device; offsets and tables are integer bookkeeping done here, the pixels are moved by
// k_scatter_sell.
static int build_sorted_layout(nid_ctx* c, int pair) {
  const int NC = c->ncell, L = c->task_px;
  c->pair_sorted[pair] = 0;
  if (!use_sorted(c)) return NID_OK;  // the natural-order kernels need none of this
  if (!c->sd0) {
    OKR(dalloc(&c->sd0, (size_t)c->n_pairs * c->sell_cap, "sd0"));
    OKR(dalloc(&c->sid, (size_t)c->n_pairs * c->sell_cap, "sid"));
  }
  if (c->sell_points && !c->sd1) {
    OKR(dalloc(&c->sd1, (size_t)c->n_pairs * c->sell_cap, "sd1"));
    OKR(dalloc(&c->sd2, (size_t)c->n_pairs * c->sell_cap, "sd2"));
  }
  const size_t n_cnt = (size_t)NC * NID_NCLS;
  if (c->h_cnt_cap < n_cnt) {
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    c->h_cnt = nullptr;
    c->h_cnt_cap = 0;
    CU(cudaMallocHost((void**)&c->h_cnt, sizeof(unsigned int) * n_cnt), "pinned cnt");
    c->h_cnt_cap = n_cnt;
  }
  const unsigned int* cnt = c->h_cnt;
  CU(cudaMemcpyAsync(c->h_cnt, c->cnt + (size_t)pair * NC * NID_NCLS, sizeof(unsigned int) * n_cnt,
                     cudaMemcpyDeviceToHost, c->stream), "D2H cnt");
  CU(cudaStreamSynchronize(c->stream), "sync cnt");
  std::vector<int> cts(NC + 1), clsts((size_t)NC * (NID_NCLS + 1));
  std::vector<int2> tasks;
  tasks.reserve(c->max_tasks);
  for (int cell = 0; cell < NC; cell++) {
    cts[cell] = (int)tasks.size();
    long long n_c = 0;
    for (int v = 0; v < 256; v++) n_c += cnt[(size_t)cell * NID_NCLS + v];
    const bool active = n_c >= NID_MIN_CELL_POINTS;  // inactive cells get no work at all
    for (int v = 0; v < NID_NCLS; v++) {
      const int len = (int)cnt[(size_t)cell * NID_NCLS + v];
      clsts[(size_t)cell * (NID_NCLS + 1) + v] = (int)tasks.size();
      for (int o = 0; active && o < len; o += L) {
        int2 t;
        t.x = o;
        t.y = std::min(L, len - o) | (v << 9) | (cell << 18);
        tasks.push_back(t);
      }
    }
    clsts[(size_t)cell * (NID_NCLS + 1) + NID_NCLS] = (int)tasks.size();
  }
  cts[NC] = (int)tasks.size();
  const int nt = (int)tasks.size();
  if (nt > c->max_tasks || NC > 0x3fff) { set_error("task table overflow"); return NID_ERR_STATE; }
  // Slices: the tasks of a cell ordered by length (longest first; counting sort, ties keep task order) and cut
  // into groups of 32. Slices never mix cells: the lanes of a warp then sample one cell-sized region of the
  // target image (texture-cache locality) and still run out of work together.
  std::vector<int> sl_off(1, 0), sl_task, sl_cell, task_pos(std::max(nt, 1)), css(NC + 1, 0);
  sl_task.reserve((size_t)nt + 32 * (size_t)NC);
  long long off = 0;
  {
    std::vector<int> order, bucket(NID_TASK_PX_MAX + 2);
    for (int cell = 0; cell < NC; cell++) {
      css[cell] = (int)sl_off.size() - 1;
      const int t0 = cts[cell], t1 = cts[cell + 1];
      if (t1 == t0) continue;
      order.assign(t1 - t0, 0);
      std::fill(bucket.begin(), bucket.end(), 0);
      for (int t = t0; t < t1; t++) bucket[NID_TASK_PX_MAX - (tasks[t].y & 0x1ff) + 1]++;
      for (int i = 1; i <= NID_TASK_PX_MAX + 1; i++) bucket[i] += bucket[i - 1];
      for (int t = t0; t < t1; t++) order[bucket[NID_TASK_PX_MAX - (tasks[t].y & 0x1ff)]++] = t;
      for (int s0 = 0; s0 < t1 - t0; s0 += 32) {
        const int longest = tasks[order[s0]].y & 0x1ff;
        for (int l = 0; l < 32; l++) {
          const int t = s0 + l < t1 - t0 ? order[s0 + l] : -1;
          sl_task.push_back(t);
          if (t >= 0) task_pos[t] = (int)off + 4 * l;
        }
        off += (long long)((longest + 3) / 4) * 128;
        sl_off.push_back((int)off);
        sl_cell.push_back(cell);
      }
    }
  }
  const int ns = (int)sl_off.size() - 1;
  css[NC] = ns;
  if ((size_t)off > c->sell_cap || ns > c->max_slices) { set_error("sliced pixel store overflow"); return NID_ERR_STATE; }
  // One pinned staging arena for all tables (the copies are then truly asynchronous; the arena is not reused before
  // the next nid_prepare's first synchronisation on the same stream).
  {
    const size_t n_i = (size_t)2 * nt + nt + sl_task.size() + sl_cell.size() + (ns + 1) + 2 * (size_t)(NC + 1) + clsts.size() + 2;
    if (c->h_stage_cap < n_i) {
      if (c->h_stage) cudaFreeHost(c->h_stage);
      c->h_stage = nullptr;
      c->h_stage_cap = 0;
      CU(cudaMallocHost((void**)&c->h_stage, sizeof(int) * (n_i + n_i / 4)), "pinned stage");
      c->h_stage_cap = n_i + n_i / 4;
    }
    int* w = c->h_stage;
    auto put = [&](void* dst, const void* src, size_t n_ints, const char* what) -> int {
      if (!n_ints) return NID_OK;
      memcpy(w, src, sizeof(int) * n_ints);
      CU(cudaMemcpyAsync(dst, w, sizeof(int) * n_ints, cudaMemcpyHostToDevice, c->stream), what);
      w += n_ints;
      return NID_OK;
    };
    OKR(put(c->tasks + (size_t)pair * c->max_tasks, tasks.data(), (size_t)2 * nt, "H2D tasks"));
    OKR(put(c->task_pos + (size_t)pair * c->max_tasks, task_pos.data(), nt, "H2D task_pos"));
    OKR(put(c->sl_task + (size_t)pair * c->max_slices * 32, sl_task.data(), nt ? sl_task.size() : 0, "H2D sl_task"));
    OKR(put(c->sl_off + (size_t)pair * (c->max_slices + 1), sl_off.data(), ns + 1, "H2D sl_off"));
    OKR(put(c->sl_cell + (size_t)pair * c->max_slices, sl_cell.data(), sl_cell.size(), "H2D sl_cell"));
    OKR(put(c->nslices + pair, &ns, 1, "H2D nslices"));
    OKR(put(c->cell_task_start + (size_t)pair * (NC + 1), cts.data(), NC + 1, "H2D cell_task_start"));
    OKR(put(c->cell_slice_start + (size_t)pair * (NC + 1), css.data(), NC + 1, "H2D cell_slice_start"));
    OKR(put(c->ntasks + pair, &nt, 1, "H2D ntasks"));
    OKR(put(c->cls_task_start + (size_t)pair * NC * (NID_NCLS + 1), clsts.data(), clsts.size(), "H2D cls_task_start"));
  }
  OKR(launch_scatter(c, pair));
  c->h_ntasks[pair] = nt;
  c->h_nslices[pair] = ns;
  c->pair_sorted[pair] = 1;
  c->max_ntasks_prepared = 0;
  for (int v : c->h_ntasks) c->max_ntasks_prepared = std::max(c->max_ntasks_prepared, v);
  c->max_nslices_prepared = 0;
  for (int v : c->h_nslices) c->max_nslices_prepared = std::max(c->max_nslices_prepared, v);
  return NID_OK;
}

int nid_prepare(nid_ctx* c, int pair, const double T_cw1[16], int* bs_counter, double* Href) {
  if (!c || pair < 0 || pair >= c->n_pairs || !T_cw1) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!c->pair_set[pair]) { set_error("pair not set"); return NID_ERR_STATE; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  memcpy(c->h_poses, T_cw1, sizeof(double) * 16);
  CU(cudaMemcpyAsync(c->poses, c->h_poses, sizeof(double) * 16, cudaMemcpyHostToDevice, c->stream), "H2D pose");
  OKR(launch_prepare(c, pair, c->poses));
  OKR(launch_href(c, pair));
  OKR(build_sorted_layout(c, pair));
  int* h_nc = (int*)c->h_out;
  double* h_href = c->h_out + c->ncell;  // ncell ints fit in ncell doubles
  CU(cudaMemcpyAsync(h_nc, c->n_c + pair * c->ncell, sizeof(int) * c->ncell, cudaMemcpyDeviceToHost, c->stream), "D2H n_c");
  CU(cudaMemcpyAsync(h_href, c->href + pair * c->ncell, sizeof(double) * c->ncell, cudaMemcpyDeviceToHost, c->stream), "D2H href");
  CU(cudaStreamSynchronize(c->stream), "sync prepare");
  if (bs_counter) memcpy(bs_counter, h_nc, sizeof(int) * c->ncell);
  if (Href) memcpy(Href, h_href, sizeof(double) * c->ncell);
  c->pair_prepared[pair] = 1;
  return NID_OK;
}

int nid_get_ref_weights(nid_ctx* c, int pair, double* bs_value, int* bs_index) {
  if (!c || pair < 0 || pair >= c->n_pairs) { set_error("bad argument"); return NID_ERR_ARG; }
  if (!c->pair_prepared[pair]) { set_error("pair not prepared"); return NID_ERR_STATE; }
  if (!c->d_bsv) { OKR(dalloc(&c->d_bsv, (size_t)4 * c->N, "d_bsv")); OKR(dalloc(&c->d_bsi, (size_t)c->N, "d_bsi")); }
  OKR(launch_ref_weights(c, pair));
  if (bs_value) CU(cudaMemcpyAsync(bs_value, c->d_bsv, sizeof(double) * 4 * c->N, cudaMemcpyDefault, c->stream), "D2H bs_value");
  if (bs_index) CU(cudaMemcpyAsync(bs_index, c->d_bsi, sizeof(int) * c->N, cudaMemcpyDefault, c->stream), "D2H bs_index");
  CU(cudaStreamSynchronize(c->stream), "sync ref_weights");
  return NID_OK;
}

int nid_stage_jobs(nid_ctx* c, int n_jobs, const int* job_pair, const double* poses) {
  if (!c || !poses) { set_error("bad argument"); return NID_ERR_ARG; }
  return stage_jobs(c, n_jobs, job_pair, poses);
}

int nid_eval_staged(nid_ctx* c, int n_jobs, int want_jac) {
  if (!c || n_jobs < 1 || n_jobs > c->max_jobs) { set_error("bad argument"); return NID_ERR_ARG; }
  return launch_eval(c, n_jobs, want_jac);
}

int nid_fetch_results(nid_ctx* c, int n_jobs, int want_jac, double* Ht, double* Hj, double* der) {
  if (!c || n_jobs < 1 || n_jobs > c->max_jobs) { set_error("bad argument"); return NID_ERR_ARG; }
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

// Lock-step LM over n problems. The problems are cut into two halves that ping-pong: while the device evaluates the
// jobs of one half (its own stream and its own range of job slots), the host absorbs the results of the other half,
// solves the 6x6 systems and stages the next poses. Each problem sees exactly the schedule of a solo solve.
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
  // (the natural-order kernels size their per-job partial buffers by the job count of a launch: one range there)
  const int nh = (n >= 8 && use_sorted(c)) ? 2 : 1;
  if (nh == 2 && !c->stream2) {
    CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking), "cudaStreamCreate (LM)");
  }
  CU(cudaStreamSynchronize(c->stream), "sync before LM");
  struct Half { int lo, hi, base, na; std::vector<int> order; cudaStream_t stream; };
  Half H[2];
  H[0] = {0, nh == 2 ? n / 2 : n, 0, 0, {}, c->stream};
  H[1] = {n / 2, n, n / 2, 0, {}, c->stream2};
  cudaStream_t const main_stream = c->stream;
  int rc = NID_OK;
  // stage and launch the jobs of half h: first every problem that needs cost+Jacobian, then every trial pose
  auto issue = [&](Half& h) -> int {
    h.order.clear();
    int nj = 0;
    for (int j = h.lo; j < h.hi; j++) if (st[j].phase == 0) { h.order.push_back(j); nj++; }
    for (int j = h.lo; j < h.hi; j++) if (st[j].phase == 1) h.order.push_back(j);
    h.na = (int)h.order.size();
    if (h.na == 0) return NID_OK;
    for (int k = 0; k < h.na; k++) {
      nidhost::pose_to_mat16(st[h.order[k]].est, c->h_poses + 16 * (size_t)(h.base + k));
      c->h_job_pair[h.base + k] = st[h.order[k]].pair;
    }
    CU(cudaMemcpyAsync(c->poses + 16 * (size_t)h.base, c->h_poses + 16 * (size_t)h.base, sizeof(double) * 16 * h.na,
                       cudaMemcpyHostToDevice, h.stream), "H2D poses");
    CU(cudaMemcpyAsync(c->job_pair + h.base, c->h_job_pair + h.base, sizeof(int) * h.na, cudaMemcpyHostToDevice, h.stream),
       "H2D job_pair");
    c->stream = h.stream;  // the launchers issue on the context's current stream
    const int r = launch_eval_mixed(c, h.base, nj, h.na - nj, delta);
    c->stream = main_stream;
    if (r != NID_OK) return r;
    CU(cudaMemcpyAsync(c->h_out + 44 * (size_t)h.base, c->gn + 44 * (size_t)h.base, sizeof(double) * 44 * h.na,
                       cudaMemcpyDeviceToHost, h.stream), "D2H gn");
    return NID_OK;
  };
  for (int i = 0; i < nh && rc == NID_OK; i++) rc = issue(H[i]);
  while (rc == NID_OK && (H[0].na > 0 || (nh == 2 && H[1].na > 0))) {
    for (int i = 0; i < nh && rc == NID_OK; i++) {
      Half& h = H[i];
      if (h.na == 0) continue;
      if (cudaStreamSynchronize(h.stream) != cudaSuccess) { rc = check_cuda(cudaGetLastError(), "sync lm"); break; }
      for (int k = 0; k < h.na; k++) lm_absorb(c, st[h.order[k]], c->h_out + 44 * (size_t)(h.base + k), n, max_iters);
      rc = issue(h);
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
  if (n_jobs < 1 || n_jobs > c->max_jobs) { set_error("n_jobs out of range"); return NID_ERR_ARG; }
  for (int j = 0; j < n_jobs; j++) {
    int pr = job_pair ? job_pair[j] : 0;
    if (pr < 0 || pr >= c->n_pairs || !c->pair_set[pr]) { set_error("job_pair invalid or pair not set"); return NID_ERR_ARG; }
    c->h_job_pair[j] = pr;
  }
  memcpy(c->h_poses, poses, sizeof(double) * 16 * n_jobs);
  CU(cudaMemcpyAsync(c->poses, c->h_poses, sizeof(double) * 16 * n_jobs, cudaMemcpyHostToDevice, c->stream), "H2D poses");
  CU(cudaMemcpyAsync(c->job_pair, c->h_job_pair, sizeof(int) * n_jobs, cudaMemcpyHostToDevice, c->stream), "H2D job_pair");
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
  memcpy(c->h_poses, T_cw1, sizeof(double) * 16);
  CU(cudaMemcpyAsync(c->poses, c->h_poses, sizeof(double) * 16, cudaMemcpyHostToDevice, c->stream), "H2D pose");
  return launch_warp_sample(c, pair, c->poses, f64);
}

int nid_warp_sample(nid_ctx* c, int pair, const double T_cw1[16], float* out) {
  OKR(warp_sample_common(c, pair, T_cw1, 0));
  if (out) CU(cudaMemcpyAsync(out, c->d_pix4, sizeof(float) * 4 * c->N, cudaMemcpyDefault, c->stream), "D2H pix4");
  CU(cudaStreamSynchronize(c->stream), "sync warp_sample");
  return NID_OK;
}

int nid_warp_sample_jobs(nid_ctx* c, int n_jobs, const int* job_pair, const double* poses, float* out) {
  if (!c || !poses) { set_error("bad argument"); return NID_ERR_ARG; }
  CU(cudaSetDevice(c->device), "cudaSetDevice");
  if (n_jobs < 1 || n_jobs > c->max_jobs) { set_error("n_jobs out of range"); return NID_ERR_ARG; }
  if (c->sell_points) { set_error("nid_warp_sample_jobs needs depth pairs (nid_set_pair), not caller-supplied points"); return NID_ERR_UNSUPPORTED; }
  if (!c->d_tex2) { set_error("nid_warp_sample_jobs needs the packed target textures (8 to 40 bins)"); return NID_ERR_UNSUPPORTED; }
  for (int j = 0; j < n_jobs; j++) {
    const int pr = job_pair ? job_pair[j] : 0;
    if (pr < 0 || pr >= c->n_pairs || !c->pair_set[pr]) { set_error("job_pair invalid or pair not set"); return NID_ERR_ARG; }
    c->h_job_pair[j] = pr;
  }
  if (!c->d_pix4_jobs) OKR(dalloc(&c->d_pix4_jobs, (size_t)4 * c->N * c->max_jobs, "d_pix4_jobs"));
  memcpy(c->h_poses, poses, sizeof(double) * 16 * n_jobs);
  CU(cudaMemcpyAsync(c->poses, c->h_poses, sizeof(double) * 16 * n_jobs, cudaMemcpyHostToDevice, c->stream), "H2D poses");
  CU(cudaMemcpyAsync(c->job_pair, c->h_job_pair, sizeof(int) * n_jobs, cudaMemcpyHostToDevice, c->stream), "H2D job_pair");
  OKR(launch_warp_sample_jobs(c, n_jobs, (float4*)c->d_pix4_jobs));
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
  if (!strcmp(key, "force_strips")) { c->opt_force_strips = value; return NID_OK; }
  if (!strcmp(key, "path")) {
    if (value < 0 || value > 2) { set_error("path must be 0 (auto), 1 (natural) or 2 (sorted)"); return NID_ERR_ARG; }
    if (value == 2 && (c->bins > NID_SORTED_MAX_BINS || c->bins < 8)) { set_error("the sorted path supports 8 to 40 bins"); return NID_ERR_UNSUPPORTED; }
    c->opt_path = value;
    return NID_OK;
  }
  if (!strcmp(key, "keep_hist")) { c->opt_keep_hist = value; return NID_OK; }
  if (!strcmp(key, "task_px")) {
    if (value < 16 || value > NID_TASK_PX_MAX || (value & 3)) { set_error("task_px must be a multiple of 4 in [16, 256]"); return NID_ERR_ARG; }
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
