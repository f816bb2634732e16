// Exact-signature replacements of the reference's three CUDA host entry points, so that the
// reference's own drivers (NID_pose_estimation.cpp, g2o's LM) link against this library unchanged:
//
//   void Calculate3Dpoint(...)        CudaPoints3d.cuh:6
//   void CudaComputeHref(...)         CudaComputeHref.cuh:6
//   void g2o::CudaComputeH(...)       g2o/g2o/core/computeH.cuh:8
//
// They keep the reference's contract (void, blocking, device 0, errors printed and swallowed,
// `-=` accumulation into caller-zeroed Href/Htarget/Hjoint, NaN for inactive cells, der untouched when
// calculate_der == false) and forward to the C-ABI. State that the reference recomputed or re-uploaded on
// every call (points, reference spline weights, images) lives in a process-global context keyed by the
// caller's `points3d` pointer, which the reference keeps alive for the whole solve
// (NID_pose_estimation.cpp:240-276). Call nid_shim_reset() when the buffers behind those pointers change.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "../../include/nid_b200.h"

namespace {

struct ShimCtx {
  nid_ctx* ctx = nullptr;
  int rows = 0, cols = 0, cell = 0, bins = 0;
  const double* im0 = nullptr;
  const double* im1 = nullptr;
  bool prepared = false;
};

std::mutex g_mu;
std::map<const double*, ShimCtx> g_ctx;  // key: points3d

void warn(const char* where) { fprintf(stderr, "[nid_b200 shim] %s failed: %s\n", where, nid_last_error()); }

ShimCtx* get_ctx(const double* points3d, int rows, int cols, int cell, int bins, int degree, const double* intr) {
  auto it = g_ctx.find(points3d);
  if (it != g_ctx.end()) {
    ShimCtx& s = it->second;
    if (s.rows == rows && s.cols == cols && s.cell == cell && s.bins == bins) return &s;
    nid_destroy(s.ctx);
    g_ctx.erase(it);
  }
  ShimCtx s;
  if (nid_create(&s.ctx, 0, rows, cols, cell, bins, degree, 1, 1) != NID_OK) { warn("nid_create"); return nullptr; }
  s.rows = rows; s.cols = cols; s.cell = cell; s.bins = bins;
  if (nid_set_pair_points(s.ctx, 0, points3d, nullptr, nullptr, intr) != NID_OK) {
    warn("nid_set_pair_points");
    nid_destroy(s.ctx);
    return nullptr;
  }
  return &(g_ctx[points3d] = s);
}

}  // namespace

extern "C" void nid_shim_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_ctx) nid_destroy(kv.second.ctx);
  g_ctx.clear();
}

// CudaPoints3d.cu:35-74. points_3d may be managed, device or host memory.
void Calculate3Dpoint(double* depth, double* pose_c2w, double* points_3d, double* camera_intrincis, int rows, int cols) {
  std::lock_guard<std::mutex> lk(g_mu);
  nid_ctx* c = nullptr;
  if (nid_create(&c, 0, rows, cols, 1, 8, 3, 1, 1) != NID_OK) { warn("Calculate3Dpoint/nid_create"); return; }
  std::vector<uint8_t> blank((size_t)rows * cols, 0);
  if (nid_set_pair(c, 0, depth, blank.data(), blank.data(), pose_c2w, camera_intrincis) != NID_OK ||
      nid_get_points3d(c, 0, points_3d) != NID_OK)
    warn("Calculate3Dpoint");
  nid_destroy(c);
  // a new set of points invalidates whatever was cached for this buffer
  auto it = g_ctx.find(points_3d);
  if (it != g_ctx.end()) { nid_destroy(it->second.ctx); g_ctx.erase(it); }
}

// CudaComputeHref.cu:139-223
void CudaComputeHref(double* im0, double* points3d, double* pose, double* camera_intrincis, int bin_num, int bs_degree,
                     int cell_num, int rows, int cols, double* bs_value, int* bs_index, int* bs_counter, double* Href) {
  std::lock_guard<std::mutex> lk(g_mu);
  ShimCtx* s = get_ctx(points3d, rows, cols, cell_num, bin_num, bs_degree, camera_intrincis);
  if (!s) return;
  if (nid_set_pair_points(s->ctx, 0, nullptr, im0, nullptr, nullptr) != NID_OK) { warn("CudaComputeHref/im0"); return; }
  s->im0 = im0;
  const int nc = cell_num * cell_num;
  std::vector<double> href(nc);
  if (nid_prepare(s->ctx, 0, pose, bs_counter, href.data()) != NID_OK) { warn("CudaComputeHref/nid_prepare"); return; }
  s->prepared = true;
  if (nid_get_ref_weights(s->ctx, 0, bs_value, bs_index) != NID_OK) { warn("CudaComputeHref/ref_weights"); return; }
  // `Href[i] -= p log2 p` into the caller-zeroed buffer (CudaComputeHref.cu:218); NaN when n_c < 300 (:206-209)
  for (int i = 0; i < nc; i++) Href[i] = isnan(href[i]) ? NAN : Href[i] + href[i];
  // side effect of the reference kernel: im0 is clamped in place where it was used (CudaComputeHref.cu:102-105)
  std::vector<uint8_t> inb((size_t)rows * cols);
  if (nid_get_inbounds(s->ctx, 0, inb.data()) == NID_OK) {
    for (size_t i = 0; i < inb.size(); i++)
      if (inb[i]) {
        if (im0[i] >= 255) im0[i] = 254.999;
        if (im0[i] < 0) im0[i] = 0;
      }
  }
}

namespace g2o {

// computeH.cu:373-502
void CudaComputeH(bool calculate_der, double* im0, double* im1, double* points3d, int* bs_counter, double* bs_ref,
                  int* bs_index_ref, double* pose, double* camera_intrincis, int bin_num, int bs_degree, int cell_num,
                  int rows, int cols, double* Href, double* pro_target, double* pro_joint, double* Htarget,
                  double* Hjoint, double* der) {
  (void)pro_target; (void)pro_joint; (void)bs_index_ref;
  std::lock_guard<std::mutex> lk(g_mu);
  ShimCtx* s = get_ctx(points3d, rows, cols, cell_num, bin_num, bs_degree, camera_intrincis);
  if (!s) return;
  if (s->im0 != im0) {
    // the clamped value 254.999 written back by CudaComputeHref maps to the same spline weights as 255
    std::vector<double> tmp(im0, im0 + (size_t)rows * cols);
    for (auto& v : tmp) if (v == 254.999) v = 255.0;
    if (nid_set_pair_points(s->ctx, 0, nullptr, tmp.data(), nullptr, nullptr) != NID_OK) { warn("CudaComputeH/im0"); return; }
    s->im0 = im0;
  }
  if (s->im1 != im1) {
    if (nid_set_pair_points(s->ctx, 0, nullptr, nullptr, im1, nullptr) != NID_OK) { warn("CudaComputeH/im1"); return; }
    s->im1 = im1;
  }
  if (!s->prepared) {
    // CudaComputeHref was not routed through this library: adopt the caller's prepare
    if (nid_import_prepare(s->ctx, 0, bs_ref, bs_counter, Href) != NID_OK) { warn("CudaComputeH/import_prepare"); return; }
    s->prepared = true;
  }
  const int nc = cell_num * cell_num;
  std::vector<double> ht(nc), hj(nc), dj(calculate_der ? 6 * nc : 0);
  if (nid_eval(s->ctx, 0, pose, calculate_der ? 1 : 0, ht.data(), hj.data(), calculate_der ? dj.data() : nullptr) != NID_OK) {
    warn("CudaComputeH/nid_eval");
    return;
  }
  for (int i = 0; i < nc; i++) {
    // `Htarget[i] -= ...` on the caller's zeros, NaN for inactive cells (computeH.cu:271-299)
    Htarget[i] = isnan(ht[i]) ? NAN : Htarget[i] + ht[i];
    Hjoint[i] = isnan(hj[i]) ? NAN : Hjoint[i] + hj[i];
  }
  if (calculate_der) memcpy(der, dj.data(), sizeof(double) * 6 * nc);
}

}  // namespace g2o

// C-linkage trampolines so that the shims can be exercised through ctypes in tests
extern "C" {
void nid_shim_Calculate3Dpoint(double* depth, double* pose_c2w, double* points_3d, double* intr, int rows, int cols) {
  Calculate3Dpoint(depth, pose_c2w, points_3d, intr, rows, cols);
}
void nid_shim_CudaComputeHref(double* im0, double* points3d, double* pose, double* intr, int bin_num, int bs_degree,
                              int cell_num, int rows, int cols, double* bs_value, int* bs_index, int* bs_counter,
                              double* Href) {
  CudaComputeHref(im0, points3d, pose, intr, bin_num, bs_degree, cell_num, rows, cols, bs_value, bs_index, bs_counter, Href);
}
void nid_shim_CudaComputeH(int calculate_der, double* im0, double* im1, double* points3d, int* bs_counter, double* bs_ref,
                           int* bs_index_ref, double* pose, double* intr, int bin_num, int bs_degree, int cell_num,
                           int rows, int cols, double* Href, double* pro_target, double* pro_joint, double* Htarget,
                           double* Hjoint, double* der) {
  g2o::CudaComputeH(calculate_der != 0, im0, im1, points3d, bs_counter, bs_ref, bs_index_ref, pose, intr, bin_num,
                    bs_degree, cell_num, rows, cols, Href, pro_target, pro_joint, Htarget, Hjoint, der);
}
}
