// TEST INFRASTRUCTURE ONLY. C-linkage trampolines onto the reference's own (unmodified)
// CUDA host entry points, which are compiled from the upstream tree by oracle/Makefile:
//   Calculate3Dpoint      CudaPoints3d.cuh:6
//   g2o::CudaComputeH     g2o/g2o/core/computeH.cuh:8
//   CudaComputeHref       CudaComputeHref.cuh:6   (with the cudaMemset fix of oracle/Makefile)
#include <cuda_runtime.h>
void CudaComputeHref(double* im0, double* points3d, double* pose, double* camera_intrincis, int bin_num, int bs_degree,
                     int cell_num, int rows, int cols, double* bs_value, int* bs_index, int* bs_counter, double* Href);
void Calculate3Dpoint(double* depth, double* pose_c2w, double* points_3d, double* camera_intrincis, int rows, int cols);
namespace g2o {
void CudaComputeH(bool calculate_der, double* im0, double* im1, double* points3d, int* bs_counter, double* bs_ref,
                  int* bs_index_ref, double* pose, double* camera_intrincis, int bin_num, int bs_degree, int cell_num,
                  int rows, int cols, double* Href, double* pro_target, double* pro_joint, double* Htarget,
                  double* Hjoint, double* der);
}
extern "C" {
void ref_Calculate3Dpoint(double* depth, double* pose_c2w, double* points_3d, double* intr, int rows, int cols) {
  Calculate3Dpoint(depth, pose_c2w, points_3d, intr, rows, cols);
}
void ref_CudaComputeH(int calculate_der, double* im0, double* im1, double* points3d, int* bs_counter, double* bs_ref,
                      int* bs_index_ref, double* pose, double* intr, int bin_num, int bs_degree, int cell_num, int rows,
                      int cols, double* Href, double* Htarget, double* Hjoint, double* der) {
  g2o::CudaComputeH(calculate_der != 0, im0, im1, points3d, bs_counter, bs_ref, bs_index_ref, pose, intr, bin_num,
                    bs_degree, cell_num, rows, cols, Href, nullptr, nullptr, Htarget, Hjoint, der);
}
void ref_CudaComputeHref(double* im0, double* points3d, double* pose, double* intr, int bin_num, int bs_degree, int cell_num,
                         int rows, int cols, double* bs_value, int* bs_index, int* bs_counter, double* Href) {
  CudaComputeHref(im0, points3d, pose, intr, bin_num, bs_degree, cell_num, rows, cols, bs_value, bs_index, bs_counter, Href);
}
// the reference keeps points3d / im0 / im1 in managed memory (NID_pose_estimation.cpp:240-242)
void* ref_managed_alloc(size_t bytes) { void* p = nullptr; cudaMallocManaged(&p, bytes); return p; }
void ref_managed_free(void* p) { cudaFree(p); }
int ref_last_error() { return (int)cudaGetLastError(); }
}
