/*
 * nid_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C++17, fp64, no Eigen / OpenCV) of the reference's CPU
 * implementation of the NID cost + 6-DoF Jacobian path and of the g2o
 * Levenberg-Marquardt loop that drives it.  Every function cites the reference
 * file:line (relative to the upstream tree) whose arithmetic it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.  The product (the CUDA library
 * behind include/nid_b200.h) never links, imports or calls it.
 *
 * Parity pin: the reference ships no tests, golden vectors or fixtures for this
 * path ("parity unpinned" by upstream).  This oracle is pinned instead against
 * outputs of the reference's own CUDA implementation (g2o/g2o/core/computeH.cu,
 * CudaPoints3d.cu) compiled unmodified into oracle/_ref/ and executed on a B200;
 * the vectors live in tests/golden/ref_gpu_*.npz next to the script that made
 * them (oracle/gen_ref_golden.py).
 *
 * Conventions
 *   pose7  = {tx,ty,tz, qx,qy,qz,qw}   (g2o SE3Quat::toVector order, se3quat.h:144-154)
 *   mat16  = column-major 4x4          (Eigen default; what the reference hands to CUDA)
 *   intr   = {fx,fy,cx,cy}
 *   cell c = ci*cell + cj, row-major   (NID_pose_estimation.cpp:283-337)
 */
#ifndef NID_ORACLE_H
#define NID_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_problem orc_problem;

/* ---- B-spline (types_six_dof_expmap.cpp:738-800; knots generalised per computeH.cu:99-134) */
double orc_bspline(int index, int order, double u, int bins);
double orc_bspline_der(int index, int order, double u, int bins);

/* ---- SE3 algebra (se3quat.h, se3_ops.hpp, Eigen quaternion conventions) */
void orc_se3_from_Rt(const double R_rowmajor[9], const double t[3], double pose7[7]);
void orc_se3_exp(const double upd6[6], double pose7[7]);
void orc_se3_mul(const double a7[7], const double b7[7], double out7[7]);
void orc_se3_inverse(const double a7[7], double out7[7]);
void orc_se3_to_mat16(const double pose7[7], double mat16[16]);
void orc_se3_map(const double pose7[7], const double p[3], double out[3]);
/* the reference's pose perturbation, NID_pose_estimation.cpp:186-208 */
void orc_reference_perturbation(const double T_wc1_mat16[16], double pose7_cw1_perturbed[7]);

/* ---- Huber (robust_kernel_impl.cpp:65-90; note `float dsqr`, robust_kernel_impl.h:84) */
void orc_huber(double chi2, double delta, double rho[3]);

/* ---- 6x6 pivoted LDLT solve (linear_solver_dense.h:65-113 / Eigen::LDLT). returns 1 if positive */
int orc_ldlt6_solve(const double H_rowmajor[36], const double b[6], double x[6]);

/* ---- problem = one frame pair, cell x cell unary NID edges on one pose vertex */
orc_problem* orc_create(const uint8_t* im0, const double* depth, const uint8_t* im1,
                        int rows, int cols, const double T_wc0_mat16[16],
                        const double intr[4], int cell, int bins, int threads);
void orc_destroy(orc_problem*);
/* jac_bound_gpu=1 switches the Jacobian bounds test to the reference GPU's `u+3<=cols`
 * (computeH.cu:164) instead of the CPU's `u+3<=cols-1` (types_six_dof_expmap.cpp:433);
 * used only to pin the oracle against the reference CUDA outputs. Default 0. */
void orc_set_quirks(orc_problem*, int jac_bound_gpu, int warp_with_matrix);

/* a1: world points (NID_pose_estimation.cpp:401-432 / CudaPoints3d.cu:5-32). out: 3*rows*cols, NaN if invalid */
void orc_points3d(const orc_problem*, double* out);
int  orc_cell_points(const orc_problem*, int c);

/* a2: computeHref at the initial pose (types_six_dof_expmap.cpp:655-725).
 * n_c[cell^2] in-bounds counts, Href[cell^2] (NaN when n_c<300). */
void orc_prepare(orc_problem*, const double pose7[7], int* n_c, double* Href);
/* per-pixel reference spline data as the reference GPU API exposes it
 * (CudaComputeHref.cu:33-135): bs_value[4*N], bs_index[N]; pixels that are invalid or
 * out of bounds at the initial pose get weight 0 / index 0 (CPU semantics, SURVEY B-6). */
void orc_ref_weights(const orc_problem*, double* bs_value, int* bs_index);

/* a6-a10: ComputeH (+ linearizeOplus when want_jac) at `pose7`
 * (types_six_dof_expmap.cpp:544-637, 381-529, .h:220-228).
 * Outputs per cell; inactive cells (n_c<300) get NaN. Any output pointer may be NULL. */
void orc_eval(orc_problem*, const double pose7[7], int want_jac,
              double* Ht, double* Hj, double* err, double* J6);
/* histograms of the last orc_eval for one cell: P_t[bins], P_j[bins*bins] (normalised) */
void orc_last_hist(const orc_problem*, int c, double* P_t, double* P_j);
/* derivative tensors of the last orc_eval(want_jac=1) for one cell:
 * dP_t[bins*6], dP_j[bins*bins*6] (normalised) */
void orc_last_dhist(const orc_problem*, int c, double* dP_t, double* dP_j);

/* per-pixel warp/sample record at `pose7` for kernel-1 parity: out[8*N] =
 * {u, v, I_c(clamped), g_x, g_y, valid_cost, valid_jac, p_z}; NaN rows for invalid depth */
void orc_pixels(const orc_problem*, const double pose7[7], double* out);

/* a11 + LM (optimization_algorithm_levenberg.cpp:61-250, sparse_optimizer.cpp:356-450,
 * block_solver.hpp:502-570, base_unary_edge.hpp:43-72). pose7 in/out.
 * trace (may be NULL): per outer iteration {chi2_after, lambda, lm_trials, tx,ty,tz,qx,qy,qz,qw} = 10 doubles.
 * returns number of outer iterations performed. counts (may be NULL): {jac_evals, cost_evals}. */
int orc_optimize(orc_problem*, double pose7[7], int max_iters, double huber_delta,
                 double* trace, int* counts);
/* GN pieces at a pose, for unit tests: chi2 (robust), H[36] row-major, b[6] */
void orc_gn_system(orc_problem*, const double pose7[7], double huber_delta,
                   double* chi2, double* H36, double* b6);

/* a12: hard-binned NID (NID_standard_property.cpp:342-485). Returns sqrt(sum nid_c^2);
 * nid_cells[cell^2] (may be NULL); sparse cells contribute 0 (SURVEY B-10). T_cw1 given as mat16. */
double orc_hard_nid(const uint8_t* im0, const double* depth, const uint8_t* im1,
                    int rows, int cols, const double T_wc0_mat16[16],
                    const double T_cw1_mat16[16], const double intr[4],
                    int cell, int bins, double* nid_cells, int threads);

#ifdef __cplusplus
}
#endif
#endif
