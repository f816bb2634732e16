/*
 * nid_oracle.cpp — TEST INFRASTRUCTURE ONLY (see nid_oracle.h).
 *
 * fp64 CPU restatement of the reference CPU path. "parity pin": outputs of the
 * reference's own CUDA kernels run on a B200 (tests/golden/ref_gpu_*.npz).
 * All citations are relative to the upstream tree (arpg/NID-Pose-Estimation).
 */
#include "nid_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double kSigma = 1e-30;  // types_six_dof_expmap.h:281
const double kNaN = std::numeric_limits<double>::quiet_NaN();

// ---------------------------------------------------------------- B-spline
// Clamped uniform knot vector t_k = clamp(k-3, 0, B-3), k = 0..B+3. This is the
// closed form of every table in computeH.cu:99-112 and of knots_[14] (B=10) in
// types_six_dof_expmap.h:287.
inline double knot(int k, int bins) {
  int v = k - 3;
  if (v < 0) v = 0;
  if (v > bins - 3) v = bins - 3;
  return (double)v;
}

// types_six_dof_expmap.cpp:738-764
double bspline(int index, int order, double u, int bins) {
  double coef1, coef2;
  if (order == 1) {
    if (index == 0)
      if ((knot(index, bins) <= u) && (u <= knot(index + 1, bins))) return 1.0;
    if ((knot(index, bins) < u) && (u <= knot(index + 1, bins))) return 1.0;
    else return 0.0;
  } else {
    if (knot(index + order - 1, bins) == knot(index, bins)) {
      coef1 = (u == knot(index, bins)) ? 1 : 0;
    } else {
      coef1 = (u - knot(index, bins)) / (knot(index + order - 1, bins) - knot(index, bins));
    }
    if (knot(index + order, bins) == knot(index + 1, bins)) {
      coef2 = (u == knot(index + order, bins)) ? 1 : 0;
    } else {
      coef2 = (knot(index + order, bins) - u) / (knot(index + order, bins) - knot(index + 1, bins));
    }
    return coef1 * bspline(index, order - 1, u, bins) + coef2 * bspline(index + 1, order - 1, u, bins);
  }
}

// types_six_dof_expmap.cpp:766-800
double bspline_der(int index, int order, double u, int bins) {
  double coef1, coef2, coef3, coef4;
  if (order == 1) return 0.0;
  if (knot(index + order - 1, bins) == knot(index, bins)) {
    coef1 = (u == knot(index, bins)) ? 1 : 0;
    coef3 = 0.0;
  } else {
    coef1 = (u - knot(index, bins)) / (knot(index + order - 1, bins) - knot(index, bins));
    coef3 = 1.0 / (knot(index + order - 1, bins) - knot(index, bins));
  }
  if (knot(index + order, bins) == knot(index + 1, bins)) {
    coef2 = (u == knot(index + order, bins)) ? 1 : 0;
    coef4 = 0.0;
  } else {
    coef2 = (knot(index + order, bins) - u) / (knot(index + order, bins) - knot(index + 1, bins));
    coef4 = -1.0 / (knot(index + order, bins) - knot(index + 1, bins));
  }
  return coef1 * bspline_der(index, order - 1, u, bins) + coef2 * bspline_der(index + 1, order - 1, u, bins) +
         coef3 * bspline(index, order - 1, u, bins) + coef4 * bspline(index + 1, order - 1, u, bins);
}

// ---------------------------------------------------------------- SE3 (quaternion + translation)
struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };

// Eigen Quaternion(Matrix3) — Eigen/src/Geometry/Quaternion.h (quaternionbase_assign_impl<.,3,3>)
Quat quat_from_R(const double m[3][3]) {
  double q[3];
  Quat out;
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    out.w = 0.5 * t;
    t = 0.5 / t;
    out.x = (m[2][1] - m[1][2]) * t;
    out.y = (m[0][2] - m[2][0]) * t;
    out.z = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    out.w = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
    out.x = q[0]; out.y = q[1]; out.z = q[2];
  }
  return out;
}

// Eigen QuaternionBase::toRotationMatrix
void quat_to_R(const Quat& q, double R[3][3]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz;       R[0][2] = txz + twy;
  R[1][0] = txy + twz;       R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy;       R[2][1] = tyz + twx;       R[2][2] = 1 - (txx + tyy);
}

// Eigen quaternion product
Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

// Eigen QuaternionBase::_transformVector: v + w*uv + q.vec x uv, uv = 2 q.vec x v
inline void quat_rot(const Quat& q, const double v[3], double out[3]) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
  out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
  out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}

// se3quat.h:280-285
void normalize_rotation(Quat& q) {
  if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

// se3quat.h:217-220
inline void se3_map(const SE3& T, const double p[3], double out[3]) {
  quat_rot(T.r, p, out);
  out[0] += T.t[0]; out[1] += T.t[1]; out[2] += T.t[2];
}

// se3quat.h:54-56 : SE3Quat(R,t)
SE3 se3_from_Rt(const double R[3][3], const double t[3]) {
  SE3 T;
  T.r = quat_from_R(R);
  T.t[0] = t[0]; T.t[1] = t[1]; T.t[2] = t[2];
  normalize_rotation(T.r);
  return T;
}

// se3quat.h:102-108 : operator*
SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r = a;
  double rt[3];
  quat_rot(a.r, b.t, rt);
  r.t[0] += rt[0]; r.t[1] += rt[1]; r.t[2] += rt[2];
  r.r = quat_mul(a.r, b.r);
  normalize_rotation(r.r);
  return r;
}

// se3quat.h:124-129
SE3 se3_inverse(const SE3& a) {
  SE3 r;
  r.r.x = -a.r.x; r.r.y = -a.r.y; r.r.z = -a.r.z; r.r.w = a.r.w;
  double nt[3] = {a.t[0] * -1., a.t[1] * -1., a.t[2] * -1.};
  quat_rot(r.r, nt, r.t);
  return r;
}

void mat3_mul(const double A[3][3], const double B[3][3], double C[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += A[i][k] * B[k][j];
      C[i][j] = s;
    }
}

// se3_ops.hpp:27-38
void skew(const double v[3], double m[3][3]) {
  m[0][0] = 0; m[0][1] = -v[2]; m[0][2] = v[1];
  m[1][0] = v[2]; m[1][1] = 0; m[1][2] = -v[0];
  m[2][0] = -v[1]; m[2][1] = v[0]; m[2][2] = 0;
}

// se3quat.h:223-257
SE3 se3_exp(const double upd[6]) {
  double omega[3] = {upd[0], upd[1], upd[2]};
  double upsilon[3] = {upd[3], upd[4], upd[5]};
  double theta = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  double Om[3][3], Om2[3][3], R[3][3], V[3][3];
  skew(omega, Om);
  mat3_mul(Om, Om, Om2);
  if (theta < 0.00001) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        R[i][j] = (i == j ? 1.0 : 0.0) + Om[i][j] + Om2[i][j];
        V[i][j] = R[i][j];
      }
  } else {
    double a = std::sin(theta) / theta;
    double b = (1 - std::cos(theta)) / (theta * theta);
    double c = (theta - std::sin(theta)) / (std::pow(theta, 3));
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        R[i][j] = (i == j ? 1.0 : 0.0) + a * Om[i][j] + b * Om2[i][j];
        V[i][j] = (i == j ? 1.0 : 0.0) + b * Om[i][j] + c * Om2[i][j];
      }
  }
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = V[i][0] * upsilon[0] + V[i][1] * upsilon[1] + V[i][2] * upsilon[2];
  // SE3Quat(const Quaterniond& q, const Vector3d& t), se3quat.h:60-62
  SE3 T;
  T.r = quat_from_R(R);
  T.t[0] = t[0]; T.t[1] = t[1]; T.t[2] = t[2];
  normalize_rotation(T.r);
  return T;
}

// se3quat.h:270-278 ; column-major like Eigen's .data()
void se3_to_mat16(const SE3& T, double m[16]) {
  double R[3][3];
  quat_to_R(T.r, R);
  for (int c = 0; c < 3; c++) {
    for (int r = 0; r < 3; r++) m[4 * c + r] = R[r][c];
    m[4 * c + 3] = 0.0;
  }
  m[12] = T.t[0]; m[13] = T.t[1]; m[14] = T.t[2]; m[15] = 1.0;
}

SE3 se3_from7(const double p[7]) {
  SE3 T;
  T.t[0] = p[0]; T.t[1] = p[1]; T.t[2] = p[2];
  T.r.x = p[3]; T.r.y = p[4]; T.r.z = p[5]; T.r.w = p[6];
  return T;
}
void se3_to7(const SE3& T, double p[7]) {
  p[0] = T.t[0]; p[1] = T.t[1]; p[2] = T.t[2];
  p[3] = T.r.x; p[4] = T.r.y; p[5] = T.r.z; p[6] = T.r.w;
}

// ---------------------------------------------------------------- Huber
// robust_kernel_impl.cpp:65-90. `dsqr` is declared float (robust_kernel_impl.h:84).
void huber(double e, double delta, double rho[3]) {
  const double dsqr = (double)(float)(delta * delta);
  if (e <= dsqr) {
    rho[0] = e; rho[1] = 1.; rho[2] = 0.;
  } else {
    double sqrte = std::sqrt(e);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e;
  }
}

// ---------------------------------------------------------------- 6x6 LDLT with diagonal pivoting
// linear_solver_dense.h:104-110 uses Eigen::LDLT (lower, symmetric pivoting on the largest
// |diagonal|) and accepts the solution only if isPositive(). Restated after the published
// algorithm of Eigen/src/Cholesky/LDLT.h (ldlt_inplace<Lower>::unblocked and solve_impl).
int ldlt6_solve(const double Hin[36], const double b[6], double x[6]) {
  const int n = 6;
  double A[6][6];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) A[i][j] = Hin[i * 6 + j];
  int transp[6];
  int sign = 0;  // 0 ZeroSign, 1 PositiveSemiDef, -1 NegativeSemiDef, 2 Indefinite
  bool found_zero_pivot = false;
  for (int k = 0; k < n; ++k) {
    int piv = k; double big = std::fabs(A[k][k]);
    for (int i = k + 1; i < n; i++) if (std::fabs(A[i][i]) > big) { big = std::fabs(A[i][i]); piv = i; }
    transp[k] = piv;
    if (piv != k) {
      // symmetric swap of rows/cols k and piv, touching the lower triangle only
      int s = n - piv - 1;
      for (int j = 0; j < k; j++) std::swap(A[k][j], A[piv][j]);
      for (int i = 0; i < s; i++) std::swap(A[piv + 1 + i][k], A[piv + 1 + i][piv]);
      std::swap(A[k][k], A[piv][piv]);
      for (int i = k + 1; i < piv; ++i) std::swap(A[i][k], A[piv][i]);
    }
    int rs = n - k - 1;
    double temp[6];
    if (k > 0) {
      for (int j = 0; j < k; j++) temp[j] = A[j][j] * A[k][j];
      double acc = 0;
      for (int j = 0; j < k; j++) acc += A[k][j] * temp[j];
      A[k][k] -= acc;
      for (int i = 0; i < rs; i++) {
        double a2 = 0;
        for (int j = 0; j < k; j++) a2 += A[k + 1 + i][j] * temp[j];
        A[k + 1 + i][k] -= a2;
      }
    }
    double realAkk = A[k][k];
    bool pivot_is_valid = (std::fabs(realAkk) > 0.0);
    if (k == 0 && !pivot_is_valid) {
      sign = 0;
      for (int j = 0; j < n; ++j) { transp[j] = j; }
      // matrix is zero
      break;
    }
    if (rs > 0 && pivot_is_valid) {
      for (int i = 0; i < rs; i++) A[k + 1 + i][k] /= realAkk;
    } else if (rs > 0) {
      bool allzero = true;
      for (int i = 0; i < rs; i++) if (A[k + 1 + i][k] != 0.0) allzero = false;
      if (!allzero) sign = 2;
    }
    if (found_zero_pivot && pivot_is_valid) sign = 2;
    else if (!pivot_is_valid) found_zero_pivot = true;
    if (sign == 1) { if (realAkk < 0) sign = 2; }
    else if (sign == -1) { if (realAkk > 0) sign = 2; }
    else if (sign == 0) { if (realAkk > 0) sign = 1; else if (realAkk < 0) sign = -1; }
  }
  bool positive = (sign == 1 || sign == 0);
  if (!positive) return 0;
  // solve: dst = P b ; L^-1 ; D^-1 (pseudo-inverse) ; L^-T ; P^T
  double y[6];
  for (int i = 0; i < n; i++) y[i] = b[i];
  for (int k = 0; k < n; k++) std::swap(y[k], y[transp[k]]);
  for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) y[i] -= A[i][j] * y[j];
  const double tolerance = std::numeric_limits<double>::min();
  for (int i = 0; i < n; i++) {
    if (std::fabs(A[i][i]) > tolerance) y[i] /= A[i][i];
    else y[i] = 0;
  }
  for (int i = n - 1; i >= 0; i--) for (int j = i + 1; j < n; j++) y[i] -= A[j][i] * y[j];
  for (int k = n - 1; k >= 0; k--) std::swap(y[k], y[transp[k]]);
  for (int i = 0; i < n; i++) x[i] = y[i];
  return 1;
}

// ---------------------------------------------------------------- edge (one cell)
struct Edge {
  // what NID_pose_estimation.cpp:289-307 stores on the edge in CPU mode
  std::vector<double> xw;          // x_world_set_ (3 per point)
  std::vector<double> meas;        // _measurement (reference intensity as double)
  std::vector<int> loc;            // pixel_location_
  // set_bspline_relates (types_six_dof_expmap.cpp:639-653)
  std::vector<double> bs_ref;      // bs_value_ref_ (4 per point), zero rows for OOB-at-init points
  std::vector<double> icur;        // intensity_current_
  std::vector<double> pro_cur, pro_ref, pro_joint;
  std::vector<double> d_pt, d_pj;  // normalised derivative tensors of the last linearizeOplus
  double H_cur = 0, H_ref = 0, H_joint = 0;
  int ob = 0;
  int level = 0;
  double err = 0;
  double J[6] = {0, 0, 0, 0, 0, 0};
};

}  // namespace

struct orc_problem {
  int rows, cols, cell, bins, degree, threads;
  double fx, fy, cx, cy;
  double T_wc0[16];
  std::vector<uint8_t> im0, im1;
  std::vector<double> depth;
  std::vector<Edge> edges;
  int jac_bound_gpu = 0;
  int warp_with_matrix = 0;
  bool prepared = false;
};

namespace {

// types_six_dof_expmap.h:310-328 (uchar image, (int) truncation). x == -1.0 or y == -1.0 exactly (the gradient
// taps of a point with u == 0 or v == 0, which the Jacobian bounds test :433 admits) index column / row -1
// upstream: an out-of-bounds cv::Mat access, undefined. Defined here as the continuous extension of the
// (-1, 0) interval: index clamped to 0, fraction -1.
inline double interp(const orc_problem& P, double x, double y) {
  int ix = (int)x;
  int iy = (int)y;
  if (ix < 0) ix = 0;
  if (iy < 0) iy = 0;
  double dx = x - ix;
  double dy = y - iy;
  double dxdy = dx * dy;
  const uint8_t* r0 = &P.im1[(size_t)iy * P.cols];
  const uint8_t* r1 = &P.im1[(size_t)(iy + 1) * P.cols];
  return double(dxdy * r1[ix + 1] + (dy - dxdy) * r1[ix] + (dx - dxdy) * r0[ix + 1] +
                (1 - dx - dy + dxdy) * r0[ix]);
}

inline void warp_point(const orc_problem& P, const SE3& T, const double m16[16], const double* xw, double pc[3]) {
  if (P.warp_with_matrix) {
    // computeH.cu:152-154
    pc[0] = m16[0] * xw[0] + m16[4] * xw[1] + m16[8] * xw[2] + m16[12];
    pc[1] = m16[1] * xw[0] + m16[5] * xw[1] + m16[9] * xw[2] + m16[13];
    pc[2] = m16[2] * xw[0] + m16[6] * xw[1] + m16[10] * xw[2] + m16[14];
  } else {
    se3_map(T, xw, pc);
  }
}

// NID_pose_estimation.cpp:401-432
void build_edge_points(const orc_problem& P, int ci, int cj, Edge& e) {
  int row_block = P.rows / P.cell;
  int col_block = P.cols / P.cell;
  int row_start = row_block * ci, col_start = col_block * cj;
  int row_end = row_block * (ci + 1), col_end = col_block * (cj + 1);
  const double* T = P.T_wc0;
  for (int i = row_start; i < row_end; i++)
    for (int j = col_start; j < col_end; j++) {
      double z_p = P.depth[(size_t)i * P.cols + j];
      if (z_p < 0.01 || z_p > 100) continue;
      double x_p = z_p * (j - P.cx) / P.fx;
      double y_p = z_p * (i - P.cy) / P.fy;
      // (T_wc * Vector4d(x,y,z,1)).head(3), column-major T
      e.xw.push_back(T[0] * x_p + T[4] * y_p + T[8] * z_p + T[12] * 1.0);
      e.xw.push_back(T[1] * x_p + T[5] * y_p + T[9] * z_p + T[13] * 1.0);
      e.xw.push_back(T[2] * x_p + T[6] * y_p + T[10] * z_p + T[14] * 1.0);
      e.meas.push_back((double)P.im0[(size_t)i * P.cols + j]);
      e.loc.push_back(i * P.cols + j);
    }
}

// types_six_dof_expmap.cpp:639-653
void set_bspline_relates(const orc_problem& P, Edge& e) {
  size_t n = e.meas.size();
  e.pro_cur.assign(P.bins, 0.0);
  e.pro_ref.assign(P.bins, 0.0);
  e.pro_joint.assign((size_t)P.bins * P.bins, 0.0);
  e.bs_ref.assign(4 * n, 0.0);
  e.icur.assign(n, 0.0);
}

// types_six_dof_expmap.cpp:655-725
void compute_href(const orc_problem& P, const SE3& T, Edge& e) {
  const int B = P.bins, d = P.degree;
  double m16[16];
  se3_to_mat16(T, m16);
  e.ob = 0;
  e.H_ref = 0.0;
  e.level = 0;
  std::fill(e.pro_ref.begin(), e.pro_ref.end(), 0.0);
  std::fill(e.bs_ref.begin(), e.bs_ref.end(), 0.0);
  size_t n = e.meas.size();
  for (size_t i = 0; i < n; i++) {
    double pc[3];
    warp_point(P, T, m16, &e.xw[3 * i], pc);
    double u = P.fx * pc[0] / pc[2] + P.cx;
    double v = P.fy * pc[1] / pc[2] + P.cy;
    if (u >= 0 && u + 3 <= P.cols && v >= 0 && v + 3 <= P.rows) {
      e.icur[i] = interp(P, u, v);
    } else {
      e.ob++;
      continue;
    }
    double obs = e.meas[i];
    if (obs >= 255) obs = 254.999;
    if (obs < 0) obs = 0.0;
    double bin_pos_ref = obs * (B - d) / 255.0;
    int k = (int)std::floor(bin_pos_ref);
    for (int m = 0; m < 4; m++) {
      double w = bspline(k + m, d + 1, bin_pos_ref, B);
      e.bs_ref[4 * i + m] = w;
      e.pro_ref[k + m] += w;
    }
  }
  if ((int)n - e.ob < 300) {
    e.level = 1;
    return;
  }
  for (int i = 0; i < B; i++) e.pro_ref[i] /= ((int)n - e.ob);
  for (int i = 0; i < B; i++) {
    if (e.pro_ref[i] < kSigma) continue;
    e.H_ref -= e.pro_ref[i] * std::log2(e.pro_ref[i]);
  }
}

// ClearPrevH (types_six_dof_expmap.cpp:727-736) + ComputeH (:544-637) + error (.h:227)
void compute_h(const orc_problem& P, const SE3& T, Edge& e) {
  const int B = P.bins, d = P.degree;
  double m16[16];
  se3_to_mat16(T, m16);
  std::fill(e.pro_cur.begin(), e.pro_cur.end(), 0.0);
  std::fill(e.pro_joint.begin(), e.pro_joint.end(), 0.0);
  e.H_joint = 0.0;
  e.H_cur = 0.0;
  size_t n = e.meas.size();
  for (size_t i = 0; i < n; i++) {
    double pc[3];
    warp_point(P, T, m16, &e.xw[3 * i], pc);
    double obs = e.meas[i];
    if (obs >= 255) obs = 254.999;
    if (obs < 0) obs = 0.0;
    double bin_pos_ref = obs * (B - d) / 255.0;
    int kr = (int)std::floor(bin_pos_ref);
    double u = P.fx * pc[0] / pc[2] + P.cx;
    double v = P.fy * pc[1] / pc[2] + P.cy;
    if (u >= 0 && u + 3 <= P.cols && v >= 0 && v + 3 <= P.rows) {
      e.icur[i] = interp(P, u, v);
    } else {
      continue;
    }
    if (e.icur[i] >= 255) e.icur[i] = 254.999;
    if (e.icur[i] < 0) e.icur[i] = 0.0;
    double bin_pos_cur = e.icur[i] * (B - 3.0) / 255.0;
    int kt = (int)std::floor(bin_pos_cur);
    double wt[4];
    for (int m = 0; m < 4; m++) wt[m] = bspline(kt + m, d + 1, bin_pos_cur, B);
    for (int m = 0; m < 4; m++) e.pro_cur[kt + m] += wt[m];
    for (int m = 0; m < d + 1; m++)
      for (int q = 0; q < d + 1; q++)
        e.pro_joint[(size_t)(kr + m) * B + kt + q] += e.bs_ref[4 * i + m] * wt[q];
  }
  int nc = (int)n - e.ob;
  if (nc < 300) {
    e.level = 1;
    return;
  }
  for (int i = 0; i < B; i++) e.pro_cur[i] /= nc;
  for (size_t i = 0; i < (size_t)B * B; i++) e.pro_joint[i] /= nc;
  for (int i = 0; i < B; i++) {
    if (e.pro_cur[i] < kSigma) continue;
    e.H_cur -= e.pro_cur[i] * std::log2(e.pro_cur[i]);
  }
  for (size_t i = 0; i < (size_t)B * B; i++) {
    if (e.pro_joint[i] < kSigma) continue;
    e.H_joint -= e.pro_joint[i] * std::log2(e.pro_joint[i]);
  }
  e.err = (2 * e.H_joint - e.H_ref - e.H_cur) / e.H_joint;
}

// types_six_dof_expmap.cpp:381-529 (CPU branch). Must follow compute_h at the same pose.
void linearize(const orc_problem& P, const SE3& T, Edge& e) {
  const int B = P.bins, d = P.degree;
  double m16[16];
  se3_to_mat16(T, m16);
  std::vector<double>& dpt = e.d_pt;
  std::vector<double>& dpj = e.d_pj;
  dpt.assign((size_t)B * 6, 0.0);
  dpj.assign((size_t)B * B * 6, 0.0);
  const double d_mi_i = (B - d) / 255.0;
  size_t n = e.meas.size();
  const int jac_cols = P.jac_bound_gpu ? P.cols : P.cols - 1;
  for (size_t i = 0; i < n; i++) {
    double pc[3];
    warp_point(P, T, m16, &e.xw[3 * i], pc);
    double obs = e.meas[i];
    if (obs >= 255) obs = 254.999;
    if (obs < 0) obs = 0.0;
    double bin_pos_ref = obs * (B - d) / 255.0;
    int kr = (int)std::floor(bin_pos_ref);
    double u_c = pc[0] / pc[2];
    double v_c = pc[1] / pc[2];
    double x = pc[0], y = pc[1];
    double invz = 1.0 / pc[2];
    double invz_2 = invz * invz;
    double u = P.fx * u_c + P.cx;
    double v = P.fy * v_c + P.cy;
    double bin_pos_cur = e.icur[i] * (B - 3.0) / 255.0;
    int kt = (int)std::floor(bin_pos_cur);
    double gpx, gpy, Ju[6], Jv[6];
    if (u >= 0 && u + 3 <= jac_cols && v >= 0 && v + 3 <= P.rows) {
      gpx = (interp(P, u + 1, v) - interp(P, u - 1, v)) / 2;
      gpy = (interp(P, u, v + 1) - interp(P, u, v - 1)) / 2;
      Ju[0] = -x * y * invz_2 * P.fx;
      Ju[1] = (1 + (x * x * invz_2)) * P.fx;
      Ju[2] = -y * invz * P.fx;
      Ju[3] = invz * P.fx;
      Ju[4] = 0;
      Ju[5] = -x * invz_2 * P.fx;
      Jv[0] = -(1 + y * y * invz_2) * P.fy;
      Jv[1] = x * y * invz_2 * P.fy;
      Jv[2] = x * invz * P.fy;
      Jv[3] = 0;
      Jv[4] = invz * P.fy;
      Jv[5] = -y * invz_2 * P.fy;
    } else {
      continue;
    }
    double dip[6];
    for (int a = 0; a < 6; a++) dip[a] = gpx * Ju[a] + gpy * Jv[a];
    double dbs[4];
    for (int m = 0; m < 4; m++) dbs[m] = bspline_der(kt + m, d + 1, bin_pos_cur, B);
    for (int m = 0; m < d + 1; m++)
      for (int a = 0; a < 6; a++) dpt[(size_t)(kt + m) * 6 + a] += dbs[m] * d_mi_i * dip[a];
    for (int k = 0; k < d + 1; k++)
      for (int m = 0; m < d + 1; m++)
        for (int a = 0; a < 6; a++)
          dpj[((size_t)(kr + k) * B + kt + m) * 6 + a] += e.bs_ref[4 * i + k] * dbs[m] * d_mi_i * dip[a];
  }
  int nc = (int)n - e.ob;
  for (auto& v : dpt) v /= nc;
  for (auto& v : dpj) v /= nc;
  double d_hj[6], d_hl[6];
  for (int a = 0; a < 6; a++) {
    double tmp = 0.0;
    for (int m = 0; m < B; m++)
      for (int q = 0; q < B; q++) {
        double pj = e.pro_joint[(size_t)m * B + q];
        if (pj < kSigma) continue;
        tmp -= (1.0 + std::log2(pj)) * dpj[((size_t)m * B + q) * 6 + a];
      }
    d_hj[a] = tmp;
  }
  for (int a = 0; a < 6; a++) {
    d_hl[a] = 0.0;
    for (int j = 0; j < B; j++) {
      if (e.pro_cur[j] < kSigma) continue;
      d_hl[a] -= (1.0 + std::log2(e.pro_cur[j])) * dpt[(size_t)j * 6 + a];
    }
  }
  double inv_square_hj = 1.0 / (e.H_joint * e.H_joint);
  for (int a = 0; a < 6; a++)
    e.J[a] = (d_hj[a] * (e.H_cur + e.H_ref) - d_hl[a] * e.H_joint) * inv_square_hj;
}

// do_h: computeError() of every active edge (ClearPrevH + ComputeH); do_jac: linearizeOplus()
// of every active edge, which reads the cache left by the ComputeH of the same pose.
void eval_all(orc_problem& P, const SE3& T, bool do_h, bool do_jac) {
  int ne = (int)P.edges.size();
#pragma omp parallel for schedule(dynamic, 1) num_threads(P.threads) if (P.threads > 1)
  for (int c = 0; c < ne; c++) {
    Edge& e = P.edges[c];
    if (e.level != 0) continue;
    if (do_h) compute_h(P, T, e);
    if (do_jac) linearize(P, T, e);
  }
}

// sparse_optimizer.cpp:102-116 over active (level-0) edges, chi2 = e^2 (information = I_1)
double robust_chi2(const orc_problem& P, double delta) {
  double chi = 0.0;
  for (const Edge& e : P.edges) {
    if (e.level != 0) continue;
    double rho[3];
    huber(e.err * e.err, delta, rho);
    chi += rho[0];
  }
  return chi;
}

// block_solver.hpp:502-570 + base_unary_edge.hpp:43-72 (robustInformation: base_edge.h:96-102)
void build_system(const orc_problem& P, double delta, double H[36], double b[6]) {
  for (int i = 0; i < 36; i++) H[i] = 0;
  for (int i = 0; i < 6; i++) b[i] = 0;
  for (const Edge& e : P.edges) {
    if (e.level != 0) continue;
    double rho[3];
    huber(e.err * e.err, delta, rho);
    for (int i = 0; i < 6; i++) b[i] -= rho[1] * e.J[i] * 1.0 * e.err;
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) H[i * 6 + j] += e.J[i] * rho[1] * e.J[j];
  }
}

}  // namespace

// ================================================================ C API
extern "C" {

double orc_bspline(int index, int order, double u, int bins) { return bspline(index, order, u, bins); }
double orc_bspline_der(int index, int order, double u, int bins) { return bspline_der(index, order, u, bins); }

void orc_se3_from_Rt(const double Rrm[9], const double t[3], double pose7[7]) {
  double R[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i][j] = Rrm[3 * i + j];
  se3_to7(se3_from_Rt(R, t), pose7);
}
void orc_se3_exp(const double upd6[6], double pose7[7]) { se3_to7(se3_exp(upd6), pose7); }
void orc_se3_mul(const double a7[7], const double b7[7], double out7[7]) {
  se3_to7(se3_mul(se3_from7(a7), se3_from7(b7)), out7);
}
void orc_se3_inverse(const double a7[7], double out7[7]) { se3_to7(se3_inverse(se3_from7(a7)), out7); }
void orc_se3_to_mat16(const double pose7[7], double mat16[16]) { se3_to_mat16(se3_from7(pose7), mat16); }
void orc_se3_map(const double pose7[7], const double p[3], double out[3]) { se3_map(se3_from7(pose7), p, out); }

// NID_pose_estimation.cpp:186-208
void orc_reference_perturbation(const double Twc1[16], double pose7[7]) {
  const double t_offset = 0.02, r_offset = 0.005;
  double t_dist[3] = {0.5 * t_offset, -t_offset, -t_offset};
  double a = r_offset * M_PI;
  double c = std::cos(a), s = std::sin(a);
  double Rx[3][3] = {{1, 0, 0}, {0, c, -s}, {0, s, c}};
  double Ry[3][3] = {{c, 0, s}, {0, 1, 0}, {-s, 0, c}};
  double Rz[3][3] = {{c, -s, 0}, {s, c, 0}, {0, 0, 1}};
  double Rxy[3][3], Rd[3][3];
  mat3_mul(Rx, Ry, Rxy);
  mat3_mul(Rxy, Rz, Rd);
  // r_cw1 = R_wc1^T ; t_cw1 = -r_cw1 * t_wc1
  double rcw[3][3], tcw[3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rcw[i][j] = Twc1[4 * i + j];  // transpose of col-major block
  for (int i = 0; i < 3; i++) {
    double acc = 0;
    for (int j = 0; j < 3; j++) acc += (-rcw[i][j]) * Twc1[12 + j];
    tcw[i] = acc;
  }
  double r2[3][3];
  mat3_mul(Rd, rcw, r2);
  for (int i = 0; i < 3; i++) tcw[i] = tcw[i] + t_dist[i];
  se3_to7(se3_from_Rt(r2, tcw), pose7);
}

void orc_huber(double chi2, double delta, double rho[3]) { huber(chi2, delta, rho); }
int orc_ldlt6_solve(const double H[36], const double b[6], double x[6]) { return ldlt6_solve(H, b, x); }

orc_problem* orc_create(const uint8_t* im0, const double* depth, const uint8_t* im1, int rows, int cols,
                        const double T_wc0[16], const double intr[4], int cell, int bins, int threads) {
  orc_problem* P = new orc_problem();
  P->rows = rows; P->cols = cols; P->cell = cell; P->bins = bins; P->degree = 3;
  P->threads = threads < 1 ? 1 : threads;
  P->fx = intr[0]; P->fy = intr[1]; P->cx = intr[2]; P->cy = intr[3];
  std::memcpy(P->T_wc0, T_wc0, sizeof(P->T_wc0));
  size_t N = (size_t)rows * cols;
  P->im0.assign(im0, im0 + N);
  P->im1.assign(im1, im1 + N);
  P->depth.assign(depth, depth + N);
  P->edges.resize((size_t)cell * cell);
  for (int i = 0; i < cell; i++)
    for (int j = 0; j < cell; j++) {
      Edge& e = P->edges[(size_t)i * cell + j];
      build_edge_points(*P, i, j, e);
      set_bspline_relates(*P, e);
    }
  return P;
}
void orc_destroy(orc_problem* P) { delete P; }
void orc_set_quirks(orc_problem* P, int jac_bound_gpu, int warp_with_matrix) {
  P->jac_bound_gpu = jac_bound_gpu;
  P->warp_with_matrix = warp_with_matrix;
}

void orc_points3d(const orc_problem* P, double* out) {
  size_t N = (size_t)P->rows * P->cols;
  for (size_t i = 0; i < 3 * N; i++) out[i] = kNaN;
  // CudaPoints3d.cu:5-32 (same formulas as the per-cell CPU extraction, all pixels)
  const double* T = P->T_wc0;
  for (int r = 0; r < P->rows; r++)
    for (int c = 0; c < P->cols; c++) {
      size_t id = (size_t)r * P->cols + c;
      double z = P->depth[id];
      if (z < 0.01 || z > 100) continue;
      double x0 = z * (c - P->cx) / P->fx;
      double y0 = z * (r - P->cy) / P->fy;
      out[3 * id] = T[0] * x0 + T[4] * y0 + T[8] * z + T[12];
      out[3 * id + 1] = T[1] * x0 + T[5] * y0 + T[9] * z + T[13];
      out[3 * id + 2] = T[2] * x0 + T[6] * y0 + T[10] * z + T[14];
    }
}
int orc_cell_points(const orc_problem* P, int c) { return (int)P->edges[c].meas.size(); }

void orc_prepare(orc_problem* P, const double pose7[7], int* n_c, double* Href) {
  SE3 T = se3_from7(pose7);
  int ne = (int)P->edges.size();
#pragma omp parallel for schedule(dynamic, 1) num_threads(P->threads) if (P->threads > 1)
  for (int c = 0; c < ne; c++) compute_href(*P, T, P->edges[c]);
  for (int c = 0; c < ne; c++) {
    const Edge& e = P->edges[c];
    if (n_c) n_c[c] = (int)e.meas.size() - e.ob;
    if (Href) Href[c] = e.level ? kNaN : e.H_ref;
  }
  P->prepared = true;
}

void orc_ref_weights(const orc_problem* P, double* bs_value, int* bs_index) {
  size_t N = (size_t)P->rows * P->cols;
  for (size_t i = 0; i < 4 * N; i++) bs_value[i] = 0.0;
  for (size_t i = 0; i < N; i++) bs_index[i] = 0;
  for (const Edge& e : P->edges)
    for (size_t i = 0; i < e.meas.size(); i++) {
      double obs = e.meas[i];
      if (obs >= 255) obs = 254.999;
      if (obs < 0) obs = 0.0;
      int k = (int)std::floor(obs * (P->bins - P->degree) / 255.0);
      bs_index[e.loc[i]] = k;
      for (int m = 0; m < 4; m++) bs_value[4 * (size_t)e.loc[i] + m] = e.bs_ref[4 * i + m];
    }
}

void orc_eval(orc_problem* P, const double pose7[7], int want_jac, double* Ht, double* Hj, double* err,
              double* J6) {
  SE3 T = se3_from7(pose7);
  eval_all(*P, T, true, want_jac != 0);
  for (size_t c = 0; c < P->edges.size(); c++) {
    const Edge& e = P->edges[c];
    bool act = (e.level == 0);
    if (Ht) Ht[c] = act ? e.H_cur : kNaN;
    if (Hj) Hj[c] = act ? e.H_joint : kNaN;
    if (err) err[c] = act ? e.err : kNaN;
    if (J6 && want_jac)
      for (int a = 0; a < 6; a++) J6[6 * c + a] = act ? e.J[a] : kNaN;
  }
}

void orc_last_hist(const orc_problem* P, int c, double* P_t, double* P_j) {
  const Edge& e = P->edges[c];
  if (P_t) std::memcpy(P_t, e.pro_cur.data(), sizeof(double) * P->bins);
  if (P_j) std::memcpy(P_j, e.pro_joint.data(), sizeof(double) * P->bins * P->bins);
}
void orc_last_dhist(const orc_problem* P, int c, double* dP_t, double* dP_j) {
  const Edge& e = P->edges[c];
  if (dP_t && !e.d_pt.empty()) std::memcpy(dP_t, e.d_pt.data(), sizeof(double) * e.d_pt.size());
  if (dP_j && !e.d_pj.empty()) std::memcpy(dP_j, e.d_pj.data(), sizeof(double) * e.d_pj.size());
}

void orc_pixels(const orc_problem* P, const double pose7[7], double* out) {
  size_t N = (size_t)P->rows * P->cols;
  for (size_t i = 0; i < 8 * N; i++) out[i] = kNaN;
  SE3 T = se3_from7(pose7);
  double m16[16];
  se3_to_mat16(T, m16);
  std::vector<double> pts(3 * N);
  orc_points3d(P, pts.data());
  const int jac_cols = P->jac_bound_gpu ? P->cols : P->cols - 1;
  for (size_t id = 0; id < N; id++) {
    if (std::isnan(pts[3 * id])) continue;
    double pc[3];
    warp_point(*P, T, m16, &pts[3 * id], pc);
    double u = P->fx * pc[0] / pc[2] + P->cx;
    double v = P->fy * pc[1] / pc[2] + P->cy;
    double* o = &out[8 * id];
    o[0] = u; o[1] = v; o[7] = pc[2];
    // the CPU edge projects a second time in linearizeOplus, as fx*(x/z)+cx (types_six_dof_expmap.cpp:407-421);
    // its bounds test (:433) and gradient samples (:434-435) use that value
    double u2 = P->fx * (pc[0] / pc[2]) + P->cx;
    double v2 = P->fy * (pc[1] / pc[2]) + P->cy;
    bool vc = (u >= 0 && u + 3 <= P->cols && v >= 0 && v + 3 <= P->rows);
    bool vj = vc && (u2 >= 0 && u2 + 3 <= jac_cols && v2 >= 0 && v2 + 3 <= P->rows);
    o[5] = vc ? 1.0 : 0.0;
    o[6] = vj ? 1.0 : 0.0;
    o[2] = 0; o[3] = 0; o[4] = 0;
    if (vc) {
      double ic = interp(*P, u, v);
      if (ic >= 255) ic = 254.999;
      if (ic < 0) ic = 0.0;
      o[2] = ic;
    }
    if (vj) {
      o[3] = (interp(*P, u2 + 1, v2) - interp(*P, u2 - 1, v2)) / 2;
      o[4] = (interp(*P, u2, v2 + 1) - interp(*P, u2, v2 - 1)) / 2;
    }
  }
}

void orc_gn_system(orc_problem* P, const double pose7[7], double delta, double* chi2, double* H36, double* b6) {
  SE3 T = se3_from7(pose7);
  eval_all(*P, T, true, true);
  double H[36], b[6];
  build_system(*P, delta, H, b);
  if (chi2) *chi2 = robust_chi2(*P, delta);
  if (H36) std::memcpy(H36, H, sizeof(H));
  if (b6) std::memcpy(b6, b, sizeof(b));
}

// optimization_algorithm_levenberg.cpp:61-250 driven by sparse_optimizer.cpp:356-450
int orc_optimize(orc_problem* P, double pose7[7], int max_iters, double delta, double* trace, int* counts) {
  SE3 est = se3_from7(pose7);
  double lambda = -1., ni = 2.;
  int nBad = 0;
  const double tau = 1e-5, goodUp = 2. / 3., goodLo = 1. / 3.;
  const int maxTrials = 10;
  int jac_evals = 0, cost_evals = 0;
  double x[6] = {0, 0, 0, 0, 0, 0};
  int it = 0;
  bool ok = true;
  for (it = 0; it < max_iters && ok; it++) {
    // ---- solve(it)
    eval_all(*P, est, true, false);  // computeActiveErrors (:115)
    double currentChi = robust_chi2(*P, delta);
    double tempChi = currentChi;
    double iniChi = currentChi;
    eval_all(*P, est, false, true);  // buildSystem -> linearizeOplus (reads the cache of the same pose)
    jac_evals++;
    double H[36], b[6];
    build_system(*P, delta, H, b);
    if (it == 0) {  // computeLambdaInit (:227-241)
      double maxDiagonal = 0.;
      for (int j = 0; j < 6; j++) maxDiagonal = std::max(std::fabs(H[j * 6 + j]), maxDiagonal);
      lambda = tau * maxDiagonal;
      ni = 2;
      nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      SE3 backup = est;  // push
      double Hl[36];
      std::memcpy(Hl, H, sizeof(H));
      for (int j = 0; j < 6; j++) Hl[j * 6 + j] += lambda;  // setLambda
      int ok2 = ldlt6_solve(Hl, b, x);
      est = se3_mul(se3_exp(x), est);  // oplusImpl, types_six_dof_expmap.h:74-77
      eval_all(*P, est, true, false);
      cost_evals++;
      tempChi = robust_chi2(*P, delta);
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = (currentChi - tempChi);
      double scale = 0.;
      for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);  // computeScale (:243-250)
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = (std::min)(alpha, goodUp);
        double scaleFactor = (std::max)(goodLo, alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        est = backup;  // pop
      }
      qmax++;
    } while (rho < 0 && qmax < maxTrials);

    bool terminate = false;
    if (qmax == maxTrials || rho == 0) terminate = true;
    if (!terminate) {
      if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
      else nBad = 0;
      if (nBad >= 3) terminate = true;
    }
    ok = !terminate;
    // verbose block of SparseOptimizer::optimize (sparse_optimizer.cpp:404-433): re-evaluate at the current estimate
    eval_all(*P, est, true, false);
    cost_evals++;
    double chi_now = robust_chi2(*P, delta);
    if (trace) {
      double* t = trace + 10 * it;
      t[0] = chi_now; t[1] = lambda; t[2] = qmax;
      se3_to7(est, t + 3);
    }
  }
  se3_to7(est, pose7);
  if (counts) { counts[0] = jac_evals; counts[1] = cost_evals; }
  return it;
}

// NID_standard_property.cpp:200-230, 342-485
double orc_hard_nid(const uint8_t* im0, const double* depth, const uint8_t* im1, int rows, int cols,
                    const double T_wc0[16], const double M[16], const double intr[4], int cell, int bins,
                    double* nid_cells, int threads) {
  orc_problem P;
  P.rows = rows; P.cols = cols; P.cell = cell; P.bins = bins; P.degree = 3; P.threads = 1;
  P.fx = intr[0]; P.fy = intr[1]; P.cx = intr[2]; P.cy = intr[3];
  std::memcpy(P.T_wc0, T_wc0, sizeof(P.T_wc0));
  size_t N = (size_t)rows * cols;
  P.im0.assign(im0, im0 + N);
  P.im1.assign(im1, im1 + N);
  P.depth.assign(depth, depth + N);
  int ne = cell * cell;
  std::vector<double> nid(ne, 0.0);
  if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
  for (int c = 0; c < ne; c++) {
    Edge e;
    build_edge_points(P, c / cell, c % cell, e);
    size_t n = e.meas.size();
    std::vector<double> pr(bins, 0.0), pc_(bins, 0.0), pj((size_t)bins * bins, 0.0);
    double Hr = 0, Hc = 0, Hj = 0;
    // ComputeHref (:342-392)
    int ob = 0;
    std::vector<double> i0(e.meas);
    for (size_t i = 0; i < n; i++) {
      const double* pw = &e.xw[3 * i];
      double px = M[0] * pw[0] + M[4] * pw[1] + M[8] * pw[2] + M[12] * 1.0;
      double py = M[1] * pw[0] + M[5] * pw[1] + M[9] * pw[2] + M[13] * 1.0;
      double pz = M[2] * pw[0] + M[6] * pw[1] + M[10] * pw[2] + M[14] * 1.0;
      double u = P.fx * px / pz + P.cx;
      double v = P.fy * py / pz + P.cy;
      if (!(u >= 0 && u + 3 <= cols && v >= 0 && v + 3 <= rows)) { ob++; continue; }
      if (i0[i] >= 255) i0[i] = 254.999;
      if (i0[i] < 0) i0[i] = 0.0;
      int kr = (int)std::floor(i0[i] * bins / 255.0);
      pr[kr] += 1.0;
    }
    for (int i = 0; i < bins; i++) pr[i] /= ((int)n - ob);
    for (int i = 0; i < bins; i++) { if (pr[i] < kSigma) continue; Hr -= pr[i] * std::log2(pr[i]); }
    // ComputeH (:395-485)
    ob = 0;
    for (size_t i = 0; i < n; i++) {
      const double* pw = &e.xw[3 * i];
      double px = M[0] * pw[0] + M[4] * pw[1] + M[8] * pw[2] + M[12] * 1.0;
      double py = M[1] * pw[0] + M[5] * pw[1] + M[9] * pw[2] + M[13] * 1.0;
      double pz = M[2] * pw[0] + M[6] * pw[1] + M[10] * pw[2] + M[14] * 1.0;
      if (i0[i] >= 255) i0[i] = 254.999;
      if (i0[i] < 0) i0[i] = 0.0;
      int kr = (int)std::floor(i0[i] * bins / 255.0);
      double u = P.fx * px / pz + P.cx;
      double v = P.fy * py / pz + P.cy;
      double ic;
      if (u >= 0 && u + 3 <= cols && v >= 0 && v + 3 <= rows) ic = interp(P, u, v);
      else { ob++; continue; }
      if (ic >= 255) ic = 254.999;
      if (ic < 0) ic = 0.0;
      int kt = (int)std::floor(ic * bins / 255.0);
      pc_[kt] += 1.0;
      pj[(size_t)kr * bins + kt] += 1.0;
    }
    if ((int)n - ob < 300) { nid[c] = 0.0; continue; }  // SURVEY B-10: sparse cell contributes 0
    for (int i = 0; i < bins; i++) pc_[i] /= ((int)n - ob);
    for (auto& v : pj) v /= ((int)n - ob);
    for (int i = 0; i < bins; i++) { if (pc_[i] < kSigma) continue; Hc -= pc_[i] * std::log2(pc_[i]); }
    for (auto& v : pj) { if (v < kSigma) continue; Hj -= v * std::log2(v); }
    double val = (2 * Hj - Hr - Hc) / Hj;
    if (Hr == 0.0 && Hc == 0.0 && Hj == 0.0) val = 0.0;
    nid[c] = val;
  }
  double total = 0.0;
  for (int c = 0; c < ne; c++) {
    total += nid[c] * nid[c];
    if (nid_cells) nid_cells[c] = nid[c];
  }
  return std::sqrt(total);
}

}  // extern "C"
