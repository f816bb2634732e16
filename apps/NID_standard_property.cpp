// NID_standard_property <config.yaml> -- the reference's second binary (NID_standard_property.cpp:69-198): hard-binned
// per-cell NID (cell = 16, bin_num = 8, no B-spline) of frame 1 against frame 0 at the ground-truth pose, printed as
// sqrt(sum nid_c^2). Same YAML keys and ETH-CVG layout. Optional extra keys: cell, bin_num, and
//   sweep: N        evaluate the cost surface of BASELINE config 5 as well: for every axis a of the 6-DoF
//                   perturbation, offsets (d_a, d_{(a+1) mod 6}) on an N x N lattice over +-sweep_range
//                   (0.05 m / rad) around the ground-truth pose, left-multiplied: T <- exp(xi) * T_cw1
//   sweep_csv: path (default nid_surface.csv) rows: axis,i,j,offset_a,offset_b,nid
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../include/nid_b200.h"
#include "../nid-pose-estimation_b200/host/nid_host_math.hpp"
#include "nid_io.hpp"

static void die(const char* what) {
  std::fprintf(stderr, "%s: %s\n", what, nid_last_error());
  std::exit(1);
}

int main(int argc, char** argv) {
  if (argc != 2) {
    std::cout << "usage './program path_to_config.yaml', image1's timestamp should be smaller than image2" << std::endl;
    return 0;
  }
  try {
    const nidio::Config cfg = nidio::read_config(argv[1]);
    const int cell = cfg.integer("cell", 16), bins = cfg.integer("bin_num", 8);
    const int sweep = cfg.integer("sweep", 0);
    const double range = cfg.num("sweep_range", 0.05);
    nidio::Pair P = nidio::load_pair(cfg);
    double T_cw1[16];
    nidio::invert_rigid(P.T_wc1.data(), T_cw1);  // the reference warps by tf_.inverse() (NID_standard_property.cpp:395-420)

    const int chunk = sweep > 0 ? 512 : 1;
    nid_ctx* ctx = nullptr;
    if (nid_create(&ctx, 0, P.rows, P.cols, cell, bins, 3, 1, chunk) != NID_OK) die("nid_create");
    if (nid_set_pair(ctx, 0, P.depth0.data(), P.im0.data(), P.im1.data(), P.T_wc0.data(), P.intr) != NID_OK) die("nid_set_pair");
    double total = 0.0;
    if (nid_hard_eval_jobs(ctx, 1, nullptr, T_cw1, &total, nullptr) != NID_OK) die("nid_hard_eval_jobs");
    std::cout << "final nid is " << total << std::endl;

    if (sweep > 0) {
      const nidhost::Pose7 gt = nidhost::pose_from_mat16(T_cw1);
      const size_t n = (size_t)6 * sweep * sweep;
      std::vector<double> poses(16 * n), out(n);
      size_t q = 0;
      for (int a = 0; a < 6; a++)
        for (int i = 0; i < sweep; i++)
          for (int j = 0; j < sweep; j++, q++) {
            double xi[6] = {0, 0, 0, 0, 0, 0};
            xi[a] = sweep > 1 ? -range + 2 * range * i / (sweep - 1) : 0.0;
            xi[(a + 1) % 6] = sweep > 1 ? -range + 2 * range * j / (sweep - 1) : 0.0;
            nidhost::pose_to_mat16(nidhost::pose_mul(nidhost::pose_exp(xi), gt), &poses[16 * q]);
          }
      for (size_t s0 = 0; s0 < n; s0 += chunk) {
        const int m = (int)std::min((size_t)chunk, n - s0);
        if (nid_hard_eval_jobs(ctx, m, nullptr, &poses[16 * s0], &out[s0], nullptr) != NID_OK) die("nid_hard_eval_jobs(sweep)");
      }
      const std::string csv = cfg.str("sweep_csv", "nid_surface.csv");
      std::ofstream of(csv);
      of << "axis,i,j,offset_a,offset_b,nid\n";
      q = 0;
      for (int a = 0; a < 6; a++)
        for (int i = 0; i < sweep; i++)
          for (int j = 0; j < sweep; j++, q++)
            of << a << "," << i << "," << j << "," << (sweep > 1 ? -range + 2 * range * i / (sweep - 1) : 0.0) << ","
               << (sweep > 1 ? -range + 2 * range * j / (sweep - 1) : 0.0) << "," << out[q] << "\n";
      std::cout << "cost surface: " << n << " poses written to " << csv << std::endl;
    }
    nid_destroy(ctx);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
