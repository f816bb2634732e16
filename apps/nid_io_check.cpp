// Test helper (tests/test_apps.py): decodes one image / config / ground-truth file with nid_io.hpp and prints what it saw.
#include <iostream>

#include "nid_io.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string mode = argv[1], path = argv[2];
  try {
    if (mode == "image") {
      const nidio::Image im = path.size() > 4 && path.substr(path.size() - 4) == ".pgm" ? nidio::load_pgm(path) : nidio::load_png(path);
      unsigned long long sum = 0, wsum = 0;
      for (size_t i = 0; i < im.px.size(); i++) { sum += im.px[i]; wsum += (unsigned long long)im.px[i] * (i % 1009 + 1); }
      std::cout << im.rows << " " << im.cols << " " << im.channels << " " << im.depth << " " << sum << " " << wsum;
      if (im.depth == 8) {
        const auto g = nidio::gray_like_reference(im, argc > 3 ? atoi(argv[3]) : 14);
        unsigned long long gs = 0, gw = 0;
        for (size_t i = 0; i < g.size(); i++) { gs += g[i]; gw += (unsigned long long)g[i] * (i % 1009 + 1); }
        std::cout << " " << gs << " " << gw;
      }
      std::cout << std::endl;
    } else if (mode == "config") {
      const nidio::Config c = nidio::read_config(path);
      for (const auto& kv : c.kv) std::cout << kv.first << "=" << kv.second << "\n";
      std::cout << "depth_factor_inv=" << 1.0 / c.integer("depth_factor") << std::endl;
    } else if (mode == "gt") {
      const auto gt = nidio::read_groundtruth(path);
      std::cout.precision(17);
      for (const auto& T : gt) {
        for (int i = 0; i < 16; i++) std::cout << T[i] << (i < 15 ? " " : "\n");
      }
    }
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
  return 0;
}
