// Input / output side of the two config-driven binaries (SURVEY 8f, f2 + f3): the OpenCV-FileStorage YAML
// subset of config_eth_cvg.yaml, PNG (zlib) / PGM decoding, the reference's grayscale conversion and the TUM
// ground-truth reader. The reference takes all of this from OpenCV (NID_pose_estimation.cpp:69-113,434-530),
// which is not available to this build; only what the ETH-CVG layout needs is implemented. Header-only, host-only.
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace nidio {

// ---------------------------------------------------------------------------------------------- YAML
// `%YAML:1.0` header, `key: value` scalars, single/double-quoted strings, `#` comments (config_eth_cvg.yaml:1-21).
struct Config {
  std::map<std::string, std::string> kv;
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  std::string str(const std::string& k, const std::string& def = "") const {
    auto it = kv.find(k);
    return it == kv.end() ? def : it->second;
  }
  double num(const std::string& k, double def = 0.0) const { return has(k) ? atof(kv.at(k).c_str()) : def; }
  int integer(const std::string& k, int def = 0) const { return has(k) ? (int)atof(kv.at(k).c_str()) : def; }  // (int)fc["key"]
};

inline std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

inline Config read_config(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open config " + path);
  Config c;
  std::string line;
  while (std::getline(f, line)) {
    std::string t = trim(line);
    if (t.empty() || t[0] == '#' || t[0] == '%' || t == "---") continue;
    size_t colon = t.find(':');
    if (colon == std::string::npos) continue;
    std::string key = trim(t.substr(0, colon)), val = trim(t.substr(colon + 1));
    if (!val.empty() && (val[0] == '\'' || val[0] == '"')) {
      size_t e = val.find(val[0], 1);
      val = val.substr(1, e == std::string::npos ? std::string::npos : e - 1);
    } else {
      size_t h = val.find(" #");
      if (h != std::string::npos) val = trim(val.substr(0, h));
    }
    c.kv[key] = val;
  }
  return c;
}

// ---------------------------------------------------------------------------------------------- images
struct Image {
  int rows = 0, cols = 0, channels = 0, depth = 8;  // depth: bits per sample (8 or 16)
  std::vector<uint16_t> px;                         // interleaved samples, file channel order (PNG: R,G,B[,A])
};

inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// Non-interlaced PNG, colour types 0/2/4/6, 8 or 16 bits per sample.
inline Image load_png(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (buf.size() < 33 || memcmp(buf.data(), sig, 8)) throw std::runtime_error(path + ": not a PNG");
  Image im;
  int ctype = -1, interlace = 0;
  std::vector<unsigned char> z;
  size_t pos = 8;
  while (pos + 12 <= buf.size()) {
    uint32_t len = be32(&buf[pos]);
    std::string type((const char*)&buf[pos + 4], 4);
    const unsigned char* d = &buf[pos + 8];
    if (pos + 12 + len > buf.size()) throw std::runtime_error(path + ": truncated chunk");
    if (type == "IHDR") {
      im.cols = (int)be32(d); im.rows = (int)be32(d + 4); im.depth = d[8]; ctype = d[9]; interlace = d[12];
    } else if (type == "IDAT") {
      z.insert(z.end(), d, d + len);
    } else if (type == "IEND") {
      break;
    }
    pos += 12 + len;
  }
  if (interlace) throw std::runtime_error(path + ": interlaced PNG not supported");
  if (im.depth != 8 && im.depth != 16) throw std::runtime_error(path + ": only 8/16-bit PNG supported");
  switch (ctype) {
    case 0: im.channels = 1; break;
    case 2: im.channels = 3; break;
    case 4: im.channels = 2; break;
    case 6: im.channels = 4; break;
    default: throw std::runtime_error(path + ": palette PNG not supported");
  }
  const size_t bpp = (size_t)im.channels * im.depth / 8, stride = bpp * im.cols;
  std::vector<unsigned char> raw((stride + 1) * im.rows);
  uLongf out_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &out_len, z.data(), (uLong)z.size()) != Z_OK || out_len != raw.size())
    throw std::runtime_error(path + ": zlib inflate failed");
  std::vector<unsigned char> prev(stride, 0), cur(stride);
  im.px.resize((size_t)im.rows * im.cols * im.channels);
  for (int y = 0; y < im.rows; y++) {
    const unsigned char* line = &raw[(stride + 1) * y];
    const int ft = line[0];
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
      int pred = 0;
      switch (ft) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: {
          const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
          pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
        } break;
        default: throw std::runtime_error(path + ": bad PNG filter");
      }
      cur[i] = (unsigned char)(line[1 + i] + pred);
    }
    uint16_t* o = &im.px[(size_t)y * im.cols * im.channels];
    const size_t ns = (size_t)im.cols * im.channels;
    if (im.depth == 8) for (size_t i = 0; i < ns; i++) o[i] = cur[i];
    else for (size_t i = 0; i < ns; i++) o[i] = (uint16_t)((cur[2 * i] << 8) | cur[2 * i + 1]);
    prev.swap(cur);
  }
  return im;
}

// binary PGM (P5), maxval < 65536: the decoder-free fallback layout (rgb/NNNN.pgm, depth/NNNN.pgm)
inline Image load_pgm(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::string magic;
  f >> magic;
  if (magic != "P5") throw std::runtime_error(path + ": not a binary PGM");
  auto next_int = [&f]() {
    int v;
    while (true) {
      f >> std::ws;
      if (f.peek() == '#') { std::string c; std::getline(f, c); continue; }
      f >> v;
      return v;
    }
  };
  Image im;
  im.cols = next_int(); im.rows = next_int();
  const int maxv = next_int();
  f.get();
  im.channels = 1; im.depth = maxv > 255 ? 16 : 8;
  const size_t n = (size_t)im.rows * im.cols;
  im.px.resize(n);
  std::vector<unsigned char> raw(n * (im.depth / 8));
  f.read((char*)raw.data(), (std::streamsize)raw.size());
  if ((size_t)f.gcount() != raw.size()) throw std::runtime_error(path + ": truncated PGM");
  for (size_t i = 0; i < n; i++) im.px[i] = im.depth == 8 ? raw[i] : (uint16_t)((raw[2 * i] << 8) | raw[2 * i + 1]);
  return im;
}

// <base>.png if present, else <base>.pgm
inline Image load_image(const std::string& base_without_ext) {
  if (std::ifstream(base_without_ext + ".png")) return load_png(base_without_ext + ".png");
  return load_pgm(base_without_ext + ".pgm");
}

// The reference reads colour frames with imread(UNCHANGED) (BGR in memory) and converts them with CV_RGB2GRAY
// (NID_pose_estimation.cpp:91-97), i.e. with the red and blue weights swapped. OpenCV's 8-bit kernel is fixed
// point with c0 the FIRST channel in memory, which is blue here:
//   OpenCV 2.4 / 3.x (the reference's stated dependency, README.md:15): (c0*4899 + c1*9617 + c2*1868 + 2^13) >> 14
//   OpenCV >= 3.4.3 / 4.x:                                               (c0*9798 + c1*19235 + c2*3735 + 2^14) >> 15
// `shift` selects the flavour (YAML key gray_shift, default 14).
inline std::vector<uint8_t> gray_like_reference(const Image& im, int shift = 14) {
  if (im.depth != 8) throw std::runtime_error("colour frames must be 8-bit");
  std::vector<uint8_t> g((size_t)im.rows * im.cols);
  for (size_t i = 0; i < g.size(); i++) {
    if (im.channels >= 3) {
      const unsigned R = im.px[i * im.channels], G = im.px[i * im.channels + 1], B = im.px[i * im.channels + 2];
      g[i] = shift == 15 ? (uint8_t)((B * 9798u + G * 19235u + R * 3735u + 16384u) >> 15)
                         : (uint8_t)((B * 4899u + G * 9617u + R * 1868u + 8192u) >> 14);
    } else {
      g[i] = (uint8_t)im.px[i * im.channels];  // already gray (the reference's cvtColor would refuse this input)
    }
  }
  return g;
}

// depth PNG (uint16) -> metres: convertTo(CV_64F, depth_factor), depth_factor = 1.0/(int)fc["depth_factor"]
inline std::vector<double> depth_metres(const Image& im, double depth_factor) {
  std::vector<double> d((size_t)im.rows * im.cols);
  for (size_t i = 0; i < d.size(); i++) d[i] = (double)im.px[i * im.channels] * depth_factor;
  return d;
}

// ---------------------------------------------------------------------------------------------- ground truth
// TUM lines `ts tx ty tz qx qy qz qw`, one pose per line, index = line number (NID_pose_estimation.cpp:434-530).
// Returns column-major 4x4 T_wc per line (quaternion -> rotation as Eigen's toRotationMatrix, no normalisation).
inline std::vector<std::vector<double>> read_groundtruth(const std::string& path) {
  std::vector<std::vector<double>> all;
  std::ifstream f(path);
  if (!f) { printf("cannot find the file that contains groundtruth \n"); return all; }
  std::string row;
  while (std::getline(f, row)) {
    std::istringstream ss(row);
    std::string tok;
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n = 0;
    while (n < 8 && std::getline(ss, tok, ' ')) {
      v[n] = n == 0 ? (double)atoi(tok.c_str()) : atof(tok.c_str());
      n++;
    }
    const double qx = v[4], qy = v[5], qz = v[6], qw = v[7];
    const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
    const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    const double R[3][3] = {{1 - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1 - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1 - (txx + tyy)}};
    std::vector<double> T(16, 0.0);
    for (int c = 0; c < 3; c++)
      for (int r = 0; r < 3; r++) T[4 * c + r] = R[r][c];
    T[12] = v[1]; T[13] = v[2]; T[14] = v[3]; T[15] = 1.0;
    all.push_back(T);
  }
  return all;
}

// rigid inverse of a column-major 4x4
inline void invert_rigid(const double* T, double* out) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) out[4 * c + r] = T[4 * r + c];
  for (int r = 0; r < 3; r++) out[12 + r] = -(out[r] * T[12] + out[4 + r] * T[13] + out[8 + r] * T[14]);
  out[3] = out[7] = out[11] = 0.0; out[15] = 1.0;
}

inline void print_mat4(const double* T) {  // Eigen's default operator<< look: rows, space separated
  for (int r = 0; r < 4; r++) printf("%12.6g %12.6g %12.6g %12.6g\n", T[r], T[4 + r], T[8 + r], T[12 + r]);
}

// ids of the ETH-CVG frames are zero-padded to four digits (NID_pose_estimation.cpp:372-387)
inline std::string frame_id(int id) {
  char b[16];
  snprintf(b, sizeof(b), "%04d", id);
  return b;
}

struct Pair {
  int rows = 0, cols = 0;
  std::vector<uint8_t> im0, im1;
  std::vector<double> depth0;
  std::vector<double> T_wc0, T_wc1;  // column-major 4x4
  double intr[5];                    // fx fy cx cy depth_factor
};

// Everything both binaries read before they start computing (NID_pose_estimation.cpp:69-160).
inline Pair load_pair(const Config& c) {
  const std::string type0 = c.str("image0_type"), type1 = c.str("image1_type"), id0 = c.str("image0_id"), id1 = c.str("image1_id");
  const std::string dataset = c.str("dataset"), im_add = c.str("im_address"), use_gt = c.str("use_groundtruth");
  if (dataset != "eth_cvg") throw std::runtime_error("dataset must be eth_cvg (the only layout the reference reads)");
  Pair p;
  p.intr[4] = 1.0 / c.integer("depth_factor");
  p.intr[0] = c.num("fx"); p.intr[1] = c.num("fy"); p.intr[2] = c.num("cx"); p.intr[3] = c.num("cy");
  const Image rgb0 = load_image(im_add + type0 + "/" + id0), rgb1 = load_image(im_add + type1 + "/" + id1);
  std::printf("dataset address %s\n", (im_add + type0 + "/" + id0 + ".png").c_str());
  p.rows = rgb0.rows; p.cols = rgb0.cols;
  if (rgb1.rows != p.rows || rgb1.cols != p.cols) throw std::runtime_error("frames differ in size");
  const int gray_shift = c.integer("gray_shift", 14);
  if (gray_shift != 14 && gray_shift != 15) throw std::runtime_error("gray_shift must be 14 (OpenCV 3) or 15 (OpenCV 4)");
  p.im0 = gray_like_reference(rgb0, gray_shift);
  p.im1 = gray_like_reference(rgb1, gray_shift);
  std::printf("image size [%d x %d],[%d x %d]\n", rgb0.cols, rgb0.rows, rgb1.cols, rgb1.rows);
  const Image d0 = load_image(im_add + "depth/" + id0);
  if (d0.rows != p.rows || d0.cols != p.cols) throw std::runtime_error("depth and colour frames differ in size");
  p.depth0 = depth_metres(d0, p.intr[4]);
  const auto gt = read_groundtruth(im_add + "groundtruth.txt");
  const int pose_id0 = atoi(id0.c_str()), pose_id1 = atoi(id1.c_str());
  if (use_gt == "1") {
    std::printf("use groundtruth pose \n");
    if ((int)gt.size() <= std::max(pose_id0, pose_id1)) throw std::runtime_error("groundtruth.txt has too few lines");
    p.T_wc0 = gt[pose_id0]; p.T_wc1 = gt[pose_id1];
  } else {
    throw std::runtime_error("use_groundtruth must be '1' (SLAM pose XML files are OpenCV FileStorage, not supported)");
  }
  return p;
}

}  // namespace nidio
