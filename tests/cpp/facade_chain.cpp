// facade_chain.cpp -- drives the C++ facade (include/coopermap.hpp) the way the reference's nodelets drive their stage objects:
// OrganisedScanRegistration -> LaserOdometry -> LaserMapping over a sequence of organised sweeps read from a file, poses
// written to a file.  tests/test_facade_gpu.py builds it with g++ against libcoopermap.so and compares the poses with the
// Python (ctypes) path, which is parity-tested against the oracle.
//   usage: facade_chain <in.bin> <out.bin>    in: int32 nframes, rows, cols; then nframes * rows * cols * 4 float32
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "coopermap.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int hdr[3];
  if (fread(hdr, 4, 3, f) != 3) return 4;
  const int nframes = hdr[0], rows = hdr[1], cols = hdr[2];
  cm_config cfg = coopermap::Context::defaults();
  cfg.filter_corner = 0.4f; cfg.filter_surf = 0.8f; cfg.map_filter_corner = 0.4f; cfg.map_filter_surf = 0.4f;
  try {
    coopermap::OrganisedScanRegistration scanReg(cfg);
    coopermap::LaserOdometry odometry(cfg);
    coopermap::LaserMapping mapping(cfg, 100000, 800000);
    coopermap::LoamPipeline pipeline(rows, cols, cfg, 100000, 800000);   // the same chain in one object, clouds staying on the device
    coopermap::PointCloud sweep((size_t)rows * cols);
    std::vector<float> out;
    int chain_mismatch = 0;
    for (int k = 0; k < nframes; k++) {
      if (fread(sweep.data(), sizeof(cm_point), sweep.size(), f) != sweep.size()) return 5;
      scanReg.process(sweep, rows, cols);
      odometry.process(scanReg.cornerPointsSharp(), scanReg.cornerPointsLessSharp(), scanReg.surfacePointsFlat(), scanReg.surfacePointsLessFlat());
      const cm_iso mapped = mapping.process(odometry.transformSum(), odometry.lastCornerCloud(), odometry.lastSurfaceCloud());
      const cm_iso mapped1 = pipeline.process(sweep);
      for (int i = 0; i < 9; i++) chain_mismatch += mapped1.R[i] != mapped.R[i];
      for (int i = 0; i < 3; i++) chain_mismatch += mapped1.t[i] != mapped.t[i];
      for (int i = 0; i < 3; i++) chain_mismatch += pipeline.odometry().t[i] != odometry.transformSum().t[i];
      for (int i = 0; i < 9; i++) out.push_back(mapped.R[i]);
      for (int i = 0; i < 3; i++) out.push_back(mapped.t[i]);
      out.push_back((float)scanReg.laserCloud().size());
      out.push_back((float)mapping.lastStatus());
    }
    fclose(f);
    FILE* g = fopen(argv[2], "wb");
    if (!g) return 6;
    fwrite(out.data(), 4, out.size(), g);
    fclose(g);
    pipeline.sync();
    printf("facade_chain: %d frames, map %zu corner + %zu surf points\n", nframes, mapping.mapCloud(0).size(), mapping.mapCloud(1).size());
    if (chain_mismatch || pipeline.mapCloud(1).size() != mapping.mapCloud(1).size()) {
      fprintf(stderr, "LoamPipeline differs from the three stage objects (%d pose entries)\n", chain_mismatch);
      return 7;
    }
  } catch (const coopermap::Error& e) {
    fprintf(stderr, "coopermap error %d: %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}
