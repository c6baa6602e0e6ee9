// cm_mapio.cu -- the on-disk form of the cube map: index.txt + one binary PCD file per non-empty (cube, class).
//
// Replaces FeatureMap::saveCloudToFiles / loadCloudFromFiles / fileNameFormat (L_SLAM/src/util/FeatureMap.h:135-143,
// 378-462).  Layout restated: <dir>/index.txt holds one line "count type i j k size" per file, files are <dir>/<count>.pcd,
// written while looping i (width), j (height), k (depth) with the corner cloud (type 0) of a cube before its surf cloud
// (type 1); a cube's cloud is in pcl::VoxelGrid output order (ascending voxel index = z, then y, then x).  PCD files are
// what pcl::io::savePCDFileBinary writes for pcl::PointXYZI: the v0.7 header with FIELDS x y z intensity, 16 packed bytes
// per point.  Loading follows the reference: every file is pushed through the map voxel filter
// (_downSizeFilterCorner / _downSizeFilterSurf) -- here by the map's own insert kernels, all files of a class in one call,
// which is the same thing because cubes are disjoint and a voxel group sums its points in file order.
// Host code only (file I/O is not on the hot path); the points travel through DeviceMap::export_points / insert.
#include "cm_ctx.h"
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

using namespace cm;
static int fail(cm_ctx* ctx, int code, const std::string& msg) { return ctx_fail(ctx, code, msg); }

static std::string file_name(const std::string& dir, int number) {   // fileNameFormat, FeatureMap.h:135-143
  std::ostringstream ss;
  ss << dir << '/' << number << ".pcd";
  return ss.str();
}

static bool write_pcd_binary(const std::string& path, const cm_point* pts, size_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  fprintf(f,
          "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
          "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n",
          n, n);
  const bool ok = fwrite(pts, sizeof(cm_point), n, f) == n;
  fclose(f);
  return ok;
}

// Reads x, y, z, intensity of an ascii or binary PCD file with float32 fields (any field order, extra fields skipped).
static bool read_pcd(const std::string& path, std::vector<cm_point>& out, std::string& why) {
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f) { why = "cannot open"; return false; }
  std::vector<std::string> fields; std::vector<int> size, count; std::vector<char> type;
  size_t npoints = 0, width = 0, height = 1;
  std::string line, data;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string key; ls >> key;
    if (key == "FIELDS" || key == "COLUMNS") { std::string t; while (ls >> t) fields.push_back(t); }
    else if (key == "SIZE") { int t; while (ls >> t) size.push_back(t); }
    else if (key == "TYPE") { char t; while (ls >> t) type.push_back(t); }
    else if (key == "COUNT") { int t; while (ls >> t) count.push_back(t); }
    else if (key == "WIDTH") ls >> width;
    else if (key == "HEIGHT") ls >> height;
    else if (key == "POINTS") ls >> npoints;
    else if (key == "DATA") { ls >> data; break; }
  }
  if (fields.empty() || size.size() != fields.size() || type.size() != fields.size()) { why = "malformed header"; return false; }
  if (count.empty()) count.assign(fields.size(), 1);
  if (npoints == 0) npoints = width * height;
  int off[4] = {-1, -1, -1, -1}, col[4] = {-1, -1, -1, -1};
  int stride = 0, ncol = 0;
  const char* want[4] = {"x", "y", "z", "intensity"};
  for (size_t i = 0; i < fields.size(); i++) {
    for (int w = 0; w < 4; w++)
      if (fields[i] == want[w]) {
        if (size[i] != 4 || type[i] != 'F') { why = "field " + fields[i] + " is not float32"; return false; }
        off[w] = stride; col[w] = ncol;
      }
    stride += size[i] * count[i]; ncol += count[i];
  }
  if (off[0] < 0 || off[1] < 0 || off[2] < 0) { why = "no x / y / z fields"; return false; }
  out.resize(npoints);
  if (data == "binary") {
    std::vector<char> buf((size_t)stride * npoints);
    f.read(buf.data(), (std::streamsize)buf.size());
    if ((size_t)f.gcount() != buf.size()) { why = "truncated data"; return false; }
    for (size_t i = 0; i < npoints; i++) {
      const char* p = buf.data() + i * stride;
      memcpy(&out[i].x, p + off[0], 4); memcpy(&out[i].y, p + off[1], 4); memcpy(&out[i].z, p + off[2], 4);
      if (off[3] >= 0) memcpy(&out[i].intensity, p + off[3], 4); else out[i].intensity = 0.f;
    }
  } else if (data == "ascii") {
    std::vector<double> v(ncol);
    for (size_t i = 0; i < npoints; i++) {
      for (int c = 0; c < ncol; c++) { std::string t; if (!(f >> t)) { why = "truncated data"; return false; } v[c] = strtod(t.c_str(), nullptr); }
      out[i].x = (float)v[col[0]]; out[i].y = (float)v[col[1]]; out[i].z = (float)v[col[2]];
      out[i].intensity = col[3] >= 0 ? (float)v[col[3]] : 0.f;
    }
  } else { why = "DATA " + data + " not supported"; return false; }
  return true;
}

extern "C" {

int cm_map_save_host(cm_ctx* ctx, int stream_index, const char* dir, int* n_files) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!dir || stream_index < 0 || stream_index >= ctx->map_streams) return fail(ctx, CM_ERR_ARG, "bad argument");
  const cm_config& cfg = ctx->cfg;
  struct Rec { int cube; int vx, vy, vz; cm_point p; };
  std::vector<Rec> recs[2];
  for (int cls = 0; cls < 2; cls++) {
    size_t n = 0;
    int rc = cm_map_export_host(ctx, stream_index, cls, nullptr, nullptr, 0, &n);
    if (rc != CM_OK) return rc;
    std::vector<cm_point> pts(n ? n : 1); std::vector<int> cube(n ? n : 1);
    rc = cm_map_export_host(ctx, stream_index, cls, pts.data(), cube.data(), n, &n);
    if (rc != CM_OK) return rc;
    const float leaf = cls == 0 ? cfg.map_filter_corner : cfg.map_filter_surf;
    const float inv = 1.0f / leaf;
    recs[cls].resize(n);
    for (size_t i = 0; i < n; i++) {
      Rec& r = recs[cls][i];
      r.cube = cube[i]; r.p = pts[i];
      r.vx = (int)floorf(pts[i].x * inv); r.vy = (int)floorf(pts[i].y * inv); r.vz = (int)floorf(pts[i].z * inv);
    }
    const int W = cfg.cube_w, H = cfg.cube_h;
    std::sort(recs[cls].begin(), recs[cls].end(), [W, H](const Rec& a, const Rec& b) {
      if (a.cube != b.cube) {   // file order: i outermost, then j, then k (FeatureMap.h:389-391); cube = i + j W + k W H
        const int ai = a.cube % W, aj = (a.cube / W) % H, ak = a.cube / (W * H), bi = b.cube % W, bj = (b.cube / W) % H, bk = b.cube / (W * H);
        if (ai != bi) return ai < bi;
        if (aj != bj) return aj < bj;
        return ak < bk;
      }
      if (a.vz != b.vz) return a.vz < b.vz;   // VoxelGrid output order inside the cube
      if (a.vy != b.vy) return a.vy < b.vy;
      return a.vx < b.vx;
    });
  }
  const std::string d(dir);
  std::ofstream fout((d + "/index.txt").c_str());
  if (!fout) return fail(ctx, CM_ERR_ARG, "save files error: cannot write " + d + "/index.txt");
  // merge the two sorted cube sequences: per cube, corner file first, then surf
  const int W = cfg.cube_w, H = cfg.cube_h;
  auto order = [W, H](int cube) { const long long i = cube % W, j = (cube / W) % H, k = cube / (W * H); return (i * 100000LL + j) * 100000LL + k; };
  size_t pos[2] = {0, 0};
  int count = 0;
  std::vector<cm_point> buf;
  while (pos[0] < recs[0].size() || pos[1] < recs[1].size()) {
    int cls;
    if (pos[0] >= recs[0].size()) cls = 1;
    else if (pos[1] >= recs[1].size()) cls = 0;
    else cls = order(recs[0][pos[0]].cube) <= order(recs[1][pos[1]].cube) ? 0 : 1;
    const int cube = recs[cls][pos[cls]].cube;
    buf.clear();
    while (pos[cls] < recs[cls].size() && recs[cls][pos[cls]].cube == cube) buf.push_back(recs[cls][pos[cls]++].p);
    if (!write_pcd_binary(file_name(d, count), buf.data(), buf.size())) return fail(ctx, CM_ERR_ARG, "cannot write " + file_name(d, count));
    fout << count << " " << cls << " " << cube % W << " " << (cube / W) % H << " " << cube / (W * H) << " " << buf.size() << std::endl;
    count++;
  }
  if (n_files) *n_files = count;
  return CM_OK;
}

// pushes two host clouds (corner, surf; world coordinates) of ONE stream through the map's insert kernels (= the map voxel filter)
static int insert_clouds(cm_ctx* ctx, int stream_index, const std::vector<cm_point> cloud[2]) {
  const cm_config& cfg = ctx->cfg;
  try {
    cudaSetDevice(cfg.device);
    cudaStream_t st = ctx->stream;
    const int S = ctx->map_streams;
    std::vector<int> n(2 * S, 0);
    n[stream_index] = (int)cloud[0].size(); n[S + stream_index] = (int)cloud[1].size();
    const int cap_c = std::max<int>(1, (int)cloud[0].size()), cap_s = std::max<int>(1, (int)cloud[1].size());
    std::vector<float> tf(12 * S, 0.f);
    for (int s = 0; s < S; s++) { tf[12 * s + 0] = tf[12 * s + 4] = tf[12 * s + 8] = 1.f; }
    ctx->m_corner_in.reserve((size_t)cap_c * sizeof(cm_point)); ctx->m_surf_in.reserve((size_t)cap_s * sizeof(cm_point));
    ctx->m_n_in.reserve(sizeof(int) * 2 * S); ctx->m_tf.reserve(sizeof(float) * 12 * S);
    if (!cloud[0].empty()) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_corner_in.p, cloud[0].data(), cloud[0].size() * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    if (!cloud[1].empty()) CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_surf_in.p, cloud[1].data(), cloud[1].size() * sizeof(cm_point), cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_n_in.p, n.data(), sizeof(int) * 2 * S, cudaMemcpyHostToDevice, st));
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->m_tf.p, tf.data(), sizeof(float) * 12 * S, cudaMemcpyHostToDevice, st));
    // only stream_index has points: bias the base pointers so that [stream_index][0] is the start of the upload
    const float4* pc = (const float4*)ctx->m_corner_in.p - (size_t)stream_index * cap_c;
    const float4* ps = (const float4*)ctx->m_surf_in.p - (size_t)stream_index * cap_s;
    ctx->map.insert(0, pc, (const int*)ctx->m_n_in.p, cap_c, cap_c, nullptr, (const float*)ctx->m_tf.p, st, nullptr, true);   // every file is filtered (:428-456)
    ctx->map.insert(1, ps, (const int*)ctx->m_n_in.p + S, cap_s, cap_s, nullptr, (const float*)ctx->m_tf.p, st, nullptr, true);
    int flags[8];
    CM_CUDA_CHECK(ctx, cudaMemcpyAsync(flags, ctx->map.flags.p, sizeof(flags), cudaMemcpyDeviceToHost, st));
    CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    CM_CUDA_CHECK(ctx, cudaGetLastError());
    if (flags[0]) return fail(ctx, CM_ERR_UNSUPPORTED, "map point outside the supported voxel range (+-65536 voxels)");
    if (flags[2] || flags[3]) return fail(ctx, CM_ERR_CAPACITY, "map capacity exhausted (raise max_*_points in cm_mapping_create)");
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  return CM_OK;
}

int cm_map_load_host(cm_ctx* ctx, int stream_index, const char* dir, int* n_files, size_t* n_points, size_t* n_misplaced) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!dir || stream_index < 0 || stream_index >= ctx->map_streams) return fail(ctx, CM_ERR_ARG, "bad argument");
  const cm_config& cfg = ctx->cfg;
  const std::string d(dir);
  std::ifstream fin((d + "/index.txt").c_str());
  if (!fin) return fail(ctx, CM_ERR_ARG, "cannot open " + d + "/index.txt");
  std::vector<cm_point> cloud[2], tmp;
  int count, type, i, j, k, files = 0;
  long long size;
  size_t misplaced = 0;
  const MappingStream& ms = ctx->mstreams[stream_index];
  while (fin >> count >> type >> i >> j >> k >> size) {
    if (type != 0 && type != 1) continue;
    std::string why;
    if (!read_pcd(file_name(d, count), tmp, why)) return fail(ctx, CM_ERR_ARG, file_name(d, count) + ": " + why);
    for (const cm_point& p : tmp) {   // the reference puts the file into cube (i, j, k); here a point's cube follows from its coordinates
      const int ci = (int)(roundf(p.x / cfg.cube_size) + (float)ms.origin[0]), cj = (int)(roundf(p.y / cfg.cube_size) + (float)ms.origin[1]),
                ck = (int)(roundf(p.z / cfg.cube_size) + (float)ms.origin[2]);
      if (ci != i || cj != j || ck != k) misplaced++;
    }
    cloud[type].insert(cloud[type].end(), tmp.begin(), tmp.end());
    files++;
  }
  { const int rc = insert_clouds(ctx, stream_index, cloud); if (rc != CM_OK) return rc; }
  if (n_files) *n_files = files;
  if (n_points) *n_points = cloud[0].size() + cloud[1].size();
  if (n_misplaced) *n_misplaced = misplaced;
  return CM_OK;
}

// ---- DynamicFeatureMap paging ---------------------------------------------------------------------------------------------------
// util/DynamicFeatureMap.h:129-161 (setupPCDFileName: <dir>/index2.txt, "count type i j k size" with GLOBAL cube indices
// i = round(x / cubeSize), ...), :504-677 (update: the cubes of a window around the sensor's cube are resident; when the sensor
// changes cube, the cubes that enter the window are read -- every file through the map voxel filter -- and the cubes that leave it
// are dropped, their slots reused).  The resident window lives in the device map; cube (i, j, k) of the catalogue is cube
// (i, j, k) + origin of the device lattice (FeatureMap::worldToCube uses the same round()).  When the window would leave the
// device lattice, the lattice is re-centred on the sensor (a plain re-labelling: the map is keyed by coordinates).
int cm_map_page_open_host(cm_ctx* ctx, int stream_index, const char* dir, int window_w, int window_h, int window_d, int* n_entries) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!dir || stream_index < 0 || stream_index >= ctx->map_streams || window_w < 1 || window_h < 1 || window_d < 1 || !(window_w & 1) ||
      !(window_h & 1) || !(window_d & 1))
    return fail(ctx, CM_ERR_ARG, "bad argument (window sizes are odd: the sensor's cube is the centre)");
  if (window_w > ctx->cfg.cube_w || window_h > ctx->cfg.cube_h || window_d > ctx->cfg.cube_d)
    return fail(ctx, CM_ERR_ARG, "paging window larger than the cube lattice of cm_config");
  if (ctx->dist.on) return fail(ctx, CM_ERR_UNSUPPORTED, "paging on a sharded map");
  const std::string d(dir);
  std::ifstream fin((d + "/index2.txt").c_str());
  if (!fin) return fail(ctx, CM_ERR_ARG, "cannot open " + d + "/index2.txt");
  if (ctx->pages.size() != (size_t)ctx->map_streams) ctx->pages.assign(ctx->map_streams, PageState());
  PageState& pg = ctx->pages[stream_index];
  pg = PageState();
  pg.dir = d; pg.win[0] = window_w; pg.win[1] = window_h; pg.win[2] = window_d;
  int count, type, i, j, k; long long size; int n = 0;
  while (fin >> count >> type >> i >> j >> k >> size) {   // :141-155 (a later line for the same cube replaces the earlier one)
    if (type != 0 && type != 1) type = 1;                // `if (!type) corner else surf`
    pg.files[type][{i, j, k}] = count;
    n++;
  }
  pg.open = true;
  if (n_entries) *n_entries = n;
  return CM_OK;
}

int cm_map_page_update_host(cm_ctx* ctx, int stream_index, const float* sensor, int* n_files_loaded, int* n_cubes_evicted, size_t* n_points_loaded) {
  if (!ctx || ctx->map_streams <= 0) return fail(ctx, CM_ERR_ARG, "cm_mapping_create has not been called");
  if (!sensor || stream_index < 0 || stream_index >= ctx->map_streams || ctx->pages.size() != (size_t)ctx->map_streams ||
      !ctx->pages[stream_index].open)
    return fail(ctx, CM_ERR_ARG, "bad argument / cm_map_page_open_host has not been called for this stream");
  PageState& pg = ctx->pages[stream_index];
  const cm_config& cfg = ctx->cfg;
  MappingStream& ms = ctx->mstreams[stream_index];
  const int dims[3] = {cfg.cube_w, cfg.cube_h, cfg.cube_d};
  int g[3];
  for (int a = 0; a < 3; a++) g[a] = (int)roundf(sensor[a] / cfg.cube_size);   // Glo2GloIdx, :318-323
  int files = 0, evicted = 0; size_t points = 0;
  if (n_files_loaded) *n_files_loaded = 0;
  if (n_cubes_evicted) *n_cubes_evicted = 0;
  if (n_points_loaded) *n_points_loaded = 0;
  if (!pg.first && g[0] == pg.sensor[0] && g[1] == pg.sensor[1] && g[2] == pg.sensor[2]) return CM_OK;   // same cube: nothing moves (:556-558)
  const int half[3] = {pg.win[0] / 2, pg.win[1] / 2, pg.win[2] / 2};
  auto in_window = [&](const int c[3], const int centre[3]) {   // !OutRange, :265-273
    for (int a = 0; a < 3; a++) if (c[a] < centre[a] - half[a] || c[a] > centre[a] + half[a]) return false;
    return true;
  };
  try {
    cudaSetDevice(cfg.device);
    cudaStream_t st = ctx->stream;
    // does the new window fit the device lattice with the current origin?  if not, re-centre the lattice on the sensor's cube
    // (and keep the sensor inside the central cubes FeatureMap::update wants, so that the stage entries never shift a paged map)
    bool fits = true;
    for (int a = 0; a < 3; a++)
      if (g[a] - half[a] + ms.origin[a] < 0 || g[a] + half[a] + ms.origin[a] >= dims[a] || g[a] + ms.origin[a] < 3 || g[a] + ms.origin[a] > dims[a] - 4)
        fits = false;
    int new_origin[3] = {ms.origin[0], ms.origin[1], ms.origin[2]}, d[3] = {0, 0, 0};
    if (!fits) for (int a = 0; a < 3; a++) { new_origin[a] = dims[a] / 2 - g[a]; d[a] = new_origin[a] - ms.origin[a]; }
    // cubes of the old window that are outside the new one are dropped (:592-611); on the first call the map keeps what it holds
    const size_t ncubes = (size_t)dims[0] * dims[1] * dims[2];
    std::vector<unsigned char> drop;
    if (!pg.first) {
      drop.assign(ncubes, 0);
      for (int i = pg.sensor[0] - half[0]; i <= pg.sensor[0] + half[0]; i++)
        for (int j = pg.sensor[1] - half[1]; j <= pg.sensor[1] + half[1]; j++)
          for (int k = pg.sensor[2] - half[2]; k <= pg.sensor[2] + half[2]; k++) {
            const int c[3] = {i, j, k};
            if (in_window(c, g)) continue;
            const int li = i + new_origin[0], lj = j + new_origin[1], lk = k + new_origin[2];   // lattice index AFTER the re-centring
            if (li < 0 || li >= dims[0] || lj < 0 || lj >= dims[1] || lk < 0 || lk >= dims[2]) { evicted++; continue; }   // falls off the lattice anyway
            drop[li + lj * dims[0] + lk * dims[0] * dims[1]] = 1;
            evicted++;
          }
    }
    if (!fits || evicted) {
      const unsigned char* d_drop = nullptr;
      if (!drop.empty()) {
        ctx->d_drop.reserve(ncubes);
        CM_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_drop.p, drop.data(), ncubes, cudaMemcpyHostToDevice, st));
        d_drop = (const unsigned char*)ctx->d_drop.p;
      }
      ctx->map.shift(stream_index, d, new_origin, st, false, d_drop);
      CM_CUDA_CHECK(ctx, cudaStreamSynchronize(st));   // `drop` goes out of scope
      for (int a = 0; a < 3; a++) ms.origin[a] = new_origin[a];
    }
    // cubes that enter the window are read (:524-553 first call, :559-590 + 636-672 afterwards), in the reference's loop order
    std::vector<cm_point> cloud[2], tmp;
    for (int i = g[0] - half[0]; i <= g[0] + half[0]; i++)
      for (int j = g[1] - half[1]; j <= g[1] + half[1]; j++)
        for (int k = g[2] - half[2]; k <= g[2] + half[2]; k++) {
          const int c[3] = {i, j, k};
          if (!pg.first && in_window(c, pg.sensor)) continue;
          for (int type = 0; type < 2; type++) {
            auto it = pg.files[type].find({i, j, k});
            if (it == pg.files[type].end()) continue;
            std::string why;
            if (!read_pcd(file_name(pg.dir, it->second), tmp, why)) continue;   // `if(!file) continue;`
            cloud[type].insert(cloud[type].end(), tmp.begin(), tmp.end());
            files++; points += tmp.size();
          }
        }
    if (files) { const int rc = insert_clouds(ctx, stream_index, cloud); if (rc != CM_OK) return rc; }
  } catch (const CudaError& e) {
    return fail(ctx, CM_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.code));
  }
  pg.first = false;
  for (int a = 0; a < 3; a++) pg.sensor[a] = g[a];
  if (n_files_loaded) *n_files_loaded = files;
  if (n_cubes_evicted) *n_cubes_evicted = evicted;
  if (n_points_loaded) *n_points_loaded = points;
  return CM_OK;
}

}  // extern "C"
