// crixus_b200 -- the reference's process surface behind the C-ABI: .ini reader, binary STL reader, particle-record
// assembly, VTU / h5sph writers, and cx_run_ini = what `Crixus file.ini` does
// (/root/reference/src/crixus.cu:57-73,126-221,310,394,433,647-649,698-707,737-795,973-1135; src/crixus.cpp:12-289).
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#include "ini.h"
#include "output.h"

static double wall_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct cx_ini {
  CxIni ini;
  explicit cx_ini(const char *fn) : ini(fn) {}
};

extern "C" {

int cx_ini_open(const char *filename, cx_ini **out)
{
  if (!filename || !out) return CX_WRONG_INPUT;
  *out = new cx_ini(filename);
  return (*out)->ini.ParseError();
}
void cx_ini_close(cx_ini *ini) { delete ini; }
int cx_ini_get(const cx_ini *ini, const char *section, const char *name, const char *def, char *buf, int buflen)
{
  if (!ini || !section || !name || !buf || buflen <= 0) return CX_WRONG_INPUT;
  const std::string v = ini->ini.Get(section, name, def ? def : "");
  snprintf(buf, (size_t)buflen, "%s", v.c_str());
  return (int)v.size();
}
long cx_ini_get_integer(const cx_ini *ini, const char *section, const char *name, long def) { return ini->ini.GetInteger(section, name, def); }
double cx_ini_get_real(const cx_ini *ini, const char *section, const char *name, double def) { return ini->ini.GetReal(section, name, def); }
int cx_ini_get_boolean(const cx_ini *ini, const char *section, const char *name, int def) { return ini->ini.GetBoolean(section, name, def != 0) ? 1 : 0; }

void cx_free(void *p) { free(p); }

int cx_stl_read(const char *filename, float **corners, float **normals, int64_t *nbe)
{
  if (!filename || !corners || !normals || !nbe) return CX_WRONG_INPUT;
  *corners = *normals = nullptr;
  *nbe = 0;
  std::vector<float> c, n;
  const int rc = cx_read_stl(filename, c, n);
  if (rc != CX_OK) return rc;
  *nbe = (int64_t)(n.size() / 3);
  *corners = (float *)malloc(std::max<size_t>(c.size() * 4, 4));
  *normals = (float *)malloc(std::max<size_t>(n.size() * 4, 4));
  if (!*corners || !*normals) return CX_INTERNAL_ERROR;
  memcpy(*corners, c.data(), c.size() * 4);
  memcpy(*normals, n.data(), n.size() * 4);
  return CX_OK;
}

int64_t cx_assemble_records(const cx_job *job, int64_t nfluid_for_indices, void *records, int64_t capacity)
{
  if (!job || !job->pos || !job->fpos) return CX_WRONG_INPUT;
  const std::vector<CxOutBuf> buf = cx_assemble(*job, nfluid_for_indices);
  if (records) {
    if ((int64_t)buf.size() > capacity) return CX_WRONG_INPUT;
    memcpy(records, buf.data(), buf.size() * sizeof(CxOutBuf));
  }
  return (int64_t)buf.size();
}

static std::vector<CxOutBuf> as_vector(const void *records, int64_t n)
{
  std::vector<CxOutBuf> v((size_t)(n > 0 ? n : 0));
  if (n > 0) memcpy(v.data(), records, (size_t)n * sizeof(CxOutBuf));
  return v;
}
int cx_write_vtu_records(const void *records, int64_t n, const char *filename)
{
  if ((!records && n > 0) || !filename) return CX_WRONG_INPUT;
  return cx_write_vtu(as_vector(records, n), filename);
}
int cx_write_h5sph_records(const void *records, int64_t n, const char *filename)
{
  if ((!records && n > 0) || !filename) return CX_WRONG_INPUT;
  return cx_write_h5sph(as_vector(records, n), filename);
}

// [output] split=true, crixus.cu:1137-1216: the fluid records, then one file per special-boundary id holding the vertex
// and segment records with that KENT (record content unchanged, AbsoluteIndex stays the index in the unsplit list)
int cx_write_split_records(const void *records, int64_t n, int64_t nfluid_particles, int32_t num_kents, const char *outname, int format)
{
  if ((!records && n > 0) || !outname || nfluid_particles < 0 || nfluid_particles > n || num_kents < 1 || (format != 1 && format != 2))
    return CX_WRONG_INPUT;
  const CxOutBuf *buf = (const CxOutBuf *)records;
  auto write = [&](const std::vector<CxOutBuf> &v, const std::string &name) {   // generic_output, crixus.cpp:261-289
    const std::string fn = format == 2 ? "0." + name + ".h5sph" : name + ".vtu";
    const int rc = format == 2 ? cx_write_h5sph(v, fn) : cx_write_vtu(v, fn);
    printf("Writing output to file %s ... %s\n", fn.c_str(), rc == CX_OK ? "[OK]" : "[FAILED]");
    return rc == CX_OK ? CX_OK : CX_WRITE_FAIL;
  };
  std::vector<std::vector<CxOutBuf>> per((size_t)num_kents);
  for (int64_t i = nfluid_particles; i < n; i++) {
    if (buf[i].kent < 0 || buf[i].kent >= num_kents) return CX_INTERNAL_ERROR;
    per[(size_t)buf[i].kent].push_back(buf[i]);
  }
  printf("Counted:\n - Fluid particles: %lld\n", (long long)nfluid_particles);
  for (int32_t k = 0; k < num_kents; k++)   // the reference's kent-0 count includes the fluid particles (crixus.cu:1155-1156)
    printf(" - Boundary particles, kent %d: %lld\n", k, (long long)(per[(size_t)k].size() + (k == 0 ? nfluid_particles : 0)));
  int rc = write(std::vector<CxOutBuf>(buf, buf + nfluid_particles), std::string(outname) + ".fluid");
  if (rc != CX_OK) return rc;
  for (int32_t k = 0; k < num_kents; k++) {
    if (per[(size_t)k].empty()) continue;
    std::ostringstream nm;
    nm << outname << ".boundary.kent" << k;
    rc = write(per[(size_t)k], nm.str());
    if (rc != CX_OK) return rc;
  }
  return CX_OK;
}

int cx_run_ini(const char *ini_path)
{
  printf("\n\tcrixus_b200 -- B200-native Crixus hot path (%s)\n\n", cx_version());
  if (!ini_path) {
    printf("No configuration file specified.\nCorrect use: crixus filename\nExample use: crixus box.ini\n");
    return CX_NO_FILE;
  }
  const std::string cfg = ini_path;
  CxIni ini(cfg);
  if (ini.ParseError() < 0) { printf("Can't load configuration file %s\n", cfg.c_str()); return CX_CANT_READ_CONFIG; }
  const std::string stl = ini.Get("mesh", "stlfile", "UNKNOWN");
  const float dr = (float)ini.GetReal("mesh", "dr", -1);
  printf("Opening file %s ...", stl.c_str());
  std::vector<float> corners, normals, fs_corners, fs_normals;
  const double t_start = wall_s();
  int rc = cx_read_stl(stl, corners, normals);
  const double t_read = wall_s() - t_start;
  if (rc != CX_OK) { printf(" [FAILED]\n"); return rc; }
  printf(" [OK]\nMesh size: %g\nRead %zu facets\n", dr, normals.size() / 3);
  if (!(dr > 0) || normals.empty()) return CX_WRONG_INPUT;

  cx_job job;
  memset(&job, 0, sizeof(job));
  job.corners = corners.data(); job.normals = normals.data(); job.nbe = (int64_t)(normals.size() / 3); job.dr = dr;
  job.swap_normals = ini.GetBoolean("mesh", "swap_normals", false);
  job.zero_open = ini.GetBoolean("mesh", "zeroOpen", false);
  const char *ax[3] = {"x", "y", "z"};
  for (int d = 0; d < 3; d++) job.periodic[d] = ini.GetBoolean("periodicity", ax[d], false);
  job.use_container = ini.GetBoolean("fluid_container", "use", false);
  if (job.use_container)
    for (int d = 0; d < 3; d++) {
      job.cmin[d] = (float)ini.GetReal("fluid_container", std::string(ax[d]) + "min", 1e9);
      job.cmax[d] = (float)ini.GetReal("fluid_container", std::string(ax[d]) + "max", -1e9);
    }
  // free-surface shape: default <ini minus 4 chars>_fshape.stl, absent file is fine (crixus.cu:647-659)
  std::string fsname = cfg.substr(0, cfg.size() >= 4 ? cfg.size() - 4 : 0) + "_fshape.stl";
  fsname = ini.Get("mesh", "fshape", fsname);
  const int frc = cx_read_stl(fsname, fs_corners, fs_normals);
  printf("Checking whether fluid geometry (%s) is available ... %s\n", fsname.c_str(), frc == CX_OK ? "[YES]" : "[NO]");
  if (frc == CX_OK) { job.fs_corners = fs_corners.data(); job.fs_normals = fs_normals.data(); job.fs_nbe = (int64_t)(fs_normals.size() / 3); }
  // special-boundary grids: <ini minus 4 chars>_sbgrid_<i>.stl or [special_boundary_grids] mesh<i>, i = 1, 2, ... until a
  // file is missing or not binary (crixus.cu:477-510); a short file is a read error (:564-567)
  std::vector<std::vector<float>> sb_corners;
  for (int sbi = 1;; sbi++) {
    std::ostringstream ss;
    ss << sbi;
    std::string name = cfg.substr(0, cfg.size() >= 4 ? cfg.size() - 4 : 0) + "_sbgrid_" + ss.str() + ".stl";
    name = ini.Get("special_boundary_grids", "mesh" + ss.str(), name);
    std::vector<float> c, nrm;
    const int src = cx_read_stl(name, c, nrm);
    printf("Checking whether special boundary grid #%d (%s) is available ... %s\n", sbi, name.c_str(),
           src == CX_FILE_NOT_OPEN ? "[NO]" : (src == CX_STL_NOT_BINARY ? "[YES] but not binary" : "[YES]"));
    if (src == CX_FILE_NOT_OPEN || src == CX_STL_NOT_BINARY) break;
    if (src != CX_OK) return src;
    printf("Read %zu facets of special boundary geometry #%d\n", nrm.size() / 3, sbi);
    sb_corners.push_back(std::move(c));
  }
  std::vector<const float *> sb_ptr;
  std::vector<int64_t> sb_n;
  for (const auto &c : sb_corners) { sb_ptr.push_back(c.data()); sb_n.push_back((int64_t)(c.size() / 9)); }
  job.sb_corners = sb_ptr.data(); job.sb_nbe = sb_n.data(); job.nsb = (int32_t)sb_ptr.size();
  std::vector<cx_fill_spec> fills;
  for (int n = 0;; n++) {  // crixus.cu:734-961: fill_0 is always attempted, then while fill_(n+1).option exists
    std::ostringstream sec;
    sec << "fill_" << n;
    if (n > 0 && ini.Get(sec.str(), "option", "UNKNOWN") == "UNKNOWN") break;
    cx_fill_spec f;
    memset(&f, 0, sizeof(f));
    const std::string opt = ini.Get(sec.str(), "option", "box");
    f.option = opt == "geometry" ? 2 : 1;
    printf("Option for fill #%d: %s\n", n, opt.c_str());
    for (int d = 0; d < 3; d++) {
      if (f.option == 1) {
        f.lo[d] = (float)ini.GetReal(sec.str(), std::string(ax[d]) + "min", 1e9);
        f.hi[d] = (float)ini.GetReal(sec.str(), std::string(ax[d]) + "max", -1e9);
      } else f.seed[d] = (float)ini.GetReal(sec.str(), std::string(ax[d]) + "seed", 1e9);
    }
    f.dr_wall = (float)ini.GetReal(sec.str(), "dr_wall", dr);
    fills.push_back(f);
  }
  job.fills = fills.data(); job.nfills = (int32_t)fills.size();
  // a free-surface file that exists but is shorter than its facet count is a READ_ERROR at the first geometry fill
  // (crixus.cu:869-872); it is reported here, before any GPU work (an input error of an earlier box fill would have come
  // first in the reference: fix one, see the other)
  if (frc == CX_READ_ERROR)
    for (const cx_fill_spec &f : fills)
      if (f.option == 2) { printf("Reading facets of fluid geometry ... [FAILED]\n"); return CX_READ_ERROR; }

  cx_ctx *ctx = nullptr;
  const double t_c0 = wall_s();
  rc = cx_create((int)std::max(0L, ini.GetInteger("system", "gpu-id", 0)), &ctx);
  if (rc != CX_OK) { printf("No usable CUDA device (crixus_b200 has no CPU path)\n"); return rc; }
  const double t_create = wall_s() - t_c0;
  rc = cx_preprocess_host(ctx, &job);
  // the reference leaves this check to its stand-alone tool resources/test-triangle-size.c; here it runs before the first stage, so
  // the diagnosis is available on the failure path too (oversized triangles are what makes calc_vert_volume return
  // CX_FAN_INCOMPLETE, where the reference reads uninitialised memory and carries on)
  if (job.tri_failed > 0)
    printf("WARNING: %lld vertex-barycentre distances exceed dr - 1e-4*dr (max %g): triangles too large for dr, results are unreliable\n",
           (long long)job.tri_failed, job.tri_maxd);
  if (rc != CX_OK) {
    printf("Preprocessing failed: %d %s\n", rc, cx_last_error(ctx));
    if (rc == CX_NO_PER_VERT) printf("(the reference skips such vertices silently, src/crixus_d.cu:229-232; see INTEGRATION.md section 4b)\n");
    cx_destroy(ctx);
    return rc;
  }
  printf("\n\tInformation:\n\tNumber of vertices:         \t%lld\n\tNumber of boundary elements:\t%lld\n\n", (long long)job.nvert, (long long)job.nbe);
  printf("\tNormals information:\n\tPositive (n.(0,0,1)) minimum z: %g (%g)\n\tNegative (n.(0,0,1)) minimum z: %g (%g)\n\n", job.zstat[0],
         job.zstat[2], job.zstat[1], job.zstat[3]);
  printf("Creation of %lld fluid particles completed. [OK]\n", (long long)job.nfluid);
  printf("Timing: weld %.3f s, upload %.3f s, GPU %.3f s, download %.3f s\n", job.t_weld_s, job.t_h2d_s, job.t_gpu_s, job.t_d2h_s);
  // particle records (crixus.cu:973-1097) and the output file: assembled and laid out on the device, streamed to the file
  // through pinned staging (CX_HOST_ASSEMBLY=1 selects the host restatement of the same steps, kept as its checker)
  const double t_a0 = wall_s();
  const std::string fmt = ini.Get("output", "format", "vtu");
  std::string outname = ini.Get("output", "name", cfg.substr(0, cfg.size() >= 4 ? cfg.size() - 4 : 0));
  const bool split = ini.GetBoolean("output", "split", false);   // crixus.cu:1135
  const char *hostenv = getenv("CX_HOST_ASSEMBLY");
  const bool on_host = hostenv && hostenv[0] == '1';
  std::vector<CxOutBuf> buf;
  void *rec_d = nullptr;
  int64_t nrec = 0;
  if (on_host) {
    buf = cx_assemble(job, job.nfluid);
    nrec = (int64_t)buf.size();
  } else {
    const int64_t cap = job.nfluid + job.nvert + job.nbe;
    if (cudaMalloc(&rec_d, 64 * (size_t)std::max<int64_t>(cap, 1)) != cudaSuccess) { cx_destroy(ctx); return CX_CUDA_ERROR; }
    rc = cx_assemble_records_device(ctx, &job.params_fill, job.dimg, job.d_fpos, job.d_pos, job.d_norm, job.d_ep, job.d_surf, job.d_vol,
                                    job.d_sbid, job.nvert, job.nbe, job.nfluid, rec_d, cap, &nrec);
    if (rc == CX_OK) rc = cx_synchronize(ctx);
    if (rc != CX_OK) { printf("Record assembly failed: %d %s\n", rc, cx_last_error(ctx)); cudaFree(rec_d); cx_destroy(ctx); return rc; }
    if (split) {   // the per-id files are cut on the host
      buf.resize((size_t)nrec);
      if (cudaMemcpy(buf.data(), rec_d, 64 * (size_t)nrec, cudaMemcpyDeviceToHost) != cudaSuccess) rc = CX_CUDA_ERROR;
    }
  }
  const double t_asm = wall_s() - t_a0;
  if (rc == CX_OK && split) {   // crixus.cu:1137-1207
    rc = cx_write_split_records(buf.data(), nrec, job.nfluid, job.nsb + 1, outname.c_str(), fmt == "h5sph" ? 2 : 1);
  } else if (rc == CX_OK) {
    outname = fmt == "h5sph" ? "0." + outname + ".h5sph" : outname + ".vtu";   // generic_output, crixus.cpp:261-289
    if (on_host) rc = fmt == "h5sph" ? cx_write_h5sph(buf, outname) : cx_write_vtu(buf, outname);
    else rc = fmt == "h5sph" ? cx_write_h5sph_records_device(ctx, rec_d, nrec, outname.c_str())
                             : cx_write_vtu_records_device(ctx, rec_d, nrec, outname.c_str());
    printf("Writing output to file %s ... %s\n", outname.c_str(), rc == CX_OK ? "[OK]" : "[FAILED]");
  }
  if (rec_d) cudaFree(rec_d);
  printf("Timing: read STL %.3f s, CUDA context %.3f s, assemble %.3f s (%s), write %.3f s, total %.3f s\n", t_read, t_create, t_asm,
         on_host ? "host" : "device", wall_s() - t_a0 - t_asm, wall_s() - t_start);
  cx_job_free(&job);
  cx_destroy(ctx);
  return rc;
}

}  // extern "C"
