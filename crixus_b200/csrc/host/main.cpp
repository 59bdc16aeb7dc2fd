// crixus_b200 -- command-line front end with the reference's process surface: `crixus <file.ini>`
// (/root/reference/src/crixus.cu:57-73,126-134,310,394,433,647-649,698-707,737-795,1125-1135, README.md:92-234).
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#include "ini.h"
#include "output.h"

static int run(int argc, char **argv)
{
  printf("\n\tcrixus_b200 -- B200-native Crixus hot path (%s)\n\n", cx_version());
  if (argc == 1) {
    printf("No configuration file specified.\nCorrect use: crixus filename\nExample use: crixus box.ini\n");
    return CX_NO_FILE;
  }
  if (argc > 2) printf("Ignoring additional arguments after configuration file.\n");
  const std::string cfg = argv[1];
  CxIni ini(cfg);
  if (ini.ParseError() < 0) { printf("Can't load configuration file %s\n", cfg.c_str()); return CX_CANT_READ_CONFIG; }
  const std::string stl = ini.Get("mesh", "stlfile", "UNKNOWN");
  const float dr = (float)ini.GetReal("mesh", "dr", -1);
  printf("Opening file %s ...", stl.c_str());
  std::vector<float> corners, normals, fs_corners, fs_normals;
  int rc = cx_read_stl(stl, corners, normals);
  if (rc != CX_OK) { printf(" [FAILED]\n"); return rc; }
  printf(" [OK]\nMesh size: %g\nRead %zu facets\n", dr, normals.size() / 3);

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
  int frc = cx_read_stl(fsname, fs_corners, fs_normals);
  printf("Checking whether fluid geometry (%s) is available ... %s\n", fsname.c_str(), frc == CX_OK ? "[YES]" : "[NO]");
  if (frc == CX_OK) { job.fs_corners = fs_corners.data(); job.fs_normals = fs_normals.data(); job.fs_nbe = (int64_t)(fs_normals.size() / 3); }
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

  cx_ctx *ctx = nullptr;
  rc = cx_create((int)std::max(0L, ini.GetInteger("system", "gpu-id", 0)), &ctx);
  if (rc != CX_OK) { printf("No usable CUDA device (crixus_b200 has no CPU path)\n"); return rc; }
  rc = cx_preprocess_host(ctx, &job);
  if (rc != CX_OK) { printf("Preprocessing failed: %d %s\n", rc, cx_last_error(ctx)); cx_destroy(ctx); return rc; }
  printf("\n\tInformation:\n\tNumber of vertices:         \t%lld\n\tNumber of boundary elements:\t%lld\n\n", (long long)job.nvert, (long long)job.nbe);
  printf("\tNormals information:\n\tPositive (n.(0,0,1)) minimum z: %g (%g)\n\tNegative (n.(0,0,1)) minimum z: %g (%g)\n\n", job.zstat[0],
         job.zstat[2], job.zstat[1], job.zstat[3]);
  printf("Creation of %lld fluid particles completed. [OK]\n", (long long)job.nfluid);
  printf("Timing: weld %.3f s, upload %.3f s, GPU %.3f s, download %.3f s\n", job.t_weld_s, job.t_h2d_s, job.t_gpu_s, job.t_d2h_s);
  std::vector<CxOutBuf> buf = cx_assemble(job, job.nfluid);
  const std::string fmt = ini.Get("output", "format", "vtu");
  std::string outname = ini.Get("output", "name", cfg.substr(0, cfg.size() >= 4 ? cfg.size() - 4 : 0));
  if (fmt == "h5sph") {
    outname = "0." + outname + ".h5sph";
    rc = cx_write_h5sph(buf, outname);
  } else {
    outname += ".vtu";
    rc = cx_write_vtu(buf, outname);
  }
  printf("Writing output to file %s ... %s\n", outname.c_str(), rc == CX_OK ? "[OK]" : "[FAILED]");
  cx_job_free(&job);
  cx_destroy(ctx);
  return rc;
}

int main(int argc, char **argv)
{
  // the reference's main() discards crixus_main's return code (crixus.cpp:7-10); we return it
  return run(argc, argv) == CX_OK ? 0 : 1;
}
