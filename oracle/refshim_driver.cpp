// TEST INFRASTRUCTURE -- not part of the product.
//
// Host driver for the reference's own device code.  oracle/build_ref.sh compiles this file as host C++
// with -I<shim cuda.h> -I/root/reference/src, so the #include below pulls in the reference's
// src/crixus_d.cu *unmodified* (all 15 kernels + device functions).  Every "launch" runs the kernel
// function on OpenMP threads with blockDim=1, gridDim=#threads, so the grid-stride loops do the work and
// the shared-memory tree reductions degenerate.
//
// What this file adds is only glue: setting the crixus_d:: "constant" globals the way
// src/crixus.cu:262-283,428,464-467,711-721,797 does, and C entry points for ctypes.
#include <omp.h>
#include <vector>
#include <cuda.h>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

#include "crixus_d.cu"   // the reference's kernels, found through -I/root/reference/src

int crixus_main(int, char **) { return 0; }   // crixus.cpp's (renamed) main refers to it

static int g_threads = 0;

template <class F> static void launch(int nthreads, F f)
{
  if (nthreads <= 0) nthreads = g_threads > 0 ? g_threads : omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
  {
    threadIdx = uint3{0, 0, 0};
    blockDim = dim3{1, 1, 1};
    blockIdx = uint3{(unsigned)omp_get_thread_num(), 0, 0};
    gridDim = dim3{(unsigned)omp_get_num_threads(), 1, 1};
    f();
  }
}

extern "C" {

void refshim_set_threads(int n) { g_threads = n; }

// mirrors the cudaMemcpyToSymbol calls of src/crixus.cu (278-283, 428, 467, 711-721, 797)
void refshim_set_consts(const float *griddr4, const int *gridn4, float eps, const float *dmin4, const float *dmax4,
                        const float *fcmin4, const float *fcmax4, const int *per3, float dr, float dr_wall)
{
  for (int j = 0; j < 4; j++) {
    crixus_d::griddr.a[j] = griddr4[j];
    crixus_d::gridn.a[j] = gridn4[j];
    crixus_d::dmin.a[j] = dmin4[j];
    crixus_d::dmax.a[j] = dmax4[j];
    crixus_d::fcmin.a[j] = fcmin4[j];
    crixus_d::fcmax.a[j] = fcmax4[j];
  }
  for (int j = 0; j < 3; j++) crixus_d::per[j] = per3[j] != 0;
  crixus_d::eps = eps;
  crixus_d::dr = dr;
  crixus_d::dr_wall = dr_wall;
}

void refshim_swap_normals(float *norm, int nbe)
{
  launch(0, [&] { crixus_d::swap_normals((uf4 *)norm, nbe); });
}

// out4 = {xminp, xminn, nminp, nminn}, initialised as in src/crixus.cu:325-326
void refshim_set_bound_elem(float *pos, float *norm, float *surf, int *ep, unsigned nbe, int nvert, float *out4)
{
  out4[0] = 1e10f; out4[1] = 1e10f; out4[2] = 0.f; out4[3] = 0.f;
  Lock lock;
  launch(0, [&] {
    crixus_d::set_bound_elem((uf4 *)pos, (uf4 *)norm, surf, (ui4 *)ep, nbe, out4 + 0, out4 + 1, out4 + 2, out4 + 3,
                             lock, nvert);
  });
}

// init_cell_idx + count_cell_bes + add_up_indices + sort_bes (src/crixus.cu:355-361), run on ONE thread so that the
// atomicAdd scatter of sort_bes yields the stable (original-index) order inside each cell.
void refshim_bin(float *pre_pos, float *pos, float *pre_norm, float *norm, int *pre_ep, int *ep, float *pre_surf,
                 float *surf, int *cell_idx, unsigned nbe, unsigned nvert)
{
  std::vector<int> cur(crixus_d::gridn.a[3] + 1, 0);
  launch(1, [&] { crixus_d::init_cell_idx(cell_idx, cur.data()); });
  launch(1, [&] { crixus_d::count_cell_bes((uf4 *)pre_pos, cell_idx, nbe, nvert); });
  launch(1, [&] { crixus_d::add_up_indices(cell_idx, cur.data()); });
  launch(1, [&] {
    crixus_d::sort_bes((uf4 *)pre_pos, (uf4 *)pos, (uf4 *)pre_norm, (uf4 *)norm, (ui4 *)pre_ep, (ui4 *)ep, pre_surf,
                       surf, cell_idx, cur.data(), nbe, nvert);
  });
}

void refshim_find_links(float *pos, int nvert, int *newlink, int idim)
{
  launch(0, [&] { crixus_d::find_links((uf4 *)pos, nvert, newlink, idim); });
}

void refshim_periodicity_links(float *pos, int *ep, int nvert, int nbe, int *newlink, int idim)
{
  launch(0, [&] { crixus_d::periodicity_links((uf4 *)pos, (ui4 *)ep, nvert, nbe, newlink, idim); });
}

// identifySpecialBoundarySegments (crixus_d.cu:1085-1112); eps must have been set through refshim_set_consts
void refshim_identify_sb(float *pos, int *ep, int nvert, int nbe, float *sbpos, int *sbep, int sbnbe, int *sbid, int sbi)
{
  launch(0, [&] { crixus_d::identifySpecialBoundarySegments((uf4 *)pos, (ui4 *)ep, nvert, nbe, (uf4 *)sbpos, (ui4 *)sbep, sbnbe, sbid, sbi); });
}

void refshim_calc_trisize(int *ep, int *trisize, int nbe)
{
  launch(0, [&] { crixus_d::calc_trisize((ui4 *)ep, trisize, nbe); });
}

void refshim_calc_vert_volume(float *pos, float *norm, int *ep, float *vol, int *trisize, int *cell_idx, int nvert,
                              int nbe, int zeroOpen)
{
  launch(0, [&] {
    crixus_d::calc_vert_volume((uf4 *)pos, (uf4 *)norm, (ui4 *)ep, vol, trisize, cell_idx, nvert, nbe, zeroOpen != 0);
  });
}

unsigned refshim_fill_fluid(unsigned *fpos, float xmin, float xmax, float ymin, float ymax, float zmin, float zmax)
{
  unsigned nfi = 0;
  Lock lock;
  launch(0, [&] { crixus_d::fill_fluid(fpos, &nfi, xmin, xmax, ymin, ymax, zmin, zmax, lock); });
  return nfi;
}

// one sweep of fill_fluid_complex (src/crixus.cu:934-945 runs these until nfi == 0)
unsigned refshim_fill_fluid_complex_sweep(unsigned *fpos, float *norm, int *ep, float *pos, int nbe, int sInd, int cnbe,
                                          int iteration, int *cell_idx)
{
  unsigned nfi = 0;
  Lock lock;
  launch(0, [&] {
    crixus_d::fill_fluid_complex(fpos, &nfi, (uf4 *)norm, (ui4 *)ep, (uf4 *)pos, nbe, sInd, lock, cnbe, iteration,
                                 cell_idx);
  });
  return nfi;
}

// the reference's host weld (src/crixus.cpp:343-455).  corners: ntri*3 points (xyz); out_verts has room for
// ntri*3 points; out_ids: ntri*3 new ids.  Returns the number of welded vertices.
int refshim_weld(const float *corners, int ntri, float dr, float *out_verts, unsigned *out_ids)
{
  std::vector<std::vector<float> > all_verts(3 * (size_t)ntri, std::vector<float>(3)), new_verts;
  std::vector<std::vector<unsigned> > old_ids((size_t)ntri, std::vector<unsigned>(3)), new_ids;
  for (size_t c = 0; c < 3 * (size_t)ntri; c++)
    for (int k = 0; k < 3; k++) all_verts[c][k] = corners[3 * c + k];
  for (size_t t = 0; t < (size_t)ntri; t++)
    for (unsigned k = 0; k < 3; k++) old_ids[t][k] = 3 * t + k;
  remove_duplicate_vertices(all_verts, old_ids, new_verts, new_ids, dr);
  for (size_t v = 0; v < new_verts.size(); v++)
    for (int k = 0; k < 3; k++) out_verts[3 * v + k] = new_verts[v][k];
  for (size_t t = 0; t < (size_t)ntri; t++)
    for (int k = 0; k < 3; k++) out_ids[3 * t + k] = new_ids[t][k];
  return (int)new_verts.size();
}

// the reference's own VTU writer (src/crixus.cpp:62-259) on an array of 64-byte OutBuf records
int refshim_vtk_output(void *records, int len, const char *filename) { return vtk_output((OutBuf *)records, len, filename); }

// the reference's own .ini reader (vendored inih + INIReader, src/ini/)
void *refshim_ini_open(const char *filename, int *parse_error)
{
  INIReader *r = new INIReader(filename);
  *parse_error = r->ParseError();
  return r;
}
void refshim_ini_close(void *r) { delete (INIReader *)r; }
int refshim_ini_get(void *r, const char *section, const char *name, const char *def, char *buf, int buflen)
{
  const std::string v = ((INIReader *)r)->Get(section, name, def);
  snprintf(buf, (size_t)buflen, "%s", v.c_str());
  return (int)v.size();
}
long refshim_ini_get_integer(void *r, const char *section, const char *name, long def) { return ((INIReader *)r)->GetInteger(section, name, def); }
double refshim_ini_get_real(void *r, const char *section, const char *name, double def) { return ((INIReader *)r)->GetReal(section, name, def); }
int refshim_ini_get_boolean(void *r, const char *section, const char *name, int def) { return ((INIReader *)r)->GetBoolean(section, name, def != 0) ? 1 : 0; }

}  // extern "C"
