// crixus_b200 -- particle records and writers (the reference's OutBuf, src/crixus.h:33-36, and its two writers).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../../include/crixus_b200.h"

struct CxOutBuf {  // 64 bytes: 8 float + 8 int, member order of src/crixus.h:33-36
  float x, y, z, nx, ny, nz, vol, surf;
  int32_t kpar, kfluid, kent, kparmob, iref, ep1, ep2, ep3;
};

// crixus.cu:973-1097: fluid particles in bit order, alive vertices in index order, segments in sorted order
std::vector<CxOutBuf> cx_assemble(const cx_job &job, int64_t nfluid_for_indices);
int cx_write_vtu(const std::vector<CxOutBuf> &buf, const std::string &filename);     // crixus.cpp:62-259
int cx_write_h5sph(const std::vector<CxOutBuf> &buf, const std::string &filename);   // crixus.cpp:12-60 (own HDF5 writer)
// the parts of the two file formats that do not depend on the particle data (shared with the device-side writers)
std::string cx_vtu_header(int len);
const char *cx_vtu_trailer();
int cx_h5sph_header(uint64_t len, std::vector<uint8_t> &out);
// binary STL, crixus.cu:126-221.  Returns CX_OK / CX_FILE_NOT_OPEN / CX_STL_NOT_BINARY / CX_READ_ERROR
int cx_read_stl(const std::string &filename, std::vector<float> &corners, std::vector<float> &normals);
