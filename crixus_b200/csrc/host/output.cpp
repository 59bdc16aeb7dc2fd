#include "output.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <string>

// ---------------------------------------------------------------------------------------------- STL
// Bulk reader with the outcome of the reference's per-value iostream loop (crixus.cu:137-221), malformed files included
// (restated call by call in oracle/stl_reader.cpp and compared on truncated files in tests/test_frontend.py):
//   * the loop runs min(count, floor(payload / 50) + 1) times -- the iteration in which the end of the file is hit still
//     appends a facet, made of the bytes that were still there and, for the rest, the PREVIOUS facet's values;
//   * READ_ERROR only if that differs from the count in the header (crixus.cu:218-221), i.e. a file whose last facet is cut
//     short is accepted by the reference, and so it is here.
int cx_read_stl(const std::string &filename, std::vector<float> &corners, std::vector<float> &normals)
{
  corners.clear();
  normals.clear();
  FILE *f = fopen(filename.c_str(), "rb");
  if (!f) return CX_FILE_NOT_OPEN;
  unsigned char head[84];
  memset(head, 0, sizeof(head));
  const size_t got = fread(head, 1, sizeof(head), f);
  if (got >= 5 && memcmp(head, "solid", 5) == 0) { fclose(f); return CX_STL_NOT_BINARY; }  // crixus.cu:137-153
  uint32_t n = 0;
  memcpy(&n, head + 80, 4);                       // a short read leaves the missing bytes untouched (zero here)
  uint64_t avail = 0;
  if (got == sizeof(head)) {
    const long at = ftell(f);
    if (fseek(f, 0, SEEK_END) == 0) {
      const long end = ftell(f);
      if (end > at) avail = (uint64_t)(end - at);
    }
    fseek(f, at, SEEK_SET);
  }
  const uint64_t m = got == sizeof(head) ? std::min<uint64_t>(n, avail / 50 + 1) : 0;   // facets the reference's loop appends
  corners.resize((size_t)9 * m);
  normals.resize((size_t)3 * m);
  const size_t CH = 1 << 16;                      // facets per fread
  std::vector<unsigned char> buf(CH * 50);
  unsigned char rec[48];
  memset(rec, 0, sizeof(rec));                    // m_v_floats before the first facet (uninitialised in the reference)
  for (uint64_t i0 = 0; i0 < m; i0 += CH) {
    const size_t cnt = (size_t)std::min<uint64_t>(CH, m - i0);
    const size_t have = fread(buf.data(), 1, cnt * 50, f);
    for (size_t i = 0; i < cnt; i++) {
      const size_t off = i * 50, left = have > off ? have - off : 0;
      memcpy(rec, buf.data() + off, std::min<size_t>(48, left));
      memcpy(&normals[3 * (size_t)(i0 + i)], rec, 12);
      memcpy(&corners[9 * (size_t)(i0 + i)], rec + 12, 36);
    }
  }
  fclose(f);
  if (m != n) return CX_READ_ERROR;               // crixus.cu:218-221
  return CX_OK;
}

// ---------------------------------------------------------------------------------------------- assembly
std::vector<CxOutBuf> cx_assemble(const cx_job &job, int64_t nfluid)
{
  const int64_t nvert = job.nvert, nbe = job.nbe;
  const cx_params &p = job.params_fill;
  const float dr = job.dr;
  std::vector<CxOutBuf> buf;
  buf.reserve((size_t)(nvert + nbe + job.nfluid));
  const float fluid_vol = (float)pow(dr, 3);  // crixus.cu:986
  const int64_t d0 = job.dimg[0], d1 = job.dimg[1], G = job.dimg[0] * job.dimg[1] * job.dimg[2];
  CxOutBuf z;
  memset(&z, 0, sizeof(z));
  for (int64_t j = 0; j < G; j++) {  // crixus.cu:991-1019
    if (!(job.fpos[j >> 5] >> (j & 31) & 1u)) continue;
    CxOutBuf b = z;
    const int64_t k = j / (d1 * d0), n = j % (d1 * d0);
    b.z = p.fcmin[2] + dr * (float)k;
    b.y = p.fcmin[1] + dr * (float)(n / d0);
    b.x = p.fcmin[0] + dr * (float)(n % d0);
    b.vol = fluid_vol;
    b.kpar = 1; b.kfluid = 1;
    b.iref = (int32_t)buf.size();
    buf.push_back(b);
  }
  std::vector<int32_t> nvshift((size_t)nvert, 0);
  int32_t ishift = 0;
  for (int64_t i = 0; i < nvert; i++) {  // crixus.cu:1026-1053
    if (job.pos[4 * i] < -1e9f) { ishift++; continue; }
    nvshift[i] = ishift;
    CxOutBuf b = z;
    b.x = job.pos[4 * i]; b.y = job.pos[4 * i + 1]; b.z = job.pos[4 * i + 2];
    b.vol = job.vol[i];
    b.kpar = 2; b.kfluid = 1;
    b.kent = job.sbid ? job.sbid[i] : 0;   // crixus.cu:1042-1045
    b.iref = (int32_t)buf.size();
    buf.push_back(b);
  }
  // segments: crixus.cu:1056-1097.  With special boundaries they are stably grouped by KENT (counting sort over the ids).
  const int32_t *sbid = job.sbid;
  int32_t nk = 1;
  if (sbid) for (int64_t i = 0; i < nbe; i++) nk = std::max(nk, sbid[nvert + i] + 1);
  std::vector<int64_t> isbe((size_t)nk + 1, 0);
  if (sbid) for (int64_t i = 0; i < nbe; i++) isbe[(size_t)sbid[nvert + i] + 1]++;
  for (int32_t k = 0; k < nk; k++) isbe[(size_t)k + 1] += isbe[(size_t)k];
  const int64_t ncur = (int64_t)buf.size();
  buf.resize((size_t)(ncur + nbe), z);
  for (int64_t i = 0; i < nbe; i++) {
    const int32_t kent = sbid ? sbid[nvert + i] : 0;
    const int64_t l = ncur + isbe[(size_t)kent]++;
    CxOutBuf &b = buf[(size_t)l];
    const float *q = job.pos + 4 * (nvert + i);
    b.x = q[0]; b.y = q[1]; b.z = q[2];
    b.nx = job.norm[4 * i]; b.ny = job.norm[4 * i + 1]; b.nz = job.norm[4 * i + 2];
    b.surf = job.surf[i];
    b.kpar = 3; b.kfluid = 1; b.kent = kent;
    b.iref = (int32_t)l;
    const int32_t *e = job.ep + 4 * i;
    b.ep1 = (int32_t)nfluid + e[0] - nvshift[e[0]];
    b.ep2 = (int32_t)nfluid + e[1] - nvshift[e[1]];
    b.ep3 = (int32_t)nfluid + e[2] - nvshift[e[2]];
  }
  return buf;
}

// ---------------------------------------------------------------------------------------------- VTU
// XML part of the file up to and including the '_' that starts the appended data (crixus.cpp:70-147)
std::string cx_vtu_header(int len)
{
  std::string h;
  char line[256];
  auto add = [&](const char *fmt, auto... a) { snprintf(line, sizeof(line), fmt, a...); h += line; };
  auto scalar = [&](const char *type, const char *name, size_t off) {
    add("  <DataArray type='%s' Name='%s' format='appended' offset='%zu'/>\n", type, name, off);
  };
  auto vec = [&](const char *type, const char *name, unsigned dim, size_t off) {
    if (name) add("  <DataArray type='%s' Name='%s' NumberOfComponents='%u' format='appended' offset='%zu'/>\n", type, name, dim, off);
    else add("  <DataArray type='%s' NumberOfComponents='%u' format='appended' offset='%zu'/>\n", type, dim, off);
  };
  h += "<?xml version='1.0'?>\n";
  add("<VTKFile type= 'UnstructuredGrid'  version= '0.1'  byte_order= '%s'>\n", "LittleEndian");
  h += " <UnstructuredGrid>\n";
  add("  <Piece NumberOfPoints='%d' NumberOfCells='%d'>\n", len, len);
  h += "   <PointData Scalars='Volume'>\n";
  size_t off = 0;
  scalar("Float32", "Volume", off); off += sizeof(float) * len + sizeof(int);
  scalar("Float32", "Surface", off); off += sizeof(float) * len + sizeof(int);
  scalar("UInt32", "ParticleType", off); off += 4 * (size_t)len + sizeof(int);
  scalar("UInt32", "FluidType", off); off += 4 * (size_t)len + sizeof(int);
  scalar("UInt32", "KENT", off); off += 4 * (size_t)len + sizeof(int);
  scalar("UInt32", "MovingBoundary", off); off += 4 * (size_t)len + sizeof(int);
  scalar("UInt32", "AbsoluteIndex", off); off += 4 * (size_t)len + sizeof(int);
  vec("Float32", "Normal", 3, off); off += 12 * (size_t)len + sizeof(int);
  vec("UInt32", "VertexParticle", 3, off); off += 12 * (size_t)len + sizeof(int);
  h += "   </PointData>\n";
  h += "   <Points>\n";
  vec("Float64", nullptr, 3, off); off += 24 * (size_t)len + sizeof(int);
  h += "   </Points>\n";
  h += "   <Cells>\n";
  scalar("Int32", "connectivity", off); off += 4 * (size_t)len + sizeof(int);
  scalar("Int32", "offsets", off); off += 4 * (size_t)len + sizeof(int);
  h += "  <DataArray type='Int32' Name='types' format='ascii'>\n";
  h.reserve(h.size() + 2 * (size_t)len + 256);
  for (int i = 0; i < len; i++) h += "1\t";   // crixus.cpp:133-137
  h += "\n";
  h += "  </DataArray>\n";
  h += "   </Cells>\n";
  h += "  </Piece>\n";
  h += " </UnstructuredGrid>\n";
  h += " <AppendedData encoding='raw'>\n_";
  return h;
}
const char *cx_vtu_trailer() { return " </AppendedData>\n</VTKFile>"; }

int cx_write_vtu(const std::vector<CxOutBuf> &buf, const std::string &filename)
{
  FILE *fid = fopen(filename.c_str(), "w");
  if (!fid) return CX_NO_WRITE_FILE;
  setvbuf(fid, nullptr, _IOFBF, 8 << 20);
  const int len = (int)buf.size();
  {
    const std::string h = cx_vtu_header(len);
    fwrite(h.data(), 1, h.size(), fid);
  }
  // appended arrays: int32 byte count, then the values; converted AoS -> SoA through a fixed 4 MB staging buffer
  const int CH = 1 << 16;
  std::vector<char> tmp((size_t)CH * 24);
  auto emit = [&](size_t elem, auto getter) {
    int nbytes = (int)(elem * (size_t)len);
    fwrite(&nbytes, sizeof(nbytes), 1, fid);
    for (int i0 = 0; i0 < len; i0 += CH) {
      const int m = std::min(CH, len - i0);
      for (int i = 0; i < m; i++) getter(buf[(size_t)i0 + i], tmp.data() + elem * (size_t)i);
      fwrite(tmp.data(), elem, (size_t)m, fid);
    }
  };
  emit(4, [](const CxOutBuf &b, char *o) { memcpy(o, &b.vol, 4); });
  emit(4, [](const CxOutBuf &b, char *o) { memcpy(o, &b.surf, 4); });
  emit(4, [](const CxOutBuf &b, char *o) { memcpy(o, &b.kpar, 4); });
  emit(4, [](const CxOutBuf &b, char *o) { memcpy(o, &b.kfluid, 4); });
  emit(4, [](const CxOutBuf &b, char *o) { memcpy(o, &b.kent, 4); });
  emit(4, [](const CxOutBuf &b, char *o) { memcpy(o, &b.kparmob, 4); });
  emit(4, [](const CxOutBuf &b, char *o) { memcpy(o, &b.iref, 4); });
  emit(12, [](const CxOutBuf &b, char *o) { memcpy(o, &b.nx, 12); });
  emit(12, [](const CxOutBuf &b, char *o) { memcpy(o, &b.ep1, 12); });
  emit(24, [](const CxOutBuf &b, char *o) { double v[3] = {(double)b.x, (double)b.y, (double)b.z}; memcpy(o, v, 24); });
  for (int pass = 0; pass < 2; pass++) {   // connectivity 0..len-1, offsets 1..len
    int nbytes = 4 * len;
    fwrite(&nbytes, sizeof(nbytes), 1, fid);
    uint32_t *idx = (uint32_t *)tmp.data();
    for (int i0 = 0; i0 < len; i0 += CH) {
      const int m = std::min(CH, len - i0);
      for (int i = 0; i < m; i++) idx[i] = (uint32_t)(i0 + i + pass);
      fwrite(idx, 4, (size_t)m, fid);
    }
  }
  fputs(cx_vtu_trailer(), fid);
  fclose(fid);
  return CX_OK;
}

// ---------------------------------------------------------------------------------------------- minimal HDF5
// What the reference asks libhdf5 for (crixus.cpp:12-60): one contiguous 1-D dataset "Compound" of `len` 64-byte
// compound records (16 members, names/offsets below) in the root group.  libhdf5 is absent from this image, so the
// file is produced here following the HDF5 File Format Specification (superblock v0, v1 object headers, v1 group
// B-tree + local heap + symbol-table node, dataspace v1, compound datatype v1, layout v3 contiguous).  A matching
// independent parser lives in tests/h5mini.py.
namespace {
struct Bytes {
  std::vector<uint8_t> b;
  void u8(uint8_t v) { b.push_back(v); }
  void u16(uint16_t v) { for (int i = 0; i < 2; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void u32(uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void u64(uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void raw(const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; b.insert(b.end(), q, q + n); }
  void pad_to(size_t n) { while (b.size() < n) b.push_back(0); }
  void align8() { while (b.size() % 8) b.push_back(0); }
  size_t size() const { return b.size(); }
};
const uint64_t UNDEF = 0xffffffffffffffffull;

void datatype_f32(Bytes &o)
{
  o.u8(0x11); o.u8(0x20); o.u8(0x1f); o.u8(0x00);  // class 1 (float) v1; LE, implied-1 mantissa, sign bit 31
  o.u32(4);
  o.u16(0); o.u16(32); o.u8(23); o.u8(8); o.u8(0); o.u8(23); o.u32(127);
}
void datatype_i32(Bytes &o)
{
  o.u8(0x10); o.u8(0x08); o.u8(0x00); o.u8(0x00);  // class 0 (fixed point) v1; LE, two's complement
  o.u32(4);
  o.u16(0); o.u16(32);
}
void message(Bytes &o, uint16_t type, const Bytes &data)
{
  size_t n = (data.size() + 7) & ~(size_t)7;
  o.u16(type); o.u16((uint16_t)n); o.u8(0); o.u8(0); o.u8(0); o.u8(0);
  o.raw(data.b.data(), data.size());
  for (size_t i = data.size(); i < n; i++) o.u8(0);
}
}  // namespace

// everything of the file that precedes the raw records (they start at byte .size() of the result)
int cx_h5sph_header(uint64_t len, std::vector<uint8_t> &out)
{
  static const char *names[16] = {"Coords_0", "Coords_1", "Coords_2", "Normal_0", "Normal_1", "Normal_2", "Volume", "Surface",
                                  "ParticleType", "FluidType", "KENT", "MovingBoundary", "AbsoluteIndex", "VertexParticle1",
                                  "VertexParticle2", "VertexParticle3"};  // crixus.cpp:23-38
  // ---- layout of the metadata blocks
  const uint64_t a_super = 0, a_root_hdr = 96;
  // root object header: 16-byte prefix + one symbol-table message (8 + 16)
  const uint64_t a_heap = a_root_hdr + 16 + 24;                  // 136
  const uint64_t heap_data_size = 40;
  const uint64_t a_heap_data = a_heap + 32;                      // 168
  const uint64_t a_btree = a_heap_data + heap_data_size;         // 208
  const uint64_t btree_size = 24 + (2 * 16 + 1) * 8 + (2 * 16) * 8;  // 544
  const uint64_t a_snod = a_btree + btree_size;                  // 752
  const uint64_t snod_size = 8 + 8 * 40;                         // 328
  const uint64_t a_dset_hdr = a_snod + snod_size;                // 1080
  // dataset messages
  Bytes m_space, m_type, m_fill, m_layout, msgs;
  m_space.u8(1); m_space.u8(1); m_space.u8(0); m_space.u8(0); m_space.u32(0); m_space.u64(len);
  m_type.u8(0x16); m_type.u8(16); m_type.u8(0); m_type.u8(0); m_type.u32(64);  // class 6 (compound) v1, 16 members, size 64
  for (int i = 0; i < 16; i++) {
    const size_t nl = strlen(names[i]) + 1, npad = (nl + 7) & ~(size_t)7;
    m_type.raw(names[i], nl);
    for (size_t k = nl; k < npad; k++) m_type.u8(0);
    m_type.u32((uint32_t)(4 * i));                     // byte offset of member
    m_type.u8(0); m_type.u8(0); m_type.u8(0); m_type.u8(0);  // dimensionality + reserved
    m_type.u32(0); m_type.u32(0);                      // permutation, reserved
    for (int d = 0; d < 4; d++) m_type.u32(0);         // dimension sizes
    if (i < 8) datatype_f32(m_type); else datatype_i32(m_type);
  }
  m_fill.u8(2); m_fill.u8(2); m_fill.u8(2); m_fill.u8(0);   // fill value v2: alloc late, write if set, undefined
  const uint64_t dset_msgs_size = (8 + ((m_space.size() + 7) & ~7ull)) + (8 + ((m_type.size() + 7) & ~7ull)) + (8 + 8) + (8 + 24);
  const uint64_t a_data = (a_dset_hdr + 16 + dset_msgs_size + 7) & ~7ull;
  m_layout.u8(3); m_layout.u8(1); m_layout.u64(a_data); m_layout.u64(len * 64);
  message(msgs, 0x0001, m_space);
  message(msgs, 0x0003, m_type);
  message(msgs, 0x0005, m_fill);
  message(msgs, 0x0008, m_layout);
  if (msgs.size() != dset_msgs_size) return CX_INTERNAL_ERROR;
  const uint64_t a_eof = a_data + len * 64;

  Bytes o;
  // ---- superblock v0
  const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  o.raw(sig, 8);
  o.u8(0); o.u8(0); o.u8(0); o.u8(0); o.u8(0); o.u8(8); o.u8(8); o.u8(0);
  o.u16(4); o.u16(16); o.u32(0);
  o.u64(0); o.u64(UNDEF); o.u64(a_eof); o.u64(UNDEF);
  // root group symbol table entry
  o.u64(0); o.u64(a_root_hdr); o.u32(1); o.u32(0); o.u64(a_btree); o.u64(a_heap);
  o.pad_to(a_root_hdr);
  (void)a_super;
  // ---- root group object header (v1): symbol table message
  o.u8(1); o.u8(0); o.u16(1); o.u32(1); o.u32(24); o.u32(0);
  { Bytes st; st.u64(a_btree); st.u64(a_heap); message(o, 0x0011, st); }
  o.pad_to(a_heap);
  // ---- local heap
  o.raw("HEAP", 4); o.u8(0); o.u8(0); o.u8(0); o.u8(0);
  o.u64(heap_data_size); o.u64(24); o.u64(a_heap_data);
  o.pad_to(a_heap_data);
  o.u64(0);                                   // offset 0: "" (root)
  o.raw("Compound\0\0\0\0\0\0\0", 16);        // offset 8
  o.u64(1); o.u64(16);                        // free block at 24: next = H5HL_FREE_NULL, size 16
  o.pad_to(a_btree);
  // ---- B-tree node (group, leaf level)
  o.raw("TREE", 4); o.u8(0); o.u8(0); o.u16(1); o.u64(UNDEF); o.u64(UNDEF);
  o.u64(0); o.u64(a_snod); o.u64(8);          // key0 = "", child0, key1 = "Compound"
  o.pad_to(a_snod);
  // ---- symbol table node
  o.raw("SNOD", 4); o.u8(1); o.u8(0); o.u16(1);
  o.u64(8); o.u64(a_dset_hdr); o.u32(0); o.u32(0); o.u64(0); o.u64(0);
  o.pad_to(a_dset_hdr);
  // ---- dataset object header (v1)
  o.u8(1); o.u8(0); o.u16(4); o.u32(1); o.u32((uint32_t)dset_msgs_size); o.u32(0);
  o.raw(msgs.b.data(), msgs.size());
  o.pad_to(a_data);
  out.swap(o.b);
  return CX_OK;
}

int cx_write_h5sph(const std::vector<CxOutBuf> &buf, const std::string &filename)
{
  const uint64_t len = buf.size();
  std::vector<uint8_t> head;
  const int rc = cx_h5sph_header(len, head);
  if (rc != CX_OK) return rc;
  FILE *f = fopen(filename.c_str(), "wb");
  if (!f) return CX_NO_WRITE_FILE;
  bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
  if (len) ok = ok && fwrite(buf.data(), 64, (size_t)len, f) == (size_t)len;
  fclose(f);
  return ok ? CX_OK : CX_WRITE_FAIL;
}
