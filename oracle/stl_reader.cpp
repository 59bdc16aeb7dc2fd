// TEST INFRASTRUCTURE -- not part of the product.
// Restatement of the reference's binary-STL ingest loop with the SAME iostream calls in the same order
// (/root/reference/src/crixus.cu:126-221: is-binary check :137-153, 20 header floats :160-164, facet count :166,
// per facet 12 float reads + one short read :188-213, count check :218-221), so that the behaviour on malformed files --
// which follows from how std::istream::read treats short reads -- is the reference's by construction:
//   * a read that hits the end stores the bytes it got, sets eofbit|failbit; every later read stores nothing;
//   * m_v_floats keeps the previous facet's values where nothing new was stored;
//   * the iteration in which the end is hit still appends a facet and increments `through`;
//   * READ_ERROR only if num_of_facets != number of appended facets.
// Used by tests/test_frontend.py to pin cx_stl_read (the product's bulk reader) on truncated files.
#include <cstring>
#include <fstream>
#include <vector>

extern "C" int orc_stl_read(const char *filename, float *corners, float *normals, long capacity, long *nfacets)
{
  using namespace std;
  *nfacets = 0;
  ifstream stl_file(filename, ios::in);
  if (!stl_file.is_open()) return -3;                       // FILE_NOT_OPEN
  bool issolid = true;
  char header[6] = "solid";
  for (int i = 0; i < 5; i++) {
    char dum;
    stl_file.read((char *)&dum, sizeof(char));
    if (dum != header[i]) { issolid = false; break; }
  }
  stl_file.close();
  if (issolid) return -1;                                   // STL_NOT_BINARY
  stl_file.open(filename, ios::in | ios::binary);
  for (int i = 0; i < 20; i++) {
    float dum;
    stl_file.read((char *)&dum, sizeof(float));
  }
  int num_of_facets = 0;                                    // (uninitialised in the reference; a file without a count is UB there)
  stl_file.read((char *)&num_of_facets, sizeof(int));
  float m_v_floats[12];
  memset(m_v_floats, 0, sizeof(m_v_floats));                // (uninitialised in the reference)
  short attribute;
  unsigned int through = 0;
  vector<float> n, c;
  while ((through < (unsigned int)num_of_facets) & (!stl_file.eof())) {
    for (int i = 0; i < 12; i++) stl_file.read((char *)&m_v_floats[i], sizeof(float));
    for (int i = 0; i < 3; i++) n.push_back(m_v_floats[i]);
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) c.push_back(m_v_floats[i + 3 * (j + 1)]);
    stl_file.read((char *)&attribute, sizeof(short));
    through++;
  }
  stl_file.close();
  *nfacets = (long)(n.size() / 3);
  if ((size_t)num_of_facets != n.size() / 3) return -5;     // READ_ERROR
  if ((long)(n.size() / 3) > capacity) return -18;
  memcpy(normals, n.data(), n.size() * sizeof(float));
  memcpy(corners, c.data(), c.size() * sizeof(float));
  return 0;
}
