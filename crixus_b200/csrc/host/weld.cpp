// crixus_b200 -- host vertex weld: outcome-exact restatement of remove_duplicate_vertices,
// /root/reference/src/crixus.cpp:343-455 (std::map<mint3, list<intb>> replaced by one sort).
//
// What the reference's code amounts to (verified against the reference in tests/):
//   * corners are bucketed by cell ((int)(x/dr), (int)(y/dr), (int)(z/dr)) -- float divide, truncation (:354);
//   * cells are visited in lexicographic (x,y,z) order, corners of a cell in insertion (= corner index) order;
//   * a corner that is not yet claimed becomes the next vertex and claims every LATER corner of its cell whose squared
//     distance (float) is < 1e-10*dr*dr (double), including corners that were already claimed (:317-329);
//   * the 13 neighbour-cell searches run on a by-value copy of the map (:331), so the `found` flags they set are lost
//     and the ids they write are overwritten when the neighbour cell is visited later: no lasting effect.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../../include/crixus_b200.h"

namespace {
struct Key {
  int32_t cx, cy, cz;
  int64_t idx;
};
inline bool key_less(const Key &a, const Key &b)
{
  if (a.cx != b.cx) return a.cx < b.cx;
  if (a.cy != b.cy) return a.cy < b.cy;
  if (a.cz != b.cz) return a.cz < b.cz;
  return a.idx < b.idx;
}
}  // namespace

extern "C" int64_t cx_weld_host(const float *corners, int64_t ncorner, float dr, float *out_verts, int32_t *out_ids)
{
  std::vector<Key> keys((size_t)ncorner);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncorner; c++) {
    keys[c].cx = (int32_t)(corners[3 * c + 0] / dr);
    keys[c].cy = (int32_t)(corners[3 * c + 1] / dr);
    keys[c].cz = (int32_t)(corners[3 * c + 2] / dr);
    keys[c].idx = c;
  }
  std::sort(keys.begin(), keys.end(), key_less);
  std::vector<char> found((size_t)ncorner, 0);
  const double thr = 1e-10 * dr * dr;
  int64_t nvert = 0;
  for (int64_t a = 0; a < ncorner;) {
    int64_t b = a + 1;
    while (b < ncorner && keys[b].cx == keys[a].cx && keys[b].cy == keys[a].cy && keys[b].cz == keys[a].cz) b++;
    for (int64_t i = a; i < b; i++) {
      if (found[i]) continue;
      found[i] = 1;
      const float *pi = corners + 3 * keys[i].idx;
      out_verts[3 * nvert + 0] = pi[0];
      out_verts[3 * nvert + 1] = pi[1];
      out_verts[3 * nvert + 2] = pi[2];
      out_ids[keys[i].idx] = (int32_t)nvert;
      for (int64_t k = i + 1; k < b; k++) {
        const float *pk = corners + 3 * keys[k].idx;
        const float xij = pk[0] - pi[0], yij = pk[1] - pi[1], zij = pk[2] - pi[2];
        const float dist = xij * xij + yij * yij + zij * zij;
        if (dist < thr) {
          found[k] = 1;
          out_ids[keys[k].idx] = (int32_t)nvert;
        }
      }
      nvert++;
    }
    a = b;
  }
  return nvert;
}
