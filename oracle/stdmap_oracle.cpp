// TEST INFRASTRUCTURE: the iteration order of the reference's hash containers.
// vo::Map::map_points_ (include/my_slam/vo/map.h:21) and Frame::inliers_to_mappt_connections_
// (include/my_slam/vo/frame.h:47) are std::unordered_map<int, ...>; the reference walks them in container order
// (src/vo/vo.cpp:24, :435, :493), so the order of the PnP candidates and of the BA edges is whatever libstdc++'s
// hashtable yields.  This file does not restate that: it wraps the real container behind a C interface for the
// Python oracle (oracle/vo_pipeline_oracle.py).
#include <stdint.h>
#include <unordered_map>

using IntMap = std::unordered_map<int, int>;

extern "C" {
void *orc_umap_new() { return new IntMap(); }
void orc_umap_free(void *h) { delete (IntMap *)h; }
int orc_umap_size(void *h) { return (int)((IntMap *)h)->size(); }
// insert(): keeps an existing entry (like unordered_map::insert); returns 1 if inserted
int orc_umap_insert(void *h, int key) { return ((IntMap *)h)->insert({key, 0}).second ? 1 : 0; }
// operator[]: creates the entry if missing
void orc_umap_touch(void *h, int key) { (*(IntMap *)h)[key] = 0; }
int orc_umap_erase(void *h, int key) { return (int)((IntMap *)h)->erase(key); }
int orc_umap_contains(void *h, int key) { return ((IntMap *)h)->count(key) ? 1 : 0; }
int orc_umap_keys(void *h, int32_t *out, int cap) {
  int n = 0;
  for (auto &kv : *(IntMap *)h) { if (n < cap) out[n] = kv.first; ++n; }
  return n;
}
}
