// ORACLE — TEST INFRASTRUCTURE ONLY.
// Thin C wrapper compiled against the REFERENCE's own vendored kd-tree where it lies:
//   /root/reference/mloam_loop/include/mloam_loop/scan_context/nanoflann.hpp  (nanoflann 1.3.2, STL-only;
//   same single-kd-tree family as the FLANN KDTreeSingleIndex behind pcl::KdTreeFLANN, leaf 15).
// Built by `make -C oracle ref` into oracle/_ref/libref_knn.so (git-ignored; travels to the GPU box as a
// binary).  Used (a) to validate orc::KdTree in tests and (b) as a kNN cross-check.  No reference source is
// copied into this repository: the header is included from the read-only mount at build time.
#include REF_NANOFLANN_HEADER
#include <cmath>
#include <cstddef>

namespace {
struct PC {
  const float *p;
  size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return p[idx * 4 + dim]; }
  template <class BBOX> bool kdtree_get_bbox(BBOX &) const { return false; }
};
typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, PC>, PC, 3> Tree;
}  // namespace

extern "C" {
// Persistent tree for the timed CPU arm: build once per setInputCloud, query per feature (what pcl::KdTreeFLANN does).
struct RefTree {
  PC pc;
  Tree tree;
  RefTree(const float *p, int m) : pc{p, (size_t)m}, tree(3, pc, nanoflann::KDTreeSingleIndexAdaptorParams(15)) { tree.buildIndex(); }
};
void *ref_tree_create(const float *map, int m) { return new RefTree(map, m); }
void ref_tree_destroy(void *t) { delete static_cast<RefTree *>(t); }
int ref_tree_knn(void *t, float qx, float qy, float qz, int k, int *idx, float *sqd) {
  size_t ids[64];
  float ds[64];
  if (k > 64) k = 64;
  nanoflann::KNNResultSet<float> rs(k);
  rs.init(ids, ds);
  const float qq[3] = {qx, qy, qz};
  static_cast<RefTree *>(t)->tree.findNeighbors(rs, qq, nanoflann::SearchParams());
  const int got = (int)rs.size();
  for (int j = 0; j < got; j++) idx[j] = (int)ids[j], sqd[j] = ds[j];
  return got;
}
// map: float32 [m,4]; q: float32 [nq,4]; idx/sqd: [nq,k]. Missing slots: idx -1 / +inf.
void ref_knn(const float *map, int m, const float *q, int nq, int k, int *idx, float *sqd) {
  PC pc{map, (size_t)m};
  Tree tree(3, pc, nanoflann::KDTreeSingleIndexAdaptorParams(15));
  tree.buildIndex();
  for (int i = 0; i < nq; i++) {
    size_t ids[64];
    float ds[64];
    nanoflann::KNNResultSet<float> rs(k);
    rs.init(ids, ds);
    float qq[3] = {q[i * 4], q[i * 4 + 1], q[i * 4 + 2]};
    tree.findNeighbors(rs, qq, nanoflann::SearchParams());
    int got = (int)rs.size();
    for (int j = 0; j < k; j++) {
      idx[(size_t)i * k + j] = j < got ? (int)ids[j] : -1;
      sqd[(size_t)i * k + j] = j < got ? ds[j] : INFINITY;
    }
  }
}
}
