// TEST INFRASTRUCTURE -- not part of the product.
//
// Working definition of the instance grouping step.  The reference calls
// ML::dbscanFit from the non-vendored cuML fork tomsal/cuml, branch
// dbscan-sizefilter (no commit pin; singularity_recipe:108-131; call site
// InstanceStixels/src/Stixels.cu:660-666).  Its source is not available, so
// parity with cuML itself is UNPINNED.  This header states the semantics both
// the oracle (linked into oracle/_ref as ML::dbscanFit) and the CUDA product
// implement; it follows the classic DBSCAN definition with the fork's
// `core_candidates` mask, cross-checked against sklearn.cluster.DBSCAN in
// tests/ (for all-true masks) and against the in-tree legacy statement
// tools/visualization/clustering_visualization.py:894-979.
//
//   d2(i,j)   = (xi-xj)*(xi-xj) + (yi-yj)*(yi-yj)   fp32, no FMA contraction
//   nbr(i,j)  = d2(i,j) <= eps*eps                  (self included)
//   core(i)   = core_candidate[i] && |{j : nbr(i,j)}| >= min_pts
//   clusters  = connected components of core points under nbr
//   cluster id= rank of the component by its lowest-index core point, 0..k-1
//   border i  = non-core with >=1 core neighbour -> cluster of the LOWEST-INDEX
//               core neighbour;  everything else -> -1 (noise)
#pragma once
#include <cstdint>
#include <vector>

namespace isx_oracle {

inline float dbscan_d2(float xi, float yi, float xj, float yj) {
  const float dx = xi - xj, dy = yi - yj;
  volatile float a = dx * dx;  // volatile: forbid contraction into fma
  volatile float b = dy * dy;
  return a + b;
}

inline void dbscan_sizefilter(const float *xy, int n, float eps, int min_pts, const uint8_t *core_candidate,
                              int *labels) {
  const float eps2 = eps * eps;
  std::vector<uint8_t> core(n, 0);
  for (int i = 0; i < n; i++) {
    int deg = 0;
    for (int j = 0; j < n; j++) deg += dbscan_d2(xy[2 * i], xy[2 * i + 1], xy[2 * j], xy[2 * j + 1]) <= eps2;
    core[i] = core_candidate[i] && deg >= min_pts;
  }
  for (int i = 0; i < n; i++) labels[i] = -1;
  int next = 0;
  std::vector<int> stack;
  for (int s = 0; s < n; s++) {
    if (!core[s] || labels[s] != -1) continue;
    labels[s] = next;
    stack.push_back(s);
    while (!stack.empty()) {
      const int i = stack.back();
      stack.pop_back();
      for (int j = 0; j < n; j++) {
        if (core[j] && labels[j] == -1 &&
            dbscan_d2(xy[2 * i], xy[2 * i + 1], xy[2 * j], xy[2 * j + 1]) <= eps2) {
          labels[j] = next;
          stack.push_back(j);
        }
      }
    }
    next++;
  }
  for (int i = 0; i < n; i++) {
    if (core[i]) continue;
    for (int j = 0; j < n; j++) {
      if (core[j] && dbscan_d2(xy[2 * i], xy[2 * i + 1], xy[2 * j], xy[2 * j + 1]) <= eps2) {
        labels[i] = labels[j];
        break;
      }
    }
  }
}

}  // namespace isx_oracle
