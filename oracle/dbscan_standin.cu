// TEST INFRASTRUCTURE -- not part of the product.
//
// Provides the single symbol the reference library leaves unresolved:
// ML::dbscanFit (cuML fork, call site InstanceStixels/src/Stixels.cu:660-666).
// Device buffers in/out like cuML; the clustering itself runs on the host with
// the definition in dbscan_def.h.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <vector>

#include <cuml/cluster/dbscan.hpp>

#include "dbscan_def.h"

// Host seconds spent inside the stand-in (its three blocking copies + the O(n^2) host clustering) since the last
// reset: bench.py --impl reference reports them, so that the reference's time can be read with and without the part
// that is NOT the reference's own code (the cuML fork it links is not available).
static double g_standin_seconds = 0.0;
extern "C" double ref_dbscan_standin_seconds(int reset) {
  const double s = g_standin_seconds;
  if (reset) g_standin_seconds = 0.0;
  return s;
}

namespace ML {
void dbscanFit(const cumlHandle &, float *input, int n_rows, int n_cols, float eps, int min_pts, int *labels,
               size_t, bool, bool *core_candidates) {
  if (n_rows <= 0) return;
  const auto t0 = std::chrono::steady_clock::now();
  if (n_cols != 2) {
    std::fprintf(stderr, "dbscan stand-in: n_cols must be 2\n");
    return;
  }
  std::vector<float> xy((size_t)n_rows * 2);
  std::vector<uint8_t> cand(n_rows);
  std::vector<int> lab(n_rows);
  static_assert(sizeof(bool) == 1, "bool size");
  cudaMemcpy(xy.data(), input, xy.size() * sizeof(float), cudaMemcpyDeviceToHost);
  cudaMemcpy(cand.data(), core_candidates, cand.size(), cudaMemcpyDeviceToHost);
  isx_oracle::dbscan_sizefilter(xy.data(), n_rows, eps, min_pts, cand.data(), lab.data());
  cudaMemcpy(labels, lab.data(), lab.size() * sizeof(int), cudaMemcpyHostToDevice);
  g_standin_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
}  // namespace ML
