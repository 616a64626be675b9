// Shim declaration of the cuML-fork entry point called at Stixels.cu:660-666.
#pragma once
#include <cstddef>
#include <cuml/cuml.hpp>
namespace ML {
// input: n_rows x n_cols row-major device floats; labels: n_rows device ints;
// core_candidates: n_rows device bools (size filter of the fork).
void dbscanFit(const cumlHandle &handle, float *input, int n_rows, int n_cols,
               float eps, int min_pts, int *labels,
               size_t max_bytes_per_batch, bool verbose,
               bool *core_candidates);
}  // namespace ML
