// Shim declaration for the cuML handle type the reference constructs at
// Stixels.cu:652. The real library (tomsal/cuml@dbscan-sizefilter) is not
// vendored in the reference tree; oracle/dbscan_standin.cpp provides the
// one symbol the reference needs.
#pragma once
namespace ML {
class cumlHandle {
 public:
  cumlHandle() {}
  ~cumlHandle() {}
};
}  // namespace ML
