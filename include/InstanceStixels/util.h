// Drop-in replacement header for the helpers apps may pick up through Stixels.hpp
// (InstanceStixels/include/InstanceStixels/util.h:24-46).
#ifndef ISX_DROPIN_UTIL_H_
#define ISX_DROPIN_UTIL_H_

#include <cstdlib>
#include <iostream>

constexpr int WARP_SIZE = 32;

#ifdef __CUDACC__
#define ISX_HD __host__ __device__
#else
#define ISX_HD
#endif
ISX_HD inline int divUp(int total, int grain) { return (total + grain - 1) / grain; }

#if defined(__CUDACC__) || defined(CUDART_VERSION)
// print + exit(1) like the reference's CUDA_CHECK_RETURN
#define CUDA_CHECK_RETURN(value)                                                                 \
    do {                                                                                         \
        cudaError_t isx_e_ = (value);                                                            \
        if (isx_e_ != cudaSuccess) {                                                             \
            std::cerr << #value << " returned " << cudaGetErrorString(isx_e_) << "(" << isx_e_   \
                      << ") at " << __FILE__ << ":" << __LINE__ << std::endl;                    \
            std::exit(1);                                                                        \
        }                                                                                        \
    } while (0)
#endif

#endif  // ISX_DROPIN_UTIL_H_
