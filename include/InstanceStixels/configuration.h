// Drop-in replacement header: compile-time constants of the stixel library.
// Same names and values as InstanceStixels/include/InstanceStixels/configuration.h:29-34
// (they are part of the API surface: apps include "configuration.h" for pixel_t).
#ifndef ISX_DROPIN_CONFIGURATION_H_
#define ISX_DROPIN_CONFIGURATION_H_

#include <limits>

typedef float pixel_t;

#ifndef MAX_LOGPROB
#define MAX_LOGPROB (std::numeric_limits<float>::infinity())
#endif
constexpr int DOWNSAMPLE_FACTOR = 8;
constexpr int MAX_STIXELS_PER_COLUMN = 200;
constexpr int LOG_LUT_SIZE = 1000000;

#endif  // ISX_DROPIN_CONFIGURATION_H_
