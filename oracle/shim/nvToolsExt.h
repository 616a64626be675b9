// Shim: CUDA 12.9 ships only the header-only nvtx3. The reference includes
// <nvToolsExt.h> (Stixels.cu:28) but never calls into it.
#pragma once
#include <nvtx3/nvToolsExt.h>
