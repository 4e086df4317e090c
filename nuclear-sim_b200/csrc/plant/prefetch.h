// Generated per-struct prefetch overloads behind NPS_PREFETCH (see hd.h): device compilation only.
#pragma once
#include "hd.h"
#include "state.h"
#if defined(__CUDACC__)
namespace nps {
#include "live_gen.inc"
}  // namespace nps
#endif
