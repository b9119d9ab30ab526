// acq_geom.h -- sizes shared by the host API and the device FFT code.
#pragma once
#include <stddef.h>

namespace acq {

constexpr int kN = 16384;
constexpr int kSub = 4096;       // points per sub-FFT
constexpr int kThreads = 256;    // threads per FFT team

// shared-memory footprint (in float2 elements) of the pieces
constexpr int kT2Elems = 4 * 15 * 16;       // W1024^{(4c+k2)*n1}, n1 = 1..15
constexpr int kBaseElems = 4 * 256;         // W16384^{4t+k2}: stage-A twiddle bases (global memory)
constexpr int kS1Elems = 4096;              // exchange A->B


}  // namespace acq
