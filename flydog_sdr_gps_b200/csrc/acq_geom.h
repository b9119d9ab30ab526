// acq_geom.h -- sizes shared by the host API and the device FFT code.
#pragma once
#include <stddef.h>

namespace acq {

constexpr int kN = 16384;
constexpr int kSub = 4096;       // points per sub-FFT
constexpr int kThreads = 256;    // threads per FFT team
constexpr int kS2Stride = 272;   // padded row stride of the second exchange buffer (16*17)

// shared-memory footprint (in float2 elements) of the pieces
constexpr int kT1Elems = 15 * 256;          // W4096^{t*n0}, n0 = 1..15
constexpr int kT2Elems = 4 * 15 * 16;       // W1024^{(4c+k2)*n1}, n1 = 1..15
constexpr int kS1Elems = 4096;              // exchange A->B
constexpr int kS2Elems = 16 * kS2Stride;    // exchange B->C (padded)


}  // namespace acq
