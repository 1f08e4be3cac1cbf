// device_once.cuh -- per-device one-time configuration of kernels (used by the launchers, not by device code).
#pragma once
#include <cuda_runtime.h>

// Function attributes (the opt-in to more than 48 KB of dynamic shared memory) belong to a device, not to the process: a
// C++ caller that drives several GPUs from one process (one sfe_ctx per GPU) needs them set on each.
// Returns true the first time it is called for the current device with this flag array.
struct PerDeviceOnce {
  unsigned char seen[64];
};
inline bool first_use_on_device(PerDeviceOnce& o) {
  int d = 0;
  cudaGetDevice(&d);
  d &= 63;
  if (o.seen[d]) return false;
  o.seen[d] = 1;
  return true;
}

