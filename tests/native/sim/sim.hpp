// sim.hpp -- shared by sim_cuda.cpp and sim_kernels.cpp (TEST INFRASTRUCTURE; see cuda_runtime.h here).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <functional>
#include <vector>

namespace sim {
[[noreturn]] void die(const char *what);
bool device_range(const void *p, size_t bytes);        // [p, p + bytes) lies inside one device allocation
void enqueue(std::function<void()> fn);                // run now, or record into the stream capture in progress
bool capturing();
void violation(const char *what);                      // a kernel contract was broken: sticky, reported by every later call
cudaError_t status();
}  // namespace sim
