// sim.hpp -- shared by sim_cuda.cpp and sim_kernels.cpp (TEST INFRASTRUCTURE; see cuda_runtime.h here).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <functional>
#include <vector>

namespace sim {
[[noreturn]] void die(const char *what);
bool device_range(const void *p, size_t bytes);        // [p, p + bytes) lies inside one device allocation
void enqueue(cudaStream_t st, std::function<void()> fn);   // one operation of stream st: run now, or record into the capture in progress
// Called from inside an operation: it touches [p, p + bytes) of device memory.  Two operations on different streams
// that touch the same bytes, at least one writing, with no event / synchronisation path between them, are a race.
void access(const void *p, size_t bytes, bool write, const char *what);
bool capturing();
void violation(const char *what);                      // a kernel contract was broken: sticky, reported by every later call
cudaError_t status();
}  // namespace sim
