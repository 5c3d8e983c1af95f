// Heap-only HostAlloc for CPU-side tools that link the parser without the CUDA runtime
// (tools/, oracle checks).  The product library uses rt/host_mem.cu instead.
#include <cstdlib>
#include "../host/parsed_frame.h"
namespace vp8r {
void *HostAlloc(size_t bytes, bool) { return std::malloc(bytes); }
void HostFree(void *p, bool) { std::free(p); }
void DeviceFree(void *, int) {}
}  // namespace vp8r
