// microclimf_b200 — host-side services shared between translation units (not part of the C ABI).
#pragma once
#include <cstddef>

namespace mcf {

struct HostXfer {
    void* dst;
    const void* src;
    size_t bytes;
};

// Checks the device (sm_100a) and sets the default memory pool's release threshold so that stream-ordered scratch is
// kept between calls.  Returns an MCF_* code.
int host_prepare_device(char* err, size_t errlen);

// Copies between host and device buffers through the pool of copy threads and pinned slots in mcf_api.cu (pageable
// host memory), or directly (pinned host memory).  Blocks until every copy has completed.  The device side must be
// idle or ordered by the caller: the copies run on the pool's own non-blocking streams.
int host_transfer(const HostXfer* jobs, int n, bool to_device, char* err, size_t errlen);

} // namespace mcf
