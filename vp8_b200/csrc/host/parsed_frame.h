// Storage behind one vp8r_frame_desc: a single host blob
//     [ vp8r_mb_info x n_mb | payload blocks (32 B each) ... ]
// so that one host->device copy moves a whole frame.  The blob lives either on the heap or in
// CUDA pinned memory (vp8r_frame_create(pinned)).  An optional device-resident copy is attached
// by vp8r_frame_upload().
#ifndef VP8R_HOST_PARSED_FRAME_H_
#define VP8R_HOST_PARSED_FRAME_H_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "vp8r.h"

namespace vp8r {
// Implemented next to the CUDA runtime glue (rt/host_mem.cc): pinned uses cudaHostAlloc.
void *HostAlloc(size_t bytes, bool pinned);
void HostFree(void *p, bool pinned);
void DeviceFree(void *p, int device);
}  // namespace vp8r

struct vp8r_frame {
  vp8r_frame_hdr hdr{};
  bool pinned = false;
  uint8_t *blob = nullptr;  // host blob
  size_t blob_cap = 0;      // bytes
  size_t n_mb = 0;
  // device-resident copy (owned); valid when d_blob != nullptr
  void *d_blob = nullptr;
  // A resident frame whose parse is deferred is REWRITTEN by the parse kernels every time it is reconstructed: the
  // engine records here the event (and its own id) after which the previous reconstruction has read the records,
  // and lets the next parse of the same frame wait for it (replays with several batches in flight).
  void *busy_event = nullptr;
  unsigned long long busy_engine = 0;
  size_t d_bytes = 0;
  int d_device = -1;
  // what the engine needs on the host once the arrays live on the device (vp8r_frame_release_host):
  // the number of intra macroblocks per dependency level and the number of DCT partitions
  std::vector<uint32_t> level_counts;
  uint32_t n_token_parts = 0;

  size_t mb_bytes() const { return n_mb * sizeof(vp8r_mb_info); }
  size_t used_bytes() const { return mb_bytes() + size_t(hdr.n_payload_blocks) * 32; }
  vp8r_mb_info *mbs() { return reinterpret_cast<vp8r_mb_info *>(blob); }
  const vp8r_mb_info *mbs() const { return reinterpret_cast<const vp8r_mb_info *>(blob); }
  int16_t *payload() { return reinterpret_cast<int16_t *>(blob + mb_bytes()); }
  const int16_t *payload() const { return reinterpret_cast<const int16_t *>(blob + mb_bytes()); }

  // Makes room for `bytes` in total, keeping the first `keep` bytes.
  bool Reserve(size_t bytes, size_t keep) {
    if (bytes <= blob_cap) return true;
    size_t cap = blob_cap ? blob_cap : 4096;
    while (cap < bytes) cap += cap / 2 + 4096;
    uint8_t *nb = static_cast<uint8_t *>(vp8r::HostAlloc(cap, pinned));
    if (!nb) return false;
    if (blob && keep) std::memcpy(nb, blob, keep);
    if (blob) vp8r::HostFree(blob, pinned);
    blob = nb;
    blob_cap = cap;
    return true;
  }

  void DropDeviceCopy() {
    if (d_blob) vp8r::DeviceFree(d_blob, d_device);
    d_blob = nullptr;
    d_bytes = 0;
    d_device = -1;
  }

  ~vp8r_frame() {
    DropDeviceCopy();
    if (blob) vp8r::HostFree(blob, pinned);
  }
};

#endif  // VP8R_HOST_PARSED_FRAME_H_
