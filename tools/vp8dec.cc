// vp8dec -- `./decode in.ivf out.yuv` of the reference (src/decode.cc:13-79) on top of libvp8r:
// IVF demux on the host, frame reconstruction on the GPU, cropped I420 frames appended to the
// output file.  Exits non-zero with a message instead of the reference's assert()/exit(1)/throw.
//
//   vp8dec [--device N] in.ivf out.yuv
//   vp8dec [--device N] [--host-parse] --batch out_dir a.ivf b.ivf ...
//
// --batch: every input is a stream of its own, all decoded in lock-step (one frame of each per time step) with the
// throughput form of the C ABI: frame headers parsed on host threads, macroblock headers and DCT tokens on the GPU
// (--host-parse: everything on the host), reconstruction batched over the streams, shown frames cropped and packed
// on the device and read back with one copy per time step; four time steps in flight.  Writes out_dir/<name>.yuv.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vp8r.h"

#include <memory>
#include <thread>

static uint32_t Le32(const uint8_t *p);

namespace {

struct IvfStream {
  std::string name;
  std::vector<uint8_t> file;
  std::vector<std::pair<size_t, size_t>> frames;  // offset, size of every frame payload
  FILE *out = nullptr;
};

bool LoadIvf(const std::string &path, IvfStream *s) {
  FILE *in = std::fopen(path.c_str(), "rb");
  if (!in) return false;
  uint8_t buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, in)) > 0) s->file.insert(s->file.end(), buf, buf + n);
  std::fclose(in);
  if (s->file.size() < 32 || std::memcmp(s->file.data(), "DKIF", 4) != 0 || std::memcmp(s->file.data() + 8, "VP80", 4) != 0) return false;
  size_t at = size_t(s->file[6]) | (size_t(s->file[7]) << 8);
  const uint32_t n_frames = Le32(s->file.data() + 24);
  for (uint32_t k = 0; k < n_frames && at + 12 <= s->file.size(); ++k) {
    const size_t size = Le32(s->file.data() + at);
    at += 12;
    if (at + size > s->file.size()) return false;
    s->frames.emplace_back(at, size);
    at += size;
  }
  const size_t slash = path.find_last_of('/');
  s->name = path.substr(slash == std::string::npos ? 0 : slash + 1);
  return true;
}

#define TRY(call)                                                        \
  do {                                                                   \
    if ((call) != VP8R_OK) {                                             \
      std::fprintf(stderr, "vp8dec: %s: %s\n", #call, vp8r_last_error()); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

// The lock-step loop of INTEGRATION.md ("Throughput-oriented use"), as a host program writes it.
int DecodeBatch(int device, bool host_parse, const std::string &out_dir, const std::vector<std::string> &inputs) {
  constexpr int kDepth = 4;
  const int n = int(inputs.size());
  std::vector<IvfStream> in(n);
  size_t steps = 0;
  for (int i = 0; i < n; ++i) {
    if (!LoadIvf(inputs[i], &in[i])) {
      std::fprintf(stderr, "vp8dec: %s is not a readable VP8 IVF file\n", inputs[i].c_str());
      return 1;
    }
    steps = std::max(steps, in[i].frames.size());
    in[i].out = std::fopen((out_dir + "/" + in[i].name + ".yuv").c_str(), "wb");
    if (!in[i].out) {
      std::fprintf(stderr, "vp8dec: cannot create %s/%s.yuv\n", out_dir.c_str(), in[i].name.c_str());
      return 1;
    }
  }
  vp8r_engine *eng = nullptr;
  TRY(vp8r_engine_create(device, nullptr, &eng));
  std::vector<vp8r_parser *> parsers(n);
  std::vector<vp8r_stream *> streams(n);
  std::vector<vp8r_frame *> slots(size_t(kDepth) * n);
  for (int i = 0; i < n; ++i) {
    parsers[i] = vp8r_parser_create();
    if (!host_parse) vp8r_parser_set_defer_modes(parsers[i], 1);
    TRY(vp8r_stream_open(eng, &streams[i]));
  }
  for (auto &f : slots) f = vp8r_frame_create(/*pinned=*/1);

  // what a time step left in the ring: which streams, how many bytes each
  struct Pending {
    uint64_t ticket = 0;
    std::vector<int> shown;
    std::vector<size_t> bytes;
    size_t stride = 0;
    bool live = false;
  } pending[kDepth];
  std::vector<uint8_t> ring[kDepth];
  size_t stride = 0;
  auto drain = [&](Pending &p, const std::vector<uint8_t> &buf) -> int {
    if (!p.live) return 0;
    TRY(vp8r_engine_wait(eng, p.ticket));
    for (size_t k = 0; k < p.shown.size(); ++k) std::fwrite(buf.data() + k * p.stride, 1, p.bytes[k], in[p.shown[k]].out);
    p.live = false;
    return 0;
  };

  for (size_t t = 0; t < steps; ++t) {
    const int slot = int(t % kDepth);
    if (drain(pending[slot], ring[slot])) return 1;  // frames and ring entry of step t - kDepth are free again
    std::vector<int> live;
    std::vector<vp8r_parser *> lp;
    std::vector<vp8r_stream *> ls;
    std::vector<vp8r_frame *> lf;
    std::vector<const uint8_t *> data;
    std::vector<size_t> sizes;
    for (int i = 0; i < n; ++i)
      if (t < in[i].frames.size()) {
        live.push_back(i);
        lp.push_back(parsers[i]);
        ls.push_back(streams[i]);
        lf.push_back(slots[size_t(slot) * n + i]);
        data.push_back(in[i].file.data() + in[i].frames[t].first);
        sizes.push_back(in[i].frames[t].second);
      }
    const int m = int(live.size());
    TRY(vp8r_parse_batch(m, lp.data(), data.data(), sizes.data(), lf.data(), int(std::max(1u, std::thread::hardware_concurrency())), nullptr));
    TRY(vp8r_reconstruct_batch(eng, m, ls.data(), lf.data()));
    Pending &p = pending[slot];
    p.shown.clear();
    p.bytes.clear();
    std::vector<vp8r_stream *> shown_streams;
    for (int k = 0; k < m; ++k) {
      vp8r_frame_desc d;
      TRY(vp8r_frame_get_desc(lf[k], &d));
      if (!d.hdr.show_frame) continue;  // hidden frames are decoded, not written (src/decode.cc:76)
      p.shown.push_back(live[k]);
      p.bytes.push_back(vp8r_stream_frame_bytes(ls[k]));
      stride = std::max(stride, p.bytes.back());
      shown_streams.push_back(ls[k]);
    }
    if (!shown_streams.empty()) {
      if (ring[slot].size() < stride * size_t(n)) {
        // (a production caller allocates pinned memory once; a pageable ring keeps this tool free of CUDA calls)
        for (int q = 0; q < kDepth; ++q)
          if (q != slot && drain(pending[q], ring[q])) return 1;
        for (auto &r : ring) r.resize(stride * size_t(n));
      }
      p.stride = stride;
      TRY(vp8r_read_batch_packed(eng, int(shown_streams.size()), shown_streams.data(), ring[slot].data(), stride, /*async=*/1));
    }
    TRY(vp8r_engine_fence(eng, &p.ticket));
    p.live = true;
  }
  for (size_t t = steps; t < steps + kDepth; ++t)
    if (drain(pending[t % kDepth], ring[t % kDepth])) return 1;
  TRY(vp8r_engine_sync(eng));  // reports a partition the device-side parse read past
  for (auto &f : slots) vp8r_frame_destroy(f);
  for (int i = 0; i < n; ++i) {
    std::fclose(in[i].out);
    vp8r_stream_close(streams[i]);
    vp8r_parser_destroy(parsers[i]);
  }
  vp8r_engine_destroy(eng);
  return 0;
}

}  // namespace

static uint32_t Le32(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

int main(int argc, char **argv) {
  int device = 0;
  bool batch = false, host_parse = false;
  std::vector<std::string> pos;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
    else if (a == "--batch") batch = true;
    else if (a == "--host-parse") host_parse = true;
    else pos.push_back(a);
  }
  if (batch) {
    if (pos.size() < 2) {
      std::fprintf(stderr, "[Usage] vp8dec [--device N] [--host-parse] --batch [output directory] [inputs...]\n");
      return 1;
    }
    return DecodeBatch(device, host_parse, pos[0], std::vector<std::string>(pos.begin() + 1, pos.end()));
  }
  if (pos.size() != 2) {
    std::fprintf(stderr, "[Usage] vp8dec [--device N] [input] [output]\n");
    return 1;
  }
  FILE *in = std::fopen(pos[0].c_str(), "rb");
  if (!in) {
    std::fprintf(stderr, "cannot open %s\n", pos[0].c_str());
    return 1;
  }
  std::vector<uint8_t> file;
  uint8_t buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, in)) > 0) file.insert(file.end(), buf, buf + n);
  std::fclose(in);
  // IVF header: 'DKIF', version 0, header length 32, 'VP80' (src/decode.cc:25-38)
  if (file.size() < 32 || std::memcmp(file.data(), "DKIF", 4) != 0 || std::memcmp(file.data() + 8, "VP80", 4) != 0) {
    std::fprintf(stderr, "%s is not a VP8 IVF file\n", pos[0].c_str());
    return 1;
  }
  const size_t hdr_len = size_t(file[6]) | (size_t(file[7]) << 8);
  const uint32_t n_frames = Le32(file.data() + 24);

  vp8r_engine *eng = nullptr;
  vp8r_stream *st = nullptr;
  if (vp8r_engine_create(device, nullptr, &eng) != VP8R_OK || vp8r_stream_open(eng, &st) != VP8R_OK) {
    std::fprintf(stderr, "vp8dec: %s\n", vp8r_last_error());
    return 1;
  }
  FILE *out = std::fopen(pos[1].c_str(), "wb");
  if (!out) {
    std::fprintf(stderr, "cannot create %s\n", pos[1].c_str());
    return 1;
  }
  std::vector<uint8_t> i420;
  size_t at = hdr_len;
  int rc = 0;
  for (uint32_t k = 0; k < n_frames && at + 12 <= file.size(); ++k) {
    const size_t size = Le32(file.data() + at);
    at += 12;
    if (at + size > file.size()) {
      std::fprintf(stderr, "vp8dec: frame %u exceeds the file\n", k);
      rc = 1;
      break;
    }
    int shown = 0;
    if (vp8r_stream_decode(st, file.data() + at, size, &shown) != VP8R_OK) {
      std::fprintf(stderr, "vp8dec: frame %u: %s\n", k, vp8r_last_error());
      rc = 1;
      break;
    }
    at += size;
    if (shown) {  // hidden frames are decoded, not written (src/decode.cc:76)
      i420.resize(vp8r_stream_frame_bytes(st));
      if (vp8r_stream_read_frame(st, i420.data(), i420.size()) != VP8R_OK) {
        std::fprintf(stderr, "vp8dec: %s\n", vp8r_last_error());
        rc = 1;
        break;
      }
      std::fwrite(i420.data(), 1, i420.size(), out);
    }
  }
  std::fclose(out);
  vp8r_stream_close(st);
  vp8r_engine_destroy(eng);
  return rc;
}
