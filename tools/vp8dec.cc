// vp8dec -- `./decode in.ivf out.yuv` of the reference (src/decode.cc:13-79) on top of libvp8r:
// IVF demux on the host, frame reconstruction on the GPU, cropped I420 frames appended to the
// output file.  Exits non-zero with a message instead of the reference's assert()/exit(1)/throw.
//
//   vp8dec [--device N] [--md5] in.ivf out.yuv
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vp8r.h"

static uint32_t Le32(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

int main(int argc, char **argv) {
  int device = 0;
  std::vector<std::string> pos;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
    else pos.push_back(a);
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
