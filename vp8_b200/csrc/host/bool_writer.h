// Boolean entropy ENCODER of RFC 6386 section 7.3 (the reference's own is unfinished: src/bool_encoder.h:72-73 has
// empty WriteByte / AddOne).  Used by the key-frame writer (host/frame_writer.cc) and by tools/vp8synth.cc.
#ifndef VP8R_HOST_BOOL_WRITER_H_
#define VP8R_HOST_BOOL_WRITER_H_

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace vp8r {

// RFC 6386 section 7.3 arithmetic coder, carry propagated into already written bytes.
class BoolWriter {
 public:
  void Put(int prob, int bit) {
    uint32_t split = 1 + (((range_ - 1) * uint32_t(prob)) >> 8);
    if (bit) {
      low_ += split;
      range_ -= split;
    } else {
      range_ = split;
    }
    while (range_ < 128) {
      range_ <<= 1;
      if (low_ & 0x80000000u) Carry();
      low_ <<= 1;
      if (--pending_ == 0) {
        bytes_.push_back(uint8_t(low_ >> 24));
        low_ &= 0x00FFFFFFu;
        pending_ = 8;
      }
    }
  }
  void Lit(int bits, uint32_t v) {
    for (int i = bits - 1; i >= 0; --i) Put(128, int((v >> i) & 1));
  }
  void SignedLit(int bits, int v) {  // magnitude then sign
    Lit(bits, uint32_t(v < 0 ? -v : v));
    Put(128, v < 0);
  }
  // Writes `value` with a RFC-style tree (positive = node index, <= 0 = negated leaf).
  void Tree(const int8_t *tree, int n_entries, const uint8_t *probs, int value) {
    int path[32], bits[32], depth = 0;
    bool ok = Find(tree, n_entries, 0, value, path, bits, &depth);
    if (!ok) {
      std::fprintf(stderr, "vp8r: value %d not in tree\n", value);
      std::exit(2);
    }
    for (int i = 0; i < depth; ++i) Put(probs[path[i] >> 1], bits[i]);
  }
  std::vector<uint8_t> Finish() {
    int c = pending_;
    uint32_t v = low_;
    if (v & (1u << (32 - c))) Carry();
    v <<= c & 7;
    c >>= 3;
    while (--c >= 0) v <<= 8;
    for (int i = 0; i < 4; ++i) {
      bytes_.push_back(uint8_t(v >> 24));
      v <<= 8;
    }
    return bytes_;
  }

 private:
  static bool Find(const int8_t *tree, int n, int node, int value, int *path, int *bits, int *depth) {
    for (int b = 0; b < 2; ++b) {
      int e = tree[node + b];
      path[*depth] = node;
      bits[*depth] = b;
      ++*depth;
      if (e <= 0) {
        if (-e == value) return true;
      } else if (e < n && Find(tree, n, e, value, path, bits, depth)) {
        return true;
      }
      --*depth;
    }
    return false;
  }
  void Carry() {
    size_t i = bytes_.size();
    while (i > 0 && bytes_[i - 1] == 255) bytes_[--i] = 0;
    if (i > 0) ++bytes_[i - 1];
  }
  std::vector<uint8_t> bytes_;
  uint32_t range_ = 255, low_ = 0;
  int pending_ = 24;
};

}  // namespace vp8r

#endif  // VP8R_HOST_BOOL_WRITER_H_
