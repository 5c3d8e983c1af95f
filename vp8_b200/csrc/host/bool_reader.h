// Boolean entropy decoder of RFC 6386 section 7, windowed form.
//
// The reference (src/bool_decoder.cc:13-41) keeps a 16-bit window and renormalises one bit at a
// time, pulling a byte every 8 shifts.  This reader produces the identical bit sequence with a
// 64-bit left-aligned window that is refilled up to 7 bytes at a time and renormalised with one
// count-leading-zeros per symbol.  `shifts_` counts renormalisation shifts so that the number of
// bytes the reference's reader would have pulled (2 + shifts/8) can be compared with the
// partition length: the reference throws std::out_of_range (src/utils.h:62-66) when it runs
// past the end; we report VP8R_ERR_TRUNCATED for the same streams.
#ifndef VP8R_HOST_BOOL_READER_H_
#define VP8R_HOST_BOOL_READER_H_

#include <cstddef>
#include <cstdint>

namespace vp8r {

class BoolReader {
 public:
  BoolReader() = default;

  void Init(const uint8_t *data, size_t size) {
    cur_ = data;
    end_ = data + size;
    size_ = size;
    window_ = 0;
    avail_ = 0;
    range_ = 255;
    shifts_ = 0;
    used_ = false;
    Refill();
  }

  // One boolean with probability prob/256 of being 0.
  inline int Bit(int prob) {
    uint32_t split = 1 + (((range_ - 1) * uint32_t(prob)) >> 8);
    used_ = true;
    if (avail_ < 8) Refill();
    uint64_t big = uint64_t(split) << 56;
    int bit;
    if (window_ >= big) {
      window_ -= big;
      range_ -= split;
      bit = 1;
    } else {
      range_ = split;
      bit = 0;
    }
    // range_ is in [1,255]; bring it back to [128,255].
    int sh = __builtin_clz(range_) - 24;
    range_ <<= sh;
    window_ <<= sh;
    avail_ -= sh;
    shifts_ += uint32_t(sh);
    return bit;
  }

  inline int Bit128() { return Bit(128); }

  inline uint32_t Literal(int n) {
    uint32_t v = 0;
    while (n-- > 0) v = (v << 1) | uint32_t(Bit(128));
    return v;
  }

  // Walks a RFC 6386 style tree: positive entries are next-node indices, entries <= 0 are
  // negated leaf values; node i uses probs[i >> 1].
  inline int Tree(const int8_t *tree, const uint8_t *probs) {
    int i = 0;
    do {
      i = tree[i + Bit(probs[i >> 1])];
    } while (i > 0);
    return -i;
  }

  // Bytes the reference's byte-at-a-time reader would have consumed so far.
  size_t BytesConsumed() const { return 2 + size_t(shifts_ >> 3); }
  // The reference initialises lazily (src/bool_decoder.cc:14-17): an unread partition never throws.
  bool Overrun() const { return used_ && BytesConsumed() > size_; }
  size_t size() const { return size_; }

 private:
  inline void Refill() {
    // Keep the window's top bits valid: append whole bytes below the `avail_` valid bits.
    while (avail_ <= 56) {
      uint64_t b = (cur_ < end_) ? uint64_t(*cur_) : 0;  // zeros past the end; Overrun() tells
      if (cur_ < end_) ++cur_;
      window_ |= b << (56 - avail_);
      avail_ += 8;
    }
  }

  const uint8_t *cur_ = nullptr;
  const uint8_t *end_ = nullptr;
  size_t size_ = 0;
  uint64_t window_ = 0;  // next bits of the stream, left aligned
  int avail_ = 0;        // number of valid bits in window_
  uint32_t range_ = 255;
  uint32_t shifts_ = 0;
  bool used_ = false;
};

}  // namespace vp8r

#endif  // VP8R_HOST_BOOL_READER_H_
