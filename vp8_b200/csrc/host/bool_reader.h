// Boolean entropy decoder of RFC 6386 section 7, windowed form.
//
// The reference (src/bool_decoder.cc:13-41) keeps a 16-bit window and renormalises one bit at a
// time, pulling a byte every 8 shifts.  This reader produces the identical bit sequence with a
// 64-bit left-aligned window: refills append up to 7 whole bytes with one big-endian 64-bit load,
// the interval update is branch-free and renormalisation is one count-leading-zeros.
//
// Over-read detection: the reference's reader has consumed 2 + shifts/8 bytes after `shifts`
// renormalisation shifts and throws std::out_of_range (src/utils.h:62-66) when that exceeds the
// partition.  Here shifts == 8 * bytes_loaded - bits_available at any time, so nothing is counted
// per symbol; Overrun() reports VP8R_ERR_TRUNCATED for the streams the reference would reject.
#ifndef VP8R_HOST_BOOL_READER_H_
#define VP8R_HOST_BOOL_READER_H_

#include <cstddef>
#include <cstdint>
#include <cstring>

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
    loaded_ = 0;
    range_ = 255;
    used_ = false;
    Refill();
  }

  // The reference initialises lazily (src/bool_decoder.cc:14-17): an unread partition never
  // throws.  Callers mark a reader before its first symbol.
  void MarkUsed() { used_ = true; }

  // One boolean with probability prob/256 of being 0.
  inline int Bit(int prob) {
    const uint32_t split = 1 + (((range_ - 1) * uint32_t(prob)) >> 8);
    if (__builtin_expect(avail_ < 8, 0)) Refill();
    const uint64_t big = uint64_t(split) << 56;
    const uint64_t take = uint64_t(0) - uint64_t(window_ >= big);  // all ones when the bit is 1
    window_ -= big & take;
    range_ = split + ((range_ - 2 * split) & uint32_t(take));  // 1: range - split, 0: split
    const int sh = __builtin_clz(range_) - 24;                 // back into [128, 255]
    range_ <<= sh;
    window_ <<= sh;
    avail_ -= sh;
    return int(take & 1);
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

  // Hand-over of the decoder state to another implementation (the device-side decoder): after
  // Prime(), the window is `Top8()` followed by the raw stream bits from bit `Shifts() + 8` on
  // (subtractions only ever touch the top 8 bits, so everything below them is still raw stream).
  void Prime() {
    if (avail_ < 8) Refill();
  }
  uint32_t Top8() const { return uint32_t(window_ >> 56); }
  uint32_t Range() const { return range_; }
  size_t Shifts() const { return 8 * loaded_ - size_t(avail_); }

  // Bytes the reference's byte-at-a-time reader would have consumed so far.
  size_t BytesConsumed() const { return 2 + ((8 * loaded_ - size_t(avail_)) >> 3); }
  bool Overrun() const { return used_ && BytesConsumed() > size_; }
  size_t size() const { return size_; }

 private:
  inline void Refill() {
    const int k = (64 - avail_) >> 3;  // whole bytes that fit below the valid bits
    if (end_ - cur_ >= 8) {
      uint64_t w;
      std::memcpy(&w, cur_, 8);
      w = __builtin_bswap64(w);
      // Bits of a partially fitting byte land below the counted ones; the next refill ORs the
      // same values over them, so they are harmless.
      window_ |= avail_ ? (w >> avail_) : w;
      cur_ += k;
    } else {
      for (int i = 0; i < k; ++i) {
        const uint64_t b = cur_ < end_ ? uint64_t(*cur_++) : 0;  // zeros past the end; Overrun() tells
        window_ |= b << (56 - avail_ - 8 * i);
      }
    }
    avail_ += 8 * k;
    loaded_ += size_t(k);
  }

  const uint8_t *cur_ = nullptr;
  const uint8_t *end_ = nullptr;
  size_t size_ = 0;
  uint64_t window_ = 0;  // next bits of the stream, left aligned
  int avail_ = 0;        // number of counted valid bits in window_
  size_t loaded_ = 0;    // bytes appended so far (virtual zeros past the end included)
  uint32_t range_ = 255;
  bool used_ = false;
};

}  // namespace vp8r

#endif  // VP8R_HOST_BOOL_READER_H_
