#include "frame_writer.h"

#include <cstring>

#include "bool_writer.h"
#include "vp8r.h"

namespace vp8r {
namespace {

#include "vp8_prob_tables.inc"

enum { DC_PRED = 0, V_PRED, H_PRED, TM_PRED, B_PRED };
const int8_t kTreeYModeKey[8] = {-B_PRED, 2, 4, 6, -DC_PRED, -V_PRED, -H_PRED, -TM_PRED};
const int8_t kTreeUvMode[6] = {-DC_PRED, 2, -V_PRED, 4, -H_PRED, -TM_PRED};
const int8_t kTreeBMode[18] = {0, 2, -1, 4, -2, 6, 8, 12, -3, 10, -5, -6, -4, 14, -7, 16, -8, -9};
const uint8_t kProbYModeKey[4] = {145, 156, 163, 128};
const uint8_t kProbUvModeKey[3] = {142, 114, 183};
const uint8_t kZigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
const uint8_t kBand[17] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0};
const uint8_t kCat1[] = {159, 0}, kCat2[] = {165, 145, 0}, kCat3[] = {173, 148, 140, 0};
const uint8_t kCat4[] = {176, 155, 140, 135, 0}, kCat5[] = {180, 157, 141, 134, 130, 0};
const uint8_t kCat6[] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0};
const uint8_t *const kCatProbs[6] = {kCat1, kCat2, kCat3, kCat4, kCat5, kCat6};
const int kCatBase[7] = {5, 7, 11, 19, 35, 67, 2115};
const int kCatBits[6] = {1, 2, 3, 4, 5, 11};

// One block of tokens (RFC 6386 section 13; the mirror of FrameParser::ReadCoefTokens): `blk` holds the quantised
// coefficients in raster order, scan position n is blk[kZigzag[n]].  No end-of-block right after a zero token.
// Returns whether any coefficient from `first` on is non-zero.  A magnitude above DCT_CAT6's range cannot be coded.
bool WriteBlock(BoolWriter &bw, int type, int ctx, int first, const int16_t *blk, bool *too_large) {
  int last = -1;
  for (int n = first; n < 16; ++n)
    if (blk[kZigzag[n]]) last = n;
  const uint8_t *base = &kCoefDefault[size_t(type) * 8 * 3 * 11];
  auto probs = [&](int n, int cx) { return base + (size_t(kBand[n]) * 3 + size_t(cx)) * 11; };
  bool prev_zero = false;
  int n = first;
  for (; n <= last; ++n) {
    const uint8_t *p = probs(n, ctx);
    const int v = blk[kZigzag[n]], a = v < 0 ? -v : v;
    if (!prev_zero) bw.Put(p[0], 1);
    if (a == 0) {
      bw.Put(p[1], 0);
      prev_zero = true;
      ctx = 0;
      continue;
    }
    bw.Put(p[1], 1);
    if (a == 1) {
      bw.Put(p[2], 0);
    } else {
      bw.Put(p[2], 1);
      if (a <= 4) {
        bw.Put(p[3], 0);
        if (a == 2) {
          bw.Put(p[4], 0);
        } else {
          bw.Put(p[4], 1);
          bw.Put(p[5], a == 4);
        }
      } else {
        if (a >= kCatBase[6]) {
          *too_large = true;
          return true;
        }
        bw.Put(p[3], 1);
        int cat = 0;
        while (a >= kCatBase[cat + 1]) ++cat;
        if (cat < 2) {
          bw.Put(p[6], 0);
          bw.Put(p[7], cat == 1);
        } else {
          bw.Put(p[6], 1);
          if (cat < 4) {
            bw.Put(p[8], 0);
            bw.Put(p[9], cat == 3);
          } else {
            bw.Put(p[8], 1);
            bw.Put(p[10], cat == 5);
          }
        }
        const int extra = a - kCatBase[cat];
        for (int i = 0; i < kCatBits[cat]; ++i) bw.Put(kCatProbs[cat][i], (extra >> (kCatBits[cat] - 1 - i)) & 1);
      }
    }
    bw.Put(128, v < 0);
    prev_zero = false;
    ctx = a > 1 ? 2 : 1;
  }
  if (n < 16) bw.Put(probs(n, ctx)[0], 0);  // end of block (never right after a zero: `last` is non-zero)
  return last >= 0;
}

}  // namespace

int WriteKeyFrame(const vp8r_frame &f, std::vector<uint8_t> *out, std::string *err) {
  const vp8r_frame_hdr &h = f.hdr;
  auto fail = [&](const char *what) {
    if (err) *err = what;
    return VP8R_ERR_INVALID_ARG;
  };
  if (!f.blob || !h.key_frame) return fail("vp8r_frame_write_bitstream: only key frames with host arrays can be written");
  if (h.tokens_deferred || h.modes_deferred) return fail("vp8r_frame_write_bitstream: the frame holds raw partitions, not macroblock records");
  const int cols = h.mb_cols, rows = h.mb_rows;
  if (cols <= 0 || rows <= 0 || size_t(cols) * rows != f.n_mb) return fail("vp8r_frame_write_bitstream: inconsistent frame size");
  const vp8r_mb_info *mbs = f.mbs();
  const int16_t *payload = f.payload();

  // How often is a macroblock skipped?  prob_skip_false = P(skip flag == 0) out of 256.
  size_t n_skip = 0;
  for (size_t i = 0; i < f.n_mb; ++i) n_skip += mbs[i].coef_mask == 0;
  int prob_skip = int(256 * (f.n_mb - n_skip) / (f.n_mb ? f.n_mb : 1));
  prob_skip = prob_skip < 1 ? 1 : (prob_skip > 255 ? 255 : prob_skip);

  BoolWriter hdr, tok;
  hdr.Lit(1, 0);  // color space
  hdr.Lit(1, 0);  // clamping type: the reconstruction clamps
  hdr.Lit(1, 0);  // segmentation_enabled
  hdr.Lit(1, h.filter_type ? 1u : 0u);
  hdr.Lit(6, h.loop_filter_level);
  hdr.Lit(3, h.sharpness_level);
  hdr.Lit(1, 0);  // loop_filter_adj_enable
  hdr.Lit(2, 0);  // one DCT partition
  hdr.Lit(7, h.q_index);
  for (int i = 0; i < 5; ++i) hdr.Lit(1, 0);  // no quantiser deltas
  hdr.Lit(1, 1);                              // refresh_entropy_probs
  for (int i = 0; i < 1056; ++i) hdr.Put(kCoefUpdate[i], 0);  // keep the default token probabilities
  hdr.Lit(1, 1);                                             // mb_no_skip_coeff
  hdr.Lit(8, uint32_t(prob_skip));

  std::vector<uint8_t> above_b(size_t(cols) * 4, 0), nz_ay(size_t(cols) * 4, 0), nz_au(size_t(cols) * 2, 0), nz_av(size_t(cols) * 2, 0),
      nz_ay2(size_t(cols), 0);
  bool too_large = false;
  for (int r = 0; r < rows; ++r) {
    uint8_t left_b[4] = {0, 0, 0, 0};
    uint8_t nz_ly[4] = {0, 0, 0, 0}, nz_lu[2] = {0, 0}, nz_lv[2] = {0, 0}, nz_ly2 = 0;
    for (int c = 0; c < cols; ++c) {
      const vp8r_mb_info &mb = mbs[size_t(r) * cols + c];
      if (mb.flags & VP8R_MB_IS_INTER) return fail("vp8r_frame_write_bitstream: inter macroblock in a key frame");
      const int ymode = int((mb.flags >> VP8R_MB_MODE_SHIFT) & 7), uvmode = int((mb.flags >> VP8R_MB_UVMODE_SHIFT) & 3);
      const bool has_y2 = ymode != B_PRED;
      if (has_y2 != ((mb.flags & VP8R_MB_HAS_Y2) != 0) || ymode > B_PRED) return fail("vp8r_frame_write_bitstream: inconsistent macroblock mode");
      const bool skip = mb.coef_mask == 0;
      hdr.Put(prob_skip, skip);
      hdr.Tree(kTreeYModeKey, 8, kProbYModeKey, ymode);
      if (ymode == B_PRED) {
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j) {
            const int b = i * 4 + j, m = int((mb.aux[b >> 3] >> ((b & 7) * 4)) & 15);
            if (m > 9) return fail("vp8r_frame_write_bitstream: invalid sub-block mode");
            hdr.Tree(kTreeBMode, 18, &kKfBmode[(size_t(above_b[size_t(c) * 4 + j]) * 10 + left_b[i]) * 9], m);
            above_b[size_t(c) * 4 + j] = left_b[i] = uint8_t(m);
          }
      } else {
        static const uint8_t implied[4] = {0, 2, 3, 1};  // DC->B_DC, V->B_VE, H->B_HE, TM->B_TM (RFC 6386 11.3)
        for (int i = 0; i < 4; ++i) above_b[size_t(c) * 4 + i] = left_b[i] = implied[ymode];
      }
      hdr.Tree(kTreeUvMode, 6, kProbUvModeKey, uvmode);

      if (!skip) {
        // stored blocks follow each other in increasing block number; absent blocks are all zero
        static const int16_t kZeroBlock[16] = {0};
        const int16_t *next = payload + size_t(mb.coef_offset) * 16;
        const int16_t *blk[25];
        for (int b = 0; b < 25; ++b) {
          if ((mb.coef_mask >> b) & 1) {
            blk[b] = next;
            next += 16;
          } else {
            blk[b] = kZeroBlock;
          }
        }
        uint32_t nz = 0;  // bit 0: Y2, 1..16: Y, 17..20: U, 21..24: V
        if (has_y2) {
          if (WriteBlock(tok, 1, nz_ay2[size_t(c)] + nz_ly2, 0, blk[0], &too_large)) nz |= 1;
          nz_ay2[size_t(c)] = nz_ly2 = uint8_t(nz & 1);
        }
        for (int b = 0; b < 16; ++b) {
          const int i = b >> 2, j = b & 3;
          const int a = i ? int((nz >> (b - 3)) & 1) : nz_ay[size_t(c) * 4 + j];
          const int l = j ? int((nz >> b) & 1) : nz_ly[i];
          if (WriteBlock(tok, has_y2 ? 0 : 3, a + l, has_y2 ? 1 : 0, blk[1 + b], &too_large)) nz |= 2u << b;
        }
        for (int pl = 0; pl < 2; ++pl) {
          uint8_t *na = pl ? &nz_av[size_t(c) * 2] : &nz_au[size_t(c) * 2];
          uint8_t *nl = pl ? nz_lv : nz_lu;
          const int base = 17 + 4 * pl;
          for (int b = 0; b < 4; ++b) {
            const int i = b >> 1, j = b & 1;
            const int a = i ? int((nz >> (base + b - 2)) & 1) : na[j];
            const int l = j ? int((nz >> (base + b - 1)) & 1) : nl[i];
            if (WriteBlock(tok, 2, a + l, 0, blk[base + b], &too_large)) nz |= 1u << (base + b);
          }
        }
        if (too_large) return fail("vp8r_frame_write_bitstream: a coefficient exceeds the largest token (2048 + 66)");
        for (int j = 0; j < 4; ++j) nz_ay[size_t(c) * 4 + j] = uint8_t((nz >> (13 + j)) & 1);
        for (int i = 0; i < 4; ++i) nz_ly[i] = uint8_t((nz >> (4 + i * 4)) & 1);
        for (int j = 0; j < 2; ++j) {
          nz_au[size_t(c) * 2 + j] = uint8_t((nz >> (19 + j)) & 1);
          nz_av[size_t(c) * 2 + j] = uint8_t((nz >> (23 + j)) & 1);
        }
        for (int i = 0; i < 2; ++i) {
          nz_lu[i] = uint8_t((nz >> (18 + i * 2)) & 1);
          nz_lv[i] = uint8_t((nz >> (22 + i * 2)) & 1);
        }
      } else {
        if (has_y2) nz_ay2[size_t(c)] = nz_ly2 = 0;
        for (int j = 0; j < 4; ++j) nz_ay[size_t(c) * 4 + j] = nz_ly[j] = 0;
        for (int j = 0; j < 2; ++j) nz_au[size_t(c) * 2 + j] = nz_av[size_t(c) * 2 + j] = nz_lu[j] = nz_lv[j] = 0;
      }
    }
  }

  const std::vector<uint8_t> first = hdr.Finish(), tokens = tok.Finish();
  if (first.size() >= (1u << 19)) return fail("vp8r_frame_write_bitstream: first partition too large for the frame tag");
  out->clear();
  const uint32_t tag = 0u /* key frame */ | (0u << 1) /* version 0 */ | (1u << 4) /* shown */ | (uint32_t(first.size()) << 5);
  out->push_back(uint8_t(tag));
  out->push_back(uint8_t(tag >> 8));
  out->push_back(uint8_t(tag >> 16));
  const uint8_t sc[3] = {0x9d, 0x01, 0x2a};
  out->insert(out->end(), sc, sc + 3);
  out->push_back(uint8_t(h.width));
  out->push_back(uint8_t((h.width >> 8) & 0x3f));
  out->push_back(uint8_t(h.height));
  out->push_back(uint8_t((h.height >> 8) & 0x3f));
  out->insert(out->end(), first.begin(), first.end());
  out->insert(out->end(), tokens.begin(), tokens.end());
  return VP8R_OK;
}

}  // namespace vp8r
