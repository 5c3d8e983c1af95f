// oracle/ref_enc_harness.cc -- TEST INFRASTRUCTURE.  Generates golden vectors for the encoder-side functions of
// the reference (row f4 of SURVEY.md section 8) by CALLING the reference's own code, compiled from the sources
// where they lie under /root/reference/src (recipe: oracle/Makefile, target `encgold`):
//   vp8::DCT / vp8::WHT              src/dct.cc:5-65
//   vp8::Quantize                    src/quantizer.cc:5-8
//   vp8::internal::PickIntraModeChroma                        src/encode_frame.cc:30-76
//   luma 16x16 candidates: the reference's predictors VPredLuma / HPredLuma / DCPredLuma / TMPredLuma
//   (src/intra_predict.cc) in the order and with the strict "<" of PickIntraModeLuma (src/encode_frame.cc:204-238),
//   error = sum of squared differences (src/encode_frame.cc:6-17).  PickIntraModeLuma itself cannot be run: its
//   B_PRED branch walks "mode <= B_HU_PRED; ++mode" and operator++ aborts on B_HU_PRED (src/encode_frame.cc:79-107,
//   115), so the whole function exits the process.  The sub-block decision is therefore pinned one level down:
//   the reference's BPredSubBlock for each of the ten modes in that order, same error, same "<"
//   (src/encode_frame.cc:109-124).
// Output: one JSON document on stdout (committed as tests/golden/enc/enc_goldens.json; seed 7122, the seed the
// reference's own tests use, test/dct_test.h:19).
#include <cstdio>
#include <random>
#include <string>

#include "dct.h"
#include "encode_frame.h"
#include "quantizer.h"

using namespace vp8;

static std::mt19937 rng(7122);
static int Rand(int lo, int hi) { return lo + int(rng() % uint32_t(hi - lo + 1)); }

static void PrintBlock(const std::array<std::array<int16_t, 4>, 4> &b) {
  std::printf("[");
  for (int i = 0; i < 16; ++i) std::printf("%s%d", i ? "," : "", int(b[size_t(i) >> 2][size_t(i) & 3]));
  std::printf("]");
}

template <size_t C>
static void FillMb(MacroBlock<C> &mb, int base, int spread) {
  for (size_t r = 0; r < C * 4; ++r)
    for (size_t c = 0; c < C * 4; ++c) {
      int v = base + Rand(-spread, spread);
      mb.SetPixel(r, c, int16_t(v < 0 ? 0 : (v > 255 ? 255 : v)));
    }
}
template <size_t C>
static void PrintMb(const MacroBlock<C> &mb) {
  std::printf("[");
  for (size_t r = 0; r < C * 4; ++r)
    for (size_t c = 0; c < C * 4; ++c) std::printf("%s%d", (r || c) ? "," : "", int(mb.GetPixel(r, c)));
  std::printf("]");
}

int main() {
  std::printf("{\n \"dct\": [");
  for (int t = 0; t < 240; ++t) {
    std::array<std::array<int16_t, 4>, 4> b{};
    const int amp = t < 120 ? 255 : (t < 180 ? 20 : 2000);
    for (auto &row : b)
      for (auto &v : row) v = int16_t(Rand(-amp, amp));
    if (t == 0) for (auto &row : b) for (auto &v : row) v = 255;
    if (t == 1) for (auto &row : b) for (auto &v : row) v = -255;
    std::printf("%s\n  {\"in\": ", t ? "," : "");
    PrintBlock(b);
    auto d = b, w = b;
    DCT(d);
    WHT(w);
    std::printf(", \"dct\": ");
    PrintBlock(d);
    std::printf(", \"wht\": ");
    PrintBlock(w);
    std::printf("}");
  }
  std::printf("\n ],\n \"quantize\": [");
  for (int t = 0; t < 200; ++t) {
    std::array<int16_t, 16> c{};
    for (auto &v : c) v = int16_t(Rand(-2500, 2500));
    const QuantFactor qf(int16_t(Rand(4, 157)), int16_t(Rand(4, 284)));
    std::printf("%s\n  {\"dc\": %d, \"ac\": %d, \"in\": [", t ? "," : "", int(qf.first), int(qf.second));
    for (int i = 0; i < 16; ++i) std::printf("%s%d", i ? "," : "", int(c[size_t(i)]));
    Quantize(c, qf);
    std::printf("], \"out\": [");
    for (int i = 0; i < 16; ++i) std::printf("%s%d", i ? "," : "", int(c[size_t(i)]));
    std::printf("]}");
  }
  // Mode pickers: a 2 x 2 macroblock plane whose first three macroblocks hold "already coded" pixels; the
  // macroblock at (r, c) is picked for; (0,0), (0,1), (1,0), (1,1) cover the frame-edge rules.
  std::printf("\n ],\n \"pick\": [");
  for (int t = 0; t < 48; ++t) {
    const size_t r = size_t(t) & 1, c = (size_t(t) >> 1) & 1;
    Plane<4> y(2, 2);
    Plane<2> u(2, 2), v(2, 2);
    const int style = (t >> 2) % 4;  // flat, textured, gradient-like, noisy
    const int spread = style == 0 ? 3 : (style == 1 ? 25 : (style == 2 ? 8 : 90));
    for (size_t i = 0; i < 2; ++i)
      for (size_t j = 0; j < 2; ++j) {
        FillMb(y.at(i).at(j), Rand(20, 235), spread);
        FillMb(u.at(i).at(j), Rand(60, 200), spread / 2 + 1);
        FillMb(v.at(i).at(j), Rand(60, 200), spread / 2 + 1);
      }
    if (style == 2) {  // a vertical ramp continues the row above: favours V / TM
      for (size_t i = 0; i < 16; ++i)
        for (size_t j = 0; j < 16; ++j) y.at(0).at(c).SetPixel(i, j, int16_t(40 + 10 * int(j)));
    }
    MacroBlock<4> ty;
    MacroBlock<2> tu, tv;
    FillMb(ty, Rand(20, 235), spread);
    FillMb(tu, Rand(60, 200), spread / 2 + 1);
    FillMb(tv, Rand(60, 200), spread / 2 + 1);
    if (style == 2)
      for (size_t i = 0; i < 16; ++i)
        for (size_t j = 0; j < 16; ++j) ty.SetPixel(i, j, int16_t(40 + 10 * int(j) + Rand(-2, 2)));
    std::printf("%s\n  {\"r\": %zu, \"c\": %zu, \"y\": [", t ? "," : "", r, c);
    for (size_t i = 0; i < 2; ++i)
      for (size_t j = 0; j < 2; ++j) {
        std::printf("%s", (i || j) ? "," : "");
        PrintMb(y.at(i).at(j));
      }
    std::printf("], \"u\": [");
    for (size_t i = 0; i < 2; ++i)
      for (size_t j = 0; j < 2; ++j) {
        std::printf("%s", (i || j) ? "," : "");
        PrintMb(u.at(i).at(j));
      }
    std::printf("], \"v\": [");
    for (size_t i = 0; i < 2; ++i)
      for (size_t j = 0; j < 2; ++j) {
        std::printf("%s", (i || j) ? "," : "");
        PrintMb(v.at(i).at(j));
      }
    std::printf("], \"ty\": ");
    PrintMb(ty);
    std::printf(", \"tu\": ");
    PrintMb(tu);
    std::printf(", \"tv\": ");
    PrintMb(tv);
    auto sse = [&](const MacroBlock<4> &p) {
      uint32_t e = 0;
      for (size_t i = 0; i < 16; ++i)
        for (size_t j = 0; j < 16; ++j) {
          const int d = ty.GetPixel(i, j) - p.GetPixel(i, j);
          e += uint32_t(d * d);
        }
      return e;
    };
    uint32_t best = UINT_MAX, errs[4];
    MacroBlockMode ym = DC_PRED;
    internal::VPredLuma(r, c, y);
    errs[0] = sse(y.at(r).at(c));
    if (errs[0] < best) best = errs[0], ym = V_PRED;
    internal::HPredLuma(r, c, y);
    errs[1] = sse(y.at(r).at(c));
    if (errs[1] < best) best = errs[1], ym = H_PRED;
    internal::DCPredLuma(r, c, y);
    errs[2] = sse(y.at(r).at(c));
    if (errs[2] < best) best = errs[2], ym = DC_PRED;
    internal::TMPredLuma(r, c, y);
    errs[3] = sse(y.at(r).at(c));
    if (errs[3] < best) best = errs[3], ym = TM_PRED;
    const MacroBlockMode uvm = internal::PickIntraModeChroma(r, c, tu, tv, u, v);
    std::printf(", \"ymode\": %d, \"uvmode\": %d, \"yerr\": [%u,%u,%u,%u]}", int(ym), int(uvm), errs[0], errs[1], errs[2], errs[3]);
  }
  std::printf("\n ],\n \"subpick\": [");
  for (int t = 0; t < 150; ++t) {
    std::array<int16_t, 8> above{};
    std::array<int16_t, 4> left{};
    const int base = Rand(10, 245), spread = (t % 3 == 0) ? 4 : ((t % 3 == 1) ? 30 : 120);
    auto px = [&]() { int v = base + Rand(-spread, spread); return int16_t(v < 0 ? 0 : (v > 255 ? 255 : v)); };
    for (auto &v : above) v = px();
    for (auto &v : left) v = px();
    const int16_t pp = px();
    SubBlock target, predict;
    for (size_t i = 0; i < 4; ++i)
      for (size_t j = 0; j < 4; ++j) target.at(i).at(j) = px();
    uint32_t best = UINT_MAX;
    int best_mode = 0;
    uint32_t errs[10];
    for (int m = 0; m < 10; ++m) {
      internal::BPredSubBlock(above, left, pp, SubBlockMode(m), predict);
      uint32_t e = 0;
      for (size_t i = 0; i < 4; ++i)
        for (size_t j = 0; j < 4; ++j) {
          const int d = target.at(i).at(j) - predict.at(i).at(j);
          e += uint32_t(d * d);
        }
      errs[m] = e;
      if (e < best) best = e, best_mode = m;
    }
    std::printf("%s\n  {\"above\": [", t ? "," : "");
    for (int i = 0; i < 8; ++i) std::printf("%s%d", i ? "," : "", int(above[size_t(i)]));
    std::printf("], \"left\": [%d,%d,%d,%d], \"p\": %d, \"target\": [", int(left[0]), int(left[1]), int(left[2]), int(left[3]), int(pp));
    for (int i = 0; i < 16; ++i) std::printf("%s%d", i ? "," : "", int(target.at(size_t(i) >> 2).at(size_t(i) & 3)));
    std::printf("], \"mode\": %d, \"err\": [", best_mode);
    for (int m = 0; m < 10; ++m) std::printf("%s%u", m ? "," : "", errs[m]);
    std::printf("]}");
  }
  std::printf("\n ]\n}\n");
  return 0;
}
