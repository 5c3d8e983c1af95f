// CPU check of vp8_b200/csrc/cuda/lf_swar.h (four pixel lines per register) against the scalar loop-filter
// edge as the reference states it (src/filter.cc:7-67, limits src/filter.cc:119-149).  The header is
// compiled here with its host emulation of the device SIMD instructions; the kernels compile the very same
// functions with the native ones.
//
// usage: lf_swar_test [million_lines]      exit code 0 = all equal
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "cuda/lf_swar.h"

namespace {
int clamp255(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }
int clamp128(int x) { return x < -128 ? -128 : (x > 127 ? 127 : x); }
int imin(int a, int b) { return a < b ? a : b; }
int imax(int a, int b) { return a > b ? a : b; }

struct Limits {
  int interior, hev, edge_mb, edge_sb;
};
// src/filter.cc:119-149
Limits MakeLimits(int level, int sharp, bool key) {
  Limits l;
  int in = level;
  if (sharp) {
    in >>= (sharp > 4) ? 2 : 1;
    in = imin(in, 9 - sharp);
  }
  l.interior = imax(in, 1);
  if (key) l.hev = level >= 40 ? 2 : (level >= 15 ? 1 : 0);
  else l.hev = level >= 40 ? 3 : (level >= 20 ? 2 : (level >= 15 ? 1 : 0));
  l.edge_mb = (level + 2) * 2 + l.interior;
  l.edge_sb = level * 2 + l.interior;
  return l;
}

// kind 0: normal macroblock edge, 1: normal sub-block edge, 2: simple (edge limit given)
void ScalarEdge(int *v, const Limits &lim, int kind, bool mb_edge) {
  const int edge = mb_edge ? lim.edge_mb : lim.edge_sb;
  const int p3 = v[0], p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6], q3 = v[7];
  const bool edge_ok = std::abs(q0 - p0) * 2 + (std::abs(p1 - q1) >> 1) <= edge;
  const int s = clamp128(p1 - q1);
  if (kind == 2) {
    const int a = clamp128(s + 3 * (q0 - p0));
    const int f1 = imin(a + 4, 127) >> 3, f2 = imin(a + 3, 127) >> 3;
    if (edge_ok) {
      v[3] = clamp255(p0 + f2);
      v[4] = clamp255(q0 - f1);
    }
    return;
  }
  const int interior = imax(imax(imax(std::abs(p3 - p2), std::abs(p2 - p1)), std::abs(p1 - p0)),
                            imax(imax(std::abs(q1 - q0), std::abs(q2 - q1)), std::abs(q3 - q2)));
  const bool on = edge_ok && interior <= lim.interior;
  const bool hev = imax(std::abs(p1 - p0), std::abs(q1 - q0)) > lim.hev;
  if (!on) return;
  if (kind == 0) {
    const int w = clamp128(s + 3 * (q0 - p0));
    if (hev) {
      const int f1 = imin(w + 4, 127) >> 3, f2 = imin(w + 3, 127) >> 3;
      v[3] = clamp255(p0 + f2);
      v[4] = clamp255(q0 - f1);
    } else {
      const int a27 = (27 * w + 63) >> 7, a18 = (18 * w + 63) >> 7, a9 = (9 * w + 63) >> 7;
      v[1] = clamp255(p2 + a9);
      v[2] = clamp255(p1 + a18);
      v[3] = clamp255(p0 + a27);
      v[4] = clamp255(q0 - a27);
      v[5] = clamp255(q1 - a18);
      v[6] = clamp255(q2 - a9);
    }
  } else {
    const int a = clamp128((hev ? s : 0) + 3 * (q0 - p0));
    const int f1 = imin(a + 4, 127) >> 3, f2 = imin(a + 3, 127) >> 3;
    v[3] = clamp255(p0 + f2);
    v[4] = clamp255(q0 - f1);
    if (!hev) {
      const int a2 = (f1 + 1) >> 1;
      v[2] = clamp255(p1 + a2);
      v[5] = clamp255(q1 - a2);
    }
  }
}
}  // namespace

int main(int argc, char **argv) {
  const long long total = (argc > 1 ? std::atoll(argv[1]) : 8) * 1000000LL;
  std::mt19937 rng(7122);
  long long bad = 0, done = 0, filtered = 0;
  for (long long it = 0; done < total; ++it) {
    // filter parameters: every level / sharpness / frame type comes up; half of the time the filter of the
    // macroblock is switched off entirely (level 0 is expressed by `enabled`)
    const int level = 1 + int(rng() % 63), sharp = int(rng() % 8);
    const bool key = rng() & 1, enabled = (rng() % 8) != 0;
    const Limits lim = MakeLimits(level, sharp, key);
    const vp8r::swar::EdgeK k = vp8r::swar::MakeEdgeK(lim.interior, lim.hev, lim.edge_mb, lim.edge_sb, enabled);
    // pixel statistics: flat areas with small steps (filters fire), extremes, and plain noise
    int px[4][8];
    const int style = int(rng() % 6);
    for (int l = 0; l < 4; ++l) {
      if (style == 5) {  // saturated neighbours (differences of 255) next to values at the limits
        const int base = int(rng() % 256);
        for (int i = 0; i < 8; ++i) {
          const unsigned pick = rng() % 8;
          px[l][i] = pick < 2 ? 0 : (pick < 4 ? 255 : clamp255(base + int(rng() % (2 * lim.interior + 3)) - lim.interior - 1));
        }
        continue;
      }
      const int base = int(rng() % 256), spread = style == 0 ? 2 : (style == 1 ? 6 : (style == 2 ? 20 : (style == 3 ? 70 : 256)));
      const int step = (style <= 2) ? int(rng() % (4 * spread + 1)) - 2 * spread : 0;
      for (int i = 0; i < 8; ++i) {
        int v = base + (i >= 4 ? step : 0) + int(rng() % (spread + 1)) - spread / 2;
        if (style == 4 && (rng() % 4) == 0) v = (rng() & 1) ? 255 : 0;
        px[l][i] = clamp255(v);
      }
    }
    for (int kind = 0; kind < 3; ++kind) {
      for (int mb = 0; mb < 2; ++mb) {
        if (kind == 0 && !mb) continue;
        if (kind == 1 && mb) continue;
        uint32_t w[8];
        for (int i = 0; i < 8; ++i) w[i] = uint32_t(px[0][i]) | (uint32_t(px[1][i]) << 8) | (uint32_t(px[2][i]) << 16) | (uint32_t(px[3][i]) << 24);
        if (kind == 0) vp8r::swar::NormalMbEdge(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], k);
        else if (kind == 1) vp8r::swar::NormalInner(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], k);
        else vp8r::swar::SimpleEdge(w[2], w[3], w[4], w[5], mb ? k.k_mb : k.k_sb);
        for (int l = 0; l < 4; ++l) {
          int v[8];
          for (int i = 0; i < 8; ++i) v[i] = px[l][i];
          if (enabled) ScalarEdge(v, lim, kind, mb != 0);
          bool changed = false;
          for (int i = 0; i < 8; ++i) {
            const int got = int((w[i] >> (8 * l)) & 0xff);
            changed |= v[i] != px[l][i];
            if (got != v[i]) {
              if (bad < 10)
                std::fprintf(stderr, "mismatch kind %d mb %d level %d sharp %d key %d line %d pixel %d: got %d want %d (in %d)\n", kind, mb,
                             level, sharp, int(key), l, i, got, v[i], px[l][i]);
              ++bad;
            }
          }
          filtered += changed;
          ++done;
        }
      }
    }
  }
  // transposition helper
  {
    uint32_t r[4] = {0x03020100u, 0x13121110u, 0x23222120u, 0x33323130u};
    vp8r::swar::Transpose4(r[0], r[1], r[2], r[3]);
    if (r[0] != 0x30201000u || r[1] != 0x31211101u || r[2] != 0x32221202u || r[3] != 0x33231303u) {
      std::fprintf(stderr, "Transpose4 wrong\n");
      ++bad;
    }
  }
  std::printf("lines %lld filtered %lld mismatches %lld\n", done, filtered, bad);
  return bad ? 1 : 0;
}
