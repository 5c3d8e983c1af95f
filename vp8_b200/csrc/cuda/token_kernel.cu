// K_tokens: device-side decode of the DCT token partitions (frames parsed with deferred tokens).
//
// Replaces, for those frames, the residual half of the host parser: ResidualTokens / the token
// loop of src/bitstream_parser.cc:466-537,572-621 driven by the bool decoder of
// src/bool_decoder.cc:13-41, plus the non-zero context hand-over of src/decode_frame.cc:6-47,111-130.
//
// Parallelism is what the bitstream offers: a bool-coded partition is a serial chain, but the
// 1/2/4/8 DCT partitions of a frame are separate chains (MB row r lives in partition r % n) that
// only meet through the "above" non-zero context, and frames of different streams are independent.
// One CTA per frame, one warp per partition with a single working lane; rows advance as a pipeline:
// row r may decode macroblock c once row r-1 has published c.  With a few hundred frames per batch
// that is ~1000 independent chains on 148 SMs, each bound by the latency of its own dependent
// integer chain (~25 instructions per boolean), which is what a GPU thread is slow at and many
// GPU threads together are fast at.
//
// Output: coefficient blocks (de-zigzagged int16, not dequantised) into the frame's device
// coefficient area, row r compact inside its own region starting at block r * cols * 25;
// coef_mask / coef_offset / VP8R_MB_LF_INNER of every vp8r_mb_info are completed in place.
#include "recon_kernels.h"

namespace vp8r {

namespace {

__constant__ unsigned char c_band[17] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0};
__constant__ unsigned char c_zigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
// DCT_CAT extra-bit probabilities, 12 per category, zero terminated (src/bitstream_const.h:89-90).
__constant__ unsigned char c_cat[6][12] = {{159, 0},
                                           {165, 145, 0},
                                           {173, 148, 140, 0},
                                           {176, 155, 140, 135, 0},
                                           {180, 157, 141, 134, 130, 0},
                                           {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0}};
__constant__ short c_cat_base[6] = {5, 7, 11, 19, 35, 67};

// RFC 6386 section 7 boolean decoder, 64-bit left-aligned window refilled 32 bits at a time with
// aligned loads.  Produces the bit sequence of src/bool_decoder.cc:13-41.
struct BoolDec {
  const unsigned *next, *end;  // next aligned word to append / first word past the raw section
  unsigned long long win;      // upcoming bits, left aligned
  int avail;                   // valid bits in win
  unsigned range;              // 128..255
  int loaded;                  // bytes appended so far (for the over-read test)

  __device__ __forceinline__ void Refill() {
    unsigned w = 0;
    if (next < end) w = __ldg(next);
    ++next;
    w = __byte_perm(w, 0, 0x0123);  // big endian
    win |= (unsigned long long)w << (32 - avail);
    avail += 32;
    loaded += 4;
  }
  __device__ __forceinline__ void Init(const unsigned char *raw, unsigned off, const unsigned *raw_end) {
    const unsigned mis = off & 3u;
    next = reinterpret_cast<const unsigned *>(raw + (off - mis));
    end = raw_end;
    unsigned w = next < end ? __ldg(next) : 0u;
    ++next;
    w = __byte_perm(w, 0, 0x0123) << (8 * mis);
    win = (unsigned long long)w << 32;
    avail = 32 - 8 * (int)mis;
    loaded = 4 - (int)mis;
    range = 255;
    Refill();
  }
  // One boolean, probability prob/256 of being 0.
  __device__ __forceinline__ int Bit(unsigned prob) {
    const unsigned split = 1u + (((range - 1u) * prob) >> 8);
    const unsigned big = split << 24;
    unsigned hi = (unsigned)(win >> 32);
    const int bit = hi >= big;
    if (bit) {
      hi -= big;
      range -= split;
    } else {
      range = split;
    }
    const int sh = __clz(range) - 24;
    range <<= sh;
    win = (((unsigned long long)hi << 32) | (unsigned)win) << sh;
    avail -= sh;
    if (avail <= 32) Refill();
    return bit;
  }
  // Bytes the reference's byte-at-a-time reader would have consumed (see host/bool_reader.h).
  __device__ __forceinline__ int BytesConsumed() const { return 2 + ((8 * loaded - avail) >> 3); }
};

struct TokenShared {
  unsigned char probs[4 * 8 * 3 * 11];
  int progress[8];      // progress[w]: macroblocks finished by partition w, counted along its rows
  short blk[8][16];     // staging of the block being decoded, one per warp
};

// Tokens of one block whose first symbol was not end-of-block (src/bitstream_parser.cc:572-621).
// Returns bit 0: some coefficient is non-zero; bit 1: some coefficient is non-zero after the
// reference's int16 dequantisation (src/decode_frame.cc:6-47).
__device__ __forceinline__ int ReadTokens(BoolDec &bd, const unsigned char *probs, int type, int ctx, int first,
                                          int dc_f, int ac_f, short *blk) {
  const unsigned char *bands = probs + type * (8 * 3 * 11);
  int n = first;
  const unsigned char *p = bands + (c_band[n] * 3 + ctx) * 11;
  int any = 0, dq_any = 0;
  bool first_symbol = true;
  while (n < 16) {
    if (!first_symbol && !bd.Bit(p[0])) break;
    first_symbol = false;
    bool ended = false;
    while (!bd.Bit(p[1])) {
      if (++n == 16) {
        ended = true;
        break;
      }
      p = bands + (c_band[n] * 3) * 11;
    }
    if (ended) break;
    int v;
    if (!bd.Bit(p[2])) {
      v = 1;
    } else if (!bd.Bit(p[3])) {
      v = !bd.Bit(p[4]) ? 2 : 3 + bd.Bit(p[5]);
    } else {
      int cat;
      if (!bd.Bit(p[6])) cat = bd.Bit(p[7]);
      else if (!bd.Bit(p[8])) cat = 2 + bd.Bit(p[9]);
      else cat = 4 + bd.Bit(p[10]);
      int extra = 0;
      const unsigned char *q = c_cat[cat];
      for (unsigned pq = *q; pq; pq = *++q) extra = extra + extra + bd.Bit(pq);
      v = c_cat_base[cat] + extra;
    }
    const int next_ctx = v > 1 ? 2 : 1;
    if (bd.Bit(128)) v = -v;
    blk[c_zigzag[n]] = (short)v;
    any = 1;
    if ((short)(v * (n == 0 ? dc_f : ac_f)) != 0) dq_any = 2;
    ++n;
    p = bands + (c_band[n] * 3 + next_ctx) * 11;
  }
  return any | dq_any;
}

}  // namespace

constexpr int kTokenWarps = 8;

__global__ void __launch_bounds__(kTokenWarps * 32) TokenKernel(const DevFrameJob *__restrict__ jobs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DevFrameJob &job = jobs[blockIdx.x];
  if (!job.tok_hdr) return;
  TokenShared &sh = *reinterpret_cast<TokenShared *>(smem_raw);
  // above-context per macroblock column: bits 0-3 Y (column j), 4-5 U, 6-7 V, 8 Y2
  unsigned short *above = reinterpret_cast<unsigned short *>(smem_raw + ((sizeof(TokenShared) + 15) & ~15));
  const vp8r_token_hdr *th = reinterpret_cast<const vp8r_token_hdr *>(job.tok_hdr);
  const int cols = job.mb_cols, rows = job.mb_rows;
  for (int i = threadIdx.x; i < (int)sizeof(sh.probs) / 4; i += blockDim.x)
    reinterpret_cast<unsigned *>(sh.probs)[i] = __ldg(reinterpret_cast<const unsigned *>(th->coef_probs) + i);
  for (int i = threadIdx.x; i < cols; i += blockDim.x) above[i] = 0;
  if (threadIdx.x < 8) sh.progress[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < 8 * 16; i += blockDim.x) (&sh.blk[0][0])[i] = 0;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_parts = (int)__ldg(&th->n_parts);
  if (lane != 0 || warp >= n_parts) return;

  const unsigned char *raw = reinterpret_cast<const unsigned char *>(th) + sizeof(vp8r_token_hdr);
  const unsigned raw_bytes = __ldg(&th->raw_bytes);
  const unsigned part_size = __ldg(&th->part_size[warp]);
  BoolDec bd;
  bd.Init(raw, __ldg(&th->part_off[warp]), reinterpret_cast<const unsigned *>(raw + raw_bytes));
  bool used = false;

  vp8r_mb_info *mbs = const_cast<vp8r_mb_info *>(job.mbs);
  short *const blk = sh.blk[warp];
  volatile int *const prog = sh.progress;
  const int prev_warp = (warp + n_parts - 1) % n_parts;
  int16_t *const coef_area = const_cast<int16_t *>(job.payload) + (size_t)job.coef_base * 16;

  int k = 0;  // index of the row among this partition's rows
  for (int r = warp; r < rows; r += n_parts, ++k) {
    unsigned left = 0;  // bits 0-3 Y (row i), 4-5 U, 6-7 V, 8 Y2
    unsigned stored_row = 0;
    const unsigned row_base = (unsigned)r * (unsigned)cols * 25u;
    // what the previous row's partition must have reached before macroblock c of this row:
    // its progress counts macroblocks along its own rows; row r-1 is its row (warp ? k : k-1).
    const int prev_row_base = (warp ? k : k - 1) * cols;
    unsigned flags_next = __ldg(&mbs[(size_t)r * cols].flags);
    for (int c = 0; c < cols; ++c) {
      const size_t idx = (size_t)r * cols + c;
      const unsigned flags = flags_next;
      if (c + 1 < cols) flags_next = __ldg(&mbs[idx + 1].flags);
      if (r > 0 && n_parts > 1) {
        const int need = prev_row_base + c + 1;
        while (prog[prev_warp] < need) __nanosleep(100);
        __threadfence_block();
      }
      const bool has_y2 = (flags & VP8R_MB_HAS_Y2) != 0;
      unsigned abv = above[c];
      unsigned mask = 0;
      unsigned new_above, new_left;
      if (flags & VP8R_MB_SKIP_COEF) {
        // no tokens: contexts are cleared, except Y2's when the macroblock has no Y2 block
        new_above = has_y2 ? 0u : (abv & 0x100u);
        new_left = has_y2 ? 0u : (left & 0x100u);
      } else {
        used = true;
        const short *dq = job.dq[(flags >> VP8R_MB_QSEG_SHIFT) & 3];
        unsigned raw_nz = 0, dq_nz = 0;  // bit b as in coef_mask
        unsigned stored = 0;
        int16_t *out = coef_area + (size_t)(row_base + stored_row) * 16;
        const int ytype = has_y2 ? 0 : 3, yfirst = has_y2 ? 1 : 0;
        for (int b = has_y2 ? 0 : 1; b < 25; ++b) {
          // block kind, neighbour contexts (inside the macroblock: raw flags, src/bitstream_parser.cc:500-534)
          int type, first, dc_f, ac_f, a, l;
          if (b == 0) {
            type = 1; first = 0; dc_f = dq[VP8R_DQ_Y2_DC]; ac_f = dq[VP8R_DQ_Y2_AC];
            a = (int)((abv >> 8) & 1); l = (int)((left >> 8) & 1);
          } else if (b <= 16) {
            const int i = (b - 1) >> 2, j = (b - 1) & 3;
            type = ytype; first = yfirst; dc_f = dq[VP8R_DQ_Y1_DC]; ac_f = dq[VP8R_DQ_Y1_AC];
            a = (int)(((i ? raw_nz >> (b - 4) : abv >> j)) & 1);
            l = (int)(((j ? raw_nz >> (b - 1) : left >> i)) & 1);
          } else {
            const int k = (b - 17) & 3, cshift = 4 + 2 * ((b - 17) >> 2), i = k >> 1, j = k & 1;
            type = 2; first = 0; dc_f = dq[VP8R_DQ_UV_DC]; ac_f = dq[VP8R_DQ_UV_AC];
            a = (int)(((i ? raw_nz >> (b - 2) : abv >> (cshift + j))) & 1);
            l = (int)(((j ? raw_nz >> (b - 1) : left >> (cshift + i))) & 1);
          }
          const unsigned char *p = sh.probs + ((type * 8 + first) * 3 + a + l) * 11;  // band of coefficient `first` is `first`
          if (!bd.Bit(p[0])) continue;
          const int res = ReadTokens(bd, sh.probs, type, a + l, first, dc_f, ac_f, blk);
          if (res & 1) {
            raw_nz |= 1u << b;
            uint4 lo = *reinterpret_cast<const uint4 *>(blk), hi = *reinterpret_cast<const uint4 *>(blk + 8);
            *reinterpret_cast<uint4 *>(out + stored * 16) = lo;
            *reinterpret_cast<uint4 *>(out + stored * 16 + 8) = hi;
            *reinterpret_cast<uint4 *>(blk) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(blk + 8) = make_uint4(0, 0, 0, 0);
            ++stored;
          }
          if (res & 2) dq_nz |= 1u << b;
        }
        mask = raw_nz;
        // contexts handed to the neighbours: post-dequant flags
        new_above = ((dq_nz >> 13) & 0xfu) | (((dq_nz >> 19) & 3u) << 4) | (((dq_nz >> 23) & 3u) << 6);
        new_left = ((dq_nz >> 4) & 1u) | (((dq_nz >> 8) & 1u) << 1) | (((dq_nz >> 12) & 1u) << 2) | (((dq_nz >> 16) & 1u) << 3) |
                   (((dq_nz >> 18) & 1u) << 4) | (((dq_nz >> 20) & 1u) << 5) | (((dq_nz >> 22) & 1u) << 6) |
                   (((dq_nz >> 24) & 1u) << 7);
        if (has_y2) {
          new_above |= (dq_nz & 1u) << 8;
          new_left |= (dq_nz & 1u) << 8;
        } else {
          new_above |= abv & 0x100u;
          new_left |= left & 0x100u;
        }
        // complete the macroblock record
        if (mask) {
          mbs[idx].coef_mask = mask;
          mbs[idx].coef_offset = job.coef_base + row_base + stored_row;
          mbs[idx].flags = flags | VP8R_MB_LF_INNER;
        }
        stored_row += stored;
      }
      above[c] = (unsigned short)new_above;
      left = new_left;
      if (n_parts > 1) {
        __threadfence_block();
        prog[warp] = k * cols + c + 1;
      }
    }
  }
  if (used && bd.BytesConsumed() > (int)part_size && job.status) atomicOr(job.status, 1);
}

cudaError_t LaunchTokens(const DevFrameJob *jobs, int n_frames, int max_cols, cudaStream_t st) {
  const size_t smem = ((sizeof(TokenShared) + 15) & ~size_t(15)) + size_t(max_cols) * 2 + 16;
  TokenKernel<<<n_frames, kTokenWarps * 32, smem, st>>>(jobs);
  return cudaGetLastError();
}

}  // namespace vp8r
