/*
 * recon_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see recon_oracle.h).
 *
 * CPU restatement of the reference's reconstruction path on planar uint8 frames.  Every function
 * cites the reference lines it restates (paths relative to the reference repository root).
 * Parity status: PINNED against the reference's 43 golden MD5 vector files and against the
 * compiled reference decoder (tests/test_oracle_golden.py).
 */
#include "recon_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct {
  int cols, rows;   /* macroblocks */
  int ys, cs;       /* strides (= MB-aligned plane widths, src/frame.h:203-218) */
  int yh, ch;
  uint8_t *y, *u, *v;
} ofrm;

struct oracle_decoder {
  ofrm pool[5];
  ofrm *ref[4];     /* CURRENT, LAST, GOLDEN, ALTREF (src/bitstream_const.h:30-36) */
  int width, height;
  int have_frame;
};

static int clamp255(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }   /* src/utils.h:17-20 */
static int clamp128(int x) { return x < -128 ? -128 : (x > 127 ? 127 : x); } /* src/utils.h:22-25 */

/* ------------------------------------------------------------------ inverse transforms ------ */

/* src/dct.cc:67-107: vertical pass first, int16 storage between the passes. */
void oracle_idct4x4(int16_t m[16]) {
  const int c1 = 20091, c2 = 35468;
  for (int i = 0; i < 4; i++) {
    int a = m[0 + i] + m[8 + i];
    int b = m[0 + i] - m[8 + i];
    int t1 = (m[4 + i] * c2) >> 16;
    int t2 = m[12 + i] + ((m[12 + i] * c1) >> 16);
    int c = t1 - t2;
    t1 = m[4 + i] + ((m[4 + i] * c1) >> 16);
    t2 = (m[12 + i] * c2) >> 16;
    int d = t1 + t2;
    m[0 + i] = (int16_t)(a + d);
    m[12 + i] = (int16_t)(a - d);
    m[4 + i] = (int16_t)(b + c);
    m[8 + i] = (int16_t)(b - c);
  }
  for (int i = 0; i < 4; i++) {
    int16_t *r = m + 4 * i;
    int a = r[0] + r[2];
    int b = r[0] - r[2];
    int t1 = (r[1] * c2) >> 16;
    int t2 = r[3] + ((r[3] * c1) >> 16);
    int c = t1 - t2;
    t1 = r[1] + ((r[1] * c1) >> 16);
    t2 = (r[3] * c2) >> 16;
    int d = t1 + t2;
    r[0] = (int16_t)((a + d + 4) >> 3);
    r[3] = (int16_t)((a - d + 4) >> 3);
    r[1] = (int16_t)((b + c + 4) >> 3);
    r[2] = (int16_t)((b - c + 4) >> 3);
  }
}

/* src/dct.cc:109-133 */
void oracle_iwht4x4(int16_t m[16]) {
  for (int i = 0; i < 4; i++) {
    int a = m[0 + i] + m[12 + i];
    int b = m[4 + i] + m[8 + i];
    int c = m[4 + i] - m[8 + i];
    int d = m[0 + i] - m[12 + i];
    m[0 + i] = (int16_t)(a + b);
    m[4 + i] = (int16_t)(c + d);
    m[8 + i] = (int16_t)(a - b);
    m[12 + i] = (int16_t)(d - c);
  }
  for (int i = 0; i < 4; i++) {
    int16_t *r = m + 4 * i;
    int a = r[0] + r[3];
    int b = r[1] + r[2];
    int c = r[1] - r[2];
    int d = r[0] - r[3];
    r[0] = (int16_t)((a + b + 3) >> 3);
    r[1] = (int16_t)((c + d + 3) >> 3);
    r[2] = (int16_t)((a - b + 3) >> 3);
    r[3] = (int16_t)((d - c + 3) >> 3);
  }
}

/* Residual of one macroblock: 24 blocks of 16 int16 plus the "AC all zero" mask.
 * src/quantizer.cc:10-13 (int16 wrap-around), src/residual.cc:42-94,110-120. */
typedef struct {
  int16_t blk[24][16];
  uint32_t dc_only;
} mbres;

static void build_residual(const vp8r_frame_desc *f, const vp8r_mb_info *mb, mbres *out) {
  int16_t coef[25][16];
  memset(coef, 0, sizeof(coef));
  const int16_t *src = f->payload + (size_t)mb->coef_offset * 16;
  for (int b = 0; b < 25; b++)
    if ((mb->coef_mask >> b) & 1) {
      memcpy(coef[b], src, 32);
      src += 16;
    }
  const int16_t *dq = f->hdr.dq[(mb->flags >> VP8R_MB_QSEG_SHIFT) & 3];
  int has_y2 = (mb->flags & VP8R_MB_HAS_Y2) != 0;
  for (int b = 0; b < 25; b++) {
    int dc = b == 0 ? dq[VP8R_DQ_Y2_DC] : (b <= 16 ? dq[VP8R_DQ_Y1_DC] : dq[VP8R_DQ_UV_DC]);
    int ac = b == 0 ? dq[VP8R_DQ_Y2_AC] : (b <= 16 ? dq[VP8R_DQ_Y1_AC] : dq[VP8R_DQ_UV_AC]);
    coef[b][0] = (int16_t)(coef[b][0] * dc);
    for (int i = 1; i < 16; i++) coef[b][i] = (int16_t)(coef[b][i] * ac);
  }
  out->dc_only = 0;
  if (has_y2) oracle_iwht4x4(coef[0]);
  for (int p = 0; p < 24; p++) {
    int16_t *c = coef[p + 1];
    int ac_zero = 1;
    for (int i = 1; i < 16; i++)
      if (c[i]) ac_zero = 0;
    if (p < 16 && has_y2) c[0] = coef[0][p]; /* src/residual.cc:113 */
    if (ac_zero) {
      out->dc_only |= 1u << p;
      memset(out->blk[p], 0, 32);
      out->blk[p][0] = c[0];
    } else {
      oracle_idct4x4(c);
      memcpy(out->blk[p], c, 32);
    }
  }
}

/* src/residual.cc:139-157 */
static void add_block(uint8_t *dst, int stride, const int16_t *res, int dc_only) {
  if (!dc_only) {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) dst[i * stride + j] = (uint8_t)clamp255((int16_t)(dst[i * stride + j] + res[i * 4 + j]));
  } else {
    int16_t v = (int16_t)((res[0] + 4) >> 3);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) dst[i * stride + j] = (uint8_t)clamp255((int16_t)(dst[i * stride + j] + v));
  }
}

/* ------------------------------------------------------------------ intra prediction -------- */

/* 16x16 luma / 8x8 chroma predictors, src/intra_predict.cc:6-98.  n = 16 or 8. */
static void predict_mb_plane(uint8_t *p, int stride, int n, int mode, int r, int c) {
  uint8_t above[16], left[16];
  for (int i = 0; i < n; i++) {
    above[i] = r == 0 ? 127 : p[-stride + i];
    left[i] = c == 0 ? 129 : p[i * stride - 1];
  }
  switch (mode) {
    case 1: /* V_PRED */
      for (int i = 0; i < n; i++) memcpy(p + i * stride, above, (size_t)n);
      break;
    case 2: /* H_PRED */
      for (int i = 0; i < n; i++) memset(p + i * stride, left[i], (size_t)n);
      break;
    case 0: { /* DC_PRED */
      int v;
      if (r == 0 && c == 0) {
        v = 128;
      } else {
        int sum = 0, shf = n == 16 ? 3 : 2;
        if (r > 0) {
          for (int i = 0; i < n; i++) sum += above[i];
          shf++;
        }
        if (c > 0) {
          for (int i = 0; i < n; i++) sum += left[i];
          shf++;
        }
        v = (sum + (1 << (shf - 1))) >> shf;
      }
      for (int i = 0; i < n; i++) memset(p + i * stride, v, (size_t)n);
      break;
    }
    default: { /* TM_PRED */
      int P = r == 0 ? 127 : (c == 0 ? 129 : p[-stride - 1]);
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) p[i * stride + j] = (uint8_t)clamp255(left[i] + above[j] - P);
    }
  }
}

/* The ten 4x4 predictors, src/intra_predict.cc:193-351.  A = above[0..7], L = left[0..3]. */
static void predict_b(uint8_t *d, int stride, int mode, const int *A, const int *L, int P) {
  int E[9] = {L[3], L[2], L[1], L[0], P, A[0], A[1], A[2], A[3]};
  int o[4][4];
#define AVG3(x, y, z) (((x) + (y) + (y) + (z) + 2) >> 2)
#define AVG2(x, y) (((x) + (y) + 1) >> 1)
  switch (mode) {
    case 2: /* B_VE_PRED */
      for (int j = 0; j < 4; j++) {
        int v = AVG3(j == 0 ? P : A[j - 1], A[j], A[j + 1]);
        for (int i = 0; i < 4; i++) o[i][j] = v;
      }
      break;
    case 3: /* B_HE_PRED */
      for (int i = 0; i < 4; i++) {
        int v = AVG3(i == 0 ? P : L[i - 1], L[i], i == 3 ? L[3] : L[i + 1]);
        for (int j = 0; j < 4; j++) o[i][j] = v;
      }
      break;
    case 0: { /* B_DC_PRED */
      int v = 4;
      for (int i = 0; i < 4; i++) v += A[i] + L[i];
      v >>= 3;
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) o[i][j] = v;
      break;
    }
    case 1: /* B_TM_PRED */
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) o[i][j] = clamp255(L[i] + A[j] - P);
      break;
    case 4: /* B_LD_PRED */
      for (int dgl = 0; dgl < 7; dgl++) {
        int v = AVG3(A[dgl], A[dgl + 1], dgl + 2 < 8 ? A[dgl + 2] : A[7]);
        for (int i = 0; i < 4; i++) {
          int j = dgl - i;
          if (j >= 0 && j < 4) o[i][j] = v;
        }
      }
      break;
    case 5: /* B_RD_PRED */
      o[3][0] = AVG3(E[0], E[1], E[2]);
      o[3][1] = o[2][0] = AVG3(E[1], E[2], E[3]);
      o[3][2] = o[2][1] = o[1][0] = AVG3(E[2], E[3], E[4]);
      o[3][3] = o[2][2] = o[1][1] = o[0][0] = AVG3(E[3], E[4], E[5]);
      o[2][3] = o[1][2] = o[0][1] = AVG3(E[4], E[5], E[6]);
      o[1][3] = o[0][2] = AVG3(E[5], E[6], E[7]);
      o[0][3] = AVG3(E[6], E[7], E[8]);
      break;
    case 6: /* B_VR_PRED */
      o[3][0] = AVG3(E[1], E[2], E[3]);
      o[2][0] = AVG3(E[2], E[3], E[4]);
      o[3][1] = o[1][0] = AVG3(E[3], E[4], E[5]);
      o[2][1] = o[0][0] = AVG2(E[4], E[5]);
      o[3][2] = o[1][1] = AVG3(E[4], E[5], E[6]);
      o[2][2] = o[0][1] = AVG2(E[5], E[6]);
      o[3][3] = o[1][2] = AVG3(E[5], E[6], E[7]);
      o[2][3] = o[0][2] = AVG2(E[6], E[7]);
      o[1][3] = AVG3(E[6], E[7], E[8]);
      o[0][3] = AVG2(E[7], E[8]);
      break;
    case 7: /* B_VL_PRED */
      o[0][0] = AVG2(A[0], A[1]);
      o[1][0] = AVG3(A[0], A[1], A[2]);
      o[2][0] = o[0][1] = AVG2(A[1], A[2]);
      o[1][1] = o[3][0] = AVG3(A[1], A[2], A[3]);
      o[2][1] = o[0][2] = AVG2(A[2], A[3]);
      o[3][1] = o[1][2] = AVG3(A[2], A[3], A[4]);
      o[2][2] = o[0][3] = AVG2(A[3], A[4]);
      o[3][2] = o[1][3] = AVG3(A[3], A[4], A[5]);
      o[2][3] = AVG3(A[4], A[5], A[6]);
      o[3][3] = AVG3(A[5], A[6], A[7]);
      break;
    case 8: /* B_HD_PRED */
      o[3][0] = AVG2(E[0], E[1]);
      o[3][1] = AVG3(E[0], E[1], E[2]);
      o[2][0] = o[3][2] = AVG2(E[1], E[2]);
      o[2][1] = o[3][3] = AVG3(E[1], E[2], E[3]);
      o[2][2] = o[1][0] = AVG2(E[2], E[3]);
      o[2][3] = o[1][1] = AVG3(E[2], E[3], E[4]);
      o[1][2] = o[0][0] = AVG2(E[3], E[4]);
      o[1][3] = o[0][1] = AVG3(E[3], E[4], E[5]);
      o[0][2] = AVG3(E[4], E[5], E[6]);
      o[0][3] = AVG3(E[5], E[6], E[7]);
      break;
    default: /* 9: B_HU_PRED */
      o[0][0] = AVG2(L[0], L[1]);
      o[0][1] = AVG3(L[0], L[1], L[2]);
      o[0][2] = o[1][0] = AVG2(L[1], L[2]);
      o[0][3] = o[1][1] = AVG3(L[1], L[2], L[3]);
      o[1][2] = o[2][0] = AVG2(L[2], L[3]);
      o[1][3] = o[2][1] = AVG3(L[2], L[3], L[3]);
      o[2][2] = o[2][3] = o[3][0] = o[3][1] = o[3][2] = o[3][3] = L[3];
      break;
  }
#undef AVG3
#undef AVG2
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) d[i * stride + j] = (uint8_t)o[i][j];
}

/* B_PRED macroblock: edge gathering rules of src/intra_predict.cc:100-191. */
static void predict_bpred(ofrm *fr, int r, int c, const vp8r_mb_info *mb, const mbres *res) {
  int s = fr->ys;
  uint8_t *base = fr->y + (size_t)(16 * r) * s + 16 * c;
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) {
      uint8_t *d = base + 4 * i * s + 4 * j;
      int A[8], L[4], P;
      for (int k = 0; k < 4; k++) A[k] = (i == 0 && r == 0) ? 127 : d[-s + k];
      if (j == 3) {
        /* every row of the right-most column uses the macroblock row above (intra_predict.cc:127-135) */
        for (int k = 0; k < 4; k++) {
          if (r == 0) A[4 + k] = 127;
          else if (c + 1 == fr->cols) A[4 + k] = base[-s + 15];
          else A[4 + k] = base[-s + 16 + k];
        }
      } else {
        for (int k = 0; k < 4; k++) A[4 + k] = (i == 0 && r == 0) ? 127 : d[-s + 4 + k];
      }
      for (int k = 0; k < 4; k++) L[k] = (j == 0 && c == 0) ? 129 : d[k * s - 1];
      if (i > 0 && j > 0) P = d[-s - 1];
      else if (i > 0) P = c == 0 ? 129 : d[-s - 1];
      else if (j > 0) P = r == 0 ? 127 : d[-s - 1];
      else P = r == 0 ? 127 : (c == 0 ? 129 : d[-s - 1]);
      int b = i * 4 + j;
      int mode = (int)((mb->aux[b >> 3] >> ((b & 7) * 4)) & 15);
      predict_b(d, s, mode, A, L, P);
      add_block(d, s, res->blk[b], (res->dc_only >> b) & 1);
    }
  }
}

/* ------------------------------------------------------------------ inter prediction -------- */

static const int16_t k_sixtap[8][6] = { /* src/inter_predict.h:18-27 */
    {0, 0, 128, 0, 0, 0},  {0, -6, 123, 12, -1, 0},   {2, -11, 108, 36, -8, 1}, {0, -9, 93, 50, -6, 0},
    {3, -16, 77, 77, -16, 3}, {0, -6, 50, 93, -9, 0}, {1, -8, 36, 108, -11, 2}, {0, -1, 12, 123, -6, 0}};
static const int16_t k_bilinear[8][6] = { /* src/inter_predict.h:29-36 */
    {0, 0, 128, 0, 0, 0}, {0, 0, 112, 16, 0, 0}, {0, 0, 96, 32, 0, 0}, {0, 0, 80, 48, 0, 0},
    {0, 0, 64, 64, 0, 0}, {0, 0, 48, 80, 0, 0}, {0, 0, 32, 96, 0, 0}, {0, 0, 16, 112, 0, 0}};

static int ref_px(const uint8_t *pl, int stride, int h, int w, int y, int x) { /* inter_predict.cc:252-256 */
  if (y < 0) y = 0;
  if (y > h - 1) y = h - 1;
  if (x < 0) x = 0;
  if (x > w - 1) x = w - 1;
  return pl[(size_t)y * stride + x];
}

/* One 4x4 block, src/inter_predict.cc:246-333. (y0,x0) = block origin, mv in 1/8 pel. */
static void predict_inter4x4(uint8_t *dst, int dstride, const uint8_t *ref, int rstride, int h, int w,
                             int y0, int x0, int mvr, int mvc, const int16_t (*filt)[6]) {
  int fr = mvr & 7, fc = mvc & 7;
  int ty = y0 + (mvr >> 3), tx = x0 + (mvc >> 3);
  if (!(fr | fc)) {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) dst[i * dstride + j] = (uint8_t)ref_px(ref, rstride, h, w, ty + i, tx + j);
    return;
  }
  int tmp[9][4];
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 4; j++) {
      int sum = 0;
      for (int k = 0; k < 6; k++) sum += ref_px(ref, rstride, h, w, ty - 2 + i, tx - 2 + j + k) * filt[fc][k];
      tmp[i][j] = clamp255((int16_t)((sum + 64) >> 7));
    }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      int sum = 0;
      for (int k = 0; k < 6; k++) sum += tmp[i + k][j] * filt[fr][k];
      dst[i * dstride + j] = (uint8_t)clamp255((int16_t)((sum + 64) >> 7));
    }
}

static void predict_inter_mb(const vp8r_frame_desc *f, ofrm *cur, const ofrm *ref, int r, int c,
                             const vp8r_mb_info *mb, const mbres *res) {
  int16_t mv[16][2];
  int split = ((mb->flags >> VP8R_MB_MODE_SHIFT) & 7) == 4;
  if (split) {
    const int16_t *p = f->payload + (size_t)mb->aux[0] * 16;
    for (int b = 0; b < 16; b++) {
      mv[b][0] = p[2 * b];
      mv[b][1] = p[2 * b + 1];
    }
  } else {
    for (int b = 0; b < 16; b++) {
      mv[b][0] = mb->mv[0];
      mv[b][1] = mb->mv[1];
    }
  }
  const int16_t (*filt)[6] = f->hdr.version == 0 ? k_sixtap : k_bilinear; /* inter_predict.cc:348-349 */
  for (int b = 0; b < 16; b++) {
    int i = b >> 2, j = b & 3;
    uint8_t *d = cur->y + (size_t)(16 * r + 4 * i) * cur->ys + 16 * c + 4 * j;
    predict_inter4x4(d, cur->ys, ref->y, ref->ys, ref->yh, ref->ys, 16 * r + 4 * i, 16 * c + 4 * j, mv[b][0],
                     mv[b][1], filt);
    add_block(d, cur->ys, res->blk[b], (res->dc_only >> b) & 1);
  }
  /* Chroma MVs: src/inter_predict.cc:116-144 (the clamp result is discarded there). */
  for (int b = 0; b < 4; b++) {
    int i = b >> 1, j = b & 1;
    int k0 = (2 * i) * 4 + 2 * j;
    int16_t sr = (int16_t)(mv[k0][0] + mv[k0 + 1][0] + mv[k0 + 4][0] + mv[k0 + 5][0]);
    int16_t sc = (int16_t)(mv[k0][1] + mv[k0 + 1][1] + mv[k0 + 4][1] + mv[k0 + 5][1]);
    int16_t dr = (int16_t)(sr >= 0 ? (sr + 4) / 8 : (sr - 4) / 8);
    int16_t dc = (int16_t)(sc >= 0 ? (sc + 4) / 8 : (sc - 4) / 8);
    if (f->hdr.version == 3) {
      dr &= ~7;
      dc &= ~7;
    }
    size_t off = (size_t)(8 * r + 4 * i) * cur->cs + 8 * c + 4 * j;
    predict_inter4x4(cur->u + off, cur->cs, ref->u, ref->cs, ref->ch, ref->cs, 8 * r + 4 * i, 8 * c + 4 * j, dr, dc,
                     filt);
    add_block(cur->u + off, cur->cs, res->blk[16 + b], (res->dc_only >> (16 + b)) & 1);
    predict_inter4x4(cur->v + off, cur->cs, ref->v, ref->cs, ref->ch, ref->cs, 8 * r + 4 * i, 8 * c + 4 * j, dr, dc,
                     filt);
    add_block(cur->v + off, cur->cs, res->blk[20 + b], (res->dc_only >> (20 + b)) & 1);
  }
}

/* ------------------------------------------------------------------ loop filter ------------- */

/* src/filter.cc:22-35.  use_outer_taps as there. q[0] = q0, q[-step] = p0. */
static void lf_adjust(uint8_t *q, int step, int use_outer) {
  int p1 = q[-2 * step], p0 = q[-step], q0 = q[0], q1 = q[step];
  int a = clamp128((use_outer ? clamp128(p1 - q1) : 0) + 3 * (q0 - p0));
  int f1 = ((a + 4 > 127) ? 127 : a + 4) >> 3;
  int f2 = ((a + 3 > 127) ? 127 : a + 3) >> 3;
  q[-step] = (uint8_t)clamp255(p0 + f2);
  q[0] = (uint8_t)clamp255(q0 - f1);
  if (!use_outer) {
    a = (f1 + 1) >> 1;
    q[-2 * step] = (uint8_t)clamp255(p1 + a);
    q[step] = (uint8_t)clamp255(q1 - a);
  }
}
static int iabs(int x) { return x < 0 ? -x : x; }
static int lf_mask_normal(const uint8_t *q, int s, int interior, int edge) { /* src/filter.cc:7-12 */
  int p3 = q[-4 * s], p2 = q[-3 * s], p1 = q[-2 * s], p0 = q[-s], q0 = q[0], q1 = q[s], q2 = q[2 * s], q3 = q[3 * s];
  return (iabs(p0 - q0) * 2 + (iabs(p1 - q1) >> 1)) <= edge && iabs(p3 - p2) <= interior &&
         iabs(p2 - p1) <= interior && iabs(p1 - p0) <= interior && iabs(q0 - q1) <= interior &&
         iabs(q1 - q2) <= interior && iabs(q2 - q3) <= interior;
}
static int lf_hev(const uint8_t *q, int s, int thr) { /* src/filter.cc:18-20 */
  return iabs(q[-2 * s] - q[-s]) > thr || iabs(q[s] - q[0]) > thr;
}
static void lf_subblock(uint8_t *q, int s, int hev_thr, int interior, int edge) { /* src/filter.cc:37-44 */
  if (!lf_mask_normal(q, s, interior, edge)) return;
  lf_adjust(q, s, lf_hev(q, s, hev_thr));
}
static void lf_macroblock(uint8_t *q, int s, int hev_thr, int interior, int edge) { /* src/filter.cc:46-67 */
  if (!lf_mask_normal(q, s, interior, edge)) return;
  if (!lf_hev(q, s, hev_thr)) {
    int p2 = q[-3 * s], p1 = q[-2 * s], p0 = q[-s], q0 = q[0], q1 = q[s], q2 = q[2 * s];
    int w = clamp128(clamp128(p1 - q1) + 3 * (q0 - p0));
    int a = (27 * w + 63) >> 7;
    q[0] = (uint8_t)clamp255(q0 - a);
    q[-s] = (uint8_t)clamp255(p0 + a);
    a = (18 * w + 63) >> 7;
    q[s] = (uint8_t)clamp255(q1 - a);
    q[-2 * s] = (uint8_t)clamp255(p1 + a);
    a = (9 * w + 63) >> 7;
    q[2 * s] = (uint8_t)clamp255(q2 - a);
    q[-3 * s] = (uint8_t)clamp255(p2 + a);
  } else {
    lf_adjust(q, s, 1);
  }
}
static void lf_simple(uint8_t *q, int s, int edge) { /* src/filter.cc:14-16,69-71 */
  if ((iabs(q[-s] - q[0]) * 2 + (iabs(q[-2 * s] - q[s]) >> 1)) <= edge) lf_adjust(q, s, 1);
}

/* src/filter.cc:119-149 */
static void lf_limits(int level, int sharp, int key, int *interior, int *hev, int *edge_mb, int *edge_sb) {
  int in = level;
  if (sharp) {
    in >>= (sharp > 4) ? 2 : 1;
    if (in > 9 - sharp) in = 9 - sharp;
  }
  if (in < 1) in = 1;
  int h = 0;
  if (key) {
    if (level >= 40) h = 2;
    else if (level >= 15) h = 1;
  } else {
    if (level >= 40) h = 3;
    else if (level >= 20) h = 2;
    else if (level >= 15) h = 1;
  }
  *interior = in;
  *hev = h;
  *edge_mb = (level + 2) * 2 + in;
  *edge_sb = level * 2 + in;
}

/* One plane, raster MB order; n = 16 (luma) or 8 (chroma). src/filter.cc:151-318. */
static void lf_plane(const vp8r_frame_desc *f, uint8_t *pl, int stride, int n, int simple) {
  int cols = f->hdr.mb_cols, rows = f->hdr.mb_rows;
  if (f->hdr.loop_filter_level == 0) return;
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      const vp8r_mb_info *mb = &f->mbs[r * cols + c];
      int level = (int)((mb->flags >> VP8R_MB_LF_SHIFT) & 63);
      if (!level) continue;
      int inner = (mb->flags & VP8R_MB_LF_INNER) != 0;
      int in, hev, emb, esb;
      lf_limits(level, f->hdr.sharpness_level, f->hdr.key_frame, &in, &hev, &emb, &esb);
      uint8_t *p = pl + (size_t)(n * r) * stride + n * c;
      if (c > 0)
        for (int y = 0; y < n; y++) {
          if (simple) lf_simple(p + y * stride, 1, emb);
          else lf_macroblock(p + y * stride, 1, hev, in, emb);
        }
      if (inner)
        for (int x = 4; x < n; x += 4)
          for (int y = 0; y < n; y++) {
            if (simple) lf_simple(p + y * stride + x, 1, esb);
            else lf_subblock(p + y * stride + x, 1, hev, in, esb);
          }
      if (r > 0)
        for (int x = 0; x < n; x++) {
          if (simple) lf_simple(p + x, stride, emb);
          else lf_macroblock(p + x, stride, hev, in, emb);
        }
      if (inner)
        for (int y = 4; y < n; y += 4)
          for (int x = 0; x < n; x++) {
            if (simple) lf_simple(p + y * stride + x, stride, esb);
            else lf_subblock(p + y * stride + x, stride, hev, in, esb);
          }
    }
}

/* ------------------------------------------------------------------ frame driver ------------ */

static int frm_alloc(ofrm *f, int cols, int rows) {
  if (f->cols == cols && f->rows == rows && f->y) return 0;
  free(f->y);
  f->cols = cols;
  f->rows = rows;
  f->ys = cols * 16;
  f->cs = cols * 8;
  f->yh = rows * 16;
  f->ch = rows * 8;
  size_t ysz = (size_t)f->ys * f->yh, csz = (size_t)f->cs * f->ch;
  f->y = (uint8_t *)calloc(ysz + 2 * csz, 1);
  if (!f->y) return -1;
  f->u = f->y + ysz;
  f->v = f->u + csz;
  return 0;
}

oracle_decoder *oracle_create(void) { return (oracle_decoder *)calloc(1, sizeof(oracle_decoder)); }

void oracle_destroy(oracle_decoder *d) {
  if (!d) return;
  for (int i = 0; i < 5; i++) free(d->pool[i].y);
  free(d);
}

int oracle_decode_frame(oracle_decoder *d, const vp8r_frame_desc *f) {
  const vp8r_frame_hdr *h = &f->hdr;
  int cols = h->mb_cols, rows = h->mb_rows;
  if (!h->key_frame && !d->have_frame) return -1;
  /* CURRENT = a buffer that no reference points at (the reference allocates a fresh Frame,
   * src/decode.cc:70-71). */
  ofrm *cur = NULL;
  for (int i = 0; i < 5 && !cur; i++) {
    ofrm *cand = &d->pool[i];
    if (cand != d->ref[1] && cand != d->ref[2] && cand != d->ref[3]) cur = cand;
  }
  if (frm_alloc(cur, cols, rows)) return -2;
  d->ref[0] = cur;
  d->width = h->width;
  d->height = h->height;

  mbres res;
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      const vp8r_mb_info *mb = &f->mbs[r * cols + c];
      build_residual(f, mb, &res);
      if (mb->flags & VP8R_MB_IS_INTER) {
        const ofrm *ref = d->ref[(mb->flags >> VP8R_MB_REF_SHIFT) & 3];
        if (!ref || ref->cols != cols || ref->rows != rows) return -3;
        predict_inter_mb(f, cur, ref, r, c, mb, &res);
      } else { /* src/intra_predict.cc:355-428 */
        int ymode = (int)((mb->flags >> VP8R_MB_MODE_SHIFT) & 7);
        int uvmode = (int)((mb->flags >> VP8R_MB_UVMODE_SHIFT) & 3);
        uint8_t *py = cur->y + (size_t)(16 * r) * cur->ys + 16 * c;
        if (ymode == 4) {
          predict_bpred(cur, r, c, mb, &res);
        } else {
          predict_mb_plane(py, cur->ys, 16, ymode, r, c);
          for (int b = 0; b < 16; b++)
            add_block(py + (b >> 2) * 4 * cur->ys + (b & 3) * 4, cur->ys, res.blk[b], (res.dc_only >> b) & 1);
        }
        size_t coff = (size_t)(8 * r) * cur->cs + 8 * c;
        predict_mb_plane(cur->u + coff, cur->cs, 8, uvmode, r, c);
        predict_mb_plane(cur->v + coff, cur->cs, 8, uvmode, r, c);
        for (int b = 0; b < 4; b++) {
          size_t o = coff + (size_t)(b >> 1) * 4 * cur->cs + (b & 1) * 4;
          add_block(cur->u + o, cur->cs, res.blk[16 + b], (res.dc_only >> (16 + b)) & 1);
          add_block(cur->v + o, cur->cs, res.blk[20 + b], (res.dc_only >> (20 + b)) & 1);
        }
      }
    }

  /* src/filter.cc:324-340 */
  if (!h->filter_type) {
    lf_plane(f, cur->y, cur->ys, 16, 0);
    lf_plane(f, cur->u, cur->cs, 8, 0);
    lf_plane(f, cur->v, cur->cs, 8, 0);
  } else {
    lf_plane(f, cur->y, cur->ys, 16, 1);
  }

  /* src/loop.h:19-46 */
  int g2a = !h->refresh_altref && h->copy_to_altref == 2;
  int a2g = !h->refresh_golden && h->copy_to_golden == 2;
  if (g2a && a2g) {
    ofrm *t = d->ref[2];
    d->ref[2] = d->ref[3];
    d->ref[3] = t;
  } else if (g2a) {
    d->ref[3] = d->ref[2];
  } else if (a2g) {
    d->ref[2] = d->ref[3];
  }
  if (h->refresh_golden) d->ref[2] = cur;
  else if (h->copy_to_golden == 1) d->ref[2] = d->ref[1];
  if (h->refresh_altref) d->ref[3] = cur;
  else if (h->copy_to_altref == 1) d->ref[3] = d->ref[1];
  if (h->refresh_last) d->ref[1] = cur;
  d->have_frame = 1;
  return 0;
}

size_t oracle_frame_bytes(const oracle_decoder *d) {
  if (!d->have_frame) return 0;
  size_t cw = (size_t)(d->width + 1) / 2, chh = (size_t)(d->height + 1) / 2;
  return (size_t)d->width * d->height + 2 * cw * chh;
}

/* src/yuv.cc:6-28 */
size_t oracle_write_i420(const oracle_decoder *d, uint8_t *dst, size_t cap) {
  size_t need = oracle_frame_bytes(d);
  if (!need || cap < need) return 0;
  const ofrm *f = d->ref[0];
  int w = d->width, h = d->height, cw = (w + 1) / 2, chh = (h + 1) / 2;
  for (int r = 0; r < h; r++, dst += w) memcpy(dst, f->y + (size_t)r * f->ys, (size_t)w);
  for (int r = 0; r < chh; r++, dst += cw) memcpy(dst, f->u + (size_t)r * f->cs, (size_t)cw);
  for (int r = 0; r < chh; r++, dst += cw) memcpy(dst, f->v + (size_t)r * f->cs, (size_t)cw);
  return need;
}

/* ------------------------------------------------------------------ encoder side (row f4) ---- */
#include "enc_oracle.c"
