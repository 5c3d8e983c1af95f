/*
 * recon_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar C restatement of the reference decoder's frame-reconstruction path (the functions that
 * libvp8r's CUDA kernels replace), consuming the same parsed-frame arrays (include/vp8r.h).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py drives host parser -> this oracle over
 * all 43 vp8-test-vector streams shipped with the reference and checks every frame's MD5
 * against the .ivf.md5 goldens (700 frames), and against the compiled reference decoder
 * (oracle/_ref/decode) on the synthetic streams.
 */
#ifndef RECON_ORACLE_H_
#define RECON_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#include "vp8r.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_decoder oracle_decoder;

oracle_decoder *oracle_create(void);
void oracle_destroy(oracle_decoder *d);

/* DecodeFrame (src/decode_frame.cc:175-187) + RefreshRefFrames (src/loop.h:19-46). 0 on success. */
int oracle_decode_frame(oracle_decoder *d, const vp8r_frame_desc *f);

/* YUV<WRITE>::WriteFrame (src/yuv.cc:6-28) of the most recently decoded frame. Returns bytes. */
size_t oracle_frame_bytes(const oracle_decoder *d);
size_t oracle_write_i420(const oracle_decoder *d, uint8_t *dst, size_t cap);

/* Stand-alone pieces, exported for unit tests (test/dct_test.h:16-71 of the reference). */
void oracle_idct4x4(int16_t blk[16]);
void oracle_iwht4x4(int16_t blk[16]);

/* Encoder side, see enc_oracle.c.  PINNED on tests/golden/enc/enc_goldens.json. */
void oracle_fdct4x4(int16_t blk[16]);
void oracle_fwht4x4(int16_t blk[16]);
void oracle_quantize(int16_t blk[16], int dc, int ac);
int oracle_pick_mb_mode(const uint8_t *target, int ts, uint8_t *p, int stride, int n, int r, int c, uint32_t err[4]);
int oracle_pick_chroma_mode(const uint8_t *tu, const uint8_t *tv, int ts, uint8_t *pu, uint8_t *pv, int stride, int r, int c);
int oracle_pick_sub_mode(const int above[8], const int left[4], int p, const uint8_t target[16], uint32_t err[10]);
int oracle_encode_key_frame(const uint8_t *sy, const uint8_t *su, const uint8_t *sv, int cols, int rows, const int16_t dq[6],
                            int lf_level, vp8r_mb_info *mbs, int16_t *payload, uint8_t *ry, uint8_t *ru, uint8_t *rv);
int oracle_encode_key_frame2(const uint8_t *sy, const uint8_t *su, const uint8_t *sv, int cols, int rows, const int16_t dq[6],
                            int lf_level, vp8r_mb_info *mbs, int16_t *payload, uint8_t *ry, uint8_t *ru, uint8_t *rv, int bpred_search);

#ifdef __cplusplus
}
#endif
#endif
