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

#ifdef __cplusplus
}
#endif
#endif
