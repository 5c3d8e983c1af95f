// Key-frame bitstream writer: the inverse of FrameParser for the frames the encoder produces (row f4).
#ifndef VP8R_HOST_FRAME_WRITER_H_
#define VP8R_HOST_FRAME_WRITER_H_

#include <cstdint>
#include <string>
#include <vector>

#include "parsed_frame.h"

namespace vp8r {

// Serialises a parsed-frame structure (key frame, intra macroblocks with 16x16 or B_PRED luma modes, one
// quantiser, no segmentation, no loop-filter deltas, one DCT partition) into a VP8 frame (RFC 6386 sections 9,
// 19.2, 19.3, 13): frame tag + start code + dimensions, first partition (headers, "keep the default token
// probabilities", per-macroblock skip flag and modes), token partition with the default probabilities.
// Parsing the result with FrameParser gives back the same macroblock records and coefficient blocks.
// Returns VP8R_OK or an error code with *err set.
int WriteKeyFrame(const vp8r_frame &f, std::vector<uint8_t> *out, std::string *err);

}  // namespace vp8r

#endif  // VP8R_HOST_FRAME_WRITER_H_
