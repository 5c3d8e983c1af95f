// Thread-local error message behind vp8r_last_error().
#ifndef VP8R_RT_ERROR_H_
#define VP8R_RT_ERROR_H_
#include <string>
namespace vp8r {
void SetError(const std::string &msg);
const char *LastError();
}  // namespace vp8r
#endif
