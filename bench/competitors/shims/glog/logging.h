// Minimal stand-in for <glog/logging.h>, enough for RoDe's utils/matrix_utils.cu (CHECK_xx(...) << "message").
// glog is not installed in this image; the competitor harness (bench/competitors/build.py) puts this directory on the
// include path instead.  A failed check prints the message and aborts, like glog's.
#ifndef VOLTRIX_BENCH_GLOG_SHIM_H_
#define VOLTRIX_BENCH_GLOG_SHIM_H_
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace vx_glog_shim {
class Fatal {
 public:
  Fatal(const char *file, int line, const char *expr) { s_ << file << ":" << line << " Check failed: " << expr << " "; }
  [[noreturn]] ~Fatal() { std::cerr << s_.str() << std::endl; std::abort(); }
  template <typename T> Fatal &operator<<(const T &v) { s_ << v; return *this; }
 private:
  std::ostringstream s_;
};
struct Voidify { void operator&(const Fatal &) {} void operator&(std::ostream &) {} };
}  // namespace vx_glog_shim

#define VX_SHIM_CHECK(cond, text) (cond) ? (void)0 : ::vx_glog_shim::Voidify() & ::vx_glog_shim::Fatal(__FILE__, __LINE__, text)
#define CHECK(c) VX_SHIM_CHECK((c), #c)
#define CHECK_EQ(a, b) VX_SHIM_CHECK((a) == (b), #a " == " #b)
#define CHECK_NE(a, b) VX_SHIM_CHECK((a) != (b), #a " != " #b)
#define CHECK_LE(a, b) VX_SHIM_CHECK((a) <= (b), #a " <= " #b)
#define CHECK_LT(a, b) VX_SHIM_CHECK((a) < (b), #a " < " #b)
#define CHECK_GE(a, b) VX_SHIM_CHECK((a) >= (b), #a " >= " #b)
#define CHECK_GT(a, b) VX_SHIM_CHECK((a) > (b), #a " > " #b)
#define LOG(severity) std::cerr
#endif
