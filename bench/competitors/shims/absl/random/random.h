// Minimal stand-in for <absl/random/random.h>: RoDe's matrix utilities only draw uniform values from an absl::BitGen.
// (abseil is vendored in the reference as source, but building it needs its own cmake project.)
#ifndef VOLTRIX_BENCH_ABSL_RANDOM_SHIM_H_
#define VOLTRIX_BENCH_ABSL_RANDOM_SHIM_H_
#include <random>
#include <type_traits>

namespace absl {
class BitGen {
 public:
  using result_type = std::mt19937_64::result_type;
  BitGen() : eng_(0x5eedULL) {}
  static constexpr result_type min() { return std::mt19937_64::min(); }
  static constexpr result_type max() { return std::mt19937_64::max(); }
  result_type operator()() { return eng_(); }
 private:
  std::mt19937_64 eng_;
};
template <typename T, typename A, typename B>
T Uniform(BitGen &gen, A lo, B hi) {
  if constexpr (std::is_integral<T>::value) {
    return std::uniform_int_distribution<T>(static_cast<T>(lo), static_cast<T>(hi) - 1)(gen);
  } else {
    return static_cast<T>(std::uniform_real_distribution<double>(static_cast<double>(lo), static_cast<double>(hi))(gen));
  }
}
}  // namespace absl
#endif
