// base/timer.h -- shim: wall-clock timer used around kernel launches for AccuProfile.
#ifndef KALDI_BASE_TIMER_H_
#define KALDI_BASE_TIMER_H_
#include <chrono>
namespace kaldi {
class Timer {
 public:
  Timer() { Reset(); }
  void Reset() { t0_ = std::chrono::steady_clock::now(); }
  double Elapsed() const {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
  }
 private:
  std::chrono::steady_clock::time_point t0_;
};
}  // namespace kaldi
#endif
