// base/kaldi-error.h -- shim: KALDI_ASSERT aborts (as in Kaldi), KALDI_ERR throws
// std::runtime_error, KALDI_WARN / KALDI_LOG print to stderr.
#ifndef KALDI_BASE_KALDI_ERROR_H_
#define KALDI_BASE_KALDI_ERROR_H_

#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>

namespace kaldi {

class MessageLogger {
 public:
  enum Severity { kError, kWarning, kInfo };
  MessageLogger(Severity s, const char *func, const char *file, int line) : sev_(s) {
    ss_ << (s == kError ? "ERROR (" : s == kWarning ? "WARNING (" : "LOG (") << func << "():"
        << Basename(file) << ":" << line << ") ";
  }
  ~MessageLogger() noexcept(false) {
    if (sev_ == kError) throw std::runtime_error(ss_.str());
    std::fprintf(stderr, "%s\n", ss_.str().c_str());
  }
  std::ostream &stream() { return ss_; }

 private:
  static const char *Basename(const char *f) {
    const char *b = f;
    for (const char *p = f; *p; ++p) if (*p == '/') b = p + 1;
    return b;
  }
  Severity sev_;
  std::ostringstream ss_;
};

// Set by the C API so a failed assertion can be reported to a foreign caller
// instead of aborting the host process (kcnn_capi.h: kcnn_last_error()).
extern bool g_assert_throws;

inline void KaldiAssertFailure_(const char *func, const char *file, int line, const char *cond) {
  std::ostringstream ss;
  ss << "KALDI_ASSERT: at " << func << ":" << file << ":" << line << ", failed: " << cond;
  if (g_assert_throws) throw std::runtime_error(ss.str());
  std::fprintf(stderr, "%s\n", ss.str().c_str());
  std::abort();
}

}  // namespace kaldi

#define KALDI_ERR ::kaldi::MessageLogger(::kaldi::MessageLogger::kError, __func__, __FILE__, __LINE__).stream()
#define KALDI_WARN ::kaldi::MessageLogger(::kaldi::MessageLogger::kWarning, __func__, __FILE__, __LINE__).stream()
#define KALDI_LOG ::kaldi::MessageLogger(::kaldi::MessageLogger::kInfo, __func__, __FILE__, __LINE__).stream()
#define KALDI_ASSERT(cond) \
  do { if (!(cond)) ::kaldi::KaldiAssertFailure_(__func__, __FILE__, __LINE__, #cond); } while (0)

#endif
