// util/text-utils.h -- shim: the string helpers the component config parsers use.
#ifndef KALDI_UTIL_TEXT_UTILS_H_
#define KALDI_UTIL_TEXT_UTILS_H_

#include <cerrno>
#include <cstdlib>
#include <string>
#include <vector>

#include "base/kaldi-common.h"

namespace kaldi {

inline void SplitStringToVector(const std::string &full, const char *delim, bool omit_empty_strings,
                                std::vector<std::string> *out) {
  size_t start = 0, found = 0, end = full.size();
  out->clear();
  while (found != std::string::npos) {
    found = full.find_first_of(delim, start);
    if (!omit_empty_strings || (found != start && start != end))
      out->push_back(full.substr(start, found - start));
    start = found + 1;
  }
}

template <class Int>
bool ConvertStringToInteger(const std::string &str, Int *out) {
  const char *this_str = str.c_str();
  char *end = NULL;
  errno = 0;
  long long i = strtoll(this_str, &end, 10);
  if (end != this_str) while (isspace(*end)) end++;
  if (end == this_str || *end != '\0' || errno != 0) return false;
  Int iInt = static_cast<Int>(i);
  if (static_cast<long long>(iInt) != i) return false;
  *out = iInt;
  return true;
}

template <class T>
bool ConvertStringToReal(const std::string &str, T *out) {
  const char *this_str = str.c_str();
  char *end = NULL;
  errno = 0;
  double d = strtod(this_str, &end);
  if (end != this_str) while (isspace(*end)) end++;
  if (end == this_str || *end != '\0' || errno != 0) return false;
  *out = static_cast<T>(d);
  return true;
}

template <class I>
bool SplitStringToIntegers(const std::string &full, const char *delim, bool omit_empty_strings,
                           std::vector<I> *out) {
  if (*(full.c_str()) == '\0') { out->clear(); return true; }
  std::vector<std::string> split;
  SplitStringToVector(full, delim, omit_empty_strings, &split);
  out->resize(split.size());
  for (size_t i = 0; i < split.size(); i++)
    if (!ConvertStringToInteger(split[i], &((*out)[i]))) return false;
  return true;
}

template <class F>
bool SplitStringToFloats(const std::string &full, const char *delim, bool omit_empty_strings,
                         std::vector<F> *out) {
  if (*(full.c_str()) == '\0') { out->clear(); return true; }
  std::vector<std::string> split;
  SplitStringToVector(full, delim, omit_empty_strings, &split);
  out->resize(split.size());
  for (size_t i = 0; i < split.size(); i++)
    if (!ConvertStringToReal(split[i], &((*out)[i]))) return false;
  return true;
}

}  // namespace kaldi
#endif
