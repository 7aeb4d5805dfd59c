// util/kaldi-io.h -- shim: ReadKaldiObject / WriteKaldiObject on plain files
// (no rxfilename pipes / offsets).  A file starting with "\0B" is binary.
#ifndef KALDI_UTIL_KALDI_IO_H_
#define KALDI_UTIL_KALDI_IO_H_
#include <fstream>
#include <string>
#include "base/kaldi-common.h"
namespace kaldi {
inline bool InitKaldiInputStream(std::istream &is, bool *binary) {
  if (is.peek() == '\0') {
    is.get();
    if (is.peek() != 'B') return false;
    is.get();
    *binary = true;
  } else {
    *binary = false;
  }
  return true;
}
inline void InitKaldiOutputStream(std::ostream &os, bool binary) {
  if (binary) { os.put('\0'); os.put('B'); }
  if (os.precision() < 7) os.precision(7);
}
template <class C>
void ReadKaldiObject(const std::string &filename, C *c) {
  std::ifstream is(filename.c_str(), std::ios::binary);
  if (!is.is_open()) KALDI_ERR << "Could not open " << filename;
  bool binary = false;
  if (!InitKaldiInputStream(is, &binary)) KALDI_ERR << "Bad header in " << filename;
  c->Read(is, binary);
}
template <class C>
void WriteKaldiObject(const C &c, const std::string &filename, bool binary) {
  std::ofstream os(filename.c_str(), std::ios::binary);
  if (!os.is_open()) KALDI_ERR << "Could not open " << filename << " for writing";
  InitKaldiOutputStream(os, binary);
  c.Write(os, binary);
}
}  // namespace kaldi
#endif
