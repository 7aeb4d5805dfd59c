// base/io-funcs.h -- shim: Kaldi's token / basic-type stream I/O, binary and text.
// Byte layout follows Kaldi (base/io-funcs{.h,-inl.h,.cc} upstream): tokens are
// written as "<Token> "; in binary mode an integer or float is a size byte
// followed by the raw little-endian value; bool is 'T' / 'F'.
#ifndef KALDI_BASE_IO_FUNCS_H_
#define KALDI_BASE_IO_FUNCS_H_

#include <cctype>
#include <cstring>
#include <iostream>
#include <limits>
#include <string>
#include <type_traits>

#include "base/kaldi-error.h"
#include "base/kaldi-types.h"

namespace kaldi {

inline void CheckToken(const char *token) {
  KALDI_ASSERT(*token != '\0');
  for (const char *p = token; *p; ++p) KALDI_ASSERT(!::isspace(*p));
}

inline void WriteToken(std::ostream &os, bool /*binary*/, const char *token) {
  CheckToken(token);
  os << token << " ";
  if (os.fail()) KALDI_ERR << "Write failure in WriteToken.";
}
inline void WriteToken(std::ostream &os, bool binary, const std::string &token) {
  WriteToken(os, binary, token.c_str());
}

inline void ReadToken(std::istream &is, bool binary, std::string *str) {
  KALDI_ASSERT(str != NULL);
  if (!binary) is >> std::ws;
  is >> *str;
  if (is.fail()) KALDI_ERR << "ReadToken, failed to read token at file position " << is.tellg();
  if (!isspace(is.peek()))
    KALDI_ERR << "ReadToken, expected space after token, saw instead "
              << static_cast<char>(is.peek()) << ", at file position " << is.tellg();
  is.get();   // consume the space
}

inline int PeekToken(std::istream &is, bool binary) {
  if (!binary) is >> std::ws;
  bool read_bracket = false;
  if (static_cast<char>(is.peek()) == '<') { read_bracket = true; is.get(); }
  int ans = is.peek();
  if (read_bracket) is.unget();
  return ans;
}

inline void ExpectToken(std::istream &is, bool binary, const char *token) {
  int pos_at_start = is.tellg();
  KALDI_ASSERT(token != NULL);
  CheckToken(token);
  if (!binary) is >> std::ws;
  std::string str;
  is >> str;
  is.get();
  if (is.fail()) KALDI_ERR << "Failed to read token [started at file position " << pos_at_start
                           << "], expected " << token;
  if (strcmp(str.c_str(), token) != 0)
    KALDI_ERR << "Expected token \"" << token << "\", got instead \"" << str << "\".";
}
inline void ExpectToken(std::istream &is, bool binary, const std::string &token) {
  ExpectToken(is, binary, token.c_str());
}

template <class T>
inline void WriteBasicType(std::ostream &os, bool binary, T t) {
  if (binary) {
    if (std::is_integral<T>::value) {
      char len_c = (std::numeric_limits<T>::is_signed ? 1 : -1) * static_cast<char>(sizeof(t));
      os.put(len_c);
    } else {
      os.put(static_cast<char>(sizeof(t)));
    }
    os.write(reinterpret_cast<const char *>(&t), sizeof(t));
  } else {
    if (sizeof(t) == 1) os << static_cast<int16>(t) << " ";
    else os << t << " ";
  }
  if (os.fail()) throw std::runtime_error("Write failure in WriteBasicType.");
}
template <>
inline void WriteBasicType<bool>(std::ostream &os, bool binary, bool b) {
  os << (b ? "T" : "F");
  if (!binary) os << " ";
  if (os.fail()) KALDI_ERR << "Write failure in WriteBasicType<bool>";
}

template <class T>
inline void ReadBasicType(std::istream &is, bool binary, T *t) {
  KALDI_ASSERT(t != NULL);
  if (binary) {
    int len_c_in = is.get();
    if (len_c_in == -1) KALDI_ERR << "ReadBasicType: encountered end of stream.";
    char len_c = static_cast<char>(len_c_in), len_c_expected;
    if (std::is_integral<T>::value)
      len_c_expected = (std::numeric_limits<T>::is_signed ? 1 : -1) * static_cast<char>(sizeof(*t));
    else
      len_c_expected = static_cast<char>(sizeof(*t));
    if (len_c != len_c_expected)
      KALDI_ERR << "ReadBasicType: did not get expected integer type, " << static_cast<int>(len_c)
                << " vs. " << static_cast<int>(len_c_expected)
                << ".  You can change this code to successfully read it later, if needed.";
    is.read(reinterpret_cast<char *>(t), sizeof(*t));
  } else {
    if (sizeof(*t) == 1) { int16 i; is >> i; *t = i; }
    else is >> *t;
  }
  if (is.fail()) KALDI_ERR << "Read failure in ReadBasicType, file position is " << is.tellg()
                           << ", next char is " << is.peek();
}
template <>
inline void ReadBasicType<bool>(std::istream &is, bool binary, bool *b) {
  KALDI_ASSERT(b != NULL);
  if (!binary) is >> std::ws;
  char c = is.peek();
  if (c == 'T') { *b = true; is.get(); }
  else if (c == 'F') { *b = false; is.get(); }
  else KALDI_ERR << "Read failure in ReadBasicType<bool>, file position is " << is.tellg()
                 << ", next char is " << c;
}

}  // namespace kaldi

#endif
