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
#include <vector>

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
// Floating-point values: a binary stream may hold them as float OR double (a Kaldi built with
// --double-precision writes 8-byte reals); either width is accepted and converted, as upstream does.
template <class Real>
inline void ReadReal(std::istream &is, bool binary, Real *t) {
  KALDI_ASSERT(t != NULL);
  if (binary) {
    const int c = is.peek();
    if (c == static_cast<int>(sizeof(float))) {
      float f;
      is.get();
      is.read(reinterpret_cast<char *>(&f), sizeof(f));
      *t = static_cast<Real>(f);
    } else if (c == static_cast<int>(sizeof(double))) {
      double d;
      is.get();
      is.read(reinterpret_cast<char *>(&d), sizeof(d));
      *t = static_cast<Real>(d);
    } else {
      KALDI_ERR << "ReadBasicType: expected float or double, saw " << c << ", at file position " << is.tellg();
    }
  } else {
    is >> *t;
  }
  if (is.fail()) KALDI_ERR << "ReadBasicType: failed to read, at file position " << is.tellg();
}
template <>
inline void ReadBasicType<float>(std::istream &is, bool binary, float *f) { ReadReal(is, binary, f); }
template <>
inline void ReadBasicType<double>(std::istream &is, bool binary, double *d) { ReadReal(is, binary, d); }

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

// std::vector of integers: binary = size byte, int32 count, raw values; text = "[ 1 2 3 ]\n".
template <class T>
inline void WriteIntegerVector(std::ostream &os, bool binary, const std::vector<T> &v) {
  if (binary) {
    char sz = sizeof(T);
    os.write(&sz, 1);
    int32 vecsz = static_cast<int32>(v.size());
    os.write(reinterpret_cast<const char *>(&vecsz), sizeof(vecsz));
    if (vecsz != 0) os.write(reinterpret_cast<const char *>(&v[0]), sizeof(T) * vecsz);
  } else {
    os << "[ ";
    for (size_t i = 0; i < v.size(); i++) os << v[i] << " ";
    os << "]\n";
  }
  if (os.fail()) KALDI_ERR << "Write failure in WriteIntegerVector.";
}

template <class T>
inline void ReadIntegerVector(std::istream &is, bool binary, std::vector<T> *v) {
  KALDI_ASSERT(v != NULL);
  v->clear();
  if (binary) {
    int sz = is.peek();
    if (sz != static_cast<int>(sizeof(T)))
      KALDI_ERR << "ReadIntegerVector: expected element size " << sizeof(T) << ", saw " << sz;
    is.get();
    int32 vecsz;
    is.read(reinterpret_cast<char *>(&vecsz), sizeof(vecsz));
    if (is.fail() || vecsz < 0) KALDI_ERR << "ReadIntegerVector: bad vector size";
    v->resize(vecsz);
    if (vecsz > 0) is.read(reinterpret_cast<char *>(&(*v)[0]), sizeof(T) * vecsz);
  } else {
    is >> std::ws;
    if (is.peek() != static_cast<int>('[')) KALDI_ERR << "ReadIntegerVector: expected [";
    is.get();
    is >> std::ws;
    while (is.peek() != static_cast<int>(']')) {
      long long x;
      is >> x >> std::ws;
      if (is.fail()) KALDI_ERR << "ReadIntegerVector: bad element";
      v->push_back(static_cast<T>(x));
    }
    is.get();
  }
  if (is.fail()) KALDI_ERR << "ReadIntegerVector: read failure at file position " << is.tellg();
}

}  // namespace kaldi

#endif
