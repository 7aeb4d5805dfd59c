// matrix/kaldi-matrix-io.cc -- shim: Kaldi's matrix / vector stream formats
// ("FM " / "FV " headers in binary mode; " [ ... ]" in text mode).
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include "matrix/matrix-lib.h"

namespace kaldi {

template <typename Real> static const char *MatToken();
template <> const char *MatToken<float>() { return "FM"; }
template <> const char *MatToken<double>() { return "DM"; }
template <typename Real> static const char *VecToken();
template <> const char *VecToken<float>() { return "FV"; }
template <> const char *VecToken<double>() { return "DV"; }

template <typename Real>
void MatrixBase<Real>::Write(std::ostream &os, bool binary) const {
  if (!os.good()) KALDI_ERR << "Failed to write matrix to stream: stream not good";
  if (binary) {
    WriteToken(os, binary, MatToken<Real>());
    int32 rows = rows_, cols = cols_;
    WriteBasicType(os, binary, rows);
    WriteBasicType(os, binary, cols);
    if (stride_ == cols_) {
      os.write(reinterpret_cast<const char *>(data_), sizeof(Real) * (size_t)rows * cols);
    } else {
      for (MatrixIndexT i = 0; i < rows; i++)
        os.write(reinterpret_cast<const char *>(RowData(i)), sizeof(Real) * cols);
    }
    if (!os.good()) KALDI_ERR << "Failed to write matrix to stream";
  } else {
    if (cols_ == 0) {
      os << " [ ]\n";
    } else {
      os << " [";
      for (MatrixIndexT i = 0; i < rows_; i++) {
        os << "\n  ";
        for (MatrixIndexT j = 0; j < cols_; j++) os << (*this)(i, j) << " ";
      }
      os << "]\n";
    }
  }
}

// One number of a text matrix / vector.  Characters are consumed one at a time up to (not including)
// white space or a bracket, so a closing bracket glued to the last number ("1.5]") stays in the stream
// without any putback (an ifstream only guarantees one character of putback).
static bool ParseReal(std::istream &is, double *v) {
  is >> std::ws;
  std::string tok;
  for (;;) {
    const int c = is.peek();
    if (c == EOF) break;
    const char ch = static_cast<char>(c);
    if (isspace(static_cast<unsigned char>(ch)) || ch == ']' || ch == '[') break;
    tok.push_back(ch);
    is.get();
  }
  if (tok.empty()) return false;
  char *end = NULL;
  *v = strtod(tok.c_str(), &end);
  return end == tok.c_str() + tok.size();
}

template <typename Real>
void Matrix<Real>::Read(std::istream &is, bool binary) {
  if (binary) {
    int peekval = is.peek();
    if (peekval == 'C') KALDI_ERR << "Compressed matrices are not supported by this shim";
    std::string token;
    ReadToken(is, binary, &token);
    bool is_double;
    if (token == "FM") is_double = false;
    else if (token == "DM") is_double = true;
    else { KALDI_ERR << "Expected \"FM\" or \"DM\", got \"" << token << "\""; return; }
    int32 rows, cols;
    ReadBasicType(is, binary, &rows);
    ReadBasicType(is, binary, &cols);
    Resize(rows, cols, kUndefined);
    size_t n = (size_t)rows * cols;
    if (is_double == (sizeof(Real) == 8)) {
      is.read(reinterpret_cast<char *>(this->data_), sizeof(Real) * n);
    } else if (is_double) {
      std::vector<double> tmp(n);
      is.read(reinterpret_cast<char *>(tmp.data()), sizeof(double) * n);
      for (size_t i = 0; i < n; i++) this->data_[i] = static_cast<Real>(tmp[i]);
    } else {
      std::vector<float> tmp(n);
      is.read(reinterpret_cast<char *>(tmp.data()), sizeof(float) * n);
      for (size_t i = 0; i < n; i++) this->data_[i] = static_cast<Real>(tmp[i]);
    }
    if (is.fail()) KALDI_ERR << "Failed to read matrix data from stream";
    return;
  }
  std::string str;
  is >> str;
  if (is.fail()) KALDI_ERR << "Failed to read matrix from stream: EOF";
  if (str == "[]") { Resize(0, 0); return; }
  if (str != "[") KALDI_ERR << "Failed to read matrix from stream: expected \"[\", got \"" << str << "\"";
  std::vector<std::vector<Real> > rows;
  std::vector<Real> cur;
  while (true) {
    int c = is.peek();
    if (c == -1) KALDI_ERR << "Failed to read matrix from stream: EOF";
    if (c == ' ' || c == '\t' || c == '\r') { is.get(); continue; }
    if (c == '\n' || c == ';') {
      is.get();
      if (!cur.empty()) { rows.push_back(cur); cur.clear(); }
      continue;
    }
    if (c == ']') {
      is.get();
      if (!cur.empty()) rows.push_back(cur);
      // consume the rest of the line as Kaldi does
      while (is.peek() == ' ' || is.peek() == '\r') is.get();
      if (is.peek() == '\n') is.get();
      break;
    }
    double v;
    if (!ParseReal(is, &v)) KALDI_ERR << "Failed to read matrix from stream: bad number";
    cur.push_back(static_cast<Real>(v));
  }
  MatrixIndexT nr = rows.size(), nc = nr ? rows[0].size() : 0;
  Resize(nr, nc, kUndefined);
  for (MatrixIndexT i = 0; i < nr; i++) {
    if ((MatrixIndexT)rows[i].size() != nc) KALDI_ERR << "Failed to read matrix from stream: ragged rows";
    for (MatrixIndexT j = 0; j < nc; j++) (*this)(i, j) = rows[i][j];
  }
}

template <typename Real>
void VectorBase<Real>::Write(std::ostream &os, bool binary) const {
  if (!os.good()) KALDI_ERR << "Failed to write vector to stream: stream not good";
  if (binary) {
    WriteToken(os, binary, VecToken<Real>());
    int32 size = dim_;
    WriteBasicType(os, binary, size);
    os.write(reinterpret_cast<const char *>(data_), sizeof(Real) * size);
  } else {
    os << " [ ";
    for (MatrixIndexT i = 0; i < dim_; i++) os << data_[i] << " ";
    os << "]\n";
  }
  if (!os.good()) KALDI_ERR << "Failed to write vector to stream";
}

template <typename Real>
void Vector<Real>::Read(std::istream &is, bool binary) {
  if (binary) {
    std::string token;
    ReadToken(is, binary, &token);
    bool is_double;
    if (token == "FV") is_double = false;
    else if (token == "DV") is_double = true;
    else { KALDI_ERR << "Expected \"FV\" or \"DV\", got \"" << token << "\""; return; }
    int32 size;
    ReadBasicType(is, binary, &size);
    Resize(size, kUndefined);
    if (is_double == (sizeof(Real) == 8)) {
      is.read(reinterpret_cast<char *>(this->data_), sizeof(Real) * size);
    } else if (is_double) {
      std::vector<double> tmp(size);
      is.read(reinterpret_cast<char *>(tmp.data()), sizeof(double) * size);
      for (int32 i = 0; i < size; i++) this->data_[i] = static_cast<Real>(tmp[i]);
    } else {
      std::vector<float> tmp(size);
      is.read(reinterpret_cast<char *>(tmp.data()), sizeof(float) * size);
      for (int32 i = 0; i < size; i++) this->data_[i] = static_cast<Real>(tmp[i]);
    }
    if (is.fail()) KALDI_ERR << "Failed to read vector data from stream";
    return;
  }
  std::string s;
  is >> s;
  if (is.fail() || (s != "[" && s != "[]")) KALDI_ERR << "Failed to read vector from stream: expected \"[\"";
  std::vector<Real> data;
  if (s == "[") {
    while (true) {
      is >> std::ws;
      int c = is.peek();
      if (c == -1) KALDI_ERR << "Failed to read vector from stream: EOF";
      if (c == ']') { is.get(); break; }
      double v;
      if (!ParseReal(is, &v)) KALDI_ERR << "Failed to read vector from stream: bad number";
      data.push_back(static_cast<Real>(v));
    }
  }
  while (is.peek() == ' ' || is.peek() == '\r') is.get();
  if (is.peek() == '\n') is.get();
  Resize(data.size(), kUndefined);
  for (size_t i = 0; i < data.size(); i++) this->data_[i] = data[i];
}

template class VectorBase<float>;
template class Vector<float>;
template class MatrixBase<float>;
template class Matrix<float>;
template class VectorBase<double>;
template class Vector<double>;
template class MatrixBase<double>;
template class Matrix<double>;

bool g_assert_throws = false;

}  // namespace kaldi
