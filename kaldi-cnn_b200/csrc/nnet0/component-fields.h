// nnet0/component-fields.h -- one description of a component's scalar members, used for
// everything that is not computation: the nnet.config line ("key=value"), the model-file token stream
// ("<Token> value", text or binary) and the Info() string.
//
// The reference spells each of these out by hand per component (nnet0/nnet-component-nnet0.cc:
// 323-421, 556-666 for the convolution; 814-867, 894-978 max-pool; 980-1119 fully connected), which is
// how its config keys, tokens and Info() fields drifted apart.  Here a component lists its members ONCE:
//
//     FieldList f;
//     f.Int("in-height", "<in_height>", &in_height_).Int("in-width", "<in_width>", &in_width_) ...
//
// and InitFromString / Read / Write / Info walk that list.  Keys, tokens and their ORDER are the
// reference's (existing nnet.config files and .mdl files must keep working); nothing else is.
#ifndef CNSL_NNET0_COMPONENT_FIELDS_H_
#define CNSL_NNET0_COMPONENT_FIELDS_H_

#include <sstream>
#include <string>
#include <vector>

#include "base/kaldi-common.h"
#include "cudamatrix/cu-matrix-lib.h"
#include "nnet2/nnet-component.h"

namespace cnsl {
namespace nnet0 {

class FieldList {
 public:
  enum Need { kRequired, kOptional };
  /// key: name on the config line (NULL: not configurable); token: name in the model file (NULL: not
  /// serialised); need: whether ParseConfig insists on the key.
  FieldList &Int(const char *key, const char *token, kaldi::int32 *v, Need need = kRequired) {
    return Push(Field(kInt, key, token, v, need));
  }
  FieldList &Float(const char *key, const char *token, kaldi::BaseFloat *v, Need need = kRequired) {
    return Push(Field(kFloat, key, token, v, need));
  }
  FieldList &Bool(const char *key, const char *token, bool *v, Need need = kRequired) {
    return Push(Field(kBool, key, token, v, need));
  }
  /// Parameter matrices / vectors: part of the token stream only.
  FieldList &Matrix(const char *token, kaldi::CuMatrix<kaldi::BaseFloat> *m) {
    return Push(Field(kMatrix, NULL, token, m, kOptional));
  }
  FieldList &Vector(const char *token, kaldi::CuVector<kaldi::BaseFloat> *v) {
    return Push(Field(kVector, NULL, token, v, kOptional));
  }

  /// Takes every "key=value" of the list out of *args.  Returns false when a required key is missing.
  bool ParseConfig(std::string *args) const {
    bool ok = true;
    for (size_t i = 0; i < fields_.size(); i++) {
      const Field &f = fields_[i];
      if (f.key == NULL) continue;
      bool found = false;
      switch (f.kind) {
        case kInt: found = kaldi::nnet2::ParseFromString(f.key, args, static_cast<kaldi::int32 *>(f.ptr)); break;
        case kFloat: found = kaldi::nnet2::ParseFromString(f.key, args, static_cast<kaldi::BaseFloat *>(f.ptr)); break;
        case kBool: found = kaldi::nnet2::ParseFromString(f.key, args, static_cast<bool *>(f.ptr)); break;
        default: break;
      }
      if (!found && f.need == kRequired) ok = false;
    }
    return ok;
  }

  /// "<Token> value" for fields [first, last) of the list (default: all), in list order.
  void Write(std::ostream &os, bool binary, size_t first = 0, size_t last = static_cast<size_t>(-1)) const {
    for (size_t i = first; i < fields_.size() && i < last; i++) {
      const Field &f = fields_[i];
      if (f.token == NULL) continue;
      kaldi::WriteToken(os, binary, f.token);
      switch (f.kind) {
        case kInt: kaldi::WriteBasicType(os, binary, *static_cast<const kaldi::int32 *>(f.ptr)); break;
        case kFloat: kaldi::WriteBasicType(os, binary, *static_cast<const kaldi::BaseFloat *>(f.ptr)); break;
        case kBool: kaldi::WriteBasicType(os, binary, *static_cast<const bool *>(f.ptr)); break;
        case kMatrix: static_cast<const kaldi::CuMatrix<kaldi::BaseFloat> *>(f.ptr)->Write(os, binary); break;
        case kVector: static_cast<const kaldi::CuVector<kaldi::BaseFloat> *>(f.ptr)->Write(os, binary); break;
      }
    }
  }

  /// Reads the same stream.  skip_first_token: the first field's token has been consumed already (the
  /// callers use ExpectOneOrTwoTokens for "<Type> <FirstToken>").
  void Read(std::istream &is, bool binary, bool skip_first_token, size_t first = 0,
            size_t last = static_cast<size_t>(-1)) const {
    for (size_t i = first; i < fields_.size() && i < last; i++) {
      const Field &f = fields_[i];
      if (f.token == NULL) continue;
      if (!(skip_first_token && i == first)) kaldi::ExpectToken(is, binary, f.token);
      ReadValue(is, binary, f);
    }
  }

  /// Value of the field whose token is `token` from the stream (the caller has read the token).
  bool ReadByToken(std::istream &is, bool binary, const std::string &token) const {
    for (size_t i = 0; i < fields_.size(); i++)
      if (fields_[i].token != NULL && token == fields_[i].token) { ReadValue(is, binary, fields_[i]); return true; }
    return false;
  }

  /// Optional tail of a stream: "<Token> value" pairs of fields [first, ...) in any subset, up to and
  /// including `end_token`.  Anything else is an error.
  void ReadTail(std::istream &is, bool binary, size_t first, const std::string &end_token) const {
    std::string tok;
    for (;;) {
      kaldi::ReadToken(is, binary, &tok);
      if (tok == end_token) return;
      bool known = false;
      for (size_t i = first; i < fields_.size() && !known; i++)
        if (fields_[i].token != NULL && tok == fields_[i].token) { ReadValue(is, binary, fields_[i]); known = true; }
      if (!known) KALDI_ERR << "Unexpected token " << tok << " (expected " << end_token << ")";
    }
  }

  const char *FirstToken() const {
    for (size_t i = 0; i < fields_.size(); i++)
      if (fields_[i].token != NULL) return fields_[i].token;
    return "";
  }

  /// "key=value" pairs of the configurable fields, separated by ", ".
  std::string Describe() const {
    std::ostringstream os;
    bool first = true;
    for (size_t i = 0; i < fields_.size(); i++) {
      const Field &f = fields_[i];
      if (f.key == NULL) continue;
      if (!first) os << ", ";
      first = false;
      os << f.key << "=";
      switch (f.kind) {
        case kInt: os << *static_cast<const kaldi::int32 *>(f.ptr); break;
        case kFloat: os << *static_cast<const kaldi::BaseFloat *>(f.ptr); break;
        case kBool: os << (*static_cast<const bool *>(f.ptr) ? "true" : "false"); break;
        default: break;
      }
    }
    return os.str();
  }

  size_t Size() const { return fields_.size(); }

 private:
  enum Kind { kInt, kFloat, kBool, kMatrix, kVector };
  struct Field {
    Kind kind; const char *key; const char *token; void *ptr; Need need;
    Field(Kind k, const char *key_, const char *token_, void *p, Need n) : kind(k), key(key_), token(token_), ptr(p), need(n) {}
  };
  FieldList &Push(const Field &f) { fields_.push_back(f); return *this; }
  static void ReadValue(std::istream &is, bool binary, const Field &f) {
    switch (f.kind) {
      case kInt: kaldi::ReadBasicType(is, binary, static_cast<kaldi::int32 *>(f.ptr)); break;
      case kFloat: kaldi::ReadBasicType(is, binary, static_cast<kaldi::BaseFloat *>(f.ptr)); break;
      case kBool: kaldi::ReadBasicType(is, binary, static_cast<bool *>(f.ptr)); break;
      case kMatrix: static_cast<kaldi::CuMatrix<kaldi::BaseFloat> *>(f.ptr)->Read(is, binary); break;
      case kVector: static_cast<kaldi::CuVector<kaldi::BaseFloat> *>(f.ptr)->Read(is, binary); break;
    }
  }
  std::vector<Field> fields_;
};

}  // namespace nnet0
}  // namespace cnsl

#endif
