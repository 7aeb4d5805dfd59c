// base/kaldi-types.h -- shim: Kaldi's basic typedefs.
#ifndef KALDI_BASE_KALDI_TYPES_H_
#define KALDI_BASE_KALDI_TYPES_H_
#include <cstdint>
namespace kaldi {
typedef float BaseFloat;   // the reference builds BaseFloat = float only
typedef int8_t int8;
typedef int16_t int16;
typedef int32_t int32;
typedef int64_t int64;
typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef uint64_t uint64;
typedef float float32;
typedef double double64;
}  // namespace kaldi
#endif
