// util/common-utils.h -- shim umbrella.
#ifndef KALDI_UTIL_COMMON_UTILS_H_
#define KALDI_UTIL_COMMON_UTILS_H_
#include "base/kaldi-common.h"
#include "util/text-utils.h"
#include "util/kaldi-io.h"
#endif
