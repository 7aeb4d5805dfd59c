// base/kaldi-common.h -- Kaldi-shaped shim (scaffolding, not product).
//
// The reference is a patch set on Kaldi trunk r4510 and ships none of Kaldi's
// base/, matrix/, util/ or cudamatrix/ trees.  This directory
// provides the MINIMUM of that API, with Kaldi's names and signatures, that the
// hot-path host code (cnslmat/conv2D.cc, nnet0/*, nnet2/nnet-component.*) needs,
// so the same host sources compile against this shim or against a real Kaldi
// tree.  GPU only: there is no CPU matrix backend behind CuMatrix here.
#ifndef KALDI_BASE_KALDI_COMMON_H_
#define KALDI_BASE_KALDI_COMMON_H_

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "base/kaldi-types.h"
#include "base/kaldi-error.h"
#include "base/io-funcs.h"
#include "base/timer.h"

#define KALDI_DISALLOW_COPY_AND_ASSIGN(type) \
  type(const type &);                        \
  void operator=(const type &)

#endif
