/* oracle/ref_shim/base/kaldi-error.h -- TEST INFRASTRUCTURE ONLY.
 * Stand-in for the Kaldi header of the same name, which the reference's kernel file
 * (src/cnslmat/cnsl-cu-kernels.h:11) includes but does not use on the device side.  It lets
 * oracle/Makefile compile the UNMODIFIED reference kernels from /root/reference into
 * oracle/_ref/ (see oracle/Makefile); nothing of the product includes it. */
#ifndef KCNN_ORACLE_REF_SHIM_KALDI_ERROR_H_
#define KCNN_ORACLE_REF_SHIM_KALDI_ERROR_H_
namespace kaldi {}
#endif
