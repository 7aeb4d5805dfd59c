// itf/options-itf.h -- shim: nothing on the hot path registers options.
#ifndef KALDI_ITF_OPTIONS_ITF_H_
#define KALDI_ITF_OPTIONS_ITF_H_
#endif
