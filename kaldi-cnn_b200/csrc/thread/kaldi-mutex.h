// thread/kaldi-mutex.h -- shim: the GPU hot path is single-threaded by contract
// (nnet2/nnet-component.cc:3611-3612 upstream comment).
#ifndef KALDI_THREAD_KALDI_MUTEX_H_
#define KALDI_THREAD_KALDI_MUTEX_H_
#endif
