// nnet2/nnet-dp.h -- the data-parallel training step, host side, in Kaldi-style C++.
//
// The reference trains its jobs independently and averages the models through the file system once
// per iteration (nnet-am-average, egs/steps/nnet0/train_conv_dropout.sh:323-341).  Here the minibatch
// rows are sharded over the GPUs of one box (one process per GPU), every rank runs the same fused
// step on its rows with the update deferred, and per layer ONE kernel (kernels_p2p.cu:
// kcnn_p2p_reduce_sgd_f32) sums the gradients over NVLink peer memory, applies the momentum /
// weight-decay step on the owner's slice and writes the new weights to every replica -- the
// reference's update with lr / N_global (nnet0/nnet-component-nnet0.cc:767, 1136), so P ranks x N/P
// rows reproduce the one-rank N-row step up to floating-point summation order.
//
// The class needs nothing from the host but addresses: its own symmetric allocation and the peers'
// mappings of theirs (kcnn_ipc_alloc / kcnn_ipc_open, cuMem handles, or any other exchange), plus
// optionally a multicast mapping for the in-switch (NVLS) form of the kernel.
#ifndef KALDI_NNET2_NNET_DP_H_
#define KALDI_NNET2_NNET_DP_H_

#include <vector>

#include "nnet2/nnet-nnet.h"

namespace kaldi {
namespace nnet2 {

class NnetDataParallel {
 public:
  /// Floats of symmetric memory a rank must provide: gradient arena + parameter arena + barrier flags.
  static size_t ArenaFloats(NnetMinibatchUpdater *updater);

  /// local_base: this rank's allocation (ArenaFloats() floats, zero-filled); peer_bases[p]: rank p's
  /// allocation as mapped into this process (peer_bases[rank] == local_base); multicast_base: 0, or the
  /// multicast mapping of the same allocation.  Moves every parameter of the network into the arena and
  /// switches the components to deferred updates.  All ranks must hold identical parameters.
  NnetDataParallel(Nnet *nnet, NnetMinibatchUpdater *updater, int32 rank, int32 world, float *local_base,
                   const unsigned long long *peer_bases, unsigned long long multicast_base);
  ~NnetDataParallel();

  /// Forward + objective of the first batch (fills the pipeline).
  void Prime(const CuMatrixBase<BaseFloat> &feats, const int32 *labels_dev);
  /// One pipelined step: backward of the batch in the pipeline, each layer's reduce + SGD + broadcast
  /// kernel issued on a communication stream as soon as that layer's backward has been issued;
  /// then the forward pass of the NEXT batch, each layer behind its own update; objective of the next
  /// batch.  Every weight is updated before the first forward pass that reads it: the numbers are those
  /// of the plain synchronous step.  rows_global: rows of the global minibatch (all ranks).
  /// The second call with the same buffers records the rotation into a CUDA graph, later calls replay it.
  void Rotate(const CuMatrixBase<BaseFloat> &feats_next, const int32 *labels_next, int32 rows_global);
  /// Backward + update of the batch still in the pipeline, no further forward pass.
  void Finish(int32 rows_global);
  /// True when a device-side barrier timed out on this rank since the last call (a peer is missing or
  /// late by more than KCNN_P2P_TIMEOUT_MS): the update of that step was skipped.  Polls a pinned word
  /// the step copies back asynchronously; synchronise = true waits for everything enqueued first.
  bool Failed(bool synchronise);
  /// Momentum is sharded: prev_grad_ is current only in the owner's slice.  Before a checkpoint is
  /// written, this makes every rank's prev_grad_ complete (one all-reduce of the masked matrices).
  void GatherMomentum();
  int32 Rank() const { return rank_; }
  int32 World() const { return world_; }
  bool LastRotateReplayed() const { return last_replayed_; }

 private:
  struct Layer {
    int32 comp;
    size_t off, len, weight_floats;       // bucket in the gradient arena (floats)
  };
  // Consecutive small layers (the convolution stack) share ONE reduce + SGD + broadcast launch: their
  // gradients all exist when the backward pass reaches the lowest of them, and the next forward pass needs
  // the lowest first.  A large layer (FC) is a group of its own, reduced as soon as its gradient exists.
  struct Group {
    int32 first, last;                    // layers_[first .. last], network order
    int32 channel;                        // 0: big buckets (FC stack), 1: small (convolutions)
    cudaEvent_t ready, ready_compute, done;   // gradients complete (branch / compute stream), update complete
  };
  void BackwardWithUpdates(int32 rows_global);
  void ForwardBehindUpdates(const CuMatrixBase<BaseFloat> &feats, const int32 *labels);
  void RotateEager(const CuMatrixBase<BaseFloat> &feats_next, const int32 *labels_next, int32 rows_global);
  void ReduceAndUpdate(const Group &g, int32 rows_global);
  void DropGraphs();

  Nnet *nnet_;
  NnetMinibatchUpdater *updater_;
  int32 rank_, world_;
  float *base_;
  std::vector<unsigned long long> peers_;
  unsigned long long multicast_;
  size_t grad_floats_, flag_off_;
  std::vector<Layer> layers_;             // network order
  std::vector<Group> groups_;             // network order
  cudaStream_t comm_[2];
  unsigned int *error_pinned_;            // [2]: the error words of the two channels, copied back per step
  bool primed_, last_replayed_;
  struct Recorded { uint64 key; cudaGraphExec_t exec; std::vector<double> count_delta; };
  std::vector<Recorded> graphs_;
  std::vector<uint64> seen_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(NnetDataParallel);
};

}  // namespace nnet2
}  // namespace kaldi
#endif
