/*
 * include/kcnn_capi.h -- C ABI over the L1 members and L2 components of
 * libkaldicnn_b200.so, for hosts that are not Kaldi C++ (tests and bench.py bind it with
 * ctypes; a cgo / JNI / N-API stub would bind the same symbols).
 *
 * A Kaldi C++ host does not need this header: it includes
 * nnet0/nnet-component-nnet0.h and gets the components through nnet2's
 * Component::NewComponentOfType (reference nnet2/nnet-component.cc:112-117), exactly as
 * with the reference.  Every function here is a thin forwarder to that C++ interface;
 * each comment names the reference interface it exposes.
 *
 * Conventions: matrices are float, row-major, (pointer, rows, cols, stride) with
 * stride >= cols in elements -- Kaldi's MatrixDim.  Unless a name ends in _host all
 * pointers are DEVICE pointers and calls are asynchronous on the stream set with
 * kcnn_set_compute_stream().  Functions returning int return 0 on success and -1 on
 * error (message: kcnn_last_error()); a Kaldi assertion or KALDI_ERR in the C++ layer
 * is reported that way instead of aborting the host process.
 */
#ifndef KCNN_CAPI_H_
#define KCNN_CAPI_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kcnn_component kcnn_component;   /* a kaldi::nnet2::Component*            */
typedef struct kcnn_nnet kcnn_nnet;             /* Nnet + NnetMinibatchUpdater            */

/* ---- device / library --------------------------------------------------------- */

/* CuDevice::Instantiate().SelectGpuId(use_gpu): "yes" | "no" | "optional". */
int kcnn_select_gpu(const char *use_gpu);
/* Stream every member / component launches on (default: legacy default stream). */
void kcnn_set_compute_stream(void *cuda_stream);
/* KCNN_MATH_FP32_SIMT (0, default) or KCNN_MATH_TF32_TC (1): CuDevice::SetMathMode. */
void kcnn_set_math_mode(int mode);
int kcnn_get_math_mode(void);
/* Seed of CuMatrix::SetRandn (parameter initialisation). */
void kcnn_set_rand_seed(unsigned long long seed);
const char *kcnn_last_error(void);
/* CuDevice::PrintProfile / AccuProfile switch (reference cnslmat/conv2D.cc:110). */
void kcnn_enable_profile(int on);
void kcnn_print_profile(void);
/* Caching allocator of the device layer: bytes held (live + cached), and bytes kept out of circulation
 * because a recorded CUDA graph may still address them (0 in steady state: the trainers size their buffers
 * before they record). */
size_t kcnn_device_bytes_pinned_by_graphs(void);
/* Bytes currently held by the caching device allocator. */
size_t kcnn_device_bytes_allocated(void);

/* ---- L1: CuMatrixBase members (reference cudamatrix/cu-matrix.h:451-482) --------- */
/* 'this' is (a, a_rows, a_cols, a_stride).  Shape rules, resize rules and assertions
 * are those of cnslmat/conv2D.cc; outputs declared CuMatrix<Real>* there must arrive
 * with the final shape here (a foreign buffer cannot be resized). */

int kcnn_mat_conv2d(const float *a, int a_rows, int a_cols, int a_stride,
                    const float *kernel, int k_rows, int k_cols, int k_stride,
                    int in_height, int in_width, int in_channel, int kernel_height,
                    int kernel_width, int group,
                    float *out, int o_rows, int o_cols, int o_stride, int concat);
int kcnn_mat_add_mat_rep_vec(float *a, int a_rows, int a_cols, int a_stride,
                             const float *vec, int vec_dim, int rep);
int kcnn_mat_flip_mat(const float *a, int a_rows, int a_cols, int a_stride,
                      int kernel_height, int kernel_width, int in_channel, int group,
                      float *flip, int f_rows, int f_cols, int f_stride);
int kcnn_mat_padding_zero(const float *a, int a_rows, int a_cols, int a_stride,
                          int orig_height, int orig_width, int orig_channel,
                          int kernel_height, int kernel_width,
                          float *pad, int p_rows, int p_cols, int p_stride);
int kcnn_mat_tp_block(const float *a, int a_rows, int a_cols, int a_stride,
                      int in_channel, int block_size,
                      float *out, int o_rows, int o_cols, int o_stride);
int kcnn_mat_tp_inside_block(const float *a, int a_rows, int a_cols, int a_stride,
                             int group, int block_size,
                             float *out, int o_rows, int o_cols, int o_stride);
int kcnn_mat_mod_permute_row(const float *a, int a_rows, int a_cols, int a_stride,
                             int in_channel, int block_size,
                             float *out, int o_rows, int o_cols, int o_stride);
int kcnn_mat_maxpool_prop(const float *a, int a_rows, int a_cols, int a_stride,
                          int in_height, int in_width, int pool_height_dim,
                          int pool_width_dim, int pool_channel_dim, int overlap,
                          int overlap2D, float *out, int o_rows, int o_cols, int o_stride);
int kcnn_mat_maxpool_backprop(const float *a, int a_rows, int a_cols, int a_stride,
                              const float *out_value, int ov_rows, int ov_cols, int ov_stride,
                              const float *out_deriv, int od_rows, int od_cols, int od_stride,
                              float *in_deriv, int id_rows, int id_cols, int id_stride,
                              int in_height, int in_width, int pool_height_dim,
                              int pool_width_dim, int pool_channel_dim, int overlap,
                              int overlap2D);

/* ---- L2: components (reference nnet2/nnet-component.h:157-269) -------------------- */

/* Component::NewFromString: "ConvolutionComponent in-height=40 ..." (one nnet.config line). */
kcnn_component *kcnn_component_new_from_string(const char *config_line);
/* Component::ReadNew from a serialised component (binary != 0: Kaldi binary mode). */
kcnn_component *kcnn_component_read(const char *data, size_t len, int binary);
/* Component::Write into a malloc'ed buffer; release with kcnn_free(). */
int kcnn_component_write(const kcnn_component *c, int binary, char **data, size_t *len);
void kcnn_free(void *p);
/* Component::Copy. */
kcnn_component *kcnn_component_copy(const kcnn_component *c);
void kcnn_component_delete(kcnn_component *c);
const char *kcnn_component_type(const kcnn_component *c);       /* Component::Type()  */
int kcnn_component_info(const kcnn_component *c, char *buf, size_t buf_len);   /* Info() */
int kcnn_component_input_dim(const kcnn_component *c);
int kcnn_component_output_dim(const kcnn_component *c);
int kcnn_component_backprop_needs_input(const kcnn_component *c);
int kcnn_component_backprop_needs_output(const kcnn_component *c);

/* Component::Propagate(in_info, out_info, in, out) with ChunkInfo(dim, num_chunks, 0, 0). */
int kcnn_component_propagate(const kcnn_component *c, int num_chunks,
                             const float *in, int in_rows, int in_cols, int in_stride,
                             float *out, int out_rows, int out_cols, int out_stride);
/* The same two calls for components whose chunks hold several frames (SpliceComponent, reference
 * nnet2/nnet-component.cc:2640-2848): in_info = ChunkInfo(input_dim, num_chunks, in_first_offset,
 * in_last_offset), out_info likewise -- `in` has num_chunks * (in_last - in_first + 1) rows, `out`
 * num_chunks * (out_last - out_first + 1).  backprop_chunks is for components that need neither
 * in_value nor out_value. */
int kcnn_component_propagate_chunks(const kcnn_component *c, int num_chunks, int in_first_offset,
                                    int in_last_offset, int out_first_offset, int out_last_offset,
                                    const float *in, int in_rows, int in_cols, int in_stride,
                                    float *out, int out_rows, int out_cols, int out_stride);
int kcnn_component_backprop_chunks(const kcnn_component *c, int num_chunks, int in_first_offset,
                                   int in_last_offset, int out_first_offset, int out_last_offset,
                                   const float *out_deriv, int od_rows, int od_stride,
                                   float *in_deriv, int id_stride);
/* Component::Backprop(in_info, out_info, in_value, out_value, out_deriv, to_update,
 * in_deriv).  in_value / out_value may be NULL when BackpropNeedsInput / Output is false.
 * to_update may be NULL (no update), c itself (ordinary SGD) or another component. */
int kcnn_component_backprop(const kcnn_component *c, int num_chunks,
                            const float *in_value, int iv_stride,
                            const float *out_value, int ov_stride,
                            const float *out_deriv, int od_rows, int od_stride,
                            kcnn_component *to_update,
                            float *in_deriv, int id_stride);

/* Updatable components: parameter access.  which: 0 = linear_params_, 1 = bias_params_
 * (rows = 1), 2 = prev_grad_.  Returns the device pointer and its shape. */
int kcnn_component_params(kcnn_component *c, int which, float **data, int *rows, int *cols,
                          int *stride);
/* UpdatableComponent::{SetLearningRate, LearningRate}; Set{WeightDecay,Momentum}. */
int kcnn_component_set_learning_rate(kcnn_component *c, float lr);
float kcnn_component_learning_rate(const kcnn_component *c);
int kcnn_component_set_weight_decay_momentum(kcnn_component *c, float weight_decay, float momentum);
int kcnn_component_get_weight_decay_momentum(const kcnn_component *c, float *weight_decay, float *momentum);
/* MaxpoolComponent::SetIndexRouting (B200 extension, see nnet-component-nnet0.h). */
int kcnn_component_set_index_routing(kcnn_component *c, int on);

/* Data-parallel split of the update (B200 extension): with deferred != 0 Backprop leaves
 * the un-normalised gradient in the component's gradient buffers; after the caller has
 * summed those over ranks, kcnn_component_apply_gradient(total_rows) performs the
 * reference's update with lr / total_rows. */
int kcnn_component_set_deferred_update(kcnn_component *c, int deferred);
size_t kcnn_component_gradient_floats(const kcnn_component *c);
int kcnn_component_set_gradient_storage(kcnn_component *c, float *base);
/* which: 0 = weight gradient, 1 = bias gradient. */
int kcnn_component_gradient(kcnn_component *c, int which, float **data, int *rows, int *cols,
                            int *stride);
int kcnn_component_apply_gradient(kcnn_component *c, int total_rows);

/* ---- the callers of the path: a component sequence and one training step ---------- */

/* nnet2 Nnet::Init from the text of an nnet.config (one component per line).
 * skip_splice != 0 drops SpliceComponent lines (the input is then the spliced window). */
kcnn_nnet *kcnn_nnet_new_from_config(const char *config_text, int skip_splice);
kcnn_nnet *kcnn_nnet_read(const char *data, size_t len, int binary);
int kcnn_nnet_write(const kcnn_nnet *n, int binary, char **data, size_t *len);
void kcnn_nnet_delete(kcnn_nnet *n);
int kcnn_nnet_num_components(const kcnn_nnet *n);
kcnn_component *kcnn_nnet_component(kcnn_nnet *n, int index);   /* borrowed, do not delete */
int kcnn_nnet_input_dim(const kcnn_nnet *n);
int kcnn_nnet_output_dim(const kcnn_nnet *n);

/* Forward through every component; feats is [rows x input_dim] on the device (rows = examples x
 * kcnn_nnet_frames_per_example()). */
int kcnn_nnet_forward(kcnn_nnet *n, const float *feats, int rows, int stride);
/* Propagate through components [first, last] only (first == 0 binds feats as the input; the
 * activations below first must come from an earlier call on the same batch). */
int kcnn_nnet_forward_range(kcnn_nnet *n, const float *feats, int rows, int stride, int first,
                            int last);
/* Cross-entropy derivative at the output for int32 device labels[rows]; the objective
 * sum_i log p[i, label_i] accumulates on the device. */
int kcnn_nnet_objf_and_deriv(kcnn_nnet *n, const int *labels);
/* Backprop (with update unless deferred) through components last..first; last < 0 = top. */
int kcnn_nnet_backward(kcnn_nnet *n, int last, int first);
/* Output / activation access (device pointers, valid until the next forward). */
int kcnn_nnet_activation(kcnn_nnet *n, int index, const float **data, int *rows, int *cols, int *stride);
int kcnn_nnet_input_deriv(kcnn_nnet *n, const float **data, int *rows, int *cols, int *stride);
/* Reads and clears the accumulated objective (synchronises the stream). */
double kcnn_nnet_objf_and_reset(kcnn_nnet *n);
/* Data-parallel: deferred updates + one gradient arena (top layer first). */
int kcnn_nnet_set_deferred_update(kcnn_nnet *n, int deferred);
size_t kcnn_nnet_gradient_floats(const kcnn_nnet *n);
int kcnn_nnet_set_gradient_arena(kcnn_nnet *n, float *base);
int kcnn_nnet_gradient_bucket(const kcnn_nnet *n, int component, size_t *offset, size_t *length);
int kcnn_nnet_apply_gradients(kcnn_nnet *n, int total_rows);

/* Input rows per training example: 1, or -- when the network starts with a SpliceComponent
 * (nnet.config line 1) -- the span of its context: an nnet2 training example carries left-context + 1 +
 * right-context frames of input_dim features for its one labelled frame, and the feature matrix holds
 * the examples' frames back to back.  For a run of consecutive offsets the library then reads the
 * convolution's [C][W][H] window straight from those rows (no copy, no launch). */
int kcnn_nnet_frames_per_example(kcnn_nnet *n);
/* The whole step from HOST memory, the call a non-CUDA host makes: copies feats
 * [rows * kcnn_nnet_frames_per_example() x input_dim, packed] and labels[rows] to the device, runs
 * forward, objective, backward + update, and returns the minibatch objective in *objf (synchronous). */
int kcnn_nnet_train_minibatch_host(kcnn_nnet *n, const float *feats_host, const int *labels_host,
                                   int rows, double *objf);
/* Pipelined form of the call above for a host that streams minibatches: the caller's buffers
 * are copied to pinned staging memory (they may be reused as soon as the call returns), the
 * host-to-device copy of this batch runs on a copy stream while the PREVIOUS batch is still
 * computing, the step is enqueued behind it, and the call returns without waiting for the
 * GPU.  The objective accumulates on the device and is also copied back to the host after
 * every step: kcnn_nnet_running_objf() returns the accumulated value as of the newest step
 * that has completed (never blocks); kcnn_nnet_objf_and_reset() waits for everything enqueued
 * and returns the total, as in Kaldi's periodic objective report. */
int kcnn_nnet_train_minibatch_host_async(kcnn_nnet *n, const float *feats_host, const int *labels_host,
                                         int rows);
double kcnn_nnet_running_objf(kcnn_nnet *n);
/* The same step on DEVICE buffers, asynchronous on the compute stream (read the objective
 * with kcnn_nnet_objf_and_reset).  Both calls go through NnetMinibatchUpdater::TrainStep: on a
 * non-default compute stream the launches of a step (33 for the benchmarked model) are recorded into a CUDA graph on the
 * second call with the same buffers / configuration and replayed from then on
 * (KCNN_NNET_GRAPH=0 keeps every step eager). */
int kcnn_nnet_train_step(kcnn_nnet *n, const float *feats, int rows, int stride, const int *labels);
/* Fusion (default on): the fused plan when the model and math mode allow it, else
 * [Convolution | FullyConnected] + ReLU as one launch.  With fusion the pre-activation
 * kcnn_nnet_activation(c + 1) of a fused pair is not filled.  0 = component by component. */
int kcnn_nnet_set_fusion(kcnn_nnet *n, int on);
/* CUDA-graph recording of kcnn_nnet_train_step / _host / _host_async (default on); 0 runs every step's
 * launches eagerly (profiling, debugging). */
int kcnn_nnet_set_graphs(kcnn_nnet *n, int on);
/* 1 when the most recent train step of this network was a graph replay, else 0. */
int kcnn_nnet_last_step_replayed(const kcnn_nnet *n);
/* 1 when the network's current configuration (model, rows, math mode) runs as the fused plan of
 * csrc/nnet2/nnet-fused.cc (channels-last activations, element-wise components inside the GEMM
 * epilogues), 0 when it runs component by component.  Valid after the first forward pass. */
int kcnn_nnet_fused_active(const kcnn_nnet *n);

/* ---- gradient all-reduce over NVLink peer memory (kernels_p2p.cu) ---------------------------
 * Replaces the reference's file-based nnet-am-average (egs/steps/nnet0/train_conv_dropout.sh:
 * 323-341) for the data-parallel step.  Every rank owns one SYMMETRIC allocation of the same
 * size (e.g. torch.distributed._symmetric_memory, cudaIpc, cuMem fabric handles): the gradient
 * arena (kcnn_nnet_set_gradient_arena) followed by kcnn_p2p_flag_floats() zero-initialised
 * floats of flags.  peer_bases[p] is the address of rank p's allocation as mapped into THIS
 * process.  kcnn_p2p_allreduce_f32 sums floats [offset, offset + count) of all ranks' arenas
 * in place on every rank (two-shot: reduce own slice from peer loads, store the sum to every
 * peer), enqueued on `stream`; all ranks must enqueue the same sequence of calls per channel
 * (channel 0 / 1: two independent flag sets, for two streams).  Offsets and counts in floats,
 * multiples of 4.  Returns 0, or -1 on bad arguments. */
size_t kcnn_p2p_flag_floats(void);
int kcnn_p2p_allreduce_f32(void *stream, const unsigned long long *peer_bases, int rank, int world,
                           size_t offset_floats, size_t count_floats, size_t flag_offset_floats, int channel);
/* The same reduction done INSIDE the NVSwitch (NVLS): multicast_base is the multicast mapping of
 * the symmetric allocation (torch symmetric memory: handle.multicast_ptr; cuMulticast*); the
 * kernel pulls the sum of its slice with multimem.ld_reduce and broadcasts it with multimem.st.
 * The flags still go through peer_bases.  Returns -1 when multicast_base is 0. */
int kcnn_p2p_allreduce_multicast_f32(void *stream, const unsigned long long *peer_bases,
                                     unsigned long long multicast_base, int rank, int world, size_t offset_floats,
                                     size_t count_floats, size_t flag_offset_floats, int channel);
/* 1 when a barrier of this rank gave up waiting for a peer (synchronises the device).  The wait is
 * bounded by wall-clock time (KCNN_P2P_TIMEOUT_MS, default 20000); a kernel whose barrier timed out
 * skips its stores, so neither gradients nor weights are overwritten with a partial sum. */
int kcnn_p2p_error(const float *local_base, size_t flag_offset_floats);
const unsigned int *kcnn_p2p_error_word(const float *local_base, size_t flag_offset_floats, int channel);

/* Reduce-scatter + momentum SGD + all-gather of one layer's bucket in ONE kernel (kernels_p2p.cu).
 * The symmetric allocation holds the gradient arena and, param_delta_floats further on, a parameter
 * arena with the same layout.  Rank r sums its slice of the bucket over all ranks (peer loads in rank
 * order, or in the NVSwitch when multicast_base != 0), applies
 *     prev = momentum*prev + decay_alpha*W + grad_alpha*g ; W += prev      (first weight_floats floats)
 *     b += grad_alpha*g                                                       (the rest: the bias)
 * -- nnet0/nnet-component-nnet0.cc:767-775, 1136-1142 -- with ITS momentum matrix prev_grad (local,
 * indexed like the bucket) and stores the new values to every rank's parameter arena.  Same results as
 * kcnn_p2p_allreduce_f32 followed by the per-rank update, without the second pass. */
int kcnn_p2p_reduce_sgd_f32(void *stream, const unsigned long long *peer_bases, unsigned long long multicast_base,
                            int rank, int world, size_t offset_floats, size_t count_floats, size_t weight_floats,
                            size_t param_delta_floats, float *prev_grad, float momentum, float decay_alpha,
                            float grad_alpha, size_t flag_offset_floats, int channel);

/* The same for up to 8 (small) buckets between one pair of peer barriers: one launch for a stack of small
 * layers whose gradients become available together. */
typedef struct {
  size_t offset_floats, count_floats, weight_floats;    /* as the arguments of kcnn_p2p_reduce_sgd_f32 */
  float *prev_grad;
  float momentum, decay_alpha, grad_alpha;
} KcnnSgdBucket;
int kcnn_p2p_reduce_sgd_multi_f32(void *stream, const unsigned long long *peer_bases, unsigned long long multicast_base,
                                  int rank, int world, int num_buckets, const KcnnSgdBucket *buckets,
                                  size_t param_delta_floats, size_t flag_offset_floats, int channel);

/* Symmetric memory from CUDA IPC handles: kcnn_ipc_alloc returns zero-filled device memory and its
 * 64-byte cudaIpcMemHandle_t; the host sends the handle to the other ranks (any transport) and each
 * maps it with kcnn_ipc_open (peer access is enabled on first use). */
int kcnn_ipc_alloc(size_t bytes, void **ptr, unsigned char *handle64);
int kcnn_ipc_open(const unsigned char *handle64, void **ptr);
int kcnn_ipc_close(void *ptr);
int kcnn_ipc_free(void *ptr);

/* ---- the data-parallel trainer (csrc/nnet2/nnet-dp.h: NnetDataParallel) -------------------------------
 * One process per GPU.  kcnn_nnet_dp_arena_floats() floats of symmetric memory per rank (zero-filled;
 * kcnn_ipc_alloc, or any allocation every rank can map), peer_bases[p] = rank p's arena as mapped here.
 * create moves the network's parameters into the arena and defers the updates; all ranks must start from
 * identical parameters.  rotate = backward of the batch in the pipeline with one fused reduce + SGD +
 * broadcast kernel per layer, forward of the next batch behind the updates, objective of the next batch;
 * recorded into a CUDA graph on the second call with the same buffers.  feats: device [rows_local *
 * kcnn_nnet_frames_per_example() x input_dim]; labels: device int32 [rows_local].
 * The host-buffer form stages the caller's buffers through pinned memory and a copy stream, two device
 * slots, like kcnn_nnet_train_minibatch_host_async: call it once per minibatch (the first call primes the
 * pipeline), then kcnn_nnet_dp_finish.  The objective accumulates per rank (kcnn_nnet_objf_and_reset). */
typedef struct kcnn_nnet_dp kcnn_nnet_dp;
size_t kcnn_nnet_dp_arena_floats(kcnn_nnet *n);
kcnn_nnet_dp *kcnn_nnet_dp_create(kcnn_nnet *n, int rank, int world, float *local_base,
                                  const unsigned long long *peer_bases, unsigned long long multicast_base);
void kcnn_nnet_dp_delete(kcnn_nnet_dp *dp);
int kcnn_nnet_dp_prime(kcnn_nnet_dp *dp, const float *feats, int rows, int stride, const int *labels);
int kcnn_nnet_dp_rotate(kcnn_nnet_dp *dp, const float *feats_next, int rows, int stride, const int *labels_next,
                        int rows_global);
int kcnn_nnet_dp_finish(kcnn_nnet_dp *dp, int rows_global);
int kcnn_nnet_dp_train_minibatch_host_async(kcnn_nnet_dp *dp, const float *feats_host, const int *labels_host,
                                            int rows_local, int rows_global);
/* 1 when a barrier timed out on this rank (a peer is missing): the update of that step was skipped and the
 * replicas may have diverged.  synchronise != 0 waits for the enqueued work first. */
int kcnn_nnet_dp_failed(kcnn_nnet_dp *dp, int synchronise);
/* Completes every rank's momentum matrices (they are updated by the owner of each slice only); call on
 * all ranks before writing a checkpoint. */
int kcnn_nnet_dp_gather_momentum(kcnn_nnet_dp *dp);
int kcnn_nnet_dp_last_rotate_replayed(const kcnn_nnet_dp *dp);

#ifdef __cplusplus
}
#endif

#endif /* KCNN_CAPI_H_ */
