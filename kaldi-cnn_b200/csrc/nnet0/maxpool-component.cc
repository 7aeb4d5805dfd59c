// nnet0/maxpool-component.cc -- MaxpoolComponent for the B200 build.
// Behaviour of reference src/nnet0/nnet-component-nnet0.cc:779-978 (config keys, model stream, geometry
// checks, Propagate / Backprop semantics); the glue is table-driven (nnet0/component-fields.h).

#include <cmath>
#include <sstream>

#include "nnet0/nnet-component-nnet0.h"
#include "nnet0/component-fields.h"
#include "util/common-utils.h"
#include "cnsl-cu-kernels.h"

namespace cnsl {
namespace nnet0 {

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }

MaxpoolComponent::~MaxpoolComponent() {
  if (index_) CuDevice::Instantiate().Free(index_);
}

// ---- geometry ------------------------------------------------------------------------------------
// Output maps per (h, w) position for the three pooling modes of the reference (:779-812, :832-850):
//   plain      C / pool_channel_dim            (windows tile the maps)
//   overlap    C - pool_channel_dim + 1        (window slides over the maps, step 1)
//   overlap2D  (sqrt(C) - pool_channel_dim + 1)^2   (maps arranged as a sqrt(C) x sqrt(C) grid, square window)
// 0 = the geometry is not valid.
static int32 PooledMaps(int32 in_channel, int32 pool_channel_dim, bool overlap, bool overlap2D) {
  if (in_channel <= 0 || pool_channel_dim <= 0 || (overlap && overlap2D)) return 0;
  if (overlap2D) {
    // as the reference computes it (double sqrt, truncation of the square), so that a config the
    // reference accepts gives the same output-dim here
    int32 maps = pow((sqrt(in_channel) - pool_channel_dim + 1), 2);
    return maps > 0 ? maps : 0;
  }
  if (overlap) return in_channel >= pool_channel_dim ? in_channel - pool_channel_dim + 1 : 0;
  return in_channel % pool_channel_dim == 0 ? in_channel / pool_channel_dim : 0;
}

FieldList MaxpoolComponent::StreamFields() {
  FieldList f;
  f.Int(NULL, "<InputDim>", &input_dim_)
      .Int("in-height", "<in_height>", &in_height_)
      .Int("in-width", "<in_width>", &in_width_)
      .Int("in-channel", "<in_channel>", &in_channel_)
      .Int(NULL, "<OutputDim>", &output_dim_)
      .Int("pool-height-dim", "<PoolHeightDim>", &pool_height_dim_)
      .Int("pool-width-dim", "<PoolWidthDim>", &pool_width_dim_)
      .Int("pool-channel-dim", "<PoolChannelDim>", &pool_channel_dim_)
      .Bool("overlap", "<Overlap>", &overlap_, FieldList::kOptional)          // [8], [9]: absent from
      .Bool("overlap2D", "<Overlap2D>", &overlap2D_, FieldList::kOptional);   // older model files
  return f;
}
static const size_t kMaxpoolOptionalFrom = 8;

// Checks what the reference's Init asserts (:791-811).
void MaxpoolComponent::Check() const {
  KALDI_ASSERT(input_dim_ > 0 && output_dim_ > 0 && in_height_ * in_width_ * in_channel_ == input_dim_);
  KALDI_ASSERT(pool_height_dim_ > 0 && pool_width_dim_ > 0 && pool_channel_dim_ > 0);
  KALDI_ASSERT(in_height_ % pool_height_dim_ == 0 && in_width_ % pool_width_dim_ == 0);
  KALDI_ASSERT(!(overlap_ && overlap2D_));
  const int32 maps = PooledMaps(in_channel_, pool_channel_dim_, overlap_, overlap2D_);
  KALDI_ASSERT(maps > 0);
  if (overlap_ || overlap2D_) {      // sliding windows run over the maps only
    KALDI_ASSERT(pool_height_dim_ == 1 && pool_width_dim_ == 1);
    KALDI_ASSERT(output_dim_ == in_height_ * in_width_ * maps);
  } else {
    KALDI_ASSERT(output_dim_ == (in_height_ / pool_height_dim_) * (in_width_ / pool_width_dim_) * maps);
  }
}

void MaxpoolComponent::Init(int32 input_dim, int32 output_dim, int32 in_height, int32 in_width,
                            int32 in_channel, int32 pool_height_dim, int32 pool_width_dim,
                            int32 pool_channel_dim, bool overlap, bool overlap2D) {
  input_dim_ = input_dim; output_dim_ = output_dim;
  in_height_ = in_height; in_width_ = in_width; in_channel_ = in_channel;
  pool_height_dim_ = pool_height_dim; pool_width_dim_ = pool_width_dim; pool_channel_dim_ = pool_channel_dim;
  overlap_ = overlap; overlap2D_ = overlap2D;
  Check();
}

// Config line (reference :814-867): the six geometry keys are required, overlap / overlap2D optional;
// input-dim and output-dim follow from them.
void MaxpoolComponent::InitFromString(std::string args) {
  const std::string line(args);
  MaxpoolComponent parsed;
  parsed.in_height_ = parsed.in_width_ = parsed.in_channel_ = 1;
  parsed.pool_height_dim_ = parsed.pool_width_dim_ = parsed.pool_channel_dim_ = 1;
  bool ok = parsed.StreamFields().ParseConfig(&args) && args.empty();
  ok = ok && parsed.pool_height_dim_ > 0 && parsed.pool_width_dim_ > 0;
  int32 positions = 0, maps = 0;
  if (ok) {
    maps = PooledMaps(parsed.in_channel_, parsed.pool_channel_dim_, parsed.overlap_, parsed.overlap2D_);
    positions = parsed.in_height_ * parsed.in_width_;
    if (!parsed.overlap_ && !parsed.overlap2D_) {
      // as the reference: one integer division of the whole input dim (the divisibility of H and W is
      // asserted by Init)
      positions = parsed.in_height_ * parsed.in_width_ / (parsed.pool_height_dim_ * parsed.pool_width_dim_);
    }
  }
  if (!ok || maps <= 0 || positions <= 0)
    KALDI_ERR << "Invalid initializer for layer of type " << Type() << ": \"" << line << "\"";
  Init(parsed.in_height_ * parsed.in_width_ * parsed.in_channel_, positions * maps, parsed.in_height_,
       parsed.in_width_, parsed.in_channel_, parsed.pool_height_dim_, parsed.pool_width_dim_,
       parsed.pool_channel_dim_, parsed.overlap_, parsed.overlap2D_);
}

// reference :869-880
void MaxpoolComponent::Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                                 const CuMatrixBase<BaseFloat> &in,
                                 CuMatrixBase<BaseFloat> *out) const {
  in_info.CheckSize(in);
  out_info.CheckSize(*out);
  if (index_routing_ && !overlap_ && !overlap2D_ &&
      pool_height_dim_ * pool_width_dim_ * pool_channel_dim_ <= 256) {
    CuDevice::Instantiate().RequireEnabled("MaxpoolComponent::Propagate");
    int32 stride = ((output_dim_ + 15) / 16) * 16;
    if (index_ == NULL || index_rows_ != in.NumRows() || index_stride_ != stride) {
      if (index_) CuDevice::Instantiate().Free(index_);
      index_ = static_cast<unsigned char *>(CuDevice::Instantiate().Malloc((size_t)stride * in.NumRows()));
      index_rows_ = in.NumRows();
      index_stride_ = stride;
    }
    cudaF_maxpool_prop_index(Str(), in.Data(), in.Dim(), out->Data(), out->Dim(), index_,
                             index_stride_, in_height_, in_width_, pool_height_dim_, pool_width_dim_,
                             pool_channel_dim_);
    CU_SAFE_CALL(cudaGetLastError());
    return;
  }
  in.Maxpool_prop(in_height_, in_width_, pool_height_dim_, pool_width_dim_, pool_channel_dim_,
                  overlap_, overlap2D_, out);
}

// reference :882-892: in_deriv->Resize(kSetZero) then Maxpool_backprop.  For the plain
// mode the zero fill and the routing are one kernel (every element of in_deriv is
// written exactly once), so in_deriv is only (re)sized here.
void MaxpoolComponent::Backprop(const ChunkInfo &, const ChunkInfo &,
                                const CuMatrixBase<BaseFloat> &in_value,
                                const CuMatrixBase<BaseFloat> &out_value,
                                const CuMatrixBase<BaseFloat> &out_deriv, Component *,
                                CuMatrix<BaseFloat> *in_deriv) const {
  KALDI_ASSERT(output_dim_ == out_value.NumCols());
  if (overlap_ || overlap2D_) {
    in_deriv->Resize(in_value.NumRows(), in_value.NumCols(), kSetZero);
    in_value.Maxpool_backprop(out_value, out_deriv, in_deriv, in_height_, in_width_,
                              pool_height_dim_, pool_width_dim_, pool_channel_dim_, overlap_,
                              overlap2D_);
    return;
  }
  CuDevice::Instantiate().RequireEnabled("MaxpoolComponent::Backprop");
  KALDI_ASSERT(in_value.NumCols() == input_dim_ && out_deriv.NumCols() == output_dim_ &&
               out_deriv.NumRows() == in_value.NumRows());
  in_deriv->Resize(in_value.NumRows(), in_value.NumCols(), kUndefined);
  if (index_routing_ && index_ != NULL && index_rows_ == in_value.NumRows()) {
    cudaF_maxpool_backprop_index(Str(), index_, index_stride_, out_deriv.Data(), out_deriv.Dim(),
                                 in_deriv->Data(), in_deriv->Dim(), in_height_, in_width_,
                                 pool_height_dim_, pool_width_dim_, pool_channel_dim_);
  } else {
    cudaF_maxpool_backprop_s(Str(), in_value.Data(), in_value.Dim(), out_value.Data(),
                             out_value.Dim(), out_deriv.Data(), out_deriv.Dim(), in_deriv->Data(),
                             in_deriv->Dim(), in_height_, in_width_, pool_height_dim_,
                             pool_width_dim_, pool_channel_dim_, KCNN_POOL_PLAIN, 1);
  }
  CU_SAFE_CALL(cudaGetLastError());
}

// Model stream (reference :894-959): eight ints, then <Overlap> / <Overlap2D> -- which files written
// before those modes existed do not have.  (The reference's reader asks for one token too many when
// only <Overlap> is present, :925-928, and can then only fail; such a stream is accepted here.)
void MaxpoolComponent::Read(std::istream &is, bool binary) {
  const FieldList f = StreamFields();
  ExpectOneOrTwoTokens(is, binary, "<" + Type() + ">", f.FirstToken());
  f.Read(is, binary, /*skip_first_token=*/true, 0, kMaxpoolOptionalFrom);
  overlap_ = overlap2D_ = false;
  f.ReadTail(is, binary, kMaxpoolOptionalFrom, "</" + Type() + ">");
}

void MaxpoolComponent::Write(std::ostream &os, bool binary) const {
  WriteToken(os, binary, "<" + Type() + ">");
  const_cast<MaxpoolComponent *>(this)->StreamFields().Write(os, binary);
  WriteToken(os, binary, "</" + Type() + ">");
}

std::string MaxpoolComponent::Info() const {
  std::ostringstream os;
  os << Type() << ", input-dim=" << input_dim_ << ", output-dim=" << output_dim_ << ", "
     << const_cast<MaxpoolComponent *>(this)->StreamFields().Describe();
  return os.str();
}

}  // namespace nnet0
}  // namespace cnsl
