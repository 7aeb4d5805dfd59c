// nnet0/maxpool-component.cc -- MaxpoolComponent for the B200 build.
// Follows reference src/nnet0/nnet-component-nnet0.cc:779-978.

#include <cmath>
#include <sstream>

#include "nnet0/nnet-component-nnet0.h"
#include "util/common-utils.h"
#include "cnsl-cu-kernels.h"

namespace cnsl {
namespace nnet0 {

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }

MaxpoolComponent::~MaxpoolComponent() {
  if (index_) CuDevice::Instantiate().Free(index_);
}

// reference :779-812
void MaxpoolComponent::Init(int32 input_dim, int32 output_dim, int32 in_height, int32 in_width,
                            int32 in_channel, int32 pool_height_dim, int32 pool_width_dim,
                            int32 pool_channel_dim, bool overlap, bool overlap2D) {
  input_dim_ = input_dim;
  output_dim_ = output_dim;
  in_height_ = in_height;
  in_width_ = in_width;
  in_channel_ = in_channel;
  pool_height_dim_ = pool_height_dim;
  pool_width_dim_ = pool_width_dim;
  pool_channel_dim_ = pool_channel_dim;
  overlap_ = overlap;
  overlap2D_ = overlap2D;

  KALDI_ASSERT((in_height_ * in_width_ * in_channel_) == input_dim_);
  KALDI_ASSERT(input_dim_ > 0 && output_dim_ > 0 && pool_height_dim_ > 0 && pool_width_dim_ > 0 &&
               pool_channel_dim_ > 0);
  KALDI_ASSERT(in_height_ % pool_height_dim_ == 0);
  KALDI_ASSERT(in_width_ % pool_width_dim_ == 0);
  KALDI_ASSERT((overlap && overlap2D) != true);

  if (overlap2D) {   // pooling region = pool_channel_dim x pool_channel_dim on a sqrt(C) x sqrt(C) map
    KALDI_ASSERT(pool_height_dim_ == 1 && pool_width_dim_ == 1);
    int32 output_channel = output_dim_ / (in_height_ * in_width_);
    int32 expected_output_channel = pow((sqrt(in_channel_) - pool_channel_dim_ + 1), 2);
    KALDI_ASSERT(output_channel == expected_output_channel);
  } else if (overlap) {
    KALDI_ASSERT(pool_height_dim_ == 1 && pool_width_dim_ == 1);
    KALDI_ASSERT(input_dim_ / in_channel_ * (in_channel_ - pool_channel_dim_ + 1) == output_dim_);
  } else {
    KALDI_ASSERT(input_dim_ % output_dim_ == 0);
    KALDI_ASSERT(in_channel_ % pool_channel_dim_ == 0);
    KALDI_ASSERT(input_dim_ / (pool_height_dim_ * pool_width_dim_ * pool_channel_dim_) == output_dim_);
  }
}

// reference :814-867
void MaxpoolComponent::InitFromString(std::string args) {
  std::string orig_args(args);
  int32 in_height = 1, in_width = 1, in_channel = 1;
  int32 pool_height_dim = 1, pool_width_dim = 1, pool_channel_dim = 1;
  bool overlap = false, overlap2D = false;

  bool ok = ParseFromString("in-height", &args, &in_height) &&
            ParseFromString("in-width", &args, &in_width) &&
            ParseFromString("in-channel", &args, &in_channel) &&
            ParseFromString("pool-height-dim", &args, &pool_height_dim) &&
            ParseFromString("pool-width-dim", &args, &pool_width_dim) &&
            ParseFromString("pool-channel-dim", &args, &pool_channel_dim);
  ParseFromString("overlap", &args, &overlap);
  ParseFromString("overlap2D", &args, &overlap2D);

  int32 input_dim = in_height * in_width * in_channel;
  int32 output_dim = 0;
  if (ok && in_channel > 0 && pool_height_dim > 0 && pool_width_dim > 0 && pool_channel_dim > 0) {
    if (overlap2D) {
      int32 output_channel = pow((sqrt(in_channel) - pool_channel_dim + 1), 2);
      output_dim = input_dim / in_channel * output_channel;
    } else if (overlap) {
      output_dim = input_dim / in_channel * (in_channel - pool_channel_dim + 1);
    } else {
      output_dim = input_dim / (pool_height_dim * pool_width_dim * pool_channel_dim);
    }
  }
  if (!ok || !args.empty() || output_dim <= 0)
    KALDI_ERR << "Invalid initializer for layer of type " << Type() << ": \"" << orig_args << "\"";
  Init(input_dim, output_dim, in_height, in_width, in_channel, pool_height_dim, pool_width_dim,
       pool_channel_dim, overlap, overlap2D);
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

// reference :894-934
void MaxpoolComponent::Read(std::istream &is, bool binary) {
  const std::string beg = "<" + Type() + ">", end = "</" + Type() + ">";
  ExpectOneOrTwoTokens(is, binary, beg, "<InputDim>");
  ReadBasicType(is, binary, &input_dim_);
  ExpectToken(is, binary, "<in_height>");
  ReadBasicType(is, binary, &in_height_);
  ExpectToken(is, binary, "<in_width>");
  ReadBasicType(is, binary, &in_width_);
  ExpectToken(is, binary, "<in_channel>");
  ReadBasicType(is, binary, &in_channel_);
  ExpectToken(is, binary, "<OutputDim>");
  ReadBasicType(is, binary, &output_dim_);
  ExpectToken(is, binary, "<PoolHeightDim>");
  ReadBasicType(is, binary, &pool_height_dim_);
  ExpectToken(is, binary, "<PoolWidthDim>");
  ReadBasicType(is, binary, &pool_width_dim_);
  ExpectToken(is, binary, "<PoolChannelDim>");
  ReadBasicType(is, binary, &pool_channel_dim_);
  std::string tok;
  ReadToken(is, binary, &tok);
  overlap_ = false;
  overlap2D_ = false;
  if (tok == "<Overlap>") {       // newer files; older ones end right here
    ReadBasicType(is, binary, &overlap_);
    ReadToken(is, binary, &tok);
    if (tok == "<Overlap2D>") {
      ReadBasicType(is, binary, &overlap2D_);
      ExpectToken(is, binary, end);
    } else {
      // the reference reads ONE MORE token here (ExpectToken after ReadToken, :925-928),
      // which can only fail; accept the closing token that was just read.
      KALDI_ASSERT(tok == end);
    }
  } else {
    KALDI_ASSERT(tok == end);
  }
}

// reference :936-959
void MaxpoolComponent::Write(std::ostream &os, bool binary) const {
  WriteToken(os, binary, "<MaxpoolComponent>");
  WriteToken(os, binary, "<InputDim>");
  WriteBasicType(os, binary, input_dim_);
  WriteToken(os, binary, "<in_height>");
  WriteBasicType(os, binary, in_height_);
  WriteToken(os, binary, "<in_width>");
  WriteBasicType(os, binary, in_width_);
  WriteToken(os, binary, "<in_channel>");
  WriteBasicType(os, binary, in_channel_);
  WriteToken(os, binary, "<OutputDim>");
  WriteBasicType(os, binary, output_dim_);
  WriteToken(os, binary, "<PoolHeightDim>");
  WriteBasicType(os, binary, pool_height_dim_);
  WriteToken(os, binary, "<PoolWidthDim>");
  WriteBasicType(os, binary, pool_width_dim_);
  WriteToken(os, binary, "<PoolChannelDim>");
  WriteBasicType(os, binary, pool_channel_dim_);
  WriteToken(os, binary, "<Overlap>");
  WriteBasicType(os, binary, overlap_);
  WriteToken(os, binary, "<Overlap2D>");
  WriteBasicType(os, binary, overlap2D_);
  WriteToken(os, binary, "</MaxpoolComponent>");
}

// reference :961-978
std::string MaxpoolComponent::Info() const {
  std::stringstream stream;
  stream << Type() << " input-dim=" << input_dim_ << " ( in-height=" << in_height_
         << ", in-width=" << in_width_ << ", in-channels=" << in_channel_
         << "), output-dim=" << output_dim_ << ", pool_height_dim_= " << pool_height_dim_
         << ", pool_width_dim_ = " << pool_width_dim_ << ", pool_channel_dim_ = " << pool_channel_dim_
         << ", max-pool-overlap_ = " << overlap_ << ", max-pool-overlap_2D = " << overlap2D_;
  return stream.str();
}

}  // namespace nnet0
}  // namespace cnsl
