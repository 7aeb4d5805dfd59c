// nnet2/nnet-component.cc -- shim of the nnet2 Component machinery (see the header).
// GPU only; every matrix operation is stream-ordered on CuDevice::Stream().

#include <algorithm>
#include <sstream>

#include "nnet2/nnet-component.h"
#include "nnet0/nnet-component-nnet0.h"
#include "util/common-utils.h"
#include "cnsl-cu-kernels.h"
#include "kcnn_common.cuh"

namespace kaldi {
namespace nnet2 {

static inline cudaStream_t Str() { return CuDevice::Instantiate().Stream(); }

// ------------------------------------------------------------------ ChunkInfo --

int32 ChunkInfo::GetIndex(int32 offset) const {
  if (offsets_.empty()) {
    KALDI_ASSERT((offset <= last_offset_) && (offset >= first_offset_));
    return offset - first_offset_;
  }
  std::vector<int32>::const_iterator iter = std::lower_bound(offsets_.begin(), offsets_.end(), offset);
  KALDI_ASSERT(iter != offsets_.end() && *iter == offset);
  return static_cast<int32>(iter - offsets_.begin());
}

int32 ChunkInfo::GetOffset(int32 index) const {
  if (offsets_.empty()) {
    int32 offset = index + first_offset_;
    KALDI_ASSERT((offset <= last_offset_) && (offset >= first_offset_));
    return offset;
  }
  KALDI_ASSERT((index >= 0) && (index < static_cast<int32>(offsets_.size())));
  return offsets_[index];
}

void ChunkInfo::Check() const {
  KALDI_ASSERT((feat_dim_ > 0) && (num_chunks_ > 0));
  if (!offsets_.empty()) {
    KALDI_ASSERT((first_offset_ == offsets_.front()) && (last_offset_ == offsets_.back()));
  } else {
    KALDI_ASSERT((first_offset_ >= 0) && (last_offset_ >= first_offset_));
  }
  KALDI_ASSERT(NumRows() % num_chunks_ == 0);
}

void ChunkInfo::CheckSize(const CuMatrixBase<BaseFloat> &mat) const {
  KALDI_ASSERT((mat.NumRows() == NumRows()) && (mat.NumCols() == NumCols()));
}

// ------------------------------------------------------------------ Component --

// static
Component *Component::NewComponentOfType(const std::string &component_type) {
  Component *ans = NULL;
  if (component_type == "SoftmaxComponent") {
    ans = new SoftmaxComponent();
  } else if (component_type == "RectifiedLinearComponent") {
    ans = new RectifiedLinearComponent();
  } else if (component_type == "AffineComponent") {
    ans = new AffineComponent();
  } else if (component_type == "DropoutComponent") {
    ans = new DropoutComponent();
  } else if (component_type == "NormalizeComponent") {
    ans = new NormalizeComponent();
  } else if (component_type == "SpliceComponent") {
    ans = new SpliceComponent();
  } else if (component_type == "ConvolutionComponent") {       // reference :112-113
    ans = new cnsl::nnet0::ConvolutionComponent();
  } else if (component_type == "MaxpoolComponent") {           // reference :114-115
    ans = new cnsl::nnet0::MaxpoolComponent();
  } else if (component_type == "FullyConnectedComponent") {    // reference :116-117
    ans = new cnsl::nnet0::FullyConnectedComponent();
  }
  return ans;
}

// static
Component *Component::ReadNew(std::istream &is, bool binary) {
  std::string token;
  ReadToken(is, binary, &token);   // e.g. "<SigmoidComponent>".
  token.erase(0, 1);
  token.erase(token.length() - 1);
  Component *ans = NewComponentOfType(token);
  if (!ans) KALDI_ERR << "Unknown component type " << token;
  ans->Read(is, binary);
  return ans;
}

// static
Component *Component::NewFromString(const std::string &initializer_line) {
  std::istringstream istr(initializer_line);
  std::string component_type;
  istr >> component_type >> std::ws;
  std::string rest_of_line;
  getline(istr, rest_of_line);
  Component *ans = NewComponentOfType(component_type);
  if (ans == NULL)
    KALDI_ERR << "Bad initializer line (no such type of Component): " << initializer_line;
  ans->InitFromString(rest_of_line);
  return ans;
}

std::string Component::Info() const {
  std::stringstream stream;
  stream << Type() << ", input-dim=" << InputDim() << ", output-dim=" << OutputDim();
  return stream.str();
}

std::string UpdatableComponent::Info() const {
  std::stringstream stream;
  stream << Type() << ", input-dim=" << InputDim() << ", output-dim=" << OutputDim()
         << ", learning-rate=" << LearningRate();
  return stream.str();
}

void ExpectOneOrTwoTokens(std::istream &is, bool binary, const std::string &token1,
                          const std::string &token2) {
  KALDI_ASSERT(token1 != token2);
  std::string temp;
  ReadToken(is, binary, &temp);
  if (temp == token1) {
    ExpectToken(is, binary, token2);
  } else if (temp != token2) {
    KALDI_ERR << "Expecting token " << token1 << " or " << token2 << " but got " << temp;
  }
}

// One scan for all the ParseFromString overloads: finds "name=value", removes it from
// *string, returns the value text.
static bool TakeOption(const std::string &name, std::string *string, std::string *value) {
  std::vector<std::string> pieces;
  SplitStringToVector(*string, " \t", true, &pieces);
  const std::string key = name + "=";
  for (size_t i = 0; i < pieces.size(); i++) {
    if (pieces[i].compare(0, key.length(), key) != 0) continue;
    *value = pieces[i].substr(key.length());
    std::string rest;
    for (size_t j = 0; j < pieces.size(); j++) {
      if (j == i) continue;
      if (!rest.empty()) rest += " ";
      rest += pieces[j];
    }
    *string = rest;
    return true;
  }
  return false;
}

bool ParseFromString(const std::string &name, std::string *string, int32 *param) {
  std::string v;
  if (!TakeOption(name, string, &v)) return false;
  if (!ConvertStringToInteger(v, param)) KALDI_ERR << "Bad option " << name << "=" << v;
  return true;
}
bool ParseFromString(const std::string &name, std::string *string, bool *param) {
  std::string v;
  if (!TakeOption(name, string, &v)) return false;
  if (v.empty()) KALDI_ERR << "Bad option " << name << "=" << v;
  if (v[0] == 'f' || v[0] == 'F') *param = false;
  else if (v[0] == 't' || v[0] == 'T') *param = true;
  else KALDI_ERR << "Bad option " << name << "=" << v;
  return true;
}
bool ParseFromString(const std::string &name, std::string *string, BaseFloat *param) {
  std::string v;
  if (!TakeOption(name, string, &v)) return false;
  if (!ConvertStringToReal(v, param)) KALDI_ERR << "Bad option " << name << "=" << v;
  return true;
}
bool ParseFromString(const std::string &name, std::string *string, std::string *param) {
  return TakeOption(name, string, param);
}
bool ParseFromString(const std::string &name, std::string *string, std::vector<int32> *param) {
  std::string v;
  if (!TakeOption(name, string, &v)) return false;
  if (!SplitStringToIntegers(v, ":", false, param)) KALDI_ERR << "Bad option " << name << "=" << v;
  return true;
}

// --------------------------------------------------------- NonlinearComponent --

namespace {
// Column sums of y and of [y > 0] over the rows, added to the double accumulators.
// A block owns 32 adjacent columns: 32 row-threads x 128-byte row segments.
__global__ void __launch_bounds__(1024)
nonlin_stats_kernel(const float *__restrict__ y, ::MatrixDim d, double *value_sum,
                    double *deriv_sum) {
  kcnn::pdl_prologue();
  __shared__ float pv[32][33], pd[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  float sv0 = 0.0f, sv1 = 0.0f, sd0 = 0.0f, sd1 = 0.0f;
  if (col < d.cols) {
    const float *p = y + col;
    int r = ty;
    for (; r + 32 < d.rows; r += 64) {
      float a = __ldg(p + (size_t)r * d.stride), b = __ldg(p + (size_t)(r + 32) * d.stride);
      sv0 += a; sd0 += a > 0.0f ? 1.0f : 0.0f;
      sv1 += b; sd1 += b > 0.0f ? 1.0f : 0.0f;
    }
    for (; r < d.rows; r += 32) {
      float a = __ldg(p + (size_t)r * d.stride);
      sv0 += a; sd0 += a > 0.0f ? 1.0f : 0.0f;
    }
  }
  pv[ty][tx] = sv0 + sv1; pd[ty][tx] = sd0 + sd1;
  __syncthreads();
  if (ty == 0 && col < d.cols) {
    float a = 0.0f, b = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) { a += pv[i][tx]; b += pd[i][tx]; }
    value_sum[col] += (double)a;
    if (deriv_sum) deriv_sum[col] += (double)b;
  }
}

// RectifiedLinearComponent::Backprop and the statistics of NonlinearComponent::UpdateStats
// (reference nnet2/nnet-component.cc:337-363, 813-827) in ONE pass over out_value:
// in_deriv = out_deriv * [y > 0], value_sum += colsum(y), deriv_sum += colsum([y > 0]).
// CTA = 128 columns (32 lanes x float4) x 64 rows (8 row-threads x 8 rows, all loads
// independent); per-CTA column sums go to the double accumulators with atomics (sums of a
// few hundred floats in double: order does not change the result in practice).
constexpr int kReluRowsPerCta = 64;
__global__ void __launch_bounds__(256)
relu_bprop_stats_kernel(const float *__restrict__ y, ::MatrixDim yd, const float *__restrict__ od,
                        int od_stride, float *__restrict__ id, int id_stride, double *value_sum,
                        double *deriv_sum) {
  kcnn::pdl_prologue();
  __shared__ float4 pv[8][33], pd[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 4;
  const int r0 = blockIdx.y * kReluRowsPerCta;
  float4 sv = make_float4(0.f, 0.f, 0.f, 0.f), sd = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < yd.cols) {
    float4 a[8], d[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int r = r0 + ty + 8 * k;
      if (r < yd.rows) {
        a[k] = __ldg(reinterpret_cast<const float4 *>(y + (size_t)r * yd.stride + col));
        d[k] = __ldg(reinterpret_cast<const float4 *>(od + (size_t)r * od_stride + col));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int r = r0 + ty + 8 * k;
      if (r < yd.rows) {
        float4 o;
        o.x = a[k].x > 0.0f ? d[k].x : 0.0f; o.y = a[k].y > 0.0f ? d[k].y : 0.0f;
        o.z = a[k].z > 0.0f ? d[k].z : 0.0f; o.w = a[k].w > 0.0f ? d[k].w : 0.0f;
        *reinterpret_cast<float4 *>(id + (size_t)r * id_stride + col) = o;
        sv.x += a[k].x; sv.y += a[k].y; sv.z += a[k].z; sv.w += a[k].w;
        sd.x += a[k].x > 0.0f ? 1.0f : 0.0f; sd.y += a[k].y > 0.0f ? 1.0f : 0.0f;
        sd.z += a[k].z > 0.0f ? 1.0f : 0.0f; sd.w += a[k].w > 0.0f ? 1.0f : 0.0f;
      }
    }
  }
  pv[ty][tx] = sv; pd[ty][tx] = sd;
  __syncthreads();
  if (ty == 0 && col < yd.cols) {
    float4 v = pv[0][tx], w = pd[0][tx];
#pragma unroll
    for (int i = 1; i < 8; i++) {
      v.x += pv[i][tx].x; v.y += pv[i][tx].y; v.z += pv[i][tx].z; v.w += pv[i][tx].w;
      w.x += pd[i][tx].x; w.y += pd[i][tx].y; w.z += pd[i][tx].z; w.w += pd[i][tx].w;
    }
    atomicAdd(value_sum + col, (double)v.x); atomicAdd(value_sum + col + 1, (double)v.y);
    atomicAdd(value_sum + col + 2, (double)v.z); atomicAdd(value_sum + col + 3, (double)v.w);
    atomicAdd(deriv_sum + col, (double)w.x); atomicAdd(deriv_sum + col + 1, (double)w.y);
    atomicAdd(deriv_sum + col + 2, (double)w.z); atomicAdd(deriv_sum + col + 3, (double)w.w);
  }
}

using kcnn::mix32;      // splitmix64 finaliser shared with the fused affine epilogue (kcnn_common.cuh)

// The mask is a pure function of (seed, element index i*cols + j).  Each thread owns up to four
// adjacent columns of one row (128-bit accesses when kVec4).
template <bool kVec4>
__global__ void __launch_bounds__(256)
dropout_fprop_kernel(const float *__restrict__ in, ::MatrixDim id, float *__restrict__ out,
                     ::MatrixDim od, float dp, float low, float high,
                     const unsigned long long *seed_dev, kcnn::FastDiv div_units) {
  kcnn::pdl_prologue();
  const int units = kVec4 ? od.cols / 4 : od.cols;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)od.rows * units) return;
  uint32_t i, u;
  div_units.divmod((uint32_t)t, i, u);
  const unsigned long long base = *seed_dev * 0x100000001B3ull + (unsigned long long)i * od.cols;
  if (kVec4) {
    const int j = 4 * (int)u;
    float4 v = __ldg(reinterpret_cast<const float4 *>(in + (size_t)i * id.stride) + u);
    float *e = &v.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float r = (mix32(base + (unsigned long long)(j + k)) >> 8) * (1.0f / 16777216.0f);
      e[k] *= (r - dp > 0.0f) ? high : low;
    }
    reinterpret_cast<float4 *>(out + (size_t)i * od.stride)[u] = v;
  } else {
    float r = (mix32(base + u) >> 8) * (1.0f / 16777216.0f);
    out[(size_t)i * od.stride + u] = ((r - dp > 0.0f) ? high : low) * __ldg(in + (size_t)i * id.stride + u);
  }
}

__global__ void bump_seed_kernel(unsigned long long *seed_dev) {
  kcnn::pdl_prologue(); *seed_dev += 1; }

// in_deriv = out_deriv .* out_value ./ in_value   (out_deriv where in_value == 0):
// Kaldi's AddMatMatDivMat, reference nnet2/nnet-component.cc:3634-3636.
template <bool kVec4>
__global__ void __launch_bounds__(256)
dropout_bprop_kernel(const float *__restrict__ iv, int iv_stride, const float *__restrict__ ov,
                     int ov_stride, const float *__restrict__ od, int od_stride,
                     float *__restrict__ id, ::MatrixDim idd, kcnn::FastDiv div_units) {
  kcnn::pdl_prologue();
  const int units = kVec4 ? idd.cols / 4 : idd.cols;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)idd.rows * units) return;
  uint32_t i, u;
  div_units.divmod((uint32_t)t, i, u);
  if (kVec4) {
    const float4 x = __ldg(reinterpret_cast<const float4 *>(iv + (size_t)i * iv_stride) + u);
    const float4 y = __ldg(reinterpret_cast<const float4 *>(ov + (size_t)i * ov_stride) + u);
    float4 d = __ldg(reinterpret_cast<const float4 *>(od + (size_t)i * od_stride) + u);
    d.x = x.x == 0.0f ? d.x : d.x * y.x / x.x;
    d.y = x.y == 0.0f ? d.y : d.y * y.y / x.y;
    d.z = x.z == 0.0f ? d.z : d.z * y.z / x.z;
    d.w = x.w == 0.0f ? d.w : d.w * y.w / x.w;
    reinterpret_cast<float4 *>(id + (size_t)i * idd.stride)[u] = d;
  } else {
    float x = __ldg(iv + (size_t)i * iv_stride + u), d = __ldg(od + (size_t)i * od_stride + u);
    id[(size_t)i * idd.stride + u] = x == 0.0f ? d : d * __ldg(ov + (size_t)i * ov_stride + u) / x;
  }
}

inline bool rows_vec4(int cols, std::initializer_list<int> strides, std::initializer_list<const void *> ptrs) {
  if (cols & 3) return false;
  for (int st : strides) if (st & 3) return false;
  for (const void *p : ptrs) if (reinterpret_cast<uintptr_t>(p) & 15u) return false;
  return true;
}
}  // namespace

NonlinearComponent::NonlinearComponent(const NonlinearComponent &other)
    : Component(), dim_(other.dim_), count_(other.count_), stats_(NULL), stats_dim_(0) {
  other.GetStats(&value_sum_host_, &deriv_sum_host_);
}

NonlinearComponent::~NonlinearComponent() {
  if (stats_) CuDevice::Instantiate().Free(stats_);
}

void NonlinearComponent::EnsureStats() {
  if (stats_ == NULL || stats_dim_ != dim_) {
    if (stats_) CuDevice::Instantiate().Free(stats_);
    stats_ = static_cast<double *>(CuDevice::Instantiate().Malloc(sizeof(double) * 2 * dim_));
    stats_dim_ = dim_;
    std::vector<double> init(2 * dim_, 0.0);
    for (int32 i = 0; i < dim_ && i < value_sum_host_.Dim(); i++) init[i] = value_sum_host_(i);
    for (int32 i = 0; i < dim_ && i < deriv_sum_host_.Dim(); i++) init[dim_ + i] = deriv_sum_host_(i);
    CU_SAFE_CALL(cudaMemcpyAsync(stats_, init.data(), sizeof(double) * 2 * dim_,
                                 cudaMemcpyHostToDevice, Str()));
    CU_SAFE_CALL(cudaStreamSynchronize(Str()));
  }
}

void NonlinearComponent::UpdateStats(const CuMatrixBase<BaseFloat> &out_value, bool relu_deriv) {
  KALDI_ASSERT(out_value.NumCols() == InputDim());
  EnsureStats();
  count_ += out_value.NumRows();
  if (out_value.NumRows() == 0) return;
  KCNN_LAUNCH(nonlin_stats_kernel, kcnn::ceil_div_u(dim_, 32), 1024, 0, Str(), out_value.Data(),
              out_value.Dim(), stats_, relu_deriv ? stats_ + dim_ : (double *)NULL);
}

void NonlinearComponent::BackpropReluWithStats(const CuMatrixBase<BaseFloat> &out_value,
                                               const CuMatrixBase<BaseFloat> &out_deriv,
                                               CuMatrix<BaseFloat> *in_deriv) {
  KALDI_ASSERT(out_value.NumCols() == InputDim());
  const bool v4 = (dim_ & 3) == 0 && (out_value.Stride() & 3) == 0 && (out_deriv.Stride() & 3) == 0 &&
                  (in_deriv->Stride() & 3) == 0 && ((uintptr_t)out_value.Data() & 15) == 0 &&
                  ((uintptr_t)out_deriv.Data() & 15) == 0 && ((uintptr_t)in_deriv->Data() & 15) == 0;
  if (!v4) {                                          // unaligned views: the two separate kernels
    cudaF_relu_bprop(Str(), out_value.Data(), out_value.Dim(), out_deriv.Data(), out_deriv.Dim(),
                     in_deriv->Data(), in_deriv->Dim());
    UpdateStats(out_value, true);
    return;
  }
  EnsureStats();
  count_ += out_value.NumRows();
  dim3 grid(kcnn::ceil_div_u(dim_, 128), kcnn::ceil_div_u(out_value.NumRows(), kReluRowsPerCta));
  KCNN_LAUNCH(relu_bprop_stats_kernel, grid, 256, 0, Str(), out_value.Data(), out_value.Dim(),
              out_deriv.Data(), out_deriv.Stride(), in_deriv->Data(), in_deriv->Stride(), stats_, stats_ + dim_);
}

void NonlinearComponent::GetStats(Vector<double> *value_sum, Vector<double> *deriv_sum) const {
  if (stats_ == NULL) {
    *value_sum = value_sum_host_;
    *deriv_sum = deriv_sum_host_;
    return;
  }
  std::vector<double> h(2 * stats_dim_);
  CU_SAFE_CALL(cudaMemcpyAsync(h.data(), stats_, sizeof(double) * 2 * stats_dim_,
                               cudaMemcpyDeviceToHost, Str()));
  CU_SAFE_CALL(cudaStreamSynchronize(Str()));
  value_sum->Resize(stats_dim_);
  deriv_sum->Resize(stats_dim_);
  for (int32 i = 0; i < stats_dim_; i++) { (*value_sum)(i) = h[i]; (*deriv_sum)(i) = h[stats_dim_ + i]; }
}

void NonlinearComponent::Read(std::istream &is, bool binary) {
  std::ostringstream ostr_beg, ostr_end;
  ostr_beg << "<" << Type() << ">";
  ostr_end << "</" << Type() << ">";
  ExpectOneOrTwoTokens(is, binary, ostr_beg.str(), "<Dim>");
  ReadBasicType(is, binary, &dim_);
  ExpectToken(is, binary, "<ValueSum>");
  value_sum_host_.Read(is, binary);
  ExpectToken(is, binary, "<DerivSum>");
  deriv_sum_host_.Read(is, binary);
  ExpectToken(is, binary, "<Count>");
  ReadBasicType(is, binary, &count_);
  ExpectToken(is, binary, ostr_end.str());
  if (stats_) { CuDevice::Instantiate().Free(stats_); stats_ = NULL; stats_dim_ = 0; }
}

void NonlinearComponent::Write(std::ostream &os, bool binary) const {
  std::ostringstream ostr_beg, ostr_end;
  ostr_beg << "<" << Type() << ">";
  ostr_end << "</" << Type() << ">";
  Vector<double> vs, ds;
  GetStats(&vs, &ds);
  WriteToken(os, binary, ostr_beg.str());
  WriteToken(os, binary, "<Dim>");
  WriteBasicType(os, binary, dim_);
  WriteToken(os, binary, "<ValueSum>");
  vs.Write(os, binary);
  WriteToken(os, binary, "<DerivSum>");
  ds.Write(os, binary);
  WriteToken(os, binary, "<Count>");
  WriteBasicType(os, binary, count_);
  WriteToken(os, binary, ostr_end.str());
}

void NonlinearComponent::InitFromString(std::string args) {
  std::string orig_args(args);
  int32 dim;
  bool ok = ParseFromString("dim", &args, &dim);
  if (!ok || !args.empty() || dim <= 0)
    KALDI_ERR << "Invalid initializer for layer of type " << Type() << ": \"" << orig_args << "\"";
  Init(dim);
}

// out = max(in, 0)                                     reference :799-806
void RectifiedLinearComponent::Propagate(const ChunkInfo &, const ChunkInfo &,
                                         const CuMatrixBase<BaseFloat> &in,
                                         CuMatrixBase<BaseFloat> *out) const {
  KALDI_ASSERT(in.NumRows() == out->NumRows() && in.NumCols() == out->NumCols());
  cudaF_relu_fprop(Str(), in.Data(), in.Dim(), out->Data(), out->Dim());
}

// in_deriv = out_deriv .* [out_value > 0], one pass     reference :808-827
void RectifiedLinearComponent::Backprop(const ChunkInfo &, const ChunkInfo &,
                                        const CuMatrixBase<BaseFloat> &,
                                        const CuMatrixBase<BaseFloat> &out_value,
                                        const CuMatrixBase<BaseFloat> &out_deriv,
                                        Component *to_update, CuMatrix<BaseFloat> *in_deriv) const {
  in_deriv->Resize(out_deriv.NumRows(), out_deriv.NumCols(), kUndefined);
  if (to_update != NULL && out_value.NumRows() > 0) {
    dynamic_cast<NonlinearComponent *>(to_update)->BackpropReluWithStats(out_value, out_deriv, in_deriv);
    return;
  }
  cudaF_relu_bprop(Str(), out_value.Data(), out_value.Dim(), out_deriv.Data(), out_deriv.Dim(),
                   in_deriv->Data(), in_deriv->Dim());
}

// reference :930-950
void SoftmaxComponent::Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                                 const CuMatrixBase<BaseFloat> &in,
                                 CuMatrixBase<BaseFloat> *out) const {
  in_info.CheckSize(in);
  out_info.CheckSize(*out);
  KALDI_ASSERT(in_info.NumChunks() == out_info.NumChunks());
  cudaF_softmax_fprop(Str(), in.Data(), in.Dim(), out->Data(), out->Dim());
}

// reference :952-1000
void SoftmaxComponent::Backprop(const ChunkInfo &, const ChunkInfo &, const CuMatrixBase<BaseFloat> &,
                                const CuMatrixBase<BaseFloat> &out_value,
                                const CuMatrixBase<BaseFloat> &out_deriv, Component *to_update,
                                CuMatrix<BaseFloat> *in_deriv) const {
  in_deriv->Resize(out_deriv.NumRows(), out_deriv.NumCols(), kUndefined);
  KALDI_ASSERT(out_value.NumRows() == out_deriv.NumRows() && out_value.NumCols() == out_deriv.NumCols());
  cudaF_softmax_bprop(Str(), out_value.Data(), out_value.Dim(), out_deriv.Data(), out_deriv.Dim(),
                      in_deriv->Data(), in_deriv->Dim());
  if (to_update != NULL)
    dynamic_cast<NonlinearComponent *>(to_update)->UpdateStats(out_value, false);
}

// --------------------------------------------------------- NormalizeComponent --

// reference :580-592 (CopyFromMat, AddDiagMat2, ApplyFloor, ApplyPow, MulRowsVec) as one kernel
void NormalizeComponent::Propagate(const ChunkInfo &, const ChunkInfo &, const CuMatrixBase<BaseFloat> &in,
                                   CuMatrixBase<BaseFloat> *out) const {
  KALDI_ASSERT(in.NumRows() == out->NumRows() && in.NumCols() == out->NumCols());
  cudaF_normalize_fprop(Str(), in.Data(), in.Dim(), out->Data(), out->Dim());
  CU_SAFE_CALL(cudaGetLastError());
}

// reference :616-639 as one kernel
void NormalizeComponent::Backprop(const ChunkInfo &, const ChunkInfo &, const CuMatrixBase<BaseFloat> &in_value,
                                  const CuMatrixBase<BaseFloat> &, const CuMatrixBase<BaseFloat> &out_deriv,
                                  Component *, CuMatrix<BaseFloat> *in_deriv) const {
  KALDI_ASSERT(in_value.NumRows() == out_deriv.NumRows() && in_value.NumCols() == out_deriv.NumCols());
  in_deriv->Resize(out_deriv.NumRows(), out_deriv.NumCols(), kUndefined);
  cudaF_normalize_bprop(Str(), in_value.Data(), in_value.Dim(), out_deriv.Data(), out_deriv.Dim(), in_deriv->Data(),
                        in_deriv->Dim());
  CU_SAFE_CALL(cudaGetLastError());
}

// ------------------------------------------------------------ SpliceComponent --

namespace {
// out[(chunk, oi)][c * dim + j] = in[(chunk, map[c][oi])][j]; the last const_dim columns come from
// input frame oi of the chunk.  One thread per output element, rows contiguous.
__global__ void __launch_bounds__(256)
splice_fprop_kernel(const float *__restrict__ in, int in_stride, float *__restrict__ out, int out_stride,
                    long long total, int out_cols, int dim, int const_dim, int num_splice, int in_chunk,
                    int out_chunk, const int *__restrict__ map, kcnn::FastDiv div_cols, kcnn::FastDiv div_dim,
                    kcnn::FastDiv div_oc) {
  kcnn::pdl_prologue();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  uint32_t row, col, chunk, oi;
  div_cols.divmod((uint32_t)t, row, col);
  div_oc.divmod(row, chunk, oi);
  int src_row, src_col;
  if ((int)col < num_splice * dim) {
    uint32_t c, j;
    div_dim.divmod(col, c, j);
    src_row = (int)chunk * in_chunk + __ldg(map + c * out_chunk + oi);
    src_col = (int)j;
  } else {
    src_row = (int)chunk * in_chunk + (int)oi;
    src_col = dim + ((int)col - num_splice * dim);
  }
  out[(size_t)row * out_stride + col] = __ldg(in + (size_t)src_row * in_stride + src_col);
}

// in_deriv[(chunk, ii)][j] = sum over (c, oi) with map[c][oi] == ii of out_deriv[(chunk, oi)][c * dim + j]
// (gather form of the reference's CopyRows + AddMat chain, :2745-2848; fixed summation order c, oi)
__global__ void __launch_bounds__(256)
splice_bprop_kernel(const float *__restrict__ od, int od_stride, float *__restrict__ id, int id_stride,
                    long long total, int in_cols, int dim, int const_dim, int num_splice, int in_chunk,
                    int out_chunk, const int *__restrict__ map, kcnn::FastDiv div_cols, kcnn::FastDiv div_ic) {
  kcnn::pdl_prologue();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  uint32_t row, col, chunk, ii;
  div_cols.divmod((uint32_t)t, row, col);
  div_ic.divmod(row, chunk, ii);
  float s = 0.0f;
  if ((int)col < dim) {
    for (int c = 0; c < num_splice; c++)
      for (int oi = 0; oi < out_chunk; oi++)
        if (__ldg(map + c * out_chunk + oi) == (int)ii)
          s += __ldg(od + ((size_t)chunk * out_chunk + oi) * od_stride + c * dim + col);
  } else if ((int)ii < out_chunk) {
    s = __ldg(od + ((size_t)chunk * out_chunk + ii) * od_stride + num_splice * dim + ((int)col - dim));
  }
  id[(size_t)row * id_stride + col] = s;
}
}  // namespace

SpliceComponent::~SpliceComponent() {
  if (map_dev_) CuDevice::Instantiate().Free(map_dev_);
}

void SpliceComponent::Init(int32 input_dim, std::vector<int32> context, int32 const_component_dim) {
  input_dim_ = input_dim;
  const_component_dim_ = const_component_dim;
  context_ = context;
  KALDI_ASSERT(context_.size() > 0);
  KALDI_ASSERT(input_dim_ > 0 && context_.front() <= 0 && context_.back() >= 0);
  for (size_t i = 1; i < context_.size(); i++) KALDI_ASSERT(context_[i] > context_[i - 1]);   // sorted, unique
  KALDI_ASSERT(const_component_dim_ >= 0 && const_component_dim_ < input_dim_);
}

// "input-dim=40 left-context=10 right-context=10 [const-component-dim=0]" or "input-dim=40 context=-2:0:2"
void SpliceComponent::InitFromString(std::string args) {
  const std::string orig_args(args);
  int32 input_dim = 0, left_context = 0, right_context = 0, const_component_dim = 0;
  std::vector<int32> context;
  const bool in_dim_ok = ParseFromString("input-dim", &args, &input_dim);
  const bool context_ok = ParseFromString("context", &args, &context);
  const bool left_ok = ParseFromString("left-context", &args, &left_context);
  const bool right_ok = ParseFromString("right-context", &args, &right_context);
  ParseFromString("const-component-dim", &args, &const_component_dim);
  if (!(in_dim_ok && (context_ok || (left_ok && right_ok))) || !args.empty() || input_dim <= 0)
    KALDI_ERR << "Invalid initializer for layer of type " << Type() << ": \"" << orig_args << "\"";
  if (left_ok && right_ok) {
    KALDI_ASSERT(context.size() == 0);
    for (int32 i = -left_context; i <= right_context; i++) context.push_back(i);
  }
  Init(input_dim, context, const_component_dim);
}

int32 SpliceComponent::OutputDim() const {
  return (input_dim_ - const_component_dim_) * static_cast<int32>(context_.size()) + const_component_dim_;
}

bool SpliceComponent::IsContiguousWindow() const {
  if (const_component_dim_ != 0 || context_.empty()) return false;
  for (size_t i = 1; i < context_.size(); i++)
    if (context_[i] != context_[i - 1] + 1) return false;
  return true;
}

std::string SpliceComponent::Info() const {
  std::stringstream stream;
  stream << Component::Info() << ", context=";
  for (size_t i = 0; i < context_.size(); i++) stream << context_[i] << " ";
  if (const_component_dim_ != 0) stream << ", const_component_dim=" << const_component_dim_;
  return stream.str();
}

const int32 *SpliceComponent::FrameMap(const ChunkInfo &in_info, const ChunkInfo &out_info) const {
  const int32 out_chunk = out_info.ChunkSize(), num_splice = static_cast<int32>(context_.size());
  std::vector<int32> map(static_cast<size_t>(num_splice) * out_chunk);
  for (int32 c = 0; c < num_splice; c++)
    for (int32 oi = 0; oi < out_chunk; oi++)
      map[c * out_chunk + oi] = in_info.GetIndex(out_info.GetOffset(oi) + context_[c]);     // reference :2670-2676
  const uint64 key = HashBytes(map.data(), sizeof(int32) * map.size(), 1469598103934665603ull) | 1;
  if (key != map_key_ || map.size() != map_len_) {
    if (map_dev_) CuDevice::Instantiate().Free(map_dev_);
    map_dev_ = static_cast<int32 *>(CuDevice::Instantiate().Malloc(sizeof(int32) * map.size()));
    CU_SAFE_CALL(cudaMemcpyAsync(map_dev_, map.data(), sizeof(int32) * map.size(), cudaMemcpyHostToDevice, Str()));
    CU_SAFE_CALL(cudaStreamSynchronize(Str()));
    map_len_ = map.size();
    map_key_ = key;
  }
  return map_dev_;
}

// reference :2640-2722: one gather kernel instead of one CopyRows per context offset
void SpliceComponent::Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                                const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) const {
  in_info.Check();
  out_info.Check();
  in_info.CheckSize(in);
  out_info.CheckSize(*out);
  KALDI_ASSERT(in_info.NumChunks() == out_info.NumChunks());
  const int32 in_chunk = in_info.ChunkSize(), out_chunk = out_info.ChunkSize();
  if (out_chunk <= 0) KALDI_ERR << "Splicing features: output will have zero dimension. Probably a code error.";
  const int32 dim = input_dim_ - const_component_dim_, num_splice = static_cast<int32>(context_.size());
  const int32 *map = FrameMap(in_info, out_info);
  const long long total = (long long)out->NumRows() * out->NumCols();
  if (total == 0) return;
  KCNN_LAUNCH(splice_fprop_kernel, kcnn::ceil_div_u(total, 256), 256, 0, Str(), in.Data(), in.Stride(), out->Data(),
              out->Stride(), total, out->NumCols(), dim, const_component_dim_, num_splice, in_chunk, out_chunk, map,
              kcnn::FastDiv((uint32_t)out->NumCols()), kcnn::FastDiv((uint32_t)dim), kcnn::FastDiv((uint32_t)out_chunk));
  CU_SAFE_CALL(cudaGetLastError());
}

// reference :2724-2848
void SpliceComponent::Backprop(const ChunkInfo &in_info, const ChunkInfo &out_info, const CuMatrixBase<BaseFloat> &,
                               const CuMatrixBase<BaseFloat> &, const CuMatrixBase<BaseFloat> &out_deriv, Component *,
                               CuMatrix<BaseFloat> *in_deriv) const {
  in_info.Check();
  out_info.Check();
  out_info.CheckSize(out_deriv);
  in_deriv->Resize(in_info.NumRows(), in_info.NumCols(), kUndefined);
  KALDI_ASSERT(in_info.NumChunks() == out_info.NumChunks() && OutputDim() == out_deriv.NumCols());
  const int32 in_chunk = in_info.ChunkSize(), out_chunk = out_info.ChunkSize();
  const int32 dim = input_dim_ - const_component_dim_, num_splice = static_cast<int32>(context_.size());
  const int32 *map = FrameMap(in_info, out_info);
  const long long total = (long long)in_deriv->NumRows() * in_deriv->NumCols();
  if (total == 0) return;
  KCNN_LAUNCH(splice_bprop_kernel, kcnn::ceil_div_u(total, 256), 256, 0, Str(), out_deriv.Data(), out_deriv.Stride(),
              in_deriv->Data(), in_deriv->Stride(), total, in_deriv->NumCols(), dim, const_component_dim_, num_splice,
              in_chunk, out_chunk, map, kcnn::FastDiv((uint32_t)in_deriv->NumCols()), kcnn::FastDiv((uint32_t)in_chunk));
  CU_SAFE_CALL(cudaGetLastError());
}

Component *SpliceComponent::Copy() const {
  SpliceComponent *ans = new SpliceComponent();
  ans->Init(input_dim_, context_, const_component_dim_);
  return ans;
}

// token stream of reference :2857-2890 (either <LeftContext> <RightContext> or <Context> on input)
void SpliceComponent::Read(std::istream &is, bool binary) {
  ExpectOneOrTwoTokens(is, binary, "<SpliceComponent>", "<InputDim>");
  ReadBasicType(is, binary, &input_dim_);
  std::string token;
  ReadToken(is, false, &token);
  context_.clear();
  if (token == "<LeftContext>") {
    int32 left_context = 0, right_context = 0;
    ReadBasicType(is, binary, &left_context);
    ExpectToken(is, binary, "<RightContext>");
    ReadBasicType(is, binary, &right_context);
    for (int32 i = -left_context; i <= right_context; i++) context_.push_back(i);
  } else if (token == "<Context>") {
    ReadIntegerVector(is, binary, &context_);
  } else {
    KALDI_ERR << "Unknown token" << token << ", the model might be corrupted";
  }
  ExpectToken(is, binary, "<ConstComponentDim>");
  ReadBasicType(is, binary, &const_component_dim_);
  ExpectToken(is, binary, "</SpliceComponent>");
}

void SpliceComponent::Write(std::ostream &os, bool binary) const {
  WriteToken(os, binary, "<SpliceComponent>");
  WriteToken(os, binary, "<InputDim>");
  WriteBasicType(os, binary, input_dim_);
  WriteToken(os, binary, "<Context>");
  WriteIntegerVector(os, binary, context_);
  WriteToken(os, binary, "<ConstComponentDim>");
  WriteBasicType(os, binary, const_component_dim_);
  WriteToken(os, binary, "</SpliceComponent>");
}

// ----------------------------------------------------------- DropoutComponent --

void DropoutComponent::Init(int32 dim, BaseFloat dropout_proportion, BaseFloat dropout_scale) {
  dim_ = dim;
  dropout_proportion_ = dropout_proportion;
  dropout_scale_ = dropout_scale;
}

void DropoutComponent::InitFromString(std::string args) {
  std::string orig_args(args);
  int32 dim;
  BaseFloat dropout_proportion = 0.5, dropout_scale = 0.0;
  bool ok = ParseFromString("dim", &args, &dim);
  ParseFromString("dropout-proportion", &args, &dropout_proportion);
  ParseFromString("dropout-scale", &args, &dropout_scale);
  if (!ok || !args.empty() || dim <= 0)
    KALDI_ERR << "Invalid initializer for layer of type DropoutComponent: \"" << orig_args << "\"";
  Init(dim, dropout_proportion, dropout_scale);
}

void DropoutComponent::Read(std::istream &is, bool binary) {
  ExpectOneOrTwoTokens(is, binary, "<DropoutComponent>", "<Dim>");
  ReadBasicType(is, binary, &dim_);
  ExpectToken(is, binary, "<DropoutScale>");
  ReadBasicType(is, binary, &dropout_scale_);
  ExpectToken(is, binary, "<DropoutProportion>");
  ReadBasicType(is, binary, &dropout_proportion_);
  ExpectToken(is, binary, "</DropoutComponent>");
}

void DropoutComponent::Write(std::ostream &os, bool binary) const {
  WriteToken(os, binary, "<DropoutComponent>");
  WriteToken(os, binary, "<Dim>");
  WriteBasicType(os, binary, dim_);
  WriteToken(os, binary, "<DropoutScale>");
  WriteBasicType(os, binary, dropout_scale_);
  WriteToken(os, binary, "<DropoutProportion>");
  WriteBasicType(os, binary, dropout_proportion_);
  WriteToken(os, binary, "</DropoutComponent>");
}

std::string DropoutComponent::Info() const {
  std::stringstream stream;
  stream << Type() << ", dim = " << dim_ << ", dropout-proportion = " << dropout_proportion_
         << ", dropout-scale = " << dropout_scale_;
  return stream.str();
}

DropoutComponent::~DropoutComponent() {
  if (seed_dev_) CuDevice::Instantiate().Free(seed_dev_);
}

unsigned long long *DropoutComponent::SeedDevice() const {
  if (seed_dev_ == NULL) {
    seed_dev_ = static_cast<unsigned long long *>(CuDevice::Instantiate().Malloc(sizeof(unsigned long long)));
    unsigned long long s0 = CuDevice::Instantiate().NextRandSeed();
    CU_SAFE_CALL(cudaMemcpyAsync(seed_dev_, &s0, sizeof(s0), cudaMemcpyHostToDevice, Str()));
    CU_SAFE_CALL(cudaStreamSynchronize(Str()));
  }
  return seed_dev_;
}

Component *DropoutComponent::Copy() const {
  return new DropoutComponent(dim_, dropout_proportion_, dropout_scale_);
}

void DropoutComponent::Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                                 const CuMatrixBase<BaseFloat> &in,
                                 CuMatrixBase<BaseFloat> *out) const {
  in_info.CheckSize(in);
  out_info.CheckSize(*out);
  KALDI_ASSERT(in_info.NumChunks() == out_info.NumChunks());
  KALDI_ASSERT(in.NumCols() == this->InputDim());
  BaseFloat dp = dropout_proportion_;
  KALDI_ASSERT(dp < 1.0 && dp >= 0.0);
  KALDI_ASSERT(dropout_scale_ <= 1.0 && dropout_scale_ >= 0.0);
  BaseFloat low_scale = dropout_scale_, high_scale = (1.0 - (dp * low_scale)) / (1.0 - dp);
  long long total = (long long)out->NumRows() * out->NumCols();
  if (total == 0) return;
  // One fused pass (the reference draws uniforms, thresholds, scales and multiplies in
  // five passes, :3600-3620).  The mask is a pure function of (seed, element); the seed
  // lives on the device and is bumped by a one-thread kernel, so a replayed CUDA graph
  // draws a fresh mask every step.
  SeedDevice();
  const bool v4 = rows_vec4(out->NumCols(), {in.Stride(), out->Stride()}, {in.Data(), out->Data()});
  const int units = v4 ? out->NumCols() / 4 : out->NumCols();
  const unsigned grid = kcnn::ceil_div_u((long long)out->NumRows() * units, 256);
  if (v4)
    KCNN_LAUNCH(dropout_fprop_kernel<true>, grid, 256, 0, Str(), in.Data(), in.Dim(), out->Data(), out->Dim(),
                dp, low_scale, high_scale, seed_dev_, kcnn::FastDiv((uint32_t)units));
  else
    KCNN_LAUNCH(dropout_fprop_kernel<false>, grid, 256, 0, Str(), in.Data(), in.Dim(), out->Data(), out->Dim(),
                dp, low_scale, high_scale, seed_dev_, kcnn::FastDiv((uint32_t)units));
  KCNN_LAUNCH(bump_seed_kernel, 1, 1, 0, Str(), seed_dev_);
}

void DropoutComponent::Backprop(const ChunkInfo &, const ChunkInfo &,
                                const CuMatrixBase<BaseFloat> &in_value,
                                const CuMatrixBase<BaseFloat> &out_value,
                                const CuMatrixBase<BaseFloat> &out_deriv, Component *,
                                CuMatrix<BaseFloat> *in_deriv) const {
  KALDI_ASSERT(in_value.NumRows() == out_value.NumRows() && in_value.NumCols() == out_value.NumCols());
  in_deriv->Resize(out_deriv.NumRows(), out_deriv.NumCols(), kUndefined);
  long long total = (long long)out_deriv.NumRows() * out_deriv.NumCols();
  if (total == 0) return;
  const bool v4 = rows_vec4(in_deriv->NumCols(),
                            {in_value.Stride(), out_value.Stride(), out_deriv.Stride(), in_deriv->Stride()},
                            {in_value.Data(), out_value.Data(), out_deriv.Data(), in_deriv->Data()});
  const int units = v4 ? in_deriv->NumCols() / 4 : in_deriv->NumCols();
  const unsigned grid = kcnn::ceil_div_u((long long)in_deriv->NumRows() * units, 256);
  if (v4)
    KCNN_LAUNCH(dropout_bprop_kernel<true>, grid, 256, 0, Str(), in_value.Data(), in_value.Stride(),
                out_value.Data(), out_value.Stride(), out_deriv.Data(), out_deriv.Stride(), in_deriv->Data(),
                in_deriv->Dim(), kcnn::FastDiv((uint32_t)units));
  else
    KCNN_LAUNCH(dropout_bprop_kernel<false>, grid, 256, 0, Str(), in_value.Data(), in_value.Stride(),
                out_value.Data(), out_value.Stride(), out_deriv.Data(), out_deriv.Stride(), in_deriv->Data(),
                in_deriv->Dim(), kcnn::FastDiv((uint32_t)units));
}

// ------------------------------------------------------------ AffineComponent --

AffineComponent::AffineComponent(const AffineComponent &component)
    : UpdatableComponent(component), linear_params_(component.linear_params_),
      bias_params_(component.bias_params_), is_gradient_(component.is_gradient_),
      deferred_(false), grad_external_(false) {}

AffineComponent::AffineComponent(const CuMatrixBase<BaseFloat> &linear_params,
                                 const CuVectorBase<BaseFloat> &bias_params, BaseFloat learning_rate)
    : UpdatableComponent(learning_rate), linear_params_(linear_params), bias_params_(bias_params),
      is_gradient_(false), deferred_(false), grad_external_(false) {
  KALDI_ASSERT(linear_params.NumRows() == bias_params.Dim() && bias_params.Dim() != 0);
}

void AffineComponent::SetZero(bool treat_as_gradient) {
  if (treat_as_gradient) SetLearningRate(1.0);
  linear_params_.SetZero();
  bias_params_.SetZero();
  if (treat_as_gradient) is_gradient_ = true;
}

void AffineComponent::SetParams(const VectorBase<BaseFloat> &bias, const MatrixBase<BaseFloat> &linear) {
  bias_params_ = bias;
  linear_params_ = linear;
  KALDI_ASSERT(bias_params_.Dim() == linear_params_.NumRows());
}

void AffineComponent::PerturbParams(BaseFloat stddev) {
  CuMatrix<BaseFloat> temp_linear_params(linear_params_);
  temp_linear_params.SetRandn();
  linear_params_.AddMat(stddev, temp_linear_params);
  CuVector<BaseFloat> temp_bias_params(bias_params_);
  temp_bias_params.SetRandn();
  bias_params_.AddVec(stddev, temp_bias_params);
}

std::string AffineComponent::Info() const {
  std::stringstream stream;
  BaseFloat linear_params_size = static_cast<BaseFloat>(linear_params_.NumRows()) *
                                 static_cast<BaseFloat>(linear_params_.NumCols());
  BaseFloat linear_stddev = std::sqrt(TraceMatMat(linear_params_, linear_params_, kTrans) / linear_params_size),
            bias_stddev = std::sqrt(VecVec(bias_params_, bias_params_) / bias_params_.Dim());
  stream << Type() << ", input-dim=" << InputDim() << ", output-dim=" << OutputDim()
         << ", linear-params-stddev=" << linear_stddev << ", bias-params-stddev=" << bias_stddev
         << ", learning-rate=" << LearningRate();
  return stream.str();
}

Component *AffineComponent::Copy() const {
  AffineComponent *ans = new AffineComponent();
  ans->learning_rate_ = learning_rate_;
  ans->linear_params_ = linear_params_;
  ans->bias_params_ = bias_params_;
  ans->is_gradient_ = is_gradient_;
  return ans;
}

BaseFloat AffineComponent::DotProduct(const UpdatableComponent &other_in) const {
  const AffineComponent *other = dynamic_cast<const AffineComponent *>(&other_in);
  return TraceMatMat(linear_params_, other->linear_params_, kTrans) +
         VecVec(bias_params_, other->bias_params_);
}

void AffineComponent::Init(BaseFloat learning_rate, int32 input_dim, int32 output_dim,
                           BaseFloat param_stddev, BaseFloat bias_stddev) {
  UpdatableComponent::Init(learning_rate);
  linear_params_.Resize(output_dim, input_dim);
  bias_params_.Resize(output_dim);
  KALDI_ASSERT(output_dim > 0 && input_dim > 0 && param_stddev >= 0.0);
  linear_params_.SetRandn();
  linear_params_.Scale(param_stddev);
  bias_params_.SetRandn();
  bias_params_.Scale(bias_stddev);
}

void AffineComponent::Init(BaseFloat learning_rate, std::string matrix_filename) {
  UpdatableComponent::Init(learning_rate);
  CuMatrix<BaseFloat> mat;
  ReadKaldiObject(matrix_filename, &mat);
  KALDI_ASSERT(mat.NumCols() >= 2);
  int32 input_dim = mat.NumCols() - 1, output_dim = mat.NumRows();
  linear_params_.Resize(output_dim, input_dim);
  bias_params_.Resize(output_dim);
  linear_params_.CopyFromMat(mat.Range(0, output_dim, 0, input_dim));
  bias_params_.CopyColFromMat(mat, input_dim);
}

void AffineComponent::InitFromString(std::string args) {
  std::string orig_args(args);
  bool ok = true;
  BaseFloat learning_rate = learning_rate_;
  std::string matrix_filename;
  int32 input_dim = -1, output_dim = -1;
  ParseFromString("learning-rate", &args, &learning_rate);
  if (ParseFromString("matrix", &args, &matrix_filename)) {
    Init(learning_rate, matrix_filename);
    if (ParseFromString("input-dim", &args, &input_dim))
      KALDI_ASSERT(input_dim == InputDim() && "input-dim mismatch vs. matrix.");
    if (ParseFromString("output-dim", &args, &output_dim))
      KALDI_ASSERT(output_dim == OutputDim() && "output-dim mismatch vs. matrix.");
  } else {
    ok = ok && ParseFromString("input-dim", &args, &input_dim);
    ok = ok && ParseFromString("output-dim", &args, &output_dim);
    BaseFloat param_stddev = 1.0 / std::sqrt(input_dim), bias_stddev = 1.0;
    ParseFromString("param-stddev", &args, &param_stddev);
    ParseFromString("bias-stddev", &args, &bias_stddev);
    Init(learning_rate, input_dim, output_dim, param_stddev, bias_stddev);
  }
  if (!args.empty()) KALDI_ERR << "Could not process these elements in initializer: " << args;
  if (!ok) KALDI_ERR << "Bad initializer " << orig_args;
}

// out = 1 bias^T + in W^T in ONE launch: the bias row copy (CopyRowsFromVec) is the
// GEMM epilogue.  reference :1216-1228.
void AffineComponent::PropagateAct(const ChunkInfo &in_info, const ChunkInfo &out_info,
                                   const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out,
                                   int act) const {
  in_info.CheckSize(in);
  out_info.CheckSize(*out);
  KALDI_ASSERT(in_info.NumChunks() == out_info.NumChunks());
  KALDI_ASSERT(in.NumCols() == InputDim() && out->NumCols() == OutputDim());
  CuDevice::Instantiate().RequireEnabled("AffineComponent::Propagate");
  cudaF_affine_fprop_act(Str(), CuDevice::Instantiate().MathMode(), in.Data(), in.Dim(),
                         linear_params_.Data(), linear_params_.Dim(), bias_params_.Data(), out->Data(),
                         out->Dim(), act);
  CU_SAFE_CALL(cudaGetLastError());
}

void AffineComponent::Propagate(const ChunkInfo &in_info, const ChunkInfo &out_info,
                                const CuMatrixBase<BaseFloat> &in,
                                CuMatrixBase<BaseFloat> *out) const {
  PropagateAct(in_info, out_info, in, out, KCNN_ACT_NONE);
}

bool AffineComponent::PropagateRelu(const ChunkInfo &in_info, const ChunkInfo &out_info,
                                    const CuMatrixBase<BaseFloat> &in,
                                    CuMatrixBase<BaseFloat> *out) const {
  PropagateAct(in_info, out_info, in, out, KCNN_ACT_RELU);
  return true;
}

void AffineComponent::EnsureGradBuffers() {
  if (grad_external_) return;
  if (w_grad_store_.NumRows() != linear_params_.NumRows() ||
      w_grad_store_.NumCols() != linear_params_.NumCols()) {
    w_grad_store_.Resize(linear_params_.NumRows(), linear_params_.NumCols(), kUndefined);
    b_grad_store_.Resize(bias_params_.Dim(), kUndefined);
  }
  w_grad_.data = w_grad_store_.Data(); w_grad_.rows = w_grad_store_.NumRows();
  w_grad_.cols = w_grad_store_.NumCols(); w_grad_.stride = w_grad_store_.Stride();
  b_grad_.data = b_grad_store_.Data(); b_grad_.rows = 1;
  b_grad_.cols = b_grad_store_.Dim(); b_grad_.stride = b_grad_store_.Dim();
}

size_t AffineComponent::GradientFloats() const {
  size_t stride = CuDevice::PitchInElements(linear_params_.NumCols(), sizeof(BaseFloat));
  return stride * linear_params_.NumRows() + CuDevice::PitchInElements(bias_params_.Dim(), sizeof(BaseFloat));
}

void AffineComponent::SetGradientStorage(float *base) {
  if (base == NULL) { grad_external_ = false; return; }
  int32 stride = CuDevice::PitchInElements(linear_params_.NumCols(), sizeof(BaseFloat));
  w_grad_.data = base; w_grad_.rows = linear_params_.NumRows();
  w_grad_.cols = linear_params_.NumCols(); w_grad_.stride = stride;
  b_grad_.data = base + (size_t)stride * linear_params_.NumRows(); b_grad_.rows = 1;
  b_grad_.cols = bias_params_.Dim(); b_grad_.stride = bias_params_.Dim();
  grad_external_ = true;
}

void AffineComponent::SetParameterStorage(float *base) {
  const int32 rows = linear_params_.NumRows(), cols = linear_params_.NumCols(), dim = bias_params_.Dim();
  const int32 stride = CuDevice::PitchInElements(cols, sizeof(BaseFloat));
  CuMatrix<BaseFloat> w(linear_params_);
  CuVector<BaseFloat> b(bias_params_);
  if (base != NULL) {
    linear_params_.Borrow(base, rows, cols, stride);
    bias_params_.Borrow(base + (size_t)stride * rows, dim);
    linear_params_.CopyFromMat(w);
    bias_params_.CopyFromVec(b);
  } else {
    linear_params_.Swap(&w);                 // w owns its copy: linear_params_ is an ordinary matrix again
    bias_params_.Borrow(NULL, 0);
    bias_params_ = b;
  }
}

std::vector<UpdatableComponent::GradBuffer> AffineComponent::GradientBuffers() {
  EnsureGradBuffers();
  std::vector<GradBuffer> v;
  v.push_back(w_grad_);
  v.push_back(b_grad_);
  return v;
}

void AffineComponent::ComputeGradient(const CuMatrixBase<BaseFloat> &in_value,
                                      const CuMatrixBase<BaseFloat> &out_deriv) {
  EnsureGradBuffers();
  ::MatrixDim gd = {w_grad_.rows, w_grad_.cols, w_grad_.stride};
  cudaF_affine_wgrad(Str(), CuDevice::Instantiate().MathMode(), in_value.Data(), in_value.Dim(),
                     out_deriv.Data(), out_deriv.Dim(), w_grad_.data, gd, b_grad_.data);
  CU_SAFE_CALL(cudaGetLastError());
}

// Stock update (no momentum): bias += lr * colsum(out_deriv); W += lr * out_deriv^T in.
// reference :1230-1235.
void AffineComponent::UpdateSimple(const CuMatrixBase<BaseFloat> &in_value,
                                   const CuMatrixBase<BaseFloat> &out_deriv) {
  ComputeGradient(in_value, out_deriv);
  cudaF_vec_axpy(Str(), bias_params_.Data(), b_grad_.data, bias_params_.Dim(), learning_rate_);
  CuSubMatrix<BaseFloat> g(w_grad_.data, w_grad_.rows, w_grad_.cols, w_grad_.stride);
  linear_params_.AddMat(learning_rate_, g, kNoTrans);
}

// reference :1237-1258
void AffineComponent::Backprop(const ChunkInfo &, const ChunkInfo &,
                               const CuMatrixBase<BaseFloat> &in_value,
                               const CuMatrixBase<BaseFloat> &,
                               const CuMatrixBase<BaseFloat> &out_deriv, Component *to_update_in,
                               CuMatrix<BaseFloat> *in_deriv) const {
  AffineComponent *to_update = dynamic_cast<AffineComponent *>(to_update_in);
  in_deriv->Resize(out_deriv.NumRows(), InputDim(), kUndefined);
  KALDI_ASSERT(out_deriv.NumCols() == OutputDim());
  // Propagate the derivative back to the input: in_deriv = out_deriv W (beta = 0).
  cudaF_affine_dgrad(Str(), CuDevice::Instantiate().MathMode(), out_deriv.Data(), out_deriv.Dim(),
                     linear_params_.Data(), linear_params_.Dim(), in_deriv->Data(), in_deriv->Dim());
  CU_SAFE_CALL(cudaGetLastError());
  if (to_update != NULL) {
    // must come second, in case this == to_update_in
    if (to_update->is_gradient_) to_update->UpdateSimple(in_value, out_deriv);
    else to_update->Update(in_value, out_deriv);
  }
}

void AffineComponent::Read(std::istream &is, bool binary) {
  std::ostringstream ostr_beg, ostr_end;
  ostr_beg << "<" << Type() << ">";
  ostr_end << "</" << Type() << ">";
  ExpectOneOrTwoTokens(is, binary, ostr_beg.str(), "<LearningRate>");
  ReadBasicType(is, binary, &learning_rate_);
  ExpectToken(is, binary, "<LinearParams>");
  linear_params_.Read(is, binary);
  ExpectToken(is, binary, "<BiasParams>");
  bias_params_.Read(is, binary);
  std::string tok;
  ReadToken(is, binary, &tok);
  if (tok == "<AvgInput>") {   // back-compatibility: discard
    CuVector<BaseFloat> avg_input;
    avg_input.Read(is, binary);
    BaseFloat avg_input_count;
    ExpectToken(is, binary, "<AvgInputCount>");
    ReadBasicType(is, binary, &avg_input_count);
    ReadToken(is, binary, &tok);
  }
  if (tok == "<IsGradient>") {
    ReadBasicType(is, binary, &is_gradient_);
    ExpectToken(is, binary, ostr_end.str());
  } else {
    is_gradient_ = false;
    KALDI_ASSERT(tok == ostr_end.str());
  }
}

void AffineComponent::Write(std::ostream &os, bool binary) const {
  std::ostringstream ostr_beg, ostr_end;
  ostr_beg << "<" << Type() << ">";
  ostr_end << "</" << Type() << ">";
  WriteToken(os, binary, ostr_beg.str());
  WriteToken(os, binary, "<LearningRate>");
  WriteBasicType(os, binary, learning_rate_);
  WriteToken(os, binary, "<LinearParams>");
  linear_params_.Write(os, binary);
  WriteToken(os, binary, "<BiasParams>");
  bias_params_.Write(os, binary);
  WriteToken(os, binary, "<IsGradient>");
  WriteBasicType(os, binary, is_gradient_);
  WriteToken(os, binary, ostr_end.str());
}

void AffineComponent::Scale(BaseFloat scale) {
  linear_params_.Scale(scale);
  bias_params_.Scale(scale);
}

void AffineComponent::Add(BaseFloat alpha, const UpdatableComponent &other_in) {
  const AffineComponent *other = dynamic_cast<const AffineComponent *>(&other_in);
  KALDI_ASSERT(other != NULL);
  linear_params_.AddMat(alpha, other->linear_params_);
  bias_params_.AddVec(alpha, other->bias_params_);
}

int32 AffineComponent::GetParameterDim() const { return (InputDim() + 1) * OutputDim(); }

void AffineComponent::Vectorize(VectorBase<BaseFloat> *params) const {
  KALDI_ASSERT(params->Dim() == GetParameterDim());
  Matrix<BaseFloat> w(linear_params_.NumRows(), linear_params_.NumCols());
  linear_params_.CopyToMat(&w);
  Vector<BaseFloat> b(bias_params_.Dim());
  bias_params_.CopyToVec(&b);
  int32 k = 0;
  for (int32 r = 0; r < w.NumRows(); r++)
    for (int32 c = 0; c < w.NumCols(); c++) (*params)(k++) = w(r, c);
  for (int32 i = 0; i < b.Dim(); i++) (*params)(k++) = b(i);
}

void AffineComponent::UnVectorize(const VectorBase<BaseFloat> &params) {
  KALDI_ASSERT(params.Dim() == GetParameterDim());
  Matrix<BaseFloat> w(linear_params_.NumRows(), linear_params_.NumCols());
  Vector<BaseFloat> b(bias_params_.Dim());
  int32 k = 0;
  for (int32 r = 0; r < w.NumRows(); r++)
    for (int32 c = 0; c < w.NumCols(); c++) w(r, c) = params(k++);
  for (int32 i = 0; i < b.Dim(); i++) b(i) = params(k++);
  linear_params_.CopyFromMat(w);
  bias_params_.CopyFromVec(b);
}

}  // namespace nnet2
}  // namespace kaldi
