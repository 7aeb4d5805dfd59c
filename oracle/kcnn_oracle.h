/*
 * oracle/kcnn_oracle.h -- prototypes of the CPU oracle (TEST INFRASTRUCTURE;
 * see kcnn_oracle_impl.h for the reference file:line each function follows).
 * The float set is listed; every oraF_<name> has an oraD_<name> twin taking
 * double.
 */
#ifndef KCNN_ORACLE_H_
#define KCNN_ORACLE_H_
#ifdef __cplusplus
extern "C" {
#endif

void oraF_gemm(int transA, int transB, int m, int n, int k, float alpha,
               const float *A, int lda, const float *B, int ldb, float beta,
               float *C, int ldc);
int oraF_conv2d(const float *x, int n_rows, int x_stride, const float *kern,
                int kern_stride, int H, int W, int C, int KH, int KW, int G,
                float *out, int out_stride, int concat);
void oraF_add_mat_rep_vec(float *m, int rows, int cols, int stride,
                          const float *vec, int rep);
void oraF_flip_mat(const float *in, int in_stride, int KH, int KW, int C, int G,
                   float *flip, int flip_stride);
void oraF_pad_zero(const float *in, int rows, int in_stride, int H, int W, int C,
                   int KH, int KW, float *pad, int pad_stride);
void oraF_tp_block(const float *in, int rows, int in_stride, int C, int bs,
                   float *out, int out_stride);
void oraF_tp_inside_block(const float *in, int rows, int in_stride, int G, int bs,
                          float *out, int out_stride);
void oraF_mod_permute_row(const float *in, int rows, int cols, int in_stride,
                          int C, int bs, float *out, int out_stride);
void oraF_maxpool_prop(const float *in, int rows, int in_stride, int H, int W,
                       int ph, int pw, int pc, int mode, float *out, int out_cols,
                       int out_stride);
void oraF_maxpool_backprop(const float *in, int rows, int in_stride,
                           const float *out_val, int ov_stride,
                           const float *out_deriv, int od_stride, int out_cols,
                           float *in_deriv, int id_stride, int H, int W, int ph,
                           int pw, int pc, int mode);
int oraF_conv_propagate(const float *in, int N, int in_stride, const float *lin,
                        int lin_stride, const float *bias, int H, int W, int C,
                        int pad_h, int pad_w, int KH, int KW, int G, float *out,
                        int out_stride);
int oraF_conv_backprop_uses_flip(int pad_h, int pad_w, int KH, int KW, int OH, int OW);
int oraF_conv_backprop(const float *out_deriv, int N, int od_stride,
                       const float *lin, int lin_stride, int H, int W, int C,
                       int pad_h, int pad_w, int KH, int KW, int G, int branch,
                       float *in_deriv, int id_stride);
int oraF_conv_update(const float *in_value, int N, int iv_stride,
                     const float *out_deriv, int od_stride, float *lin,
                     int lin_stride, float *bias, float *prev, int prev_stride,
                     int H, int W, int C, int pad_h, int pad_w, int KH, int KW,
                     int G, float learning_rate, float weight_decay, float momentum,
                     int apply, float *grad_out, float *bias_grad_out);
void oraF_fc_propagate(const float *in, int N, int in_stride, const float *W,
                       int w_stride, const float *bias, int in_dim, int out_dim,
                       float *out, int out_stride);
void oraF_fc_backprop(const float *out_deriv, int N, int od_stride, const float *W,
                      int w_stride, int in_dim, int out_dim, float *in_deriv,
                      int id_stride);
void oraF_fc_update(const float *in_value, int N, int iv_stride,
                    const float *out_deriv, int od_stride, float *W, int w_stride,
                    float *bias, float *prev, int prev_stride, int in_dim,
                    int out_dim, float learning_rate, float weight_decay,
                    float momentum);

void oraF_relu_propagate(const float *in, int rows, int cols, int in_stride,
                         float *out, int out_stride);
void oraF_relu_backprop(const float *out_value, int rows, int cols, int ov_stride,
                        const float *out_deriv, int od_stride, float *in_deriv,
                        int id_stride);
void oraF_softmax_propagate(const float *in, int rows, int cols, int in_stride,
                            float *out, int out_stride);
double oraF_xent_objf_and_deriv(const float *post, int rows, int cols, int p_stride,
                                const int *labels, float *deriv, int d_stride);
void oraF_softmax_backprop(const float *out_value, int rows, int cols, int ov_stride,
                           const float *out_deriv, int od_stride, float *in_deriv,
                           int id_stride);
void oraF_dropout_propagate(const float *in, int rows, int cols, int in_stride, const float *uniform,
                            int u_stride, float dp, float low_scale, float *out, int out_stride);
void oraF_dropout_backprop(const float *in_value, int rows, int cols, int iv_stride,
                           const float *out_value, int ov_stride, const float *out_deriv, int od_stride,
                           float *in_deriv, int id_stride);
void oraF_normalize_propagate(const float *in, int rows, int cols, int in_stride, float *out,
                              int out_stride);
void oraF_normalize_backprop(const float *in_value, int rows, int cols, int iv_stride,
                             const float *out_deriv, int od_stride, float *in_deriv, int id_stride);
void oraF_nonlin_update_stats(const float *out_value, int rows, int cols, int ov_stride, const float *deriv,
                              int d_stride, double *value_sum, double *deriv_sum, double *count);

#ifdef __cplusplus
}
#endif
#endif
