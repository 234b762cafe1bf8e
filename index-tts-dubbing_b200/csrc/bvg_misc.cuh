#pragma once
#include "bvg_common.cuh"

cudaError_t launch_pack_latent(const void* x, int in_dtype, void* y, int out_dtype, const SegDesc* seg,
                               const int* src_row, int B, int Tmax, int C, int R, cudaStream_t s);
cudaError_t launch_nct_to_c8(const float* x, void* y, int out_dtype, const SegDesc* seg, int B, int C, int T, int R,
                             cudaStream_t s);
cudaError_t launch_c8_to_nct(const void* x, int in_dtype, float* y, const SegDesc* seg, int B, int C, int T, int R,
                             cudaStream_t s);
cudaError_t launch_cond_bias(const float* bias, const float* cw, const float* cb, const float* spk, float* out, int C,
                             int D, int B, int spkB, int out_bstride, cudaStream_t s);
cudaError_t launch_conv_post_tanh(const void* x, int dtype, const float* w, const float* bias, float* wav, short* pcm,
                                  const SegDesc* seg, const int* dst_row, int hop, int B, int C, int R, int Lmax,
                                  cudaStream_t s);
cudaError_t launch_repack_conv(const float* w, float* wp, int Cout, int Cin, int k, cudaStream_t s);
cudaError_t launch_repack_convt(const float* w, float* wp, int Cin, int Cout, int k, int u, cudaStream_t s);
cudaError_t launch_snake_params(const float* la, const float* lb, float* alpha, float* inv_beta, int n,
                                cudaStream_t s);
// 16-bit UMMA weight images (bvg_conv_umma.cu); dtype 1 = bf16, 2 = fp16
// small: the 64-column n-tile images of the small-batch variant (wide layers only)
size_t umma_weight_image_bytes(int ntaps, int Cin, int N, bool small = false, bool packed = false);
bool umma_k_packed_default(int Cin, int N);   // whether bvg_forward's image of such a layer is K-packed (BVG_CONV_PACKK, default 1)
// out_scale [N] / in_scale [Cin]: optional per-channel factors folded into the image (fp32, before rounding);
// res_diag [N]: optional diagonal of the residual identity image (layers whose residual goes through the tensor core)
cudaError_t launch_repack_umma(const float* wp_tap_major, void* img, int dtype, int ntaps, int Cin, int N, float acc_img_scale,
                               bool small, cudaStream_t s, const float* out_scale = nullptr, const float* in_scale = nullptr,
                               const float* res_diag = nullptr, bool packed = false);
// fp32 tensor-core mode helpers: weights [tap][Cin][N] -> [tap][3 Cin][N] = S [2^-11 W_hi; W_lo; W_hi] (bvg_conv_umma.cu);
// packed fp32 tensor [C/8][R][8] -> split bf16 tensor [2 C/8][R][8] = [hi chunks | lo chunks] (valid rows only)
cudaError_t launch_split3_weights(const float* w_tap_major, float* w3, int ntaps, int Cin, int N, float scale, cudaStream_t s);
float split3_weight_scale(float absmax);   // the layer's power-of-two scale S for that image
cudaError_t launch_absmax(const float* x, size_t n, float* out_zeroed, cudaStream_t s);
cudaError_t launch_split_c8(const float* x, void* y_split, const SegDesc* seg, int B, int C, int R, int max_len, cudaStream_t s);
cudaError_t launch_scale_vec(const float* x, const float* scale, float* y, int n, cudaStream_t s);   // y = x * scale
cudaError_t launch_zero_guards(void* buf, int esize, const SegDesc* seg, int B, int C, int R, cudaStream_t s);
// every packed buffer of a plan in one launch
struct GuardJob { void* buf; const SegDesc* seg; int chunks, R, vec_per_row; };
struct GuardJobs { GuardJob job[3 + 11 * 8]; int n; };
cudaError_t launch_zero_guards_all(const GuardJobs& jobs, int B, cudaStream_t s);
