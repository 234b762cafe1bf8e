// ECAPA-TDNN speaker encoder (mel [B',Tm,n_mels] -> embedding [B',1,512]) on CUDA cores, fp32.
// Reference: indextts/BigVGAN/ECAPA_TDNN.py:429-581 (architecture defaults :470-481), nnet/CNN.py:305-545
// (SpeechBrain Conv1d, "same" reflect padding), nnet/normalization.py:13-108 (BatchNorm1d, eval).
// fp32 FFMA throughout (7e-6 against the reference's embedding).  The reference call site bigvgan(latent, mel_ref)
// runs it on every call, so it sits on the end-to-end path: register-tiled implicit-GEMM convolutions, warp-per-output
// GEMVs for the length-1 layers, ~0.4 ms for a 511-frame prompt (round 1: 7 ms).
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <string>
#include <vector>

struct EcapaTdnn {   // conv -> ReLU -> BatchNorm(eval)   (ECAPA_TDNN.py:126-128)
  int cin = 0, cout = 0, k = 1, d = 1;
  float *w = nullptr, *b = nullptr, *bn_w = nullptr, *bn_b = nullptr, *bn_m = nullptr, *bn_v = nullptr;
  float *scale = nullptr, *shift = nullptr;   // folded BN: y = relu(conv) * scale + shift
};
struct EcapaLin {    // 1x1 conv with bias
  int cin = 0, cout = 0;
  float *w = nullptr, *b = nullptr;
};
struct EcapaBlock {  // SERes2NetBlock (ECAPA_TDNN.py:341-426)
  int d = 1;
  EcapaTdnn tdnn1, res[7], tdnn2;
  EcapaLin se1, se2;
};
struct EcapaModel {
  int n_mels = 100, C = 512, lin = 512, att = 128, se = 128, scale = 8;
  EcapaTdnn b0;
  EcapaBlock blk[3];
  EcapaTdnn mfa, asp_tdnn;
  EcapaLin asp_conv, fc;
  float *abn_w = nullptr, *abn_b = nullptr, *abn_m = nullptr, *abn_v = nullptr, *abn_scale = nullptr, *abn_shift = nullptr;
};

// Registers every parameter under its reference state-dict name (without the "speaker_encoder." prefix).
using EcapaRegisterFn = std::function<void(const std::string&, float**, std::vector<long long>)>;
void ecapa_register(EcapaModel& m, int n_mels, int lin_neurons, const EcapaRegisterFn& reg);
// All TDNN layers / the pooling BatchNorm, for folding (scale/shift are allocated by the caller).
void ecapa_collect_bn(EcapaModel& m, std::vector<EcapaTdnn*>& tdnns);
cudaError_t ecapa_fold_bn(const float* w, const float* b, const float* mean, const float* var, float* scale, float* shift,
                          int C, cudaStream_t s);
size_t ecapa_workspace_bytes(const EcapaModel& m, int B, int T);
// mel [B,T,n_mels] fp32 (device), rel_lens [B] fp32 (device) or nullptr, emb [B,lin] fp32 (device).
cudaError_t ecapa_forward(const EcapaModel& m, const float* mel, int B, int T, const float* rel_lens, float* emb,
                          void* workspace, cudaStream_t s, int* launches);
