// Log-mel front-end of the speaker-conditioning prompt on the GPU: MelSpectrogramFeatures.forward
// (indextts/utils/feature_extractors.py:24-50, padding="center") -- torchaudio MelSpectrogram(sample_rate 24000,
// n_fft 1024, hop 256, hann periodic window, power 1, centre reflect padding, HTK mel scale, norm None, 100 mels)
// followed by safe_log = log(clip(x, 1e-7)) (utils/common.py:110-121).  One block per (frame, batch item):
// reflect-gathered, windowed frame -> 1024-point radix-2 FFT in shared memory -> magnitude of the 513 one-sided
// bins -> triangular mel filterbank -> log.  fp32; the twiddles come from sincospif (1 ulp), the filterbank is
// built on the host in double and cached on the device.  SURVEY.md section 8(f) row 4 (first half).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/b200vgan.h"
#include "bvg_common.cuh"

namespace {

constexpr int NFFT = 1024, LOGN = 10, NBINS = NFFT / 2 + 1, MAXMEL = 128;

__global__ void __launch_bounds__(256) log_mel_kernel(const float* __restrict__ audio, int N, int hop, int frames,
                                                      const float* __restrict__ fb /* [NBINS][n_mels] */, int n_mels,
                                                      float clip, float* __restrict__ out, int transposed) {
  __shared__ float re[NFFT], im[NFFT];
  __shared__ float mag[NBINS + 3];
  const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* a = audio + (size_t)b * N;
  // windowed frame, stored in bit-reversed order (decimation in time)
  for (int i = tid; i < NFFT; i += 256) {
    int t = f * hop + i - NFFT / 2;            // centre padding: frame f starts NFFT/2 before sample f*hop
    if (t < 0) t = -t;                         // reflect (torch.stft pad_mode="reflect")
    if (t >= N) t = 2 * (N - 1) - t;
    t = min(max(t, 0), N - 1);
    float s, c;
    sincospif(2.f * (float)i / (float)NFFT, &s, &c);
    const float w = 0.5f - 0.5f * c;           // periodic Hann window (torch.hann_window default)
    const int r = (int)(__brev((unsigned)i) >> (32 - LOGN));
    re[r] = a[t] * w;
    im[r] = 0.f;
  }
  __syncthreads();
#pragma unroll 1
  for (int st = 0; st < LOGN; ++st) {
    const int half = 1 << st;
    for (int k = tid; k < NFFT / 2; k += 256) {
      const int j = k & (half - 1), i0 = ((k >> st) << (st + 1)) + j, i1 = i0 + half;
      float s, c;
      sincospif(-(float)j / (float)half, &s, &c);   // exp(-2 pi i j / (2 half))
      const float xr = re[i1], xi = im[i1];
      const float tr = xr * c - xi * s, ti = xr * s + xi * c;
      const float ur = re[i0], ui = im[i0];
      re[i0] = ur + tr; im[i0] = ui + ti;
      re[i1] = ur - tr; im[i1] = ui - ti;
    }
    __syncthreads();
  }
  for (int k = tid; k < NBINS; k += 256) mag[k] = sqrtf(re[k] * re[k] + im[k] * im[k]);   // power = 1
  __syncthreads();
  for (int m = tid; m < n_mels; m += 256) {
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < NBINS; ++k) acc = fmaf(mag[k], fb[k * n_mels + m], acc);
    const float v = logf(fmaxf(acc, clip));
    if (transposed) out[((size_t)b * frames + f) * n_mels + m] = v;
    else out[((size_t)b * n_mels + m) * frames + f] = v;
  }
}

// torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm=None, mel_scale="htk")
std::vector<float> make_fbank(int n_mels, double sample_rate, double f_min, double f_max) {
  auto hz2mel = [](double f) { return 2595.0 * std::log10(1.0 + f / 700.0); };
  auto mel2hz = [](double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); };
  std::vector<double> fpts(n_mels + 2);
  const double m0 = hz2mel(f_min), m1 = hz2mel(f_max);
  for (int i = 0; i < n_mels + 2; ++i) fpts[i] = mel2hz(m0 + (m1 - m0) * i / (n_mels + 1));
  std::vector<float> fb((size_t)NBINS * n_mels);
  for (int k = 0; k < NBINS; ++k) {
    const double fr = (double)((int)sample_rate / 2) * k / (NBINS - 1);   // torch.linspace(0, sample_rate // 2, n_freqs)
    for (int m = 0; m < n_mels; ++m) {
      const double down = (fr - fpts[m]) / (fpts[m + 1] - fpts[m]), up = (fpts[m + 2] - fr) / (fpts[m + 2] - fpts[m + 1]);
      fb[(size_t)k * n_mels + m] = (float)std::max(0.0, std::min(down, up));
    }
  }
  return fb;
}

struct FbKey {
  int dev, n_mels, sr, fmin, fmax;
  bool operator<(const FbKey& o) const {
    return std::tie(dev, n_mels, sr, fmin, fmax) < std::tie(o.dev, o.n_mels, o.sr, o.fmin, o.fmax);
  }
};
std::mutex g_mu;
std::map<FbKey, float*> g_fb;

}  // namespace

extern "C" int bvg_mel_frames(int32_t n_samples, int32_t hop) { return (n_samples < 1 || hop < 1) ? 0 : 1 + n_samples / hop; }

extern "C" int bvg_set_error(const char* msg);   // bvg_api.cu

extern "C" int bvg_log_mel(const float* audio, int32_t B, int32_t n_samples, int32_t sample_rate, int32_t hop, int32_t n_mels,
                           float f_min, float f_max, float* mel, int32_t transposed, void* stream) {
  if (bvg_device_check()) return 1;
  if (!audio || !mel) return bvg_set_error("bvg_log_mel: null argument");
  if (B < 1 || hop < 1 || n_mels < 1 || n_mels > MAXMEL || sample_rate < 2) return bvg_set_error("bvg_log_mel: bad argument");
  if (n_samples <= NFFT / 2) return bvg_set_error("bvg_log_mel: reflect padding needs more than n_fft/2 = 512 samples");
  if (f_max <= 0.f) f_max = 0.5f * (float)sample_rate;   // torchaudio: f_max=None -> sample_rate // 2
  int dev = 0;
  cudaGetDevice(&dev);
  float* fb = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    const FbKey key{dev, n_mels, sample_rate, (int)lrintf(f_min * 16.f), (int)lrintf(f_max * 16.f)};
    auto it = g_fb.find(key);
    if (it == g_fb.end()) {
      std::vector<float> h = make_fbank(n_mels, sample_rate, f_min, f_max);
      if (cudaMalloc((void**)&fb, h.size() * sizeof(float)) != cudaSuccess) return bvg_set_error("bvg_log_mel: filterbank allocation failed");
      if (cudaMemcpy(fb, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return bvg_set_error("bvg_log_mel: filterbank upload failed");
      g_fb[key] = fb;
    } else {
      fb = it->second;
    }
  }
  const int frames = 1 + n_samples / hop;
  log_mel_kernel<<<dim3(frames, B), 256, 0, (cudaStream_t)stream>>>(audio, n_samples, hop, frames, fb, n_mels, 1e-7f, mel, transposed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return bvg_set_error(cudaGetErrorString(e));
  return 0;
}
