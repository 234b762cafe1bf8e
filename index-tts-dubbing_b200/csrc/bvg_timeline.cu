// Timeline merge of the dubbing tool on the GPU: the step right after the vocoder in the reference's SRT pipeline,
// AudioProcessor._time_synchronized_merge (srt_dubbing/src/audio_processor.py:133-230) --
//     merged[start_i : start_i + n_i] += audio_i   for the segments in start-time order,   then (optionally)
//     merged /= max|merged|  when the peak exceeds AUDIO.MAX_AMPLITUDE (config.py:21).
// The placement arithmetic (sorting, the previous-segment overlap rule, array growth) is host logic
// (b200vgan/timeline.py); this file sums and normalises.  One thread per OUTPUT sample gathers the segments that
// cover it in placement order, so overlapping segments are added in exactly the reference's order (bit-exact fp32,
// no atomics), and the decoded segments never leave the device between the vocoder and the finished timeline.
#include "../../include/b200vgan.h"
#include "bvg_common.cuh"

extern "C" int bvg_set_error(const char* msg);

namespace {

constexpr int MAXCOVER = 16;   // segments that may overlap one sample (more: reported as an error by the host code)

// seg_dst sorted ascending; seg_rank[i] = position of segment i in the order the reference adds in (it can differ from
// the start order when the overlap rule pushed a segment past its successor); seg_* are device arrays of nseg
__global__ void timeline_merge_kernel(const float* __restrict__ flat, const long long* __restrict__ seg_src,
                                      const long long* __restrict__ seg_dst, const int* __restrict__ seg_n,
                                      const int* __restrict__ seg_rank, int nseg,
                                      long long max_n, float* __restrict__ out, long long total, unsigned* __restrict__ peak_bits) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (t < total) {
    // last segment with dst <= t
    int lo = 0, hi = nseg;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg_dst[mid] <= t) lo = mid + 1; else hi = mid; }
    int cover[MAXCOVER], rank[MAXCOVER], nc = 0;
    for (int i = lo - 1; i >= 0 && seg_dst[i] > t - max_n; --i)
      if (t < seg_dst[i] + seg_n[i] && nc < MAXCOVER) {   // insertion sort by the reference's adding order
        const int r = seg_rank[i];
        int k = nc++;
        while (k > 0 && rank[k - 1] > r) { rank[k] = rank[k - 1]; cover[k] = cover[k - 1]; --k; }
        rank[k] = r; cover[k] = i;
      }
    for (int k = 0; k < nc; ++k) {
      const int i = cover[k];
      acc = __fadd_rn(acc, flat[seg_src[i] + (t - seg_dst[i])]);
    }
    out[t] = acc;
  }
  if (peak_bits) {   // max |x| over the block, then one atomic: non-negative floats order like their bit patterns
    float m = fabsf(acc);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float wm[8];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, wm[w]);
      atomicMax(peak_bits, __float_as_uint(m));
    }
  }
}

__global__ void timeline_normalize_kernel(float* __restrict__ out, long long total, const unsigned* __restrict__ peak_bits, float max_amp) {
  const float peak = __uint_as_float(*peak_bits);
  if (!(peak > max_amp)) return;   // audio_processor.py:229
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) out[t] = __fdiv_rn(out[t], peak);
}

}  // namespace

extern "C" int bvg_timeline_merge(const float* flat, const int64_t* seg_src, const int64_t* seg_dst, const int32_t* seg_n,
                                  const int32_t* seg_rank, int32_t nseg, int64_t max_n, float* out, int64_t total, int32_t normalize, float max_amplitude,
                                  uint32_t* peak_scratch, void* stream) {
  if (bvg_device_check()) return 1;
  if (!out || total < 0 || nseg < 0 || (nseg > 0 && (!flat || !seg_src || !seg_dst || !seg_n || !seg_rank)) || (normalize && !peak_scratch))
    return bvg_set_error("bvg_timeline_merge: bad argument");
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (normalize && cudaMemsetAsync(peak_scratch, 0, sizeof(uint32_t), s) != cudaSuccess) return bvg_set_error("bvg_timeline_merge: memset failed");
  const unsigned blocks = (unsigned)((total + 255) / 256);
  timeline_merge_kernel<<<blocks, 256, 0, s>>>(flat, (const long long*)seg_src, (const long long*)seg_dst, seg_n, seg_rank, nseg, (long long)max_n, out,
                                              (long long)total, normalize ? peak_scratch : nullptr);
  if (normalize) timeline_normalize_kernel<<<blocks, 256, 0, s>>>(out, (long long)total, peak_scratch, max_amplitude);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return bvg_set_error(cudaGetErrorString(e));
  return 0;
}
