// NSF harmonic source: F0 -> excitation waveform, and the int16 export.
//
// Replaces Generator.pitch2source (vits_decoder/generator.py:160-165) -> SourceModuleHnNSF.forward
// (vits_decoder/nsf.py:383-394) -> SineGen.forward/_f02sine/_f02uv (nsf.py:217-316).
//
// The reference computes, per (item, harmonic) over the whole utterance,
//   rad[n]  = (f0[n]*(h+1)/sr) % 1           (+ rand_ini at n = 0)
//   c1      = cumsum(rad)                     (torch CPU accumulates fp32 cumsum in double)
//   over[n] = (c1[n] % 1) - (c1[n-1] % 1) < 0 (after the fp32 cast of c1)
//   c2      = cumsum(rad - over)              (again double accumulate, fp32 result)
//   sine    = sin(c2 * 2 * pi) * 0.1
// F0 is nearest-upsampled x hop, so rad is constant inside a frame and both running sums have a
// closed form per frame: c1 = base1[f] + (i+1)*rad, #wraps so far = floor(fp32(c1)) - floor at the
// frame start, c2 = c1 + D[f] + wraps_in_frame * dd_f, where dd_f = fp32(rad - 1) - rad carries the
// fp32 rounding of the reference's `rad_values + cumsum_shift`.  A T-step sequential scan over
// frames (one thread per item x harmonic) produces base1 / D / floor-base; everything else is
// embarrassingly parallel over samples.  Valid while f0 * n_harm < sampling_rate (rad < 1 per step),
// which holds for the reference's 50..1100 Hz pitch range.
#include "common.cuh"

namespace svcb {

__device__ __forceinline__ float rad_of(float f0, int h, float sr) {
  const float fh = f0 * (float)(h + 1);
  return fmodf(fh / sr, 1.0f);
}

// scan_ws layout: [3][B][NH][T] doubles: base1, D, floor-base
__global__ void source_scan_kernel(const float* __restrict__ f0, const float* __restrict__ rand_ini,
                                   double* __restrict__ ws, int B, int T, int hop, int NH, float sr) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * NH) return;
  const int b = idx / NH, h = idx % NH;
  const long long plane = (long long)B * NH * T;
  double* base1 = ws + (long long)idx * T;
  double* Dd = base1 + plane;
  double* FB = Dd + plane;
  const float* f0b = f0 + (long long)b * T;
  const float ri = h == 0 ? 0.f : rand_ini[(long long)b * NH + h];
  float rad = rad_of(f0b[0], h, sr);
  const float rad0i = rad + ri;                 // rad_values[:, 0, :] += rand_ini (fp32 add)
  double b1 = (double)rad0i - (double)rad;      // so that base1 + 1*rad == c1[0]
  double D = 0.0;
  float fb = floorf(rad0i);                     // floor(fp32(c1[0]))
  for (int f = 0; f < T; ++f) {
    base1[f] = b1;
    Dd[f] = D;
    FB[f] = (double)fb;
    const double c1_end = b1 + (double)hop * (double)rad;
    const float fe = floorf((float)c1_end);
    const double dd = (double)(rad - 1.0f) - (double)rad;
    D += (double)(fe - fb) * dd;
    fb = fe;
    b1 = c1_end;
    if (f + 1 < T) rad = rad_of(f0b[f + 1], h, sr);
  }
}

__global__ void __launch_bounds__(256)
source_sample_kernel(const float* __restrict__ f0, const float* __restrict__ noise,
                     const float* __restrict__ merge_w, const float* __restrict__ merge_b,
                     const double* __restrict__ ws, float* __restrict__ out, int B, int T, int hop,
                     int NH, float sr) {
  const long long L = (long long)T * hop;
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= L) return;
  const int f = (int)(n / hop), i = (int)(n % hop);
  const float f0v = f0[(long long)b * T + f];
  const float uv = f0v > 0.f ? 1.f : 0.f;
  const float namp = uv * 0.003f + ((1.f - uv) * 0.1f) / 3.f;
  const long long plane = (long long)B * NH * T;
  const float* nz = noise + ((long long)b * L + n) * NH;
  float dot = 0.f;
  for (int h = 0; h < NH; ++h) {
    const long long o = ((long long)b * NH + h) * T + f;
    const float rad = rad_of(f0v, h, sr);
    const double c1 = ws[o] + (double)(i + 1) * (double)rad;
    const float wf = floorf((float)c1) - (float)ws[2 * plane + o];
    const double dd = (double)(rad - 1.0f) - (double)rad;
    const double c2 = c1 + ws[plane + o] + (double)wf * dd;
    const float ph = ((float)c2 * 2.f) * 3.14159274101257324f;
    const float sw = sinf(ph) * 0.1f;
    const float val = sw * uv + namp * nz[h];
    dot = fmaf(val, __ldg(merge_w + h), dot);
  }
  out[(long long)b * L + n] = tanhf(dot + __ldg(merge_b));
}

size_t source_scan_ws_bytes(int B, int T, int n_harm) {
  return (size_t)3 * B * n_harm * T * sizeof(double);
}

int launch_source(const float* f0, const float* rand_ini, const float* noise, const float* merge_w,
                  const float* merge_b, float* source, double* scan_ws, int B, int T, int hop,
                  int n_harm, float sampling_rate, cudaStream_t s) {
  if (B <= 0 || T <= 0) return SVCB_OK;
  const int nth = B * n_harm;
  {
    KernelScope ks("source_scan", s, 0.0, 24.0 * nth * T);
    source_scan_kernel<<<(nth + 63) / 64, 64, 0, s>>>(f0, rand_ini, scan_ws, B, T, hop, n_harm,
                                                      sampling_rate);
    SVCB_LAUNCH_CHECK("source_scan");
  }
  const long long L = (long long)T * hop;
  dim3 grid((unsigned)((L + 255) / 256), B);
  KernelScope ks("source_sample", s, 0.0, 4.0 * B * (double)L * (n_harm + 1));
  source_sample_kernel<<<grid, 256, 0, s>>>(f0, noise, merge_w, merge_b, scan_ws, source, B, T, hop,
                                            n_harm, sampling_rate);
  SVCB_LAUNCH_CHECK("source_sample");
  return SVCB_OK;
}

// Generator.source2wav (generator.py:167-173): *32768, clamp to [-32768, 32767], truncate.
__global__ void source2wav_kernel(const float* __restrict__ src, int16_t* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 32768.0f * src[i];
  v = fminf(fmaxf(v, -32768.0f), 32767.0f);
  out[i] = (int16_t)v;
}

int launch_source2wav(const float* src, int16_t* out, size_t n, cudaStream_t s) {
  if (n == 0) return SVCB_OK;
  KernelScope ks("source2wav", s, 0.0, 6.0 * (double)n);
  source2wav_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, out, n);
  SVCB_LAUNCH_CHECK("source2wav");
  return SVCB_OK;
}

}  // namespace svcb
