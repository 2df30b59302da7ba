// Stand-alone anti-aliased Snake (SnakeAlias) — fp32, one pass over [B,C,L].
//
// Replaces SnakeAlias.forward (vits_decoder/alias/act.py:124-128):
//   UpSample1d   (alias/resample.py:25-33): replicate-pad 5/5, depthwise ConvTranspose1d with the
//                12-tap Kaiser-sinc, stride 2, times 2, crop 15/15
//   SnakeBeta    (alias/act.py:79-92, log-scale): v = u + sin^2(u*e^alpha) / (e^beta + 1e-9)
//   DownSample1d (alias/filter.py:86-94): replicate-pad 5/6, depthwise 12-tap conv, stride 2
// in closed form:
//   u[2a]   = 2 * sum_{d=-3..2} x[clamp(a+d)] * fu[5-2d]
//   u[2a+1] = 2 * sum_{d=-2..3} x[clamp(a+d)] * fu[6-2d]
//   out[n]  = sum_{k=0..11} v[clamp(2n+k-5, 0, 2L-1)] * fd[k]
// The 2L-long intermediate lives only in shared memory.
#include "common.cuh"

namespace svcb {

constexpr int SA_TL = 1024;  // outputs per CTA

__global__ void __launch_bounds__(256)
snake_alias_kernel(const float* __restrict__ x, float* __restrict__ y,
                   const float* __restrict__ ea, const float* __restrict__ inv_b,
                   const float* __restrict__ fu, const float* __restrict__ fd, int C, int L) {
  __shared__ float xs[SA_TL + 12];
  __shared__ float vs[2 * SA_TL + 12];
  __shared__ float f_up[12], f_dn[12];
  const int c = blockIdx.y, b = blockIdx.z;
  const int n0 = blockIdx.x * SA_TL;
  const float* row = x + ((long long)b * C + c) * L;
  float* orow = y + ((long long)b * C + c) * L;
  const int tid = threadIdx.x;
  if (tid < 12) { f_up[tid] = __ldg(fu + tid); f_dn[tid] = __ldg(fd + tid); }
  for (int i = tid; i < SA_TL + 12; i += blockDim.x) {
    int g = n0 - 6 + i;
    g = min(max(g, 0), L - 1);
    xs[i] = __ldg(row + g);
  }
  __syncthreads();
  const float a_ = __ldg(ea + c), ib = __ldg(inv_b + c);
  const int nv = 2 * SA_TL + 10;
  for (int idx = tid; idx < nv; idx += blockDim.x) {
    int m = 2 * n0 - 5 + idx;
    m = min(max(m, 0), 2 * L - 1);
    const int a = m >> 1;
    const float* xp = xs + (a - (n0 - 6));
    float acc = 0.f;
    if ((m & 1) == 0) {
#pragma unroll
      for (int d = -3; d <= 2; ++d) acc = fmaf(xp[d], f_up[5 - 2 * d], acc);
    } else {
#pragma unroll
      for (int d = -2; d <= 3; ++d) acc = fmaf(xp[d], f_up[6 - 2 * d], acc);
    }
    const float u = 2.f * acc;
    const float sn = sinf(u * a_);
    vs[idx] = u + ib * (sn * sn);
  }
  __syncthreads();
  for (int i = tid; i < SA_TL; i += blockDim.x) {
    const int n = n0 + i;
    if (n >= L) break;
    const float* vp = vs + 2 * i;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc = fmaf(vp[k], f_dn[k], acc);
    orow[n] = acc;
  }
}

int launch_snake_alias(const float* x, float* y, const float* ea, const float* inv_b,
                       const float* fu, const float* fd, int B, int C, int L, cudaStream_t s) {
  if (B <= 0 || C <= 0 || L <= 0) return SVCB_OK;
  if (C > 65535 || B > 65535) { set_error("snake_alias: C or B exceeds grid limits"); return SVCB_E_BAD_SHAPE; }
  dim3 grid((L + SA_TL - 1) / SA_TL, C, B);
  KernelScope ks("snake_alias", s, 70.0 * B * C * (double)L, 8.0 * B * C * (double)L);
  snake_alias_kernel<<<grid, 256, 0, s>>>(x, y, ea, inv_b, fu, fd, C, L);
  SVCB_LAUNCH_CHECK("snake_alias");
  return SVCB_OK;
}

}  // namespace svcb
