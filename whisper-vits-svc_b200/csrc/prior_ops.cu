// Prior-encoder / flow operators that are not convolutions (fp32, CUDA cores).
#include <climits>

#include "common.cuh"

namespace svcb {

// ============================================================================ channel LayerNorm
// y = (v - mean) / sqrt(var + eps) * gamma + beta,  v = x (+ r), statistics over C for each (b,t).
// Replaces vits/modules.py:19-22 (LayerNorm on [B,C,T]) incl. the residual add of
// attentions.py:66,70, and SpeakerAdapter (vits_decoder/generator.py:36-47; per-batch gamma/beta).
__global__ void __launch_bounds__(256)
layernorm_c_kernel(const float* __restrict__ x, const float* __restrict__ r,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ y, int C, int T, int gb_stride, float eps) {
  extern __shared__ float sm[];
  float* vals = sm;                 // [C][32]
  float* red = sm + (size_t)C * 32; // [8][32]
  const int lane = threadIdx.x, w = threadIdx.y;
  const int b = blockIdx.y;
  const int t = blockIdx.x * 32 + lane;
  const bool ok = t < T;
  const long long base = (long long)b * C * T + t;
  float s = 0.f;
  for (int c = w; c < C; c += 8) {
    float v = 0.f;
    if (ok) {
      v = x[base + (long long)c * T];
      if (r) v += r[base + (long long)c * T];
    }
    vals[c * 32 + lane] = v;
    s += v;
  }
  red[w * 32 + lane] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i * 32 + lane];
  const float mean = tot / (float)C;
  __syncthreads();
  float q = 0.f;
  for (int c = w; c < C; c += 8) {
    const float d = vals[c * 32 + lane] - mean;
    q = fmaf(d, d, q);
  }
  red[w * 32 + lane] = q;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) var += red[i * 32 + lane];
  var /= (float)C;
  const float rstd = 1.f / sqrtf(var + eps);
  if (!ok) return;
  const float* g = gamma + (long long)b * gb_stride;
  const float* be = beta + (long long)b * gb_stride;
  for (int c = w; c < C; c += 8)
    y[base + (long long)c * T] = (vals[c * 32 + lane] - mean) * rstd * __ldg(g + c) + __ldg(be + c);
}

int launch_layernorm_c(const float* x, const float* r, const float* gamma, const float* beta,
                       float* y, int B, int C, int T, int gb_batch_stride, float eps,
                       cudaStream_t s) {
  const size_t smem = ((size_t)C * 32 + 8 * 32) * sizeof(float);
  if (smem > 200 * 1024) { set_error("layernorm_c: C too large"); return SVCB_E_UNSUPPORTED; }
  static DevSmemCache attr_cache;
  SVCB_CUDA_CHECK(ensure_dyn_smem(layernorm_c_kernel, 200 * 1024, attr_cache));
  dim3 grid((T + 31) / 32, B), block(32, 8);
  KernelScope ks("layernorm_c", s, 8.0 * B * C * (double)T, 4.0 * B * C * (double)T * (r ? 3 : 2));
  layernorm_c_kernel<<<grid, block, smem, s>>>(x, r, gamma, beta, y, C, T, gb_batch_stride, eps);
  SVCB_LAUNCH_CHECK("layernorm_c");
  return SVCB_OK;
}

// ============================================================================ relative attention
// Flash-style (scores never leave the SM) windowed relative-position self-attention.
// Replaces MultiHeadAttention.attention (vits/attentions.py:225-274) in its banded form:
//   s_ij = (q_i/sqrt(d)) . k_j + [|j-i|<=w] (q_i/sqrt(d)) . Ek[j-i+w];  masked_fill(-1e4) where
//   i or j >= len;  p = softmax_j(s);  o_i = sum_j p_ij v_j + sum_{|r|<=w} p_{i,i+r} Ev[r+w]
// which equals the reference's pad/reshape skew formulation (attentions.py:294-347) because the
// padded rows of the relative tables are zero.
constexpr int RA_BQ = 64, RA_BK = 64;

template <int D>
__global__ void __launch_bounds__(256)
rel_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ ek,
                     const float* __restrict__ ev, const long long* __restrict__ lengths,
                     float* __restrict__ out, int H, int heads, int window, int T) {
  constexpr int DP = D + 4;          // row stride of Q/K tiles: 16-byte rows, conflict-free LDS.128 over 8 rows
  constexpr int SP = RA_BK + 4;      // row stride of the score tile
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                      // [64][DP]  (also output staging)
  float* Ks = Qs + RA_BQ * DP;         // [64][DP]
  float* Vs = Ks + RA_BK * DP;         // [64][D]
  float* Ss = Vs + RA_BK * D;          // [64][SP]
  float* Ek = Ss + RA_BQ * SP;         // [2w+1][D]
  float* Ev = Ek + (2 * window + 1) * D;
  float* Rk = Ev + (2 * window + 1) * D;  // [64][2w+1]
  float* row_m = Rk + RA_BQ * (2 * window + 1);
  float* row_l = row_m + RA_BQ;
  float* row_a = row_l + RA_BQ;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, h = blockIdx.y;
  const int i0 = blockIdx.x * RA_BQ;
  const int nrel = 2 * window + 1;
  const long long len = lengths ? lengths[b] : (long long)T;
  const float* qb = qkv + ((long long)b * 3 * H + (long long)h * D) * T;
  const float* kb = qb + (long long)H * T;
  const float* vb = kb + (long long)H * T;
  const float qscale = rsqrtf((float)D);

  for (int idx = tid; idx < D * RA_BQ; idx += 256) {
    const int d = idx / RA_BQ, i = idx % RA_BQ;
    const int ig = i0 + i;
    Qs[i * DP + d] = ig < T ? qb[(long long)d * T + ig] * qscale : 0.f;
  }
  for (int idx = tid; idx < nrel * D; idx += 256) { Ek[idx] = __ldg(ek + idx); Ev[idx] = __ldg(ev + idx); }
  if (tid < RA_BQ) { row_m[tid] = -INFINITY; row_l[tid] = 0.f; }
  __syncthreads();
  for (int idx = tid; idx < RA_BQ * nrel; idx += 256) {
    const int i = idx / nrel, rr = idx % nrel;
    float a = 0.f;
    for (int d = 0; d < D; ++d) a = fmaf(Qs[i * DP + d], Ek[rr * D + d], a);
    Rk[idx] = a;
  }

  constexpr int ND = D / 16;
  float o[4][ND];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) o[ii][dd] = 0.f;

  for (int j0 = 0; j0 < T; j0 += RA_BK) {
    __syncthreads();
    for (int idx = tid; idx < D * RA_BK; idx += 256) {
      const int d = idx / RA_BK, j = idx % RA_BK;
      const int jg = j0 + j;
      const bool ok = jg < T;
      Ks[j * DP + d] = ok ? kb[(long long)d * T + jg] : 0.f;
      Vs[j * D + d] = ok ? vb[(long long)d * T + jg] : 0.f;
    }
    __syncthreads();
    // ---- S = Q K^T (+ relative-key band, mask)
    float sacc[4][4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) sacc[ii][jj] = 0.f;
    for (int d = 0; d < D; d += 4) {   // four head-dim columns per 16-byte shared load
      float4 qv[4], kv[4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) qv[ii] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * ii) * DP + d);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) kv[jj] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * jj) * DP + d);
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float a = sacc[ii][jj];
          a = fmaf(qv[ii].x, kv[jj].x, a); a = fmaf(qv[ii].y, kv[jj].y, a);
          a = fmaf(qv[ii].z, kv[jj].z, a); a = fmaf(qv[ii].w, kv[jj].w, a);
          sacc[ii][jj] = a;
        }
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int i = ty + 16 * ii, ig = i0 + i;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = tx + 16 * jj, jg = j0 + j;
        float sv = sacc[ii][jj];
        const int rel = jg - ig;
        if (rel >= -window && rel <= window) sv += Rk[i * nrel + rel + window];
        if (ig >= len || jg >= len) sv = -1e4f;
        if (jg >= T) sv = -INFINITY;
        Ss[i * SP + j] = sv;
      }
    }
    __syncthreads();
    // ---- online softmax: warp w owns rows 8w..8w+7
    {
      const int w = tid >> 5, lane = tid & 31;
      for (int rr = 0; rr < 8; ++rr) {
        const int i = w * 8 + rr;
        float s0 = Ss[i * SP + lane], s1 = Ss[i * SP + lane + 32];
        float mx = fmaxf(s0, s1);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const float m_old = row_m[i];
        const float m_new = fmaxf(m_old, mx);
        const float p0 = expf(s0 - m_new), p1 = expf(s1 - m_new);
        float ps = p0 + p1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
        Ss[i * SP + lane] = p0;
        Ss[i * SP + lane + 32] = p1;
        if (lane == 0) {
          const float alpha = expf(m_old - m_new);  // exp(-inf) = 0 on the first tile
          row_a[i] = alpha;
          row_l[i] = row_l[i] * alpha + ps;
          row_m[i] = m_new;
        }
      }
    }
    __syncthreads();
    // ---- O = O*alpha + P V (+ relative-value band)
    const bool band = (j0 <= i0 + RA_BQ - 1 + window) && (j0 + RA_BK - 1 >= i0 - window);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const float al = row_a[ty + 16 * ii];
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) o[ii][dd] *= al;
    }
    for (int j = 0; j < RA_BK; j += 4) {   // a thread owns head-dim columns ND*tx .. ND*tx+ND-1
      float4 pv[4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) pv[ii] = *reinterpret_cast<const float4*>(Ss + (ty + 16 * ii) * SP + j);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float vv[ND];
#pragma unroll
        for (int dd = 0; dd < ND; dd += 2) {
          const float2 t2 = *reinterpret_cast<const float2*>(Vs + (j + jj) * D + ND * tx + dd);
          vv[dd] = t2.x; vv[dd + 1] = t2.y;
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const float pj = jj == 0 ? pv[ii].x : jj == 1 ? pv[ii].y : jj == 2 ? pv[ii].z : pv[ii].w;
#pragma unroll
          for (int dd = 0; dd < ND; ++dd) o[ii][dd] = fmaf(pj, vv[dd], o[ii][dd]);
        }
      }
    }
    if (band) {
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const int i = ty + 16 * ii, ig = i0 + i;
        for (int rr = 0; rr < nrel; ++rr) {
          const int jg = ig + rr - window;
          if (jg < j0 || jg >= j0 + RA_BK || jg >= T || jg < 0) continue;
          const float pr = Ss[i * SP + (jg - j0)];
#pragma unroll
          for (int dd = 0; dd < ND; ++dd) o[ii][dd] = fmaf(pr, Ev[rr * D + ND * tx + dd], o[ii][dd]);
        }
      }
    }
  }
  __syncthreads();
  // ---- normalise, stage through smem, store coalesced along T
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int i = ty + 16 * ii;
    const float inv = 1.f / row_l[i];
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) Qs[i * DP + ND * tx + dd] = o[ii][dd] * inv;
  }
  __syncthreads();
  float* ob = out + ((long long)b * H + (long long)h * D) * T;
  for (int idx = tid; idx < D * RA_BQ; idx += 256) {
    const int d = idx / RA_BQ, i = idx % RA_BQ;
    const int ig = i0 + i;
    if (ig < T) ob[(long long)d * T + ig] = Qs[i * DP + d];
  }
}

int launch_rel_attention(const float* qkv, const float* ek, const float* ev,
                         const long long* lengths, float* out, int B, int H, int heads, int window,
                         int T, cudaStream_t s) {
  const int D = H / heads;
  if (D != 96 || H % heads != 0) {
    set_error("rel_attention: only head dim 96 is built (hidden_channels/heads)");
    return SVCB_E_UNSUPPORTED;
  }
  const int nrel = 2 * window + 1;
  const size_t smem = (size_t)(RA_BQ * (D + 4) + RA_BK * (D + 4) + RA_BK * D + RA_BQ * (RA_BK + 4) +
                               2 * nrel * D + RA_BQ * nrel + 3 * RA_BQ) * sizeof(float);
  static DevSmemCache attr_cache;
  SVCB_CUDA_CHECK(ensure_dyn_smem(rel_attention_kernel<96>, 200 * 1024, attr_cache));
  dim3 grid((T + RA_BQ - 1) / RA_BQ, heads, B);
  KernelScope ks("rel_attention", s, 4.0 * B * H * (double)T * T, 16.0 * B * H * (double)T);
  rel_attention_kernel<96><<<grid, 256, smem, s>>>(qkv, ek, ev, lengths, out, H, heads, window, T);
  SVCB_LAUNCH_CHECK("rel_attention");
  return SVCB_OK;
}

// ============================================================================ small ops
__global__ void linear_small_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                    const float* __restrict__ bias, float* __restrict__ y, int In,
                                    int Out) {
  const int o = blockIdx.x * blockDim.y + threadIdx.y, b = blockIdx.y;
  if (o >= Out) return;
  const float* xr = x + (long long)b * In;
  const float* wr = W + (long long)o * In;
  float a = 0.f;
  for (int i = threadIdx.x; i < In; i += 32) a = fmaf(__ldg(wr + i), __ldg(xr + i), a);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
  if (threadIdx.x == 0) y[(long long)b * Out + o] = a + (bias ? __ldg(bias + o) : 0.f);
}

int launch_linear_small(const float* x, const float* W, const float* bias, float* y, int B, int In,
                        int Out, cudaStream_t s) {
  dim3 block(32, 8), grid((Out + 7) / 8, B);
  KernelScope ks("linear_small", s, 2.0 * B * In * Out, 4.0 * ((double)In * Out + B * (In + Out)));
  linear_small_kernel<<<grid, block, 0, s>>>(x, W, bias, y, In, Out);
  SVCB_LAUNCH_CHECK("linear_small");
  return SVCB_OK;
}

// f0_to_coarse (vits/utils.py:20-33) folded into the embedding gather (vits/models.py:47).
__device__ __forceinline__ int f0_coarse(float f0) {
  // fp32 arithmetic exactly as torch evaluates it, with the one library call (log) taken in
  // double and rounded once, so the bin index does not depend on logf's last ulp.
  const float mel_min = 77.75496616579426f;            // 1127*ln(1+50/700)
  const float mel_span = 986.6532670978451f;           // 1127*ln(1+1100/700) - mel_min
  float mel = 1127.f * (float)log((double)(1.f + f0 / 700.f));
  if (mel > 0.f) mel = (mel - mel_min) * 254.f / mel_span + 1.f;
  if (mel <= 1.f) mel = 1.f;
  if (mel > 255.f) mel = 255.f;
  return (int)(mel + 0.5f);
}

__global__ void pitch_embed_add_kernel(float* __restrict__ x, const float* __restrict__ pit,
                                       const float* __restrict__ emb, int C, int T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.z;
  if (t >= T) return;
  const int bin = f0_coarse(pit[(long long)b * T + t]);
  const float* e = emb + (long long)bin * C;
  for (int c = blockIdx.y; c < C; c += gridDim.y) x[((long long)b * C + c) * T + t] += __ldg(e + c);
}

int launch_pitch_embed_add(float* x, const float* pit, const float* emb, int B, int C, int T,
                           cudaStream_t s) {
  dim3 grid((T + 127) / 128, 16, B);
  KernelScope ks("pitch_embed_add", s, 0.0, 12.0 * B * C * (double)T);
  pitch_embed_add_kernel<<<grid, 128, 0, s>>>(x, pit, emb, C, T);
  SVCB_LAUNCH_CHECK("pitch_embed_add");
  return SVCB_OK;
}

__global__ void reparam_kernel(const float* __restrict__ stats, const float* __restrict__ eps,
                               const long long* __restrict__ lengths, float* __restrict__ z, int C,
                               int T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const float m = stats[((long long)b * 2 * C + c) * T + t];
  const float lg = stats[((long long)b * 2 * C + C + c) * T + t];
  const long long o = ((long long)b * C + c) * T + t;
  const float v = m + eps[o] * expf(lg);
  z[o] = (lengths && t >= lengths[b]) ? 0.f : v;
}

int launch_reparam(const float* stats, const float* eps, const long long* lengths, float* z_p, int B,
                   int C, int T, cudaStream_t s) {
  dim3 grid((T + 127) / 128, C, B);
  KernelScope ks("reparam", s, 0.0, 16.0 * B * C * (double)T);
  reparam_kernel<<<grid, 128, 0, s>>>(stats, eps, lengths, z_p, C, T);
  SVCB_LAUNCH_CHECK("reparam");
  return SVCB_OK;
}

// Coupling layer front (vits/modules.py:289-294) with the preceding Flip (modules.py:225-229)
// folded into the channel index.
__global__ void coupling_pre_kernel(const float* __restrict__ xin, const float* __restrict__ sp,
                                    const long long* __restrict__ lengths, float* __restrict__ y,
                                    float* __restrict__ x0n, int C, int T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, b = blockIdx.z;
  const int half = C / 2;
  if (t >= T) return;
  const float x0 = xin[((long long)b * C + (C - 1 - c)) * T + t];
  y[((long long)b * C + c) * T + t] = x0;
  const float sm = __ldg(sp + (long long)b * C + c), sv = __ldg(sp + (long long)b * C + half + c);
  const float mk = (lengths && t >= lengths[b]) ? 0.f : 1.f;
  x0n[((long long)b * half + c) * T + t] = (x0 - sm) * expf(-sv) * mk;
}

int launch_coupling_pre(const float* xin, const float* sp, const long long* lengths, float* y,
                        float* x0n, int B, int C, int T, cudaStream_t s) {
  dim3 grid((T + 127) / 128, C / 2, B);
  KernelScope ks("coupling_pre", s, 0.0, 6.0 * B * C * (double)T);
  coupling_pre_kernel<<<grid, 128, 0, s>>>(xin, sp, lengths, y, x0n, C, T);
  SVCB_LAUNCH_CHECK("coupling_pre");
  return SVCB_OK;
}

// Coupling layer back, reverse branch with mean_only (vits/modules.py:313-316).
__global__ void coupling_post_kernel(const float* __restrict__ xin, const float* __restrict__ sp,
                                     const float* __restrict__ m,
                                     const long long* __restrict__ lengths, float* __restrict__ y,
                                     int C, int T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, b = blockIdx.z;
  const int half = C / 2;
  if (t >= T) return;
  const float x1 = xin[((long long)b * C + (half - 1 - c)) * T + t];  // flip(x)[half + c]
  const float sm = __ldg(sp + (long long)b * C + c), sv = __ldg(sp + (long long)b * C + half + c);
  const float mk = (lengths && t >= lengths[b]) ? 0.f : 1.f;
  float v = (x1 - m[((long long)b * half + c) * T + t]) * mk;
  v = (sm + v * expf(sv)) * mk;
  y[((long long)b * C + half + c) * T + t] = v;
}

int launch_coupling_post(const float* xin, const float* sp, const float* m, const long long* lengths,
                         float* y, int B, int C, int T, cudaStream_t s) {
  dim3 grid((T + 127) / 128, C / 2, B);
  KernelScope ks("coupling_post", s, 0.0, 6.0 * B * C * (double)T);
  coupling_post_kernel<<<grid, 128, 0, s>>>(xin, sp, m, lengths, y, C, T);
  SVCB_LAUNCH_CHECK("coupling_post");
  return SVCB_OK;
}

// WN residual/skip bookkeeping (vits/modules.py:196-203).
__global__ void wn_update_kernel(float* __restrict__ x, float* __restrict__ out,
                                 const float* __restrict__ rs, const long long* __restrict__ lengths,
                                 int H, int T, int first, int last) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const float mk = (lengths && t >= lengths[b]) ? 0.f : 1.f;
  const long long o = ((long long)b * H + c) * T + t;
  if (!last) {
    const float ra = rs[((long long)b * 2 * H + c) * T + t];
    const float sk = rs[((long long)b * 2 * H + H + c) * T + t];
    x[o] = (x[o] + ra) * mk;
    out[o] = first ? sk : out[o] + sk;
  } else {
    const float sk = rs[((long long)b * H + c) * T + t];
    out[o] = ((first ? 0.f : out[o]) + sk) * mk;
  }
}

int launch_wn_update(float* x, float* out, const float* rs, const long long* lengths, int B, int H,
                     int T, int first, int last, cudaStream_t s) {
  dim3 grid((T + 127) / 128, H, B);
  KernelScope ks("wn_update", s, 0.0, 24.0 * B * H * (double)T);
  wn_update_kernel<<<grid, 128, 0, s>>>(x, out, rs, lengths, H, T, first, last);
  SVCB_LAUNCH_CHECK("wn_update");
  return SVCB_OK;
}

}  // namespace svcb
