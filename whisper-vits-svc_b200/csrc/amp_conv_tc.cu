// AMP-block convolution on the 5th-gen tensor cores (tcgen05 + TMEM), with the anti-aliased
// Snake activation fused as the operand-staging prologue and bias / residual / stage-mean fused
// as the TMEM epilogue.
//
// Replaces, per launch, one `SnakeAlias -> Conv1d(+bias) [-> + residual]` link of
// AMPBlock.forward (vits_decoder/bigv.py:50-58; SnakeAlias = alias/act.py:124-128): 2 of the
// ~10 kernels the CUDA-core path needs per link, and the only part of the generator whose
// arithmetic is a dense contraction.
//
// Implicit GEMM, one CTA per (128 output samples, item):
//   D[t, co] = sum_tap sum_ci A[t + tap*dil, ci] * W_tap[co, ci]        M=128, N=Cp, K=Cp per tap
// * A (activations after SnakeAlias, zero outside the sequence = the conv's zero padding) is
//   produced by the CUDA cores straight into shared memory in the K-major "panel" layout of
//   tc.cuh, R = 128 + (K-1)*dil rows, so every tap is the SAME tile addressed with a descriptor
//   advanced by tap*dil rows — no im2col, no re-staging.
// * W_tap tiles (pre-packed by the host in the exact shared-memory image) stream through a
//   2-stage ring with 1-D bulk copies (TMA engine) signalled on mbarriers.
// * One thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM); tcgen05.commit frees
//   ring slots and finally signals the epilogue warps, which read TMEM with tcgen05.ld.
// * Precision: nsplit=1 plain bf16; nsplit=3 "bf16x3": A = Ah+Al, W = Wh+Wl (bf16 each) and
//   D = Ah*Wh + Al*Wh + Ah*Wl, i.e. ~16 mantissa bits per operand — the parity-grade mode
//   (measured waveform error 3e-5 vs 1.2e-2 for plain bf16).
#include "common.cuh"
#include "tc.cuh"

namespace svcb {

constexpr int TC_M = 128;
constexpr int TC_THREADS = 256;

__device__ __forceinline__ float fast_sin(float x) {
  // Cody-Waite reduction to [-pi, pi] then the SFU sine: abs error < 1e-6 for |x| < 1e3, an
  // order of magnitude below the bf16x3 operand rounding this kernel already accepts.
  const float k = rintf(x * 0.15915494309189535f);
  x = fmaf(k, -6.2831854820251465f, x);
  x = fmaf(k, 1.7484555e-7f, x);
  return __sinf(x);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
amp_conv_tc_kernel(const AmpConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_acc;
  __shared__ uint32_t tmem_slot;
  __shared__ float f_up[12], f_dn[12];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TC_M;
  const int P = p.dil * (p.K - 1) / 2;
  const int R = TC_M + (p.K - 1) * p.dil;
  const int KC = p.Cp / 8;
  const int n0 = t0 - P;  // sequence position of A row 0
  const uint32_t a_bytes = (uint32_t)KC * R * 16u;
  const uint32_t wb = (uint32_t)p.Cp * p.Cp * 2u;
  const int parts_w = p.nsplit == 3 ? 2 : 1;
  const int nch = p.K * parts_w;
  uint8_t* A_hi = smem;
  uint8_t* A_lo = smem + a_bytes;  // only when nsplit == 3
  uint8_t* W0 = smem + (p.nsplit == 3 ? 2u : 1u) * a_bytes;
  uint8_t* W1 = W0 + wb;           // also the prologue's staging area
  float* xs = reinterpret_cast<float*>(W1);        // [8][R + 12]
  float* vs = xs + 8 * (R + 12);                   // [8][2R + 12]
  const int XW = R + 12, VW = 2 * R + 12;

  if (tid == 0) {
    tc::mbar_init(&bar_full[0], 1); tc::mbar_init(&bar_full[1], 1);
    tc::mbar_init(&bar_empty[0], 1); tc::mbar_init(&bar_empty[1], 1);
    tc::mbar_init(&bar_acc, 1);
    tc::fence_barrier_init();
  }
  if (tid < 12) { f_up[tid] = __ldg(p.fu + tid); f_dn[tid] = __ldg(p.fd + tid); }
  const uint32_t ncols = tc::tmem_cols_for(p.Cp);
  __syncwarp();
  if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  // first weight tile travels while the CUDA cores build A
  if (tid == 128) {
    tc::mbar_arrive_expect_tx(&bar_full[0], wb);
    tc::bulk_g2s(W0, p.wpk, wb, &bar_full[0]);
  }

  // ------------------------------------------------------------------ prologue: A = SnakeAlias(x)
  const float* xb = p.x + (long long)b * p.C * p.L;
  for (int kc = 0; kc < KC; ++kc) {
    // (a) raw x, 8 channels x (R+12) samples, replicate-clamped at the sequence ends
    for (int idx = tid; idx < 8 * XW; idx += TC_THREADS) {
      const int c = idx / XW, i = idx - c * XW;
      const int cg = kc * 8 + c;
      int g = n0 - 6 + i;
      g = min(max(g, 0), p.L - 1);
      xs[idx] = cg < p.C ? __ldg(xb + (long long)cg * p.L + g) : 0.f;
    }
    __syncthreads();
    // (b) 2x up-sampled Snake: v[m], m = 2*n0 - 5 + idx
    for (int idx = tid; idx < 8 * (2 * R + 10); idx += TC_THREADS) {
      const int c = idx / (2 * R + 10), iv = idx - c * (2 * R + 10);
      const int cg = min(kc * 8 + c, p.C - 1);
      int m = 2 * n0 - 5 + iv;
      m = min(max(m, 0), 2 * p.L - 1);
      const int a = m >> 1;
      const float* xp = xs + c * XW + (a - (n0 - 6));
      float acc = 0.f;
      if ((m & 1) == 0) {
#pragma unroll
        for (int d = -3; d <= 2; ++d) acc = fmaf(xp[d], f_up[5 - 2 * d], acc);
      } else {
#pragma unroll
        for (int d = -2; d <= 3; ++d) acc = fmaf(xp[d], f_up[6 - 2 * d], acc);
      }
      const float u = 2.f * acc;
      const float sn = fast_sin(u * __ldg(p.ea + cg));
      vs[c * VW + iv] = u + __ldg(p.ib + cg) * (sn * sn);
    }
    __syncthreads();
    // (c) 12-tap decimation, bf16 split, one 16-byte K-chunk per row
    for (int r = tid; r < R; r += TC_THREADS) {
      const int tau = n0 + r;
      const bool inside = tau >= 0 && tau < p.L;
      __nv_bfloat16 hi[8], lo[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float o = 0.f;
        if (inside && kc * 8 + c < p.C) {
          const float* vp = vs + c * VW + 2 * r;
#pragma unroll
          for (int k = 0; k < 12; ++k) o = fmaf(vp[k], f_dn[k], o);
        }
        hi[c] = __float2bfloat16_rn(o);
        lo[c] = __float2bfloat16_rn(o - __bfloat162float(hi[c]));
      }
      *reinterpret_cast<uint4*>(A_hi + ((size_t)kc * R + r) * 16) = *reinterpret_cast<const uint4*>(hi);
      if (p.nsplit == 3)
        *reinterpret_cast<uint4*>(A_lo + ((size_t)kc * R + r) * 16) = *reinterpret_cast<const uint4*>(lo);
    }
    __syncthreads();
  }
  tc::fence_proxy_async_smem();  // generic-proxy writes of A -> visible to the tensor core (async proxy)
  __syncthreads();

  // ------------------------------------------------------------------ weight producer (1 thread)
  if (tid == 128) {
    for (int i = 1; i < nch; ++i) {
      const int st = i & 1;
      if (i >= 2) tc::mbar_wait(&bar_empty[st], (uint32_t)(((i >> 1) - 1) & 1));
      tc::mbar_arrive_expect_tx(&bar_full[st], wb);
      const int tap = i / parts_w, part = i % parts_w;
      tc::bulk_g2s(st ? W1 : W0, p.wpk + ((size_t)tap * 2 + part) * wb, wb, &bar_full[st]);
    }
  }
  // ------------------------------------------------------------------ MMA issuer (1 thread)
  if (tid == 160) {
    const uint32_t idesc = tc::idesc_bf16(TC_M, p.Cp);
    const uint32_t a_hi = tc::smem_u32(A_hi), a_lo = tc::smem_u32(A_lo);
    const uint32_t w_addr[2] = {tc::smem_u32(W0), tc::smem_u32(W1)};
    const uint32_t lbo_a = (uint32_t)R * 16u, lbo_b = (uint32_t)p.Cp * 16u;
    uint32_t accumulate = 0;
    for (int i = 0; i < nch; ++i) {
      const int st = i & 1;
      tc::mbar_wait(&bar_full[st], (uint32_t)((i >> 1) & 1));
      tc::fence_after_sync();
      const int tap = i / parts_w, part = i % parts_w;
      const uint32_t row_off = (uint32_t)(tap * p.dil) * 16u;
      const int n_a = (part == 0 && p.nsplit == 3) ? 2 : 1;  // Wh meets Ah and Al; Wl meets Ah
      for (int ap = 0; ap < n_a; ++ap) {
        const uint32_t abase = (ap == 0 ? a_hi : a_lo) + row_off;
        for (int kk = 0; kk < p.Cp / 16; ++kk) {
          const uint64_t ad = tc::smem_desc(abase + (uint32_t)kk * 2u * lbo_a, lbo_a);
          const uint64_t bd = tc::smem_desc(w_addr[st] + (uint32_t)kk * 2u * lbo_b, lbo_b);
          tc::mma_bf16(tmem, ad, bd, idesc, accumulate);
          accumulate = 1;
        }
      }
      tc::mma_commit(&bar_empty[st]);
    }
    tc::mma_commit(&bar_acc);
  }

  // ------------------------------------------------------------------ epilogue (warps 0-3 <-> TMEM lanes)
  if (warp < 4) {
    tc::mbar_wait(&bar_acc, 0);
    tc::fence_after_sync();
    const int t = t0 + warp * 32 + lane;
    const long long rowb = (long long)b * p.C * p.L;
    for (int c0 = 0; c0 < p.Cp; c0 += 16) {
      uint32_t v[16];
      tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tc::tmem_ld_wait();
      if (t < p.L) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = c0 + j;
          if (co < p.C) {
            const long long off = rowb + (long long)co * p.L + t;
            float o = __uint_as_float(v[j]) + __ldg(p.bias + co);
            if (p.res) o += p.res[off];
            if (p.accum) o += p.y[off];
            if (p.out_div != 0.f) o = o / p.out_div;
            p.y[off] = o;
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

size_t amp_conv_tc_smem_bytes(int Cp, int K, int dil, int nsplit) {
  const int R = TC_M + (K - 1) * dil;
  const size_t a = (size_t)(Cp / 8) * R * 16 * (nsplit == 3 ? 2 : 1);
  const size_t wb = (size_t)Cp * Cp * 2;
  const size_t staging = (size_t)8 * (3 * R + 24) * sizeof(float);
  return a + wb + (wb > staging ? wb : staging) + 128;
}

int launch_amp_conv_tc(const AmpConvParams& p, cudaStream_t s) {
  if (p.Cp % 16 || p.Cp < 16 || p.Cp > 256 || p.Cp < p.C || (p.nsplit != 1 && p.nsplit != 3)) {
    set_error("amp_conv_tc: bad channel padding / nsplit");
    return SVCB_E_BAD_SHAPE;
  }
  const size_t smem = amp_conv_tc_smem_bytes(p.Cp, p.K, p.dil, p.nsplit);
  if (smem > 227 * 1024 - 512) { set_error("amp_conv_tc: tile does not fit shared memory"); return SVCB_E_UNSUPPORTED; }
  static size_t attr_bytes = 0;  // dynamic limit excludes the kernel's (small) static shared memory
  if (smem > attr_bytes) {
    SVCB_CUDA_CHECK(cudaFuncSetAttribute(amp_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    attr_bytes = smem;
  }
  dim3 grid((p.L + TC_M - 1) / TC_M, p.B);
  const double macs = (double)p.B * p.L * p.C * p.C * p.K;
  KernelScope ks(p.nsplit == 3 ? "amp_conv_tc_bf16x3" : "amp_conv_tc_bf16", s, 2.0 * macs,
                 4.0 * (double)p.B * p.C * p.L * (p.res ? 3 : 2));
  amp_conv_tc_kernel<<<grid, TC_THREADS, smem, s>>>(p);
  SVCB_LAUNCH_CHECK("amp_conv_tc");
  return SVCB_OK;
}

}  // namespace svcb
