// AMP-block link on the 5th-gen tensor cores, as two kernels that each stream at their own roofline:
//
//   snake_pack   SnakeAlias(x) -> bf16 hi/lo operand image in HBM            (CUDA cores, HBM-bound)
//   amp_conv_tc  Conv1d(C->C, K, dilation) + bias (+residual, stage mean)    (tcgen05 + TMEM)
//
// Together they replace one `SnakeAlias -> Conv1d [-> + x]` link of AMPBlock.forward
// (vits_decoder/bigv.py:50-58; SnakeAlias = vits_decoder/alias/act.py:124-128), SURVEY.md §8a rows
// a9/a10.  A first version fused both into one CTA-per-tile kernel; measured on B200 it ran at the
// CUDA-core path's speed (15 TFLOP/s) because the per-tile Snake prologue (20 x load->sync->FIR->
// sync->FIR->sync with one resident CTA) was latency-bound and left the tensor pipe idle
// (profiles/r01_notes.md).  Splitting lets the Snake pass run with 32 warps/SM and lets the conv
// kernel feed its A operand with bulk (TMA-engine) copies.
//
// Operand image ("P8" layout): hi/lo bf16 [B][Cp/8][Lp][8], Lp = 32 + roundup(L,128) + 32, row =
// 32 + t; rows outside the sequence and channels >= C are zero.  One (octet, row-range) of the A
// tile is therefore ONE contiguous run of R*16 bytes = one bulk copy straight into the K-major
// panel layout of tc.cuh, the zero rows are the conv's zero padding, and every tap is the same
// tile addressed through a row-shifted descriptor.
#include "common.cuh"
#include "tc.cuh"

namespace svcb {

constexpr int TC_M = 128;
constexpr int P8_PAD = 32;

__host__ __device__ inline int p8_rows_of(int L) { return P8_PAD + (L + 127) / 128 * 128 + P8_PAD; }
int p8_rows(int L) { return p8_rows_of(L); }

__device__ __forceinline__ float fast_sin(float x) {
  // Cody-Waite reduction to [-pi, pi] then the SFU sine: abs error < 1e-6 for |x| < 1e3, an order
  // of magnitude below the bf16x3 operand rounding already accepted here.
  const float k = rintf(x * 0.15915494309189535f);
  x = fmaf(k, -6.2831854820251465f, x);
  x = fmaf(k, 1.7484555e-7f, x);
  return __sinf(x);
}

// ------------------------------------------------------------------------------------ snake_pack
constexpr int SP_TL = 512;  // image rows per CTA

__global__ void __launch_bounds__(256)
snake_pack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                  const float* __restrict__ ea, const float* __restrict__ inv_b,
                  const float* __restrict__ fu, const float* __restrict__ fd, int C, int L, int Lp) {
  extern __shared__ __align__(16) float sp_smem[];
  constexpr int XW = SP_TL + 12, VW = 2 * SP_TL + 12;
  float* xs = sp_smem;            // [8][XW]
  float* vs = sp_smem + 8 * XW;   // [8][VW]
  __shared__ float f_up[12], f_dn[12];
  const int tid = threadIdx.x;
  const int oc = blockIdx.y, b = blockIdx.z;
  const int row0 = blockIdx.x * SP_TL;      // first image row of this CTA
  const int n0 = row0 - P8_PAD;             // its sequence position
  if (tid < 12) { f_up[tid] = __ldg(fu + tid); f_dn[tid] = __ldg(fd + tid); }
  const long long img = ((long long)b * gridDim.y + oc) * Lp;
  const bool any = n0 < L && n0 + SP_TL > 0 && oc * 8 < C;  // block-uniform
  if (any) {
    const float* xb = x + (long long)b * C * L;
    for (int idx = tid; idx < 8 * XW; idx += 256) {
      const int c = idx / XW, i = idx - c * XW;
      const int cg = oc * 8 + c;
      int g = n0 - 6 + i;
      g = min(max(g, 0), L - 1);
      xs[idx] = cg < C ? __ldg(xb + (long long)cg * L + g) : 0.f;
    }
    __syncthreads();
    constexpr int NV = 2 * SP_TL + 10;
    for (int idx = tid; idx < 8 * NV; idx += 256) {
      const int c = idx / NV, iv = idx - c * NV;
      const int cg = min(oc * 8 + c, C - 1);
      int m = 2 * n0 - 5 + iv;
      m = min(max(m, 0), 2 * L - 1);
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
      const float sn = fast_sin(u * __ldg(ea + cg));
      vs[c * VW + iv] = u + __ldg(inv_b + cg) * (sn * sn);
    }
    __syncthreads();
  }
  for (int r = tid; r < SP_TL; r += 256) {
    const int row = row0 + r;
    if (row >= Lp) break;
    const int tau = n0 + r;
    __align__(16) __nv_bfloat16 h8[8], l8[8];
    const bool inside = any && tau >= 0 && tau < L;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float o = 0.f;
      if (inside && oc * 8 + c < C) {
        const float2* vp = reinterpret_cast<const float2*>(vs + c * VW + 2 * r);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const float2 v2 = vp[k];
          o = fmaf(v2.x, f_dn[2 * k], o);
          o = fmaf(v2.y, f_dn[2 * k + 1], o);
        }
      }
      h8[c] = __float2bfloat16_rn(o);
      l8[c] = __float2bfloat16_rn(o - __bfloat162float(h8[c]));
    }
    *reinterpret_cast<uint4*>(hi + (img + row) * 8) = *reinterpret_cast<const uint4*>(h8);
    if (lo) *reinterpret_cast<uint4*>(lo + (img + row) * 8) = *reinterpret_cast<const uint4*>(l8);
  }
}

size_t p8_image_bytes(int B, int C, int L) {  // one of hi / lo
  const int cp = (C + 15) / 16 * 16;
  return (size_t)B * (cp / 8) * p8_rows_of(L) * 16;
}

int launch_snake_pack(const float* x, void* hi, void* lo, const float* ea, const float* inv_b, const float* fu,
                      const float* fd, int B, int C, int L, cudaStream_t s) {
  if (B <= 0 || C <= 0 || L <= 0) return SVCB_OK;
  const int cp = (C + 15) / 16 * 16, Lp = p8_rows_of(L);
  const size_t smem = (size_t)(8 * (SP_TL + 12) + 8 * (2 * SP_TL + 12)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    SVCB_CUDA_CHECK(cudaFuncSetAttribute(snake_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid((Lp + SP_TL - 1) / SP_TL, cp / 8, B);
  KernelScope ks("snake_pack", s, 70.0 * B * C * (double)L, (lo ? 8.0 : 6.0) * B * C * (double)L);
  snake_pack_kernel<<<grid, 256, smem, s>>>(x, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), ea,
                                            inv_b, fu, fd, C, L, Lp);
  SVCB_LAUNCH_CHECK("snake_pack");
  return SVCB_OK;
}

// ------------------------------------------------------------------------------------ amp_conv_tc
__global__ void __launch_bounds__(192, 1)
amp_conv_tc_kernel(const AmpConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_a, bar_full[2], bar_empty[2], bar_acc;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TC_M;
  const int P = p.dil * (p.K - 1) / 2;
  const int R = TC_M + (p.K - 1) * p.dil;
  const int KC = p.Cp / 8;
  const uint32_t a_bytes = (uint32_t)KC * R * 16u;
  const uint32_t wb = (uint32_t)p.Cp * p.Cp * 2u;
  const int parts = p.nsplit == 3 ? 2 : 1;
  const int nch = p.K * parts;
  uint8_t* A_hi = smem;
  uint8_t* A_lo = smem + a_bytes;
  uint8_t* W0 = smem + (uint32_t)parts * a_bytes;
  uint8_t* W1 = W0 + wb;

  if (tid == 0) {
    tc::mbar_init(&bar_a, 1);
    tc::mbar_init(&bar_full[0], 1); tc::mbar_init(&bar_full[1], 1);
    tc::mbar_init(&bar_empty[0], 1); tc::mbar_init(&bar_empty[1], 1);
    tc::mbar_init(&bar_acc, 1);
    tc::fence_barrier_init();
  }
  const uint32_t ncols = tc::tmem_cols_for(p.Cp);
  __syncwarp();
  if (warp == 4) tc::tmem_alloc(&tmem_slot, ncols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (tid == 128) {
    // ---------------------------------------------------------------- producer: A image rows + weight ring
    const uint32_t run = (uint32_t)R * 16u;
    tc::mbar_arrive_expect_tx(&bar_a, run * KC * parts);
    const long long row = P8_PAD + t0 - P;
    for (int part = 0; part < parts; ++part) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(part == 0 ? p.a_hi : p.a_lo);
      uint8_t* dst = part == 0 ? A_hi : A_lo;
      for (int kc = 0; kc < KC; ++kc)
        tc::bulk_g2s(dst + (size_t)kc * run, src + (((long long)b * KC + kc) * p.Lp + row) * 16, run, &bar_a);
    }
    for (int i = 0; i < nch; ++i) {
      const int st = i & 1;
      if (i >= 2) tc::mbar_wait(&bar_empty[st], (uint32_t)(((i >> 1) - 1) & 1));
      tc::mbar_arrive_expect_tx(&bar_full[st], wb);
      const int tap = i / parts, part = i % parts;
      tc::bulk_g2s(st ? W1 : W0, p.wpk + ((size_t)tap * 2 + part) * wb, wb, &bar_full[st]);
    }
  } else if (tid == 160) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = tc::idesc_bf16(TC_M, p.Cp);
    const uint32_t a_hi = tc::smem_u32(A_hi), a_lo = tc::smem_u32(A_lo);
    const uint32_t w_addr[2] = {tc::smem_u32(W0), tc::smem_u32(W1)};
    const uint32_t lbo_a = (uint32_t)R * 16u, lbo_b = (uint32_t)p.Cp * 16u;
    uint32_t accumulate = 0;
    tc::mbar_wait(&bar_a, 0);
    for (int i = 0; i < nch; ++i) {
      const int st = i & 1;
      tc::mbar_wait(&bar_full[st], (uint32_t)((i >> 1) & 1));
      tc::fence_after_sync();
      const int tap = i / parts, part = i % parts;
      const uint32_t row_off = (uint32_t)(tap * p.dil) * 16u;
      const int n_a = (part == 0 && parts == 2) ? 2 : 1;  // Wh meets Ah and Al; Wl meets Ah
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
  } else if (warp < 4) {
    // ---------------------------------------------------------------- epilogue (TMEM lanes 32w..32w+31)
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
  if (warp == 4) tc::tmem_dealloc(tmem, ncols);
}

size_t amp_conv_tc_smem_bytes(int Cp, int K, int dil, int nsplit) {
  const int R = TC_M + (K - 1) * dil;
  const size_t a = (size_t)(Cp / 8) * R * 16 * (nsplit == 3 ? 2 : 1);
  return a + 2 * (size_t)Cp * Cp * 2 + 128;
}

int launch_amp_conv_tc(const AmpConvParams& p, cudaStream_t s) {
  if (p.Cp % 16 || p.Cp < 16 || p.Cp > 256 || p.Cp < p.C || (p.nsplit != 1 && p.nsplit != 3)) {
    set_error("amp_conv_tc: bad channel padding / nsplit");
    return SVCB_E_BAD_SHAPE;
  }
  if (p.dil * (p.K - 1) / 2 > P8_PAD || p.Lp != p8_rows_of(p.L) || !p.a_hi || (p.nsplit == 3 && !p.a_lo)) {
    set_error("amp_conv_tc: operand image does not match (halo > 32 rows or wrong Lp)");
    return SVCB_E_BAD_SHAPE;
  }
  const size_t smem = amp_conv_tc_smem_bytes(p.Cp, p.K, p.dil, p.nsplit);
  if (smem > 227 * 1024 - 512) { set_error("amp_conv_tc: tile does not fit shared memory"); return SVCB_E_UNSUPPORTED; }
  static size_t attr_bytes = 0;  // the dynamic limit excludes the kernel's (small) static shared memory
  if (smem > attr_bytes) {
    SVCB_CUDA_CHECK(cudaFuncSetAttribute(amp_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    attr_bytes = smem;
  }
  dim3 grid((p.L + TC_M - 1) / TC_M, p.B);
  const double macs = (double)p.B * p.L * p.C * p.C * p.K;
  KernelScope ks(p.nsplit == 3 ? "amp_conv_tc_bf16x3" : "amp_conv_tc_bf16", s, 2.0 * macs,
                 (double)p.B * p.C * p.L * ((p.nsplit == 3 ? 4.0 : 2.0) + 4.0 * (p.res ? 2 : 1)));
  amp_conv_tc_kernel<<<grid, 192, smem, s>>>(p);
  SVCB_LAUNCH_CHECK("amp_conv_tc");
  return SVCB_OK;
}

}  // namespace svcb
