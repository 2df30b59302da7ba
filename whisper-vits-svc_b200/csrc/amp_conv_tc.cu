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
#include <algorithm>
#include <cstdio>

#include "common.cuh"
#include "tc.cuh"

namespace svcb {

static inline unsigned tc_cols_host(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u; }

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
// One CTA = 8 channels (one K-octet) x SP_TL image rows.  Three phases over shared memory:
//  (a) x rows with a 6-sample halo, replicate-clamped at the sequence ends;
//  (b) one work item per INPUT sample a: both up-sampled Snake values v[2a], v[2a+1] from the same
//      7 inputs (12 FMA + 2 SFU sines), stored at vs[j], j = m - (2*n0 - 5);
//  (c) one thread per pair of output rows: 14 consecutive v (4 x LDS.128) -> two 12-tap
//      decimations per channel, bf16 hi/lo split, two 16-byte rows stored contiguously.
// Only the first/last CTA of a sequence needs the replicate padding of v (fix-up pass).
constexpr int SP_TL = 506;  // image rows per CTA: 506 + 6 = 512 up-sampling positions = 2 exact passes of 256 threads
constexpr int SP_XW = SP_TL + 12;
constexpr int SP_VW = 2 * SP_TL + 16;

__global__ void __launch_bounds__(256)
snake_pack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                  const float* __restrict__ ea, const float* __restrict__ inv_b,
                  const float* __restrict__ fu, const float* __restrict__ fd, int C, int L, int Lp) {
  extern __shared__ __align__(16) float sp_smem[];
  float* xs = sp_smem;                 // [8][SP_XW]
  float* vs = sp_smem + 8 * SP_XW;     // [8][SP_VW]
  __shared__ float f_up[12], f_dn[12], s_ea[8], s_ib[8];
  const int tid = threadIdx.x;
  const int oc = blockIdx.y, b = blockIdx.z;
  const int row0 = blockIdx.x * SP_TL;  // first image row of this CTA
  const int n0 = row0 - P8_PAD;         // its sequence position
  const int nvalid = min(8, C - oc * 8);
  if (tid < 12) { f_up[tid] = __ldg(fu + tid); f_dn[tid] = __ldg(fd + tid); }
  if (tid >= 32 && tid < 32 + 8 && tid - 32 < nvalid) {
    s_ea[tid - 32] = __ldg(ea + oc * 8 + tid - 32);
    s_ib[tid - 32] = __ldg(inv_b + oc * 8 + tid - 32);
  }
  const long long img = ((long long)b * gridDim.y + oc) * Lp;
  const bool any = n0 < L && n0 + SP_TL > 0 && nvalid > 0;  // block-uniform
  const int mbase = 2 * n0 - 5;
  if (any) {
    const float* xb = x + ((long long)b * C + oc * 8) * L;
    for (int idx = tid; idx < nvalid * SP_XW; idx += 256) {   // flat over (channel, position)
      const int c = idx / SP_XW, i = idx - c * SP_XW;
      const int g = min(max(n0 - 6 + i, 0), L - 1);
      xs[idx] = __ldg(xb + (long long)c * L + g);
    }
    __syncthreads();
    static_assert(SP_TL + 6 == 512, "phase (b) indexing assumes 512 up-sampling positions per channel");
    for (int idx = tid; idx < nvalid * 512; idx += 256) {     // flat over (channel, a - n0 + 3)
      const int c = idx >> 9, ar = (idx & 511) - 3;
      const float a_ = s_ea[c], ib = s_ib[c];
      const float* xp = xs + c * SP_XW + ar + 6;
      float* vc = vs + c * SP_VW;
      const float xm3 = xp[-3], xm2 = xp[-2], xm1 = xp[-1], x0 = xp[0], x1 = xp[1], x2 = xp[2], x3 = xp[3];
      float ue = xm3 * f_up[11];
      ue = fmaf(xm2, f_up[9], ue); ue = fmaf(xm1, f_up[7], ue); ue = fmaf(x0, f_up[5], ue);
      ue = fmaf(x1, f_up[3], ue); ue = fmaf(x2, f_up[1], ue);
      float uo = xm2 * f_up[10];
      uo = fmaf(xm1, f_up[8], uo); uo = fmaf(x0, f_up[6], uo); uo = fmaf(x1, f_up[4], uo);
      uo = fmaf(x2, f_up[2], uo); uo = fmaf(x3, f_up[0], uo);
      ue *= 2.f; uo *= 2.f;
      const float se = __sinf(ue * a_), so = __sinf(uo * a_);
      const int j = 2 * ar + 5;
      if (j >= 0) vc[j] = fmaf(ib, se * se, ue);
      vc[j + 1] = fmaf(ib, so * so, uo);
    }
    __syncthreads();
    if (mbase < 0 || mbase + 2 * SP_TL + 10 > 2 * L) {  // replicate padding of v at the sequence ends
      const int jlo = -mbase, jhi = 2 * L - 1 - mbase;   // positions of v[0] and v[2L-1]
      for (int c = 0; c < nvalid; ++c) {
        float* vc = vs + c * SP_VW;
        for (int j = tid; j < 2 * SP_TL + 12; j += 256) {
          if (j < jlo) vc[j] = vc[jlo];
          else if (j > jhi && jhi >= 0) vc[j] = vc[jhi];
        }
      }
      __syncthreads();
    }
  }
  // (c) rows 2*tid, 2*tid+1
  {
    const int r = 2 * tid;
    const int row = row0 + r;
    if (r < SP_TL && row < Lp) {
      const int tau = n0 + r;
      __align__(16) __nv_bfloat162 h2[8], l2[8];  // [row parity*4 + channel pair]
      float prev0 = 0.f, prev1 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float o0 = 0.f, o1 = 0.f;
        if (any && c < nvalid && tau + 1 >= 0 && tau < L) {
          const float4* vp = reinterpret_cast<const float4*>(vs + c * SP_VW + 2 * r);
          const float4 q0 = vp[0], q1 = vp[1], q2 = vp[2], q3 = vp[3];
          const float w[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w,
                               q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
          for (int k = 0; k < 12; ++k) { o0 = fmaf(w[k], f_dn[k], o0); o1 = fmaf(w[k + 2], f_dn[k], o1); }
          if (tau < 0) o0 = 0.f;
          if (tau + 1 >= L) o1 = 0.f;
        }
        if (c & 1) {  // one packed conversion per channel pair and row
          const __nv_bfloat162 ha = __floats2bfloat162_rn(prev0, o0), hb = __floats2bfloat162_rn(prev1, o1);
          const float2 fa = __bfloat1622float2(ha), fb = __bfloat1622float2(hb);
          h2[c >> 1] = ha; h2[4 + (c >> 1)] = hb;
          l2[c >> 1] = __floats2bfloat162_rn(prev0 - fa.x, o0 - fa.y);
          l2[4 + (c >> 1)] = __floats2bfloat162_rn(prev1 - fb.x, o1 - fb.y);
        } else {
          prev0 = o0; prev1 = o1;
        }
      }
      const __nv_bfloat162* h8 = h2;
      const __nv_bfloat162* l8 = l2;
      uint4* dh = reinterpret_cast<uint4*>(hi + (img + row) * 8);
      dh[0] = *reinterpret_cast<const uint4*>(h8);
      dh[1] = *reinterpret_cast<const uint4*>(h8 + 4);
      if (lo) {
        uint4* dl = reinterpret_cast<uint4*>(lo + (img + row) * 8);
        dl[0] = *reinterpret_cast<const uint4*>(l8);
        dl[1] = *reinterpret_cast<const uint4*>(l8 + 4);
      }
    }
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
  const size_t smem = (size_t)(8 * SP_XW + 8 * SP_VW) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    SVCB_CUDA_CHECK(cudaFuncSetAttribute(snake_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid((Lp + SP_TL - 1) / SP_TL, cp / 8, B);
  char kname[64];
  snprintf(kname, sizeof(kname), "snake_pack_c%d", C);
  KernelScope ks(kname, s, 70.0 * B * C * (double)L, (lo ? 8.0 : 6.0) * B * C * (double)L);
  snake_pack_kernel<<<grid, 256, smem, s>>>(x, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), ea,
                                            inv_b, fu, fd, C, L, Lp);
  SVCB_LAUNCH_CHECK("snake_pack");
  return SVCB_OK;
}

// ------------------------------------------------------------------------------------ amp_conv_tc
// Persistent: each CTA walks tiles (item, 128 samples) with a static stride.  Three roles pipeline
// across tiles through mbarriers:
//   producer thread  A image rows of tile i+1 (bulk copies, 1-2 buffers) and the weight tiles
//                    (all taps resident in shared memory when they fit, else a 2-slot ring per tile)
//   MMA thread       all taps x split parts of tile i into TMEM accumulator (i & 1)
//   4 epilogue warps tile i-1: tcgen05.ld -> +bias (+res, +stage accumulation, /3) -> coalesced stores
struct AmpPlan { int resident, nabuf, acc_stride, ncols; size_t smem; };

__global__ void __launch_bounds__(320, 1)
amp_conv_tc_kernel(const AmpConvParams p, const int resident, const int nabuf, const int acc_stride,
                   const uint32_t ncols) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2], w_full[2], w_empty[2], w_res, t_full[2], t_empty[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = p.dil * (p.K - 1) / 2;
  const int R = TC_M + (p.K - 1) * p.dil;
  const int KC = p.Cp / 8;
  const int parts = p.nsplit == 3 ? 2 : 1;
  const uint32_t a_part = (uint32_t)KC * R * 16u;
  const uint32_t a_buf = a_part * parts;
  const uint32_t wb = (uint32_t)p.Cp * p.Cp * 2u;
  const int nch = p.K * parts;
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + (size_t)nabuf * a_buf;
  const int tpi = (p.L + TC_M - 1) / TC_M;
  const int ntiles = p.B * tpi;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1);
      tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1);
      tc::mbar_init(&t_full[i], 1); tc::mbar_init(&t_empty[i], 256);
    }
    tc::mbar_init(&w_res, 1);
    tc::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 8) tc::tmem_alloc(&tmem_slot, ncols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (tid == 256) {
    // ---------------------------------------------------------------- producer
    if (resident) {
      tc::mbar_arrive_expect_tx(&w_res, wb * (uint32_t)nch);
      for (int i = 0; i < nch; ++i)
        tc::bulk_g2s(Wbase + (size_t)i * wb, p.wpk + ((size_t)(i / parts) * 2 + (i % parts)) * wb, wb, &w_res);
    }
    const uint32_t run = (uint32_t)R * 16u;
    int it = 0, wi = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it % nabuf;
      if (it >= nabuf) tc::mbar_wait(&a_empty[buf], (uint32_t)(((it / nabuf) - 1) & 1));
      const int b = tile / tpi, t0 = (tile - b * tpi) * TC_M;
      tc::mbar_arrive_expect_tx(&a_full[buf], run * KC * parts);
      const long long row = P8_PAD + t0 - P;
      for (int part = 0; part < parts; ++part) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(part == 0 ? p.a_hi : p.a_lo);
        uint8_t* dst = Abase + (size_t)buf * a_buf + (size_t)part * a_part;
        for (int kc = 0; kc < KC; ++kc)
          tc::bulk_g2s(dst + (size_t)kc * run, src + (((long long)b * KC + kc) * p.Lp + row) * 16, run, &a_full[buf]);
      }
      if (!resident) {
        for (int i = 0; i < nch; ++i, ++wi) {
          const int st = wi & 1;
          if (wi >= 2) tc::mbar_wait(&w_empty[st], (uint32_t)(((wi >> 1) - 1) & 1));
          tc::mbar_arrive_expect_tx(&w_full[st], wb);
          tc::bulk_g2s(Wbase + (size_t)st * wb, p.wpk + ((size_t)(i / parts) * 2 + (i % parts)) * wb, wb, &w_full[st]);
        }
      }
    }
  } else if (tid == 288) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = tc::idesc_bf16(TC_M, p.Cp);
    const uint32_t a0 = tc::smem_u32(Abase), w0 = tc::smem_u32(Wbase);
    const uint32_t lbo_a = (uint32_t)R * 16u, lbo_b = (uint32_t)p.Cp * 16u;
    if (resident) tc::mbar_wait(&w_res, 0);
    int it = 0, wi = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it % nabuf, acc = it & 1;
      tc::mbar_wait(&a_full[buf], (uint32_t)((it / nabuf) & 1));
      if (it >= 2) tc::mbar_wait(&t_empty[acc], (uint32_t)(((it >> 1) - 1) & 1));
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem + (uint32_t)(acc * acc_stride);
      const uint32_t a_hi = a0 + (uint32_t)buf * a_buf, a_lo = a_hi + a_part;
      // descriptors differ only in the start-address field (units of 16 B): advance by addition
      const uint64_t ad_hi0 = tc::smem_desc(a_hi, lbo_a), ad_lo0 = tc::smem_desc(a_lo, lbo_a);
      const uint32_t kstep_a = (2u * lbo_a) >> 4, kstep_b = (2u * lbo_b) >> 4;
      const int nk = p.Cp / 16;
      uint32_t accumulate = 0;
      int i = 0;
      for (int tap = 0; tap < p.K; ++tap) {
        const uint32_t tap_off = (uint32_t)(tap * p.dil);  // rows -> 16-byte units
        for (int part = 0; part < parts; ++part, ++i) {
          uint32_t wbase;
          int st = 0;
          if (resident) {
            wbase = w0 + (uint32_t)i * wb;
          } else {
            st = wi & 1;
            tc::mbar_wait(&w_full[st], (uint32_t)((wi >> 1) & 1));
            tc::fence_after_sync();
            wbase = w0 + (uint32_t)st * wb;
          }
          const uint64_t bd0 = tc::smem_desc(wbase, lbo_b);
          const int n_a = (part == 0 && parts == 2) ? 2 : 1;  // Wh meets Ah and Al; Wl meets Ah
          for (int ap = 0; ap < n_a; ++ap) {
            uint64_t ad = (ap == 0 ? ad_hi0 : ad_lo0) + tap_off;
            uint64_t bd = bd0;
            for (int kk = 0; kk < nk; ++kk) {
              tc::mma_bf16(d_tmem, ad, bd, idesc, accumulate);
              accumulate = 1;
              ad += kstep_a;
              bd += kstep_b;
            }
          }
          if (!resident) { tc::mma_commit(&w_empty[st]); ++wi; }
        }
      }
      tc::mma_commit(&a_empty[buf]);
      tc::mma_commit(&t_full[acc]);
    }
  } else if (warp < 8) {
    // ---------------------------------------------------------------- epilogue
    // Two groups of four warps (warp w reads TMEM lanes 32*(w%4)..+31) take alternate 16-column
    // strips; inside a group the residual / accumulator loads of the NEXT strip are issued before
    // the current strip is stored, so ~2 x 16 loads per thread are in flight (the epilogue is a
    // DRAM-latency pipeline: measured 1.4 us per strip when each strip waited for its own loads).
    const int grp = warp >> 2, wq = warp & 3;
    const int nstrips = p.Cp / 16;
    const bool do_div = p.out_div != 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int b = tile / tpi, t0 = (tile - b * tpi) * TC_M;
      const int t = t0 + wq * 32 + lane;
      const bool live = t < p.L;
      const long long rowb = (long long)b * p.C * p.L + t;
      auto load_adds = [&](int strip, float (&add)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = strip * 16 + j;
          float a = 0.f;
          if (live && co < p.C) {
            const long long off = rowb + (long long)co * p.L;
            if (p.res) a = p.res[off];
            if (p.accum) a += p.y[off];
          }
          add[j] = a;
        }
      };
      float cur[16], nxt[16];
      if (grp < nstrips) load_adds(grp, cur);   // independent of the accumulator: overlaps the MMA
      tc::mbar_wait(&t_full[acc], (uint32_t)((it >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * acc_stride);
      for (int strip = grp; strip < nstrips; strip += 2) {
        uint32_t v[16];
        tc::tmem_ld16(tbase + (uint32_t)(strip * 16), v);
        if (strip + 2 < nstrips) load_adds(strip + 2, nxt);
        tc::tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = strip * 16 + j;
            if (co < p.C) {
              float o = __uint_as_float(v[j]) + __ldg(p.bias + co) + cur[j];
              // a real (uniform) branch: if-converted, the division ran its x/0 slow path per element
              // on every launch without a divisor (r01 source-level capture: 30 % of all instructions)
              if (do_div) { asm volatile(""); o = o / p.out_div; }
              p.y[rowb + (long long)co * p.L] = o;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
      }
      tc::fence_before_sync();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&t_empty[acc])) : "memory");
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, ncols);
}

static AmpPlan amp_plan(int Cp, int K, int dil, int nsplit) {
  const int R = TC_M + (K - 1) * dil;
  const int parts = nsplit == 3 ? 2 : 1;
  const size_t a_buf = (size_t)(Cp / 8) * R * 16 * parts;
  const size_t wb = (size_t)Cp * Cp * 2;
  const size_t nch = (size_t)K * parts;
  const size_t limit = 227 * 1024 - 1024;
  AmpPlan pl;
  pl.acc_stride = (Cp + 31) / 32 * 32;
  pl.ncols = (int)tc_cols_host(2 * pl.acc_stride);
  // Measured (r01): resident weights + 2 A buffers (one CTA per SM) beat streaming weights with two
  // CTAs per SM on the C=40/80 stages (33.0 vs 37.1 ms per step), so residency is preferred.
  if (2 * a_buf + nch * wb + 128 <= limit) { pl.resident = 1; pl.nabuf = 2; pl.smem = 2 * a_buf + nch * wb + 128; }
  else if (2 * a_buf + 2 * wb + 128 <= limit) { pl.resident = 0; pl.nabuf = 2; pl.smem = 2 * a_buf + 2 * wb + 128; }
  else { pl.resident = 0; pl.nabuf = 1; pl.smem = a_buf + 2 * wb + 128; }
  return pl;
}

size_t amp_conv_tc_smem_bytes(int Cp, int K, int dil, int nsplit) { return amp_plan(Cp, K, dil, nsplit).smem; }

int launch_amp_conv_tc(const AmpConvParams& p, cudaStream_t s) {
  if (p.Cp % 16 || p.Cp < 16 || p.Cp > 256 || p.Cp < p.C || (p.nsplit != 1 && p.nsplit != 3)) {
    set_error("amp_conv_tc: bad channel padding / nsplit");
    return SVCB_E_BAD_SHAPE;
  }
  if (p.dil * (p.K - 1) / 2 > P8_PAD || p.Lp != p8_rows_of(p.L) || !p.a_hi || (p.nsplit == 3 && !p.a_lo)) {
    set_error("amp_conv_tc: operand image does not match (halo > 32 rows or wrong Lp)");
    return SVCB_E_BAD_SHAPE;
  }
  const AmpPlan pl = amp_plan(p.Cp, p.K, p.dil, p.nsplit);
  if (pl.smem > 227 * 1024 - 512) { set_error("amp_conv_tc: tile does not fit shared memory"); return SVCB_E_UNSUPPORTED; }
  static size_t attr_bytes = 0;  // the dynamic limit excludes the kernel's (small) static shared memory
  static int n_sm = 0;
  if (pl.smem > attr_bytes) {
    SVCB_CUDA_CHECK(cudaFuncSetAttribute(amp_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)pl.smem));
    attr_bytes = pl.smem;
  }
  if (!n_sm) {
    int dev = 0;
    SVCB_CUDA_CHECK(cudaGetDevice(&dev));
    SVCB_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  int occ = 1;
  SVCB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, amp_conv_tc_kernel, 320, pl.smem));
  occ = std::max(1, std::min(occ, 512 / pl.ncols));
  const int ntiles = p.B * ((p.L + TC_M - 1) / TC_M);
  const int grid = std::min(ntiles, n_sm * occ);
  const double macs = (double)p.B * p.L * p.C * p.C * p.K;
  char kname[64];
  snprintf(kname, sizeof(kname), "amp_conv_tc_%s_c%dk%d", p.nsplit == 3 ? "bf16x3" : "bf16", p.C, p.K);
  KernelScope ks(kname, s, 2.0 * macs,
                 (double)p.B * p.C * p.L * ((p.nsplit == 3 ? 4.0 : 2.0) + 4.0 * (p.res ? 2 : 1)));
  amp_conv_tc_kernel<<<grid, 320, pl.smem, s>>>(p, pl.resident, pl.nabuf, pl.acc_stride, (uint32_t)pl.ncols);
  SVCB_LAUNCH_CHECK("amp_conv_tc");
  return SVCB_OK;
}

}  // namespace svcb
