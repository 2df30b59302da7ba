// AMP-block link on the 5th-gen tensor cores, as two kernels that each stream at their own roofline:
//
//   snake_pack   SnakeAlias(x) -> bf16 hi/lo operand image in HBM            (CUDA cores, 8 B/element)
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
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc.cuh"

namespace svcb {

static inline unsigned tc_cols_host(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u; }

constexpr int TC_M = 128;
constexpr int P8_PAD = 32;

__host__ __device__ inline int p8_rows_of(int L) { return P8_PAD + (L + 127) / 128 * 128 + P8_PAD; }
int p8_rows(int L) { return p8_rows_of(L); }

// ------------------------------------------------------------------------------------ snake_pack
// Register-resident (same scheme as amp_block_fused's ab_snake_run; the earlier version staged x and
// the 2x-rate Snake values of an 8 x 506 tile in shared memory across three CTA barriers and was
// issue/latency-bound at 1.9 TB/s, profiles/r01_notes.md): no CTA barrier, no v buffer.  A warp owns one octet of
// channels x 32 image rows: lane = 4*channel + run, a thread computes 8 consecutive samples of one
// channel from the 24 inputs around them (six 16-byte loads), then the warp transposes its 8 x 32
// tile through 1 KB of shared memory so that every lane packs ONE image row (8 channels -> 16 bytes
// of bf16 hi and 16 of lo) and the warp stores 512 contiguous bytes per part.
constexpr int SP3_ROWS = 256;  // image rows per CTA (8 warps x 32)

template <bool VEC>
__device__ __forceinline__ void sp3_run(const float* __restrict__ xr, int n0, int L, const SnakeTapsV& tp,
                                        const float* f_up, const float* f_dn, float a_, float b_, float (&o)[8]) {
  // (the vector loads touch xr[n0-8 .. n0+16): keep them inside the row)
  if (VEC ? (n0 - 8 >= 0 && n0 + 16 <= L) : (n0 - 6 >= 0 && n0 + 13 <= L - 1)) {
    float x[24];  // xr[n0-8 .. n0+16)
    if (VEC) {
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(xr + n0 - 8) + q);
        x[4 * q] = t4.x; x[4 * q + 1] = t4.y; x[4 * q + 2] = t4.z; x[4 * q + 3] = t4.w;
      }
    } else {
      x[0] = x[1] = x[22] = x[23] = 0.f;   // never read
#pragma unroll
      for (int q = 2; q < 22; ++q) x[q] = __ldg(xr + n0 - 8 + q);
    }
    snake8_packed(x, tp, a_, 0.5f * b_, o);
  } else {  // a tap crosses a sequence end: replicate-clamped scalar path; rows outside [0, L) are zero
    const int mhi = 2 * L - 1;
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + i;
      float acc = 0.f;
      if (n >= 0 && n < L) {
        for (int k = 0; k < 12; ++k) {
          const int m = min(max(2 * n - 5 + k, 0), mhi);
          const int a = m >> 1, q = m & 1;
          float u = 0.f;
          for (int d = q; d < q + 6; ++d) u = fmaf(__ldg(xr + min(max(a - 3 + d, 0), L - 1)), f_up[11 + q - 2 * d], u);
          u *= 2.f;
          const float sn = snake_sin(u * a_);
          acc = fmaf(fmaf(b_, sn * sn, u), f_dn[k], acc);
        }
      }
      o[i] = acc;
    }
  }
}

template <bool VEC, int MINB>
__global__ void __launch_bounds__(256, MINB)
snake_pack3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                   const float* __restrict__ ea, const float* __restrict__ inv_b,
                   const float* __restrict__ fu_g, const float* __restrict__ fd_g, const SnakeTapsV tp, int C, int L, int Lp) {
  __shared__ float tile[8][32 * 9];   // per warp: [row][channel], row stride 9 -> conflict-free both ways
  __shared__ float f_up[12], f_dn[12];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int oc = blockIdx.y, b = blockIdx.z;
  if (tid < 12) { f_up[tid] = __ldg(fu_g + tid); f_dn[tid] = __ldg(fd_g + tid); }
  __syncthreads();
  const int row0 = blockIdx.x * SP3_ROWS + warp * 32;   // first image row of this warp
  if (row0 >= Lp) return;                               // warp-uniform
  const int c = lane >> 2, run = lane & 3;
  const int ch = oc * 8 + c;
  const int n0 = row0 - P8_PAD + 8 * run;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  if (ch < C && n0 < L && n0 + 8 > 0) {
    sp3_run<VEC>(x + ((long long)b * C + ch) * L, n0, L, tp, f_up, f_dn, __ldg(ea + ch), __ldg(inv_b + ch), o);
  }
  float* tw = tile[warp];
#pragma unroll
  for (int i = 0; i < 8; ++i) tw[(8 * run + i) * 9 + c] = o[i];
  __syncwarp();
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = tw[lane * 9 + k];
  __align__(16) __nv_bfloat162 h2[4], l2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    h2[k] = __floats2bfloat162_rn(r[2 * k], r[2 * k + 1]);
    const float2 f = __bfloat1622float2(h2[k]);
    l2[k] = __floats2bfloat162_rn(r[2 * k] - f.x, r[2 * k + 1] - f.y);
  }
  const long long img = ((long long)b * gridDim.y + oc) * Lp + row0 + lane;
  *reinterpret_cast<uint4*>(hi + img * 8) = *reinterpret_cast<const uint4*>(h2);
  if (lo) *reinterpret_cast<uint4*>(lo + img * 8) = *reinterpret_cast<const uint4*>(l2);
}

int snake_taps_from_device(const float* fu_dev, const float* fd_dev, SnakeTapsV* out) {
  float h[24];
  SVCB_CUDA_CHECK(cudaMemcpy(h, fu_dev, 12 * sizeof(float), cudaMemcpyDeviceToHost));
  SVCB_CUDA_CHECK(cudaMemcpy(h + 12, fd_dev, 12 * sizeof(float), cudaMemcpyDeviceToHost));
  *out = snake_taps_pack(h, h + 12);
  return SVCB_OK;
}

size_t p8_image_bytes(int B, int C, int L) {  // one of hi / lo
  const int cp = (C + 15) / 16 * 16;
  return (size_t)B * (cp / 8) * p8_rows_of(L) * 16;
}

int launch_snake_pack(const float* x, void* hi, void* lo, const float* ea, const float* inv_b, const float* fu,
                      const float* fd, int B, int C, int L, cudaStream_t s, const SnakeTapsV* taps) {
  if (B <= 0 || C <= 0 || L <= 0) return SVCB_OK;
  SnakeTapsV tp;
  if (taps) tp = *taps;
  else SVCB_TRY(snake_taps_from_device(fu, fd, &tp));
  const int cp = (C + 15) / 16 * 16, Lp = p8_rows_of(L);
  char kname[64];
  snprintf(kname, sizeof(kname), "snake_pack_c%d", C);
  KernelScope ks(kname, s, 0.0, (lo ? 8.0 : 6.0) * B * C * (double)L, 70.0 * B * C * (double)L);
  {
    dim3 grid3((Lp + SP3_ROWS - 1) / SP3_ROWS, cp / 8, B);
    auto* h = static_cast<__nv_bfloat16*>(hi);
    auto* l = static_cast<__nv_bfloat16*>(lo);
    const bool vec = (L & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    // 64 registers -> four CTAs (32 warps) per SM: each warp lives for one tile, so residency is
    // what hides its initial load latency (10.8 vs 11.4 ms/step with two CTAs)
    if (vec) snake_pack3_kernel<true, 4><<<grid3, 256, 0, s>>>(x, h, l, ea, inv_b, fu, fd, tp, C, L, Lp);
    else snake_pack3_kernel<false, 4><<<grid3, 256, 0, s>>>(x, h, l, ea, inv_b, fu, fd, tp, C, L, Lp);
    SVCB_LAUNCH_CHECK("snake_pack");
  }
  return SVCB_OK;
}

// ------------------------------------------------------------------------------------ amp_conv_tc
// Persistent: each CTA walks tiles (item, 128 samples) with a static stride.  Three roles pipeline
// across tiles through mbarriers:
//   producer thread  A image rows of tile i+1 (bulk copies, 1-2 buffers) and the weight tiles
//                    (all taps resident in shared memory when they fit, else a 2-slot ring per tile)
//   MMA warp         all taps x split parts of tile i into TMEM accumulator (i & 1); one elected lane
//                    issues (tc::elect_one) inside warp-uniform control flow
//   8 epilogue warps tile i-1: tcgen05.ld -> +bias (+res, +stage accumulation, /3) -> coalesced stores
struct AmpPlan { int resident, nabuf, acc_stride, ncols, ncat, nw; size_t smem; };
constexpr int AMP_WMAX = 8;   // deepest weight ring

// NK MMAs of one (tap, split part, A part) group: the operands' descriptor low words advance by
// kk * kstep with kk a compile-time constant.
template <int NK>
__device__ __forceinline__ void amp_issue(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t kstep_a, uint32_t kstep_b, uint32_t first_acc) {
  tc::mma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, first_acc);
#pragma unroll
  for (int kk = 1; kk < NK; ++kk) tc::mma_bf16_lohi(d_tmem, a_lo + kk * kstep_a, a_hi, b_lo + kk * kstep_b, b_hi, idesc, 1u);
}

__global__ void __launch_bounds__(320, 1)
amp_conv_tc_kernel(const AmpConvParams p, const int resident, const int nabuf, const int nw, const int acc_stride,
                   const uint32_t ncols, const int ncat) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2], w_full[AMP_WMAX], w_empty[AMP_WMAX], w_res, t_full[2], t_empty[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = tc::warp_uniform_idx();
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
      tc::mbar_init(&t_full[i], 1); tc::mbar_init(&t_empty[i], 256);
    }
    for (int i = 0; i < AMP_WMAX; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
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
    // ncat: the hi and lo weight tiles of a tap are interleaved per K-chunk ([kc][hi rows | lo rows][8])
    // so that ONE descriptor with N = 2*Cp multiplies A_hi by [W_hi | W_lo]; the hi-only view of the
    // same bytes (N = Cp, same stride) serves A_lo * W_hi.  The packed blob keeps its layout: the
    // interleave happens here, one bulk copy per (tap, part, K-chunk).
    const uint32_t wrow = (uint32_t)p.Cp * 16u;   // one K-chunk of one part
    auto load_tap_cat = [&](uint8_t* dst, int tap, uint64_t* bar) {
      for (int part = 0; part < 2; ++part)
        for (int kc = 0; kc < KC; ++kc)
          tc::bulk_g2s(dst + (size_t)kc * 2 * wrow + (size_t)part * wrow,
                       p.wpk + ((size_t)tap * 2 + part) * wb + (size_t)kc * wrow, wrow, bar);
    };
    if (resident) {
      tc::mbar_arrive_expect_tx(&w_res, wb * (uint32_t)nch);
      if (ncat) {
        for (int tap = 0; tap < p.K; ++tap) load_tap_cat(Wbase + (size_t)tap * 2 * wb, tap, &w_res);
      } else {
        for (int i = 0; i < nch; ++i)
          tc::bulk_g2s(Wbase + (size_t)i * wb, p.wpk + ((size_t)(i / parts) * 2 + (i % parts)) * wb, wb, &w_res);
      }
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
      if (!resident && ncat) {
        for (int tap = 0; tap < p.K; ++tap, ++wi) {
          const int st = wi % nw;
          if (wi >= nw) tc::mbar_wait(&w_empty[st], (uint32_t)(((wi / nw) - 1) & 1));
          tc::mbar_arrive_expect_tx(&w_full[st], 2 * wb);
          load_tap_cat(Wbase + (size_t)st * 2 * wb, tap, &w_full[st]);
        }
      } else if (!resident) {
        for (int i = 0; i < nch; ++i, ++wi) {
          const int st = wi % nw;
          if (wi >= nw) tc::mbar_wait(&w_empty[st], (uint32_t)(((wi / nw) - 1) & 1));
          tc::mbar_arrive_expect_tx(&w_full[st], wb);
          tc::bulk_g2s(Wbase + (size_t)st * wb, p.wpk + ((size_t)(i / parts) * 2 + (i % parts)) * wb, wb, &w_full[st]);
        }
      }
    }
  } else if (warp_u == 9) {
    // ---------------------------------------------------------------- MMA issuer (whole warp, uniform
    // control flow; one elected lane issues — see tc::elect_one)
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
      if (ncat) {
        // two MMA groups per tap instead of three: A_hi x [W_hi | W_lo] (N = 2*Cp, columns [0,Cp) and
        // [Cp,2Cp) of the accumulator) and A_lo x W_hi (N = Cp, into columns [0,Cp)); the epilogue adds
        // the two column halves.  An MMA costs ~110 cycles here whatever its N (r01 captures), so the
        // instruction count is what matters.
        const uint32_t idesc_cat = tc::idesc_bf16(TC_M, 2 * p.Cp);
        const uint32_t lbo_c = 2u * lbo_b, kstep_c = (2u * lbo_c) >> 4;
        for (int tap = 0; tap < p.K; ++tap) {
          const uint32_t tap_off = (uint32_t)(tap * p.dil);
          uint32_t wbase;
          int st = 0;
          if (resident) {
            wbase = w0 + (uint32_t)tap * 2u * wb;
          } else {
            st = wi % nw;
            tc::mbar_wait(&w_full[st], (uint32_t)((wi / nw) & 1));
            tc::fence_after_sync();
            wbase = w0 + (uint32_t)st * 2u * wb;
          }
          const uint64_t bd0 = tc::smem_desc(wbase, lbo_c);
          if (tc::elect_one()) {
            uint32_t acc_flag = accumulate;
            const uint32_t a_hiw = (uint32_t)(ad_hi0 >> 32), b_hiw = (uint32_t)(bd0 >> 32);
            uint32_t ad = (uint32_t)ad_hi0 + tap_off, bd = (uint32_t)bd0;
            for (int kk = 0; kk < nk; ++kk) {
              tc::mma_bf16_lohi(d_tmem, ad, a_hiw, bd, b_hiw, idesc_cat, acc_flag);
              acc_flag = 1;
              ad += kstep_a;
              bd += kstep_c;
            }
            ad = (uint32_t)ad_lo0 + tap_off; bd = (uint32_t)bd0;
            for (int kk = 0; kk < nk; ++kk) {
              tc::mma_bf16_lohi(d_tmem, ad, a_hiw, bd, b_hiw, idesc, 1u);
              ad += kstep_a;
              bd += kstep_c;
            }
            if (!resident) tc::mma_commit(&w_empty[st]);
          }
          accumulate = 1;
          if (!resident) ++wi;
        }
      }
      for (int tap = 0; tap < (ncat ? 0 : p.K); ++tap) {
        const uint32_t tap_off = (uint32_t)(tap * p.dil);  // rows -> 16-byte units
        for (int part = 0; part < parts; ++part, ++i) {
          uint32_t wbase;
          int st = 0;
          if (resident) {
            wbase = w0 + (uint32_t)i * wb;
          } else {
            st = wi % nw;
            tc::mbar_wait(&w_full[st], (uint32_t)((wi / nw) & 1));
            tc::fence_after_sync();
            wbase = w0 + (uint32_t)st * wb;
          }
          const uint64_t bd0 = tc::smem_desc(wbase, lbo_b);
          const int n_a = (part == 0 && parts == 2) ? 2 : 1;  // Wh meets Ah and Al; Wl meets Ah
          // descriptor low words of this group, computed in uniform code; the unrolled issue below adds
          // compile-time multiples of the K steps only (the per-MMA `ad += kstep` of round 1 lived in a vector
          // register inside the elected branch: IMAD + R2UR per operand per MMA, ~110 cycles per MMA where
          // the instruction itself needs 52 at N = 80 — profiles/r02_mma_probe.txt)
          const uint32_t a_hiw = (uint32_t)(ad_hi0 >> 32), b_hiw = (uint32_t)(bd0 >> 32);
          const uint32_t a0w = (uint32_t)ad_hi0 + tap_off, a1w = (uint32_t)ad_lo0 + tap_off, bw = (uint32_t)bd0;
          if (tc::elect_one()) {
            if (nk == 5) {
              amp_issue<5>(d_tmem, a0w, a_hiw, bw, b_hiw, idesc, kstep_a, kstep_b, accumulate);
              if (n_a == 2) amp_issue<5>(d_tmem, a1w, a_hiw, bw, b_hiw, idesc, kstep_a, kstep_b, 1u);
            } else if (nk == 10) {
              amp_issue<10>(d_tmem, a0w, a_hiw, bw, b_hiw, idesc, kstep_a, kstep_b, accumulate);
              if (n_a == 2) amp_issue<10>(d_tmem, a1w, a_hiw, bw, b_hiw, idesc, kstep_a, kstep_b, 1u);
            } else {
              uint32_t acc_flag = accumulate;
              for (int ap = 0; ap < n_a; ++ap) {
                uint32_t ad = ap == 0 ? a0w : a1w, bd = bw;
                for (int kk = 0; kk < nk; ++kk) {
                  tc::mma_bf16_lohi(d_tmem, ad, a_hiw, bd, b_hiw, idesc, acc_flag);
                  acc_flag = 1;
                  ad += kstep_a;
                  bd += kstep_b;
                }
              }
            }
            if (!resident) tc::mma_commit(&w_empty[st]);
          }
          accumulate = 1;
          if (!resident) ++wi;
        }
      }
      if (tc::elect_one()) {
        tc::mma_commit(&a_empty[buf]);
        tc::mma_commit(&t_full[acc]);
      }
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
        uint32_t v[16], v2[16];
        tc::tmem_ld16(tbase + (uint32_t)(strip * 16), v);
        if (ncat) tc::tmem_ld16(tbase + (uint32_t)(p.Cp + strip * 16), v2);   // the A_hi x W_lo half
        if (strip + 2 < nstrips) load_adds(strip + 2, nxt);
        tc::tmem_ld_wait();
        if (ncat) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        }
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
  // split operands, narrow tiles: W_hi | W_lo side by side along N (2 MMA groups per tap, not 3).
  // Measured: C=40 (N 48 -> 96) k=11 0.98 -> 0.61 ms per launch, k=7 0.70 -> 0.52; C=80 (N 80 -> 160)
  // got slower (k=11 0.41 -> 0.53 ms: an N=160 MMA costs twice an N=80 one and the ring needs 20
  // small weight copies per tap), so the concatenation is used up to N = 128 only.
  pl.ncat = (nsplit == 3 && 2 * Cp <= 128) ? 1 : 0;
  const size_t wslot = pl.ncat ? 2 * wb : wb;   // one ring slot (ncat: both parts of a tap)
  pl.acc_stride = ((pl.ncat ? 2 * Cp : Cp) + 31) / 32 * 32;
  pl.ncols = (int)tc_cols_host(2 * pl.acc_stride);
  // Measured (r01): resident weights + 2 A buffers (one CTA per SM) beat streaming weights with two
  // CTAs per SM on the C=40/80 stages (33.0 vs 37.1 ms per step), so residency is preferred.
  pl.nw = 2;
  if (2 * a_buf + nch * wb + 128 <= limit) { pl.resident = 1; pl.nabuf = 2; pl.smem = 2 * a_buf + nch * wb + 128; }
  else {
    // streaming weights.  A deeper ring (up to AMP_WMAX slots) was measured for C = 80 (8 x 12.8 KB instead of
    // 2): no change — these launches are not waiting for weights (nor for the issue loop: the unrolled
    // amp_issue<> path changed nothing either); their 23.5k cycles per tile against 8.6k of MMAs sit in the
    // epilogue's residual-load / store round trips.  Two slots are kept.
    pl.resident = 0;
    pl.nabuf = (2 * a_buf + 2 * wslot + 128 <= limit) ? 2 : 1;
    pl.nw = 2;
    pl.smem = pl.nabuf * a_buf + pl.nw * wslot + 128;
  }
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
  static DevSmemCache attr_cache;  // the dynamic limit excludes the kernel's (small) static shared memory
  SVCB_CUDA_CHECK(ensure_dyn_smem(amp_conv_tc_kernel, pl.smem, attr_cache));
  const int n_sm = device_sm_count();
  if (n_sm <= 0) { set_error("amp_conv_tc: cannot query the SM count"); return SVCB_E_CUDA; }
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
  amp_conv_tc_kernel<<<grid, 320, pl.smem, s>>>(p, pl.resident, pl.nabuf, pl.nw, pl.acc_stride, (uint32_t)pl.ncols, pl.ncat);
  SVCB_LAUNCH_CHECK("amp_conv_tc");
  return SVCB_OK;
}

}  // namespace svcb
