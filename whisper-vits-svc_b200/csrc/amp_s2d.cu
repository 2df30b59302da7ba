// AMP-block links of the narrow generator stages on the tensor cores, in space-to-depth form.
//
// One launch = one `Conv1d(C->C, k, dilation) [+ x] -> SnakeAlias` link of AMPBlock.forward
// (vits_decoder/bigv.py:50-58; SnakeAlias = vits_decoder/alias/act.py:124-128), SURVEY.md §8a rows a9/a10.
//
// Why space-to-depth.  A tcgen05.mma (SS form, M = 128, K = 16) costs max(N/2, 32 + N/4) cycles on B200
// (profiles/r02_mma_probe.txt): at N = C = 16..32 the tensor pipe idles on the A-operand read, which is
// why round 1 ran C = 10 / 20 on the fp32 FMA pipe (26 % of ITS roof).  Folding r consecutive samples
// into the channel dimension (C * r = 160: r = 8 for C = 20, 16 for C = 10) turns the dilated conv into
// `ntaps` dense 160 x 160 block-Toeplitz products over rows of r samples (pack.py:conv_s2d_matrices):
// 2-8x more MACs, all of them at the full-rate N = 160 shape (80 cycles per MMA = 8192 FLOP/cycle/SM).
//
// Why the Snake lives in the epilogue.  In this layout an accumulator row holds r CONSECUTIVE samples of
// each channel: the epilogue thread that owns TMEM lane tau has, per channel, exactly the register-resident
// run of samples the SnakeAlias code of round 1 works on.  It adds bias (+ residual), parks the run in a
// 4 KB shared strip so that neighbouring rows are visible, and computes the anti-aliased Snake of the
// NEXT link straight into that link's bf16 hi/lo operand image — no snake_pack pass, no fp32 round trip:
// 8-12 B of HBM traffic per element and link (was 20-24), and the CUDA-core work overlaps the MMAs of
// the next tile (two TMEM accumulators).
//
// Data layout ("S2D image"): bf16 hi and lo, [B][20 octets][Rp][8]; element (octet o, row, e) is
// snake(x)[b][c][r*(row - 16) + p] with 8*o + e = c*r + p.  Rows outside the sequence are zero (the
// conv's zero padding; the buffers are cleared once per stage and only valid rows are ever written).
// A tile = 128 consecutive rows of every octet = 20 bulk copies per split part, already in the K-major
// SWIZZLE_NONE panel layout of tc.cuh, and a Toeplitz tap is a row-shifted descriptor.
#include <algorithm>
#include <cstdint>
#include <cstdio>

#include "common.cuh"
#include "tc.cuh"

namespace svcb {

namespace s2d {
constexpr int N = 160;            // K' = N' = C * r
constexpr int KC = N / 8;         // octets
constexpr int RA = 144;           // rows of the A panel held in shared memory (128 + room for +-8 tap rows)
constexpr int A_OFF = 8;          // A-panel row of the tile's first output row
constexpr int PADR = 16;          // zero rows in front of every (item, octet) of an image
constexpr int TILE = 128;         // accumulator rows per tile
constexpr uint32_t A_PART = KC * RA * 16;       // 46,080 B
constexpr uint32_t W_SLOT = KC * N * 16;        // 51,200 B: one (tap, hi|lo) matrix
constexpr uint32_t STG_BYTES = 32768;           // Snake staging strips (4 groups)
constexpr uint32_t SMEM = 2 * A_PART + 2 * W_SLOT + STG_BYTES;
constexpr int EPI_WARPS = 16;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr uint32_t ACC_STRIDE = 256;            // TMEM columns between the two accumulators
}  // namespace s2d

int s2d_halo_rows(int r) { return r >= 8 ? 1 : 2; }                 // SnakeAlias reaches +-5 samples
int s2d_tile_stride(int r) { return s2d::TILE - 2 * s2d_halo_rows(r); }
int s2d_rows(int L, int r) {                                         // Rp of an image
  const int nrows = L / r, S = s2d_tile_stride(r);
  return ((nrows + S - 1) / S) * S + 160;
}
size_t s2d_image_bytes(int B, int L, int r) { return (size_t)B * s2d::KC * s2d_rows(L, r) * 16; }

// ------------------------------------------------------------------------------------------------ Snake
// 8 outputs n0 .. n0+7 from the 24 inputs xw[0..24) = x[n0-8 .. n0+16)  (same arithmetic as
// amp_conv_tc.cu:sp3_run; alias/resample.py:25-33, alias/act.py:79-92, alias/filter.py:86-94)
__device__ __forceinline__ void s2d_snake8(const float (&x)[24], const float (&fu)[12], const float (&fdn)[12],
                                           float a_, float b_, float (&o)[8]) {
  float vv[28];
#pragma unroll
  for (int p = 0; p < 14; ++p) {
    float ue = x[p + 2] * fu[11];
    ue = fmaf(x[p + 3], fu[9], ue); ue = fmaf(x[p + 4], fu[7], ue); ue = fmaf(x[p + 5], fu[5], ue);
    ue = fmaf(x[p + 6], fu[3], ue); ue = fmaf(x[p + 7], fu[1], ue);
    float uo = x[p + 3] * fu[10];
    uo = fmaf(x[p + 4], fu[8], uo); uo = fmaf(x[p + 5], fu[6], uo); uo = fmaf(x[p + 6], fu[4], uo);
    uo = fmaf(x[p + 7], fu[2], uo); uo = fmaf(x[p + 8], fu[0], uo);
    const float se = snake_sin(ue * a_), so = snake_sin(uo * a_);   // fu carries UpSample1d's x2 gain
    vv[2 * p] = fmaf(b_, se * se, ue);
    vv[2 * p + 1] = fmaf(b_, so * so, uo);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc = fmaf(vv[2 * i + 1 + k], fdn[k], acc);
    o[i] = acc;
  }
}

// One output sample n where a tap crosses a sequence end: replicate-clamped indices (the reference pads x
// by replication before the transposed conv and the 2x signal before the decimating conv).
// `xs(s)` returns sample s of the channel (0 <= s < L).  Rare path (first / last rows of an item).
template <typename F>
__device__ __noinline__ float s2d_snake1_edge(F xs, int n, int L, const float* f_up, const float* f_dn, float a_, float b_) {
  const int mhi = 2 * L - 1;
  float acc = 0.f;
  for (int k = 0; k < 12; ++k) {
    const int m = min(max(2 * n - 5 + k, 0), mhi);
    const int a = m >> 1, q = m & 1;
    float u = 0.f;
    for (int d = q; d < q + 6; ++d) u = fmaf(xs(min(max(a - 3 + d, 0), L - 1)), f_up[11 + q - 2 * d], u);
    u *= 2.f;
    const float sn = snake_sin(u * a_);
    acc = fmaf(fmaf(b_, sn * sn, u), f_dn[k], acc);
  }
  return acc;
}

__device__ __forceinline__ void s2d_store_octet(__nv_bfloat16* hi, __nv_bfloat16* lo, long long row_elem,
                                                const float (&o)[8]) {
  __align__(16) __nv_bfloat162 h2[4], l2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    h2[k] = __floats2bfloat162_rn(o[2 * k], o[2 * k + 1]);
    const float2 f = __bfloat1622float2(h2[k]);
    l2[k] = __floats2bfloat162_rn(o[2 * k] - f.x, o[2 * k + 1] - f.y);
  }
  *reinterpret_cast<uint4*>(hi + row_elem) = *reinterpret_cast<const uint4*>(h2);
  *reinterpret_cast<uint4*>(lo + row_elem) = *reinterpret_cast<const uint4*>(l2);
}

// ------------------------------------------------------------------------------------------------ pack
// SnakeAlias(x[B,C,L]) -> S2D image (the first activation of every AMP block: its input is the stage
// input, not a convolution result).  One thread = one run of 8 samples = one image row of one octet.
__global__ void __launch_bounds__(256, 4)
snake_pack_s2d_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                      const float* __restrict__ ea, const float* __restrict__ inv_b,
                      const float* __restrict__ fu_g, const float* __restrict__ fd_g, int C, int L, int r, int Rp) {
  __shared__ float f_up[12], f_dn[12];
  if (threadIdx.x < 12) { f_up[threadIdx.x] = __ldg(fu_g + threadIdx.x); f_dn[threadIdx.x] = __ldg(fd_g + threadIdx.x); }
  __syncthreads();
  const int run = blockIdx.x * 256 + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (run * 8 >= L) return;
  const int opr = r >> 3;                       // octets per row of this channel
  const int row = run / opr, o = c * opr + run % opr;
  const float* xr = x + ((long long)b * C + c) * L;
  const int n0 = run * 8;
  const float a_ = __ldg(ea + c), b_ = __ldg(inv_b + c);
  float out[8];
  if (n0 - 8 >= 0 && n0 + 16 <= L) {
    float xw[24], fu[12], fdn[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { fu[k] = 2.f * f_up[k]; fdn[k] = f_dn[k]; }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(xr + n0 - 8) + q);
      xw[4 * q] = t4.x; xw[4 * q + 1] = t4.y; xw[4 * q + 2] = t4.z; xw[4 * q + 3] = t4.w;
    }
    s2d_snake8(xw, fu, fdn, a_, b_, out);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = s2d_snake1_edge([&](int s) { return __ldg(xr + s); }, n0 + i, L, f_up, f_dn, a_, b_);
  }
  const long long row_elem = ((((long long)b * s2d::KC + o) * Rp) + s2d::PADR + row) * 8;
  s2d_store_octet(hi, lo, row_elem, out);
}

int launch_snake_pack_s2d(const float* x, void* hi, void* lo, const float* ea, const float* inv_b, const float* fu,
                          const float* fd, int B, int C, int L, cudaStream_t s) {
  const int r = C > 0 ? s2d::N / C : 0;
  if (B <= 0 || L <= 0 || C * r != s2d::N || (r != 8 && r != 16) || L % r || (reinterpret_cast<uintptr_t>(x) & 15)) {
    set_error("snake_pack_s2d: unsupported shape");
    return SVCB_E_BAD_SHAPE;
  }
  char kname[64];
  snprintf(kname, sizeof(kname), "snake_pack_s2d_c%d", C);
  KernelScope ks(kname, s, 0.0, 8.0 * B * C * (double)L, 70.0 * B * C * (double)L);
  dim3 grid((L / 8 + 255) / 256, C, B);
  snake_pack_s2d_kernel<<<grid, 256, 0, s>>>(x, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), ea,
                                             inv_b, fu, fd, C, L, r, s2d_rows(L, r));
  SVCB_LAUNCH_CHECK("snake_pack_s2d");
  return SVCB_OK;
}

// S2D image (hi + lo) -> fp32 [B, C, L]: the unit tests read a link's output image through this.
__global__ void s2d_unpack_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                  float* __restrict__ y, int C, int L, int r, int Rp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over C * L of item blockIdx.y
  if (i >= (long long)C * L) return;
  const int c = (int)(i / L), t = (int)(i % L);
  const int k = c * r + t % r, row = t / r;
  const long long e = ((((long long)blockIdx.y * s2d::KC + k / 8) * Rp) + s2d::PADR + row) * 8 + k % 8;
  y[(long long)blockIdx.y * C * L + i] = __bfloat162float(hi[e]) + __bfloat162float(lo[e]);
}
int launch_s2d_unpack(const void* hi, const void* lo, float* y, int B, int C, int L, cudaStream_t s) {
  const int r = s2d::N / C;
  dim3 grid((unsigned)(((long long)C * L + 255) / 256), B);
  s2d_unpack_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(hi), static_cast<const __nv_bfloat16*>(lo), y, C, L, r,
                                         s2d_rows(L, r));
  SVCB_LAUNCH_CHECK("s2d_unpack");
  return SVCB_OK;
}

// ------------------------------------------------------------------------------------------------ link
// Persistent: every CTA walks tiles (item, 126 useful rows) with a static stride.
//   producer warp  A panel of the tile (40 bulk copies) and the (tap, hi|lo) weight matrices through a
//                  2-slot ring (one 51,200-byte bulk copy each)
//   MMA warp       per tap: A_hi x W_hi, A_lo x W_hi, A_hi x W_lo — 30 MMAs (N = 160) with compile-time
//                  descriptor offsets (no per-MMA integer work in the issuing thread)
//   16 epilogue warps = 4 column groups x 4 TMEM lane quadrants: group g owns channels c = g (mod 4)
template <int R>
__global__ void __launch_bounds__(s2d::THREADS, 1)
amp_s2d_link_kernel(const AmpS2dParams p) {
  using namespace s2d;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t a_full, a_empty, w_full[2], w_empty[2], t_full[2], t_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_fu[12], s_fd[12];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = tc::warp_uniform_idx();
  constexpr int HS = R >= 8 ? 1 : 2;
  constexpr int S = TILE - 2 * HS;
  const int nrows = p.L / R;
  const int tpi = (nrows + S - 1) / S;
  const int ntiles = p.B * tpi;
  uint8_t* Abase = smem;
  uint8_t* Wbase = smem + 2 * A_PART;
  float* Stg = reinterpret_cast<float*>(smem + 2 * A_PART + 2 * W_SLOT);

  if (tid == 0) {
    tc::mbar_init(&a_full, 1); tc::mbar_init(&a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1);
      tc::mbar_init(&t_full[i], 1); tc::mbar_init(&t_empty[i], EPI_WARPS * 32);
    }
    tc::fence_barrier_init();
  }
  if (tid < 12 && p.o_hi) { s_fu[tid] = __ldg(p.fu + tid); s_fd[tid] = __ldg(p.fd + tid); }
  __syncwarp();
  if (warp == EPI_WARPS) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp_u == EPI_WARPS) {
    // ------------------------------------------------------------------------------------ producer
    if (tc::elect_one()) {
      const uint8_t* img[2] = {reinterpret_cast<const uint8_t*>(p.a_hi), reinterpret_cast<const uint8_t*>(p.a_lo)};
      int it = 0, wi = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = tile / tpi, t = tile - b * tpi;
        const long long row0 = (long long)t * S - HS - A_OFF + PADR;     // image row of A-panel row 0 (>= 6)
        if (it >= 1) tc::mbar_wait(&a_empty, (uint32_t)((it - 1) & 1));
        tc::mbar_arrive_expect_tx(&a_full, 2 * A_PART);
        for (int part = 0; part < 2; ++part)
          for (int kc = 0; kc < KC; ++kc)
            tc::bulk_g2s(Abase + (size_t)part * A_PART + (size_t)kc * RA * 16,
                         img[part] + ((((long long)b * KC + kc) * p.Rp) + row0) * 16, RA * 16, &a_full);
        for (int c = 0; c < 2 * p.ntaps; ++c, ++wi) {
          const int st = wi & 1;
          if (wi >= 2) tc::mbar_wait(&w_empty[st], (uint32_t)(((wi >> 1) - 1) & 1));
          tc::mbar_arrive_expect_tx(&w_full[st], W_SLOT);
          tc::bulk_g2s(Wbase + (size_t)st * W_SLOT, p.wpk + (size_t)c * W_SLOT, W_SLOT, &w_full[st]);
        }
      }
    }
  } else if (warp_u == EPI_WARPS + 1) {
    // ------------------------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = tc::idesc_bf16(TILE, N);
    constexpr uint32_t LBO_A = RA * 16, LBO_B = N * 16;
    constexpr uint32_t KSA = (2 * LBO_A) >> 4, KSB = (2 * LBO_B) >> 4;      // descriptor step per K = 16
    const uint64_t adh = tc::smem_desc(tc::smem_u32(Abase), LBO_A), adl = tc::smem_desc(tc::smem_u32(Abase + A_PART), LBO_A);
    const uint64_t bd0 = tc::smem_desc(tc::smem_u32(Wbase), LBO_B), bd1 = tc::smem_desc(tc::smem_u32(Wbase + W_SLOT), LBO_B);
    const uint32_t d_hi = (uint32_t)(adh >> 32);          // identical high words (SBO, version) for A and B
    int it = 0, wi = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      tc::mbar_wait(&a_full, (uint32_t)(it & 1));
      if (it >= 2) tc::mbar_wait(&t_empty[acc], (uint32_t)(((it >> 1) - 1) & 1));
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem + (uint32_t)acc * ACC_STRIDE;
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const uint32_t arow = (uint32_t)(A_OFF - p.mlo + tap);            // row shift = 16-byte units
        const uint32_t a_h = (uint32_t)adh + arow, a_l = (uint32_t)adl + arow;
        {  // W_hi of this tap: A_hi x W_hi, A_lo x W_hi
          const int st = wi & 1;
          tc::mbar_wait(&w_full[st], (uint32_t)((wi >> 1) & 1));
          tc::fence_after_sync();
          const uint32_t bw = (uint32_t)(st ? bd1 : bd0);
          if (tc::elect_one()) {
            tc::mma_bf16_lohi(d_tmem, a_h, d_hi, bw, d_hi, idesc, tap > 0 ? 1u : 0u);
#pragma unroll
            for (int kk = 1; kk < N / 16; ++kk) tc::mma_bf16_lohi(d_tmem, a_h + kk * KSA, d_hi, bw + kk * KSB, d_hi, idesc, 1u);
#pragma unroll
            for (int kk = 0; kk < N / 16; ++kk) tc::mma_bf16_lohi(d_tmem, a_l + kk * KSA, d_hi, bw + kk * KSB, d_hi, idesc, 1u);
            tc::mma_commit(&w_empty[st]);
          }
          ++wi;
        }
        {  // W_lo of this tap: A_hi x W_lo
          const int st = wi & 1;
          tc::mbar_wait(&w_full[st], (uint32_t)((wi >> 1) & 1));
          tc::fence_after_sync();
          const uint32_t bw = (uint32_t)(st ? bd1 : bd0);
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < N / 16; ++kk) tc::mma_bf16_lohi(d_tmem, a_h + kk * KSA, d_hi, bw + kk * KSB, d_hi, idesc, 1u);
            tc::mma_commit(&w_empty[st]);
          }
          ++wi;
        }
      }
      if (tc::elect_one()) {
        tc::mma_commit(&a_empty);
        tc::mma_commit(&t_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------ epilogue
    const int q = warp & 3, g = warp >> 2;
    const int row = q * 32 + lane;
    constexpr int NBUF = R == 8 ? 2 : 1;
    constexpr int STRIP = TILE * R;                              // floats per channel strip
    float* stg_g = Stg + (size_t)g * (STG_BYTES / 16);
    float fu[12], fdn[12];
    if (p.o_hi) {
#pragma unroll
      for (int k = 0; k < 12; ++k) { fu[k] = 2.f * s_fu[k]; fdn[k] = s_fd[k]; }
    }
    __nv_bfloat16* o_hi = static_cast<__nv_bfloat16*>(p.o_hi);
    __nv_bfloat16* o_lo = static_cast<__nv_bfloat16*>(p.o_lo);
    const bool do_div = p.out_div != 0.f;
    int it = 0, nbar = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int b = tile / tpi, t = tile - b * tpi;
      const int tau0 = t * S - HS;
      const int tau = tau0 + row;
      const bool valid = tau >= 0 && tau < nrows;
      const bool useful = valid && row >= HS && row < TILE - HS;
      tc::mbar_wait(&t_full[acc], (uint32_t)((it >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * ACC_STRIDE;
      for (int c = g; c < p.C; c += 4) {
        float v[R];
        {
          uint32_t u[R];
          if constexpr (R == 8) tc::tmem_ld8(tbase + (uint32_t)(c * R), u);
          else tc::tmem_ld16(tbase + (uint32_t)(c * R), u);
          tc::tmem_ld_wait();
          const float bias = __ldg(p.bias + c);
#pragma unroll
          for (int j = 0; j < R; ++j) v[j] = __uint_as_float(u[j]) + bias;
        }
        const long long xoff = ((long long)b * p.C + c) * p.L + (long long)tau * R;
        if (p.res && valid) {
#pragma unroll
          for (int j = 0; j < R / 4; ++j) {
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.res + xoff) + j);
            v[4 * j] += r4.x; v[4 * j + 1] += r4.y; v[4 * j + 2] += r4.z; v[4 * j + 3] += r4.w;
          }
        }
        if (p.y && useful) {
          float4* yp = reinterpret_cast<float4*>(p.y + xoff);
#pragma unroll
          for (int j = 0; j < R / 4; ++j) {
            float4 o4 = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (p.accum) { const float4 y4 = yp[j]; o4.x += y4.x; o4.y += y4.y; o4.z += y4.z; o4.w += y4.w; }
            // a real (uniform) branch: if-converted, the division would run its x/0 slow path per element
            if (do_div) { asm volatile(""); o4.x = o4.x / p.out_div; o4.y = o4.y / p.out_div; o4.z = o4.z / p.out_div; o4.w = o4.w / p.out_div; }
            yp[j] = o4;
          }
        }
        if (p.o_hi) {
          float* strip = stg_g + (NBUF == 2 ? (nbar & 1) * STRIP : 0);
#pragma unroll
          for (int j = 0; j < R / 4; ++j)
            *reinterpret_cast<float4*>(strip + row * R + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          ++nbar;
          if (useful) {
            const float a_ = __ldg(p.ea + c), b_ = __ldg(p.ib + c);
#pragma unroll
            for (int h = 0; h < R / 8; ++h) {
              const int n0 = row * R + 8 * h;                       // strip index of the run's first sample
              const int s0 = tau * R + 8 * h;                       // its sample index in the sequence
              float o[8];
              if (s0 - 6 >= 0 && s0 + 13 <= p.L - 1) {
                float xw[24];
#pragma unroll
                for (int k4 = 0; k4 < 6; ++k4) {
                  const float4 t4 = *reinterpret_cast<const float4*>(strip + n0 - 8 + 4 * k4);
                  xw[4 * k4] = t4.x; xw[4 * k4 + 1] = t4.y; xw[4 * k4 + 2] = t4.z; xw[4 * k4 + 3] = t4.w;
                }
                s2d_snake8(xw, fu, fdn, a_, b_, o);
              } else {
                const int sh = tau0 * R;                            // sample index of strip position 0
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  o[i] = s2d_snake1_edge([&](int s) { return strip[s - sh]; }, s0 + i, p.L, s_fu, s_fd, a_, b_);
              }
              const int oc = c * (R / 8) + h;
              const long long row_elem = ((((long long)b * KC + oc) * p.Rp) + PADR + tau) * 8;
              s2d_store_octet(o_hi, o_lo, row_elem, o);
            }
          }
          if (NBUF == 1) asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        }
      }
      tc::fence_before_sync();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&t_empty[acc])) : "memory");
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == s2d::EPI_WARPS) tc::tmem_dealloc(tmem, 512);
}

int launch_amp_s2d_link(const AmpS2dParams& p, cudaStream_t s) {
  const int r = p.C > 0 ? s2d::N / p.C : 0;
  if (p.B <= 0 || p.L <= 0 || p.C * r != s2d::N || (r != 8 && r != 16) || p.L % r) {
    set_error("amp_s2d_link: unsupported shape (need C * r = 160 with r in {8, 16} and L % r == 0)");
    return SVCB_E_BAD_SHAPE;
  }
  if (p.ntaps < 1 || p.mlo < 0 || p.mlo > s2d::A_OFF || p.ntaps - 1 - p.mlo > s2d::RA - s2d::TILE - s2d::A_OFF ||
      p.Rp != s2d_rows(p.L, r) || !p.a_hi || !p.a_lo || !p.wpk || !p.bias || (p.o_hi && (!p.o_lo || !p.ea || !p.ib || !p.fu || !p.fd))) {
    set_error("amp_s2d_link: tap range exceeds the A panel, wrong image rows or missing operand");
    return SVCB_E_BAD_SHAPE;
  }
  static DevSmemCache c8, c16;
  const int n_sm = device_sm_count();
  if (n_sm <= 0) { set_error("amp_s2d_link: cannot query the SM count"); return SVCB_E_CUDA; }
  const int S = s2d_tile_stride(r);
  const int ntiles = p.B * ((p.L / r + S - 1) / S);
  const int grid = std::min(ntiles, n_sm);
  char kname[64];
  snprintf(kname, sizeof(kname), "amp_s2d_link_c%dk%dr%d", p.C, p.K, r);
  const double el = (double)p.B * p.C * p.L;
  KernelScope ks(kname, s, 2.0 * p.C * p.K * el, el * (4.0 + (p.o_hi ? 4.0 : 0.0) + (p.res ? 4.0 : 0.0) + (p.y ? (p.accum ? 8.0 : 4.0) : 0.0)),
                 p.o_hi ? 70.0 * el : 0.0);
  if (r == 8) {
    SVCB_CUDA_CHECK(ensure_dyn_smem(amp_s2d_link_kernel<8>, s2d::SMEM, c8));
    amp_s2d_link_kernel<8><<<grid, s2d::THREADS, s2d::SMEM, s>>>(p);
  } else {
    SVCB_CUDA_CHECK(ensure_dyn_smem(amp_s2d_link_kernel<16>, s2d::SMEM, c16));
    amp_s2d_link_kernel<16><<<grid, s2d::THREADS, s2d::SMEM, s>>>(p);
  }
  SVCB_LAUNCH_CHECK("amp_s2d_link");
  return SVCB_OK;
}

}  // namespace svcb
