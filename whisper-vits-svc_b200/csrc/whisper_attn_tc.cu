// Whisper encoder self-attention on the 5th-gen tensor cores (head dim 64, no mask).
//
// Replaces MultiHeadAttention.qkv_attention (whisper/model.py:88-101): q, k scaled by d^-1/4 each
// (= scores / 8), softmax in fp32, w @ v — SURVEY.md §8a row a15.  Round 1 ran this on `mma.sync`
// (HMMA) at 247 TFLOP/s, 18 of the encoder's 43.7 ms; here both products are tcgen05.mma with the
// accumulators in tensor memory.
//
// Operands come straight from the QKV GEMM, whose epilogue 4 writes the head-major layout of
// common.cuh qkv_heads_off: per (item, q|k|v, head) the positions are tiled ([.][8 octets][128 or 64 rows][8]
// bf16, items padded to 128 positions), so every operand tile is one contiguous K-major SWIZZLE_NONE panel and
// one bulk copy (a first version read the flat [M/128][3D/64] tile image with 16 copies of <= 1 KB per tile:
// the producer's issue rate, ~800 cycles per tile, bounded the kernel).
//   S = Q K^T   A = Q panel [8][128][8], B = K panel [8][64 keys][8]       (K-major, 4 MMAs of K = 16)
//   O += P V    A = P panel [8][128][8] written by the softmax warps,
//               B = V panel [8 (d octets)][64 keys][8] read as an MN-major operand (no transpose pass)
// One CTA = (item, head, 128 queries), two passes over the keys (row maximum, then p = exp(s - max), sum(p)
// and P V); 8 softmax warps (two per TMEM lane quadrant), a producer warp (one bulk copy per 128-key unit into
// a 3-slot ring), one MMA warp.  256 TMEM columns and 97 KB of shared memory, so two CTAs share an SM.  The output is written as the tile image the out-projection consumes.
#include <cstdint>

#include "common.cuh"
#include "tc.cuh"

namespace svcb {

namespace wa {
constexpr int D = 64, TQ = 128, TK = 64, TU = 128, NS = 5;
constexpr uint32_t Q_BYTES = (D / 8) * TQ * 16;    // 16,384
constexpr uint32_t U_BYTES = (D / 8) * TU * 16;    // 16,384  (one ring slot: the K or the V rows of a 128-key unit)
constexpr uint32_t OFF_Q = 0, OFF_R = OFF_Q + Q_BYTES, OFF_ML = OFF_R + NS * U_BYTES, SMEM = OFF_ML + 2 * TQ * 4;
// tensor memory: O (0..63), two 64-column score buffers (64, 128) that P (bf16, the A operand of P V) overwrites
// in place, Q as the bf16 A operand of Q K^T (192..223); pass 1 has no O yet and rotates three score buffers
constexpr uint32_t COL_O = 0, COL_S = 64, COL_Q = 192, TMEM_COLS = 256;
constexpr int SM_WARPS = 8, THREADS = (SM_WARPS + 2) * 32;
static_assert(2 * (SMEM + 1024 + 256) <= 228 * 1024, "two CTAs must share an SM");
}  // namespace wa

// kind::f16, bf16 A (K-major) x bf16 B (MN-major when b_mn), fp32 D
__host__ __device__ constexpr uint32_t wa_idesc(int M, int N, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t wa_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ float wa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Pass 1 over the keys finds each query's score maximum (S = Q K^T read from tensor memory, no exponentials;
// 128-key units, N = 128 MMAs); pass 2 recomputes S in 64-key tiles, writes p = 2^((s - m) c) as the bf16 A
// panel of O += P V and sums p in fp32; O is divided by the sum in the epilogue.  One exponential per score:
// the MUFU unit (16 / clock / SM) is this kernel's floor.
__global__ void __launch_bounds__(wa::THREADS, 2)
whisper_attn_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out_img, int T, int Dm,
                       int vswap, long long* trace) {
  using namespace wa;
  const bool tr = trace != nullptr && blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0;
#define WA_T0 const long long t0_ = tr ? clock64() : 0
#define WA_T1(slot) if (tr) acc_[slot] += clock64() - t0_
  long long acc_[6] = {0, 0, 0, 0, 0, 0};
  const long long tstart_ = tr ? clock64() : 0;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t q_full, q_tmem, r_full[NS], r_empty[NS], s1_full[3], s1_empty[3], p1_done, s_full[2], p_full[2],
      o_full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = tc::warp_uniform_idx();
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int heads = Dm / 64, Tp = qkv_heads_tp(T);
  const int nu = Tp / TU, ntile = 2 * nu;         // 128-key units, 64-key tiles
  const int m_item = b * T;                       // first flat row of the item
  if (tid == 0) {
    tc::mbar_init(&q_full, 1); tc::mbar_init(&o_full, 1); tc::mbar_init(&p1_done, SM_WARPS * 32); tc::mbar_init(&q_tmem, TQ);
    for (int i = 0; i < 3; ++i) { tc::mbar_init(&s1_full[i], 1); tc::mbar_init(&s1_empty[i], SM_WARPS * 32); }
    for (int i = 0; i < NS; ++i) { tc::mbar_init(&r_full[i], 1); tc::mbar_init(&r_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&p_full[i], SM_WARPS * 32);
    }
    tc::fence_barrier_init();
  }
  __syncwarp();
  if (warp == SM_WARPS) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp_u == SM_WARPS) {
    // ---------------------------------------------------------------------- producer: one bulk copy per unit
    // ring order: K_0 .. K_{nu-1} (pass 1), then K_0 V_0 K_1 V_1 .. (pass 2); a unit = 128 keys = 16 KB
    if (tc::elect_one()) {
      const __nv_bfloat16* qb = qkv + qkv_heads_off(b, 0, h, 0, 0, heads, Tp);
      const __nv_bfloat16* kb = qkv + qkv_heads_off(b, 1, h, 0, 0, heads, Tp);
      const __nv_bfloat16* vb = qkv + qkv_heads_off(b, 2, h, 0, 0, heads, Tp);
      tc::mbar_arrive_expect_tx(&q_full, Q_BYTES);
      tc::bulk_g2s(smem + OFF_Q, qb + (size_t)qt * (TQ * D), Q_BYTES, &q_full);
      int slot = 0;
      uint32_t ph = 0;
      const int nload = 3 * nu;
      for (int n = 0; n < nload; ++n) {
        const int r = n - nu;
        const int j = n < nu ? n : r >> 1;
        const __nv_bfloat16* src = ((n >= nu && (r & 1)) ? vb : kb) + (size_t)j * (TU * D);
        { WA_T0; if (n >= NS) tc::mbar_wait_parked(&r_empty[slot], ph ^ 1u); WA_T1(0); }
        tc::mbar_arrive_expect_tx(&r_full[slot], U_BYTES);
        tc::bulk_g2s(smem + OFF_R + (size_t)slot * U_BYTES, src, U_BYTES, &r_full[slot]);
        if (++slot == NS) { slot = 0; ph ^= 1u; }
      }
      if (tr) { trace[16] = acc_[0]; trace[17] = clock64() - tstart_; }
    }
  } else if (warp_u == SM_WARPS + 1) {
    // ------------------------------------------------------------------------------------ MMA issuer
    const uint32_t sb = tc::smem_u32(smem);
    const uint64_t dq = wa_desc(sb + OFF_Q, TQ * 16, 128);
    const uint32_t hiw = (uint32_t)(dq >> 32);     // SBO = 128 B: shared by the Q, K and P descriptors
    constexpr uint32_t KS = (2 * TQ * 16) >> 4;     // K = 16 step of a 128-row panel (Q, K unit, P)
    constexpr uint32_t id_s = wa_idesc(TQ, TK, false), id_o = wa_idesc(TQ, D, true);
    // V rows of a unit: two tiles [8 d-octets][64 keys][8 d], each the MN-major B operand of O += P V: 8 keys x
    // 16 B = one core matrix; groups of 8 keys 128 B apart (LBO), d-octets TK * 16 B apart (SBO)
    const uint32_t v_lbo = vswap ? TK * 16 : 128, v_sbo = vswap ? 128 : TK * 16;
    const uint32_t vhw = (uint32_t)(wa_desc(0, v_lbo, v_sbo) >> 32);   // the low word carries address and LBO
    tc::mbar_wait_parked(&q_tmem, 0);   // Q sits in tensor memory: the S MMAs read only K from shared memory
    tc::fence_after_sync();
    for (int i = 0; i < ntile; ++i) {   // pass 1: S1[i % 3] (64 columns) = Q K^T of keys 64 i ..
      const int j = i >> 1, sub = i & 1, slot = j % NS, buf = i % 3;
      { WA_T0; if (!sub) tc::mbar_wait_parked(&r_full[slot], (uint32_t)((j / NS) & 1)); WA_T1(0); }
      { WA_T0; if (i >= 3) tc::mbar_wait_parked(&s1_empty[buf], (uint32_t)((i / 3 - 1) & 1)); WA_T1(1); }
      tc::fence_after_sync();
      const uint32_t k_lo = (uint32_t)wa_desc(sb + OFF_R + slot * U_BYTES + sub * (TK * 16), TU * 16, 128);
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          tc::mma_bf16_ts(tmem + (uint32_t)buf * TK, tmem + COL_Q + kk * 8, k_lo + kk * KS, hiw, id_s, kk ? 1u : 0u);
        if (sub) tc::mma_commit(&r_empty[slot]);
        tc::mma_commit(&s1_full[buf]);
      }
      __syncwarp();
    }
    { WA_T0; tc::mbar_wait_parked(&p1_done, 0); WA_T1(4); }   // pass 2 reuses the score columns
    // pass 2: S_0, S_1, then per tile O += P_i V_i and S_{i+2}.  P_i (bf16) is written by the softmax warps over the
    // score columns it came from, so S_{i+2} is issued behind P_i V_i (tcgen05.mma executes in issue order).
    auto issue_s = [&](int i) {
      const int buf = i & 1, j = i >> 1;
      const int n = nu + 2 * j;
      const int slot = n % NS;
      { WA_T0; if (!buf) tc::mbar_wait_parked(&r_full[slot], (uint32_t)((n / NS) & 1)); WA_T1(0); }
      tc::fence_after_sync();
      const uint32_t k_lo = (uint32_t)wa_desc(sb + OFF_R + slot * U_BYTES + buf * (TK * 16), TU * 16, 128);
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          tc::mma_bf16_ts(tmem + COL_S + (uint32_t)buf * TK, tmem + COL_Q + kk * 8, k_lo + kk * KS, hiw, id_s, kk ? 1u : 0u);
        if (buf) tc::mma_commit(&r_empty[slot]);
        tc::mma_commit(&s_full[buf]);
      }
      __syncwarp();
    };
    issue_s(0);
    issue_s(1);
    for (int kt = 0; kt < ntile; ++kt) {
      const int pb = kt & 1, j = kt >> 1;
      const int n = nu + 2 * j + 1;
      const int slot = n % NS;
      { WA_T0; tc::mbar_wait_parked(&p_full[pb], (uint32_t)((kt >> 1) & 1)); WA_T1(2); }
      { WA_T0; if (!pb) tc::mbar_wait_parked(&r_full[slot], (uint32_t)((n / NS) & 1)); WA_T1(3); }
      tc::fence_after_sync();
      const uint32_t v_lo = (uint32_t)wa_desc(sb + OFF_R + slot * U_BYTES + pb * (TK * D * 2), v_lbo, v_sbo);
      const uint32_t p_col = tmem + COL_S + (uint32_t)pb * TK;   // keys 0-31 in columns 0-15, keys 32-63 in columns 32-47
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk)   // 16 keys per MMA: 8 columns of P, two key groups = 256 B of the V tile
          tc::mma_bf16_ts(tmem + COL_O, p_col + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), v_lo + kk * 16u, vhw, id_o,
                          (kt > 0 || kk > 0) ? 1u : 0u);
        if (pb) tc::mma_commit(&r_empty[slot]);
        if (kt == ntile - 1) tc::mma_commit(&o_full);
      }
      __syncwarp();
      if (kt + 2 < ntile) issue_s(kt + 2);
    }
    if (tr && lane == 0) { for (int j = 0; j < 5; ++j) trace[8 + j] = acc_[j]; trace[13] = clock64() - tstart_; }
  } else {
    // ------------------------------------------------------------------------------------ softmax (warps 0-7)
    const int row = tid & (TQ - 1), half = tid >> 7;
    const int ti = qt * TQ + row;                 // position inside the item
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* ml_s = reinterpret_cast<float*>(smem + OFF_ML);
    constexpr float kC = 0.125f * 1.4426950408889634f;   // d^-1/4 on q and on k, and log2(e)
    if (half == 0) {   // Q row -> tensor memory: 64 bf16 = 32 columns of this thread's lane
      tc::mbar_wait_parked(&q_full, 0);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t qv[16];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const uint4 x = *reinterpret_cast<const uint4*>(smem + OFF_Q + (size_t)((c * 4 + o) * TQ + row) * 16);
          qv[4 * o] = x.x; qv[4 * o + 1] = x.y; qv[4 * o + 2] = x.z; qv[4 * o + 3] = x.w;
        }
        tc::tmem_st16(lane_base + COL_Q + (uint32_t)(c * 16), qv);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&q_tmem)) : "memory");
    }
    float m = -INFINITY;
    for (int i = 0; i < ntile; ++i) {   // pass 1: this thread's 32 of the tile's 64 columns
      const int buf = i % 3;
      { WA_T0; tc::mbar_wait_parked(&s1_full[buf], (uint32_t)((i / 3) & 1)); WA_T1(0); }
      tc::fence_after_sync();
      const long long tc0_ = tr ? clock64() : 0;
      uint32_t u[32];
      tc::tmem_ld16(lane_base + (uint32_t)(buf * TK + half * 32), reinterpret_cast<uint32_t(&)[16]>(u[0]));
      tc::tmem_ld16(lane_base + (uint32_t)(buf * TK + half * 32 + 16), reinterpret_cast<uint32_t(&)[16]>(u[16]));
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&s1_empty[buf])) : "memory");
      const int nvalid = T - (i * TK + half * 32);   // >= 32 except at the end of the item
      float a0 = m, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
      if (nvalid >= 32) {
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4) {
          a0 = fmaxf(a0, __uint_as_float(u[jj])); a1 = fmaxf(a1, __uint_as_float(u[jj + 1]));
          a2 = fmaxf(a2, __uint_as_float(u[jj + 2])); a3 = fmaxf(a3, __uint_as_float(u[jj + 3]));
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) a0 = fmaxf(a0, jj < nvalid ? __uint_as_float(u[jj]) : -INFINITY);
      }
      m = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
      if (tr) acc_[3] += clock64() - tc0_;
    }
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&p1_done)) : "memory");
    ml_s[half * TQ + row] = m;                    // the other half of the key columns of this row
    asm volatile("bar.sync 1, 256;" ::: "memory");
    m = fmaxf(m, ml_s[(half ^ 1) * TQ + row]);
    const float mc = m * kC;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int kt = 0; kt < ntile; ++kt) {   // pass 2: this thread's 32 of the tile's 64 columns
      const int buf = kt & 1;
      { WA_T0; tc::mbar_wait_parked(&s_full[buf], (uint32_t)((kt >> 1) & 1)); WA_T1(1); }
      tc::fence_after_sync();
      const long long tc0_ = tr ? clock64() : 0;
      uint32_t u[32];
      tc::tmem_ld16(lane_base + COL_S + (uint32_t)(buf * TK + half * 32), reinterpret_cast<uint32_t(&)[16]>(u[0]));
      tc::tmem_ld16(lane_base + COL_S + (uint32_t)(buf * TK + half * 32 + 16), reinterpret_cast<uint32_t(&)[16]>(u[16]));
      tc::tmem_ld_wait();
      if (tr) acc_[5] += clock64() - tc0_;
      const int nvalid = T - (kt * TK + half * 32);        // >= 32 except at the end of the item
      float p[32];
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) p[jj] = wa_ex2(fmaf(__uint_as_float(u[jj]), kC, -mc));
      if (nvalid < 32) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) p[jj] = jj < nvalid ? p[jj] : 0.f;
      }
#pragma unroll
      for (int jj = 0; jj < 32; jj += 4) { l0 += p[jj]; l1 += p[jj + 1]; l2 += p[jj + 2]; l3 += p[jj + 3]; }
      // P over this thread's own score columns: 32 bf16 = 16 columns at buf * 64 + half * 32
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(p[2 * e], p[2 * e + 1]);
        pk[e] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      tc::tmem_st16(lane_base + COL_S + (uint32_t)(buf * TK + half * 32), pk);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&p_full[buf])) : "memory");
      if (tr) acc_[4] += clock64() - tc0_;
    }
    // epilogue: O / sum(p) -> bf16, written as the A tile image of the out-projection ([M/128][D/64][8][128][8])
    const float l = (l0 + l1) + (l2 + l3);
    asm volatile("bar.sync 1, 256;" ::: "memory");   // ml_s is read above by the other half
    ml_s[half * TQ + row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv_l = 1.f / (l + ml_s[(half ^ 1) * TQ + row]);
    tc::mbar_wait_parked(&o_full, 0);
    tc::fence_after_sync();
    uint32_t u[32];
    tc::tmem_ld16(lane_base + COL_O + (uint32_t)(half * 32), reinterpret_cast<uint32_t(&)[16]>(u[0]));
    tc::tmem_ld16(lane_base + COL_O + (uint32_t)(half * 32 + 16), reinterpret_cast<uint32_t(&)[16]>(u[16]));
    tc::tmem_ld_wait();
    if (ti < T) {
      const int mrow = m_item + ti;
      __nv_bfloat16* dst = out_img + ((size_t)(mrow >> 7) * heads + h) * 8192 + (size_t)(mrow & 127) * 8;
#pragma unroll
      for (int oc = 0; oc < 4; ++oc) {
        __align__(16) __nv_bfloat162 h2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          h2[e] = __floats2bfloat162_rn(__uint_as_float(u[oc * 8 + 2 * e]) * inv_l, __uint_as_float(u[oc * 8 + 2 * e + 1]) * inv_l);
        *reinterpret_cast<uint4*>(dst + (size_t)(half * 4 + oc) * 1024) = *reinterpret_cast<const uint4*>(h2);
      }
    }
    if (tr && tid == 0) { for (int j = 0; j < 6; ++j) trace[j] = acc_[j]; trace[6] = clock64() - tstart_; }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == wa::SM_WARPS) tc::tmem_dealloc(tmem, wa::TMEM_COLS);
}

// qkv: head-major layout of common.cuh (B x 3 x heads blocks of Tp x 64, pad rows finite);
// out_img: tile image of [B*T, D]
int launch_whisper_attention_tc(const void* qkv_img, void* out_img, int B, int T, int D, int heads, int vswap,
                                cudaStream_t s) {
  if (D != heads * 64 || D % 64) { set_error("whisper_attention_tc: head dim must be 64"); return SVCB_E_UNSUPPORTED; }
  if (B <= 0 || T <= 0) return SVCB_OK;
  static DevSmemCache attr_cache;
  SVCB_CUDA_CHECK(ensure_dyn_smem(whisper_attn_tc_kernel, wa::SMEM, attr_cache));
  KernelScope ks("whisper_attn_tc", s, 4.0 * B * (double)T * T * D, 2.0 * 4 * B * (double)T * D);
  whisper_attn_tc_kernel<<<dim3((T + wa::TQ - 1) / wa::TQ, heads, B), wa::THREADS, wa::SMEM, s>>>(
      static_cast<const __nv_bfloat16*>(qkv_img), static_cast<__nv_bfloat16*>(out_img), T, D, vswap, s2d_get_trace());
  SVCB_LAUNCH_CHECK("whisper_attn_tc");
  return SVCB_OK;
}

// row-major [B*T, 3D] (q | k | v) -> the head-major layout  (unit tests; the encoder's QKV GEMM writes it directly)
__global__ void qkv_rowmajor_to_heads_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int T, int Dm) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int oct_per_row = 3 * Dm / 8;
  if (idx >= (size_t)B * T * oct_per_row) return;
  const int m = (int)(idx / oct_per_row), n = (int)(idx % oct_per_row) * 8;
  const int w = n / Dm, hd = (n - w * Dm) >> 6, d = n & 63, b = m / T, t = m - b * T;
  *reinterpret_cast<uint4*>(dst + qkv_heads_off(b, w, hd, t, d, Dm >> 6, qkv_heads_tp(T))) =
      *reinterpret_cast<const uint4*>(src + (size_t)m * 3 * Dm + n);
}
int launch_qkv_rowmajor_to_heads(const void* src, void* dst, int B, int T, int D, cudaStream_t s) {
  const size_t n = (size_t)B * T * (3 * D / 8);
  qkv_rowmajor_to_heads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src),
                                                                          static_cast<__nv_bfloat16*>(dst), B, T, D);
  SVCB_LAUNCH_CHECK("qkv_rowmajor_to_heads");
  return SVCB_OK;
}

// tile image ([R/128][K/64][8][128][8]) -> row-major [R, K]  (unit tests)
__global__ void image_to_rowmajor_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int kc_per_row = K / 8;
  if (idx >= (size_t)R * kc_per_row) return;
  const int m = (int)(idx / kc_per_row), k = (int)(idx % kc_per_row) * 8;
  const size_t off = ((size_t)(m >> 7) * (K / 64) + (k >> 6)) * 8192 + (size_t)((k & 63) >> 3) * 1024 + (size_t)(m & 127) * 8;
  *reinterpret_cast<uint4*>(dst + (size_t)m * K + k) = *reinterpret_cast<const uint4*>(src + off);
}
int launch_image_to_rowmajor(const void* src, void* dst, int R, int K, cudaStream_t s) {
  const size_t n = (size_t)R * (K / 8);
  image_to_rowmajor_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src),
                                                                       static_cast<__nv_bfloat16*>(dst), R, K);
  SVCB_LAUNCH_CHECK("image_to_rowmajor");
  return SVCB_OK;
}

}  // namespace svcb
