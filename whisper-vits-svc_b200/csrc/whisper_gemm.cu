// Dense bf16 GEMM on tcgen05/TMEM for the Whisper encoder's linear layers, with fused epilogues.
//
//   C[M,N] = A[M,K] . W[N,K]^T + bias            A = LayerNorm output / attention output (bf16)
//                                               W = nn.Linear weight (bf16, [out,in] as stored)
// Replaces the F.linear calls of whisper/model.py:66-82 (query/key/value/out) and :116
// (mlp = Linear -> GELU -> Linear) including their bias, GELU and residual adds.
//
// One CTA = one 128 x BN output tile.  Four producer warps stream A/W k-tiles (64 wide) into a
// STAGES-deep ring with 16-byte cp.async, scattering each K-chunk into the K-major panel layout of
// tc.cuh; completion is published with cp.async.wait_group + fence.proxy.async + mbarrier.arrive.
// One thread issues tcgen05.mma (M=128, N=BN, K=16 x4 per tile) and frees slots with
// tcgen05.commit.  The producer warps then become the epilogue: tcgen05.ld -> bias / GELU /
// residual -> vector stores.
#include "common.cuh"
#include "tc.cuh"

namespace svcb {

enum GemmEpi : int { EPI_BF16 = 0, EPI_GELU_BF16 = 1, EPI_RESID_F32 = 2 };

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = tc::smem_u32(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

template <int BN, int EPI>
__global__ void __launch_bounds__(160, 1)
gemm_tc_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ W,
               const float* __restrict__ bias, void* out, const float* res, int M, int N, int K) {
  constexpr int BM = 128, BK = 64, KC = BK / 8;
  constexpr int STAGES = BN == 256 ? 4 : 4;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, ST_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_acc;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nk = K / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&bar_full[s], 128); tc::mbar_init(&bar_empty[s], 1); }
    tc::mbar_init(&bar_acc, 1);
    tc::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 4) tc::tmem_alloc(&tmem_slot, BN);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (tid < 128) {
    // ------------------------------------------------------------------ producers
    auto issue = [&](int st, int kt) {
      uint8_t* As = smem + (size_t)st * ST_BYTES;
      uint8_t* Bs = As + A_BYTES;
      const int k0 = kt * BK;
#pragma unroll
      for (int i = 0; i < (BM * KC) / 128; ++i) {
        const int c = tid + 128 * i;
        const int r = (c & 7) + 8 * ((c >> 5) & 15), kc = ((c >> 3) & 3) + 4 * (c >> 9);
        const int gm = m0 + r;
        const bool ok = gm < M;
        cp_async16(As + ((size_t)kc * BM + r) * 16, A + (size_t)(ok ? gm : 0) * K + k0 + kc * 8, ok);
      }
#pragma unroll
      for (int i = 0; i < (BN * KC) / 128; ++i) {
        const int c = tid + 128 * i;
        constexpr int RH = BN / 8;  // row groups of 8
        const int r = (c & 7) + 8 * ((c >> 5) % RH), kc = ((c >> 3) & 3) + 4 * ((c >> 5) / RH);
        cp_async16(Bs + ((size_t)kc * BN + r) * 16, W + (size_t)(n0 + r) * K + k0 + kc * 8, true);
      }
    };
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nk) issue(s, s);
      cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) {
        const int st = nxt % STAGES;
        if (nxt >= STAGES) tc::mbar_wait(&bar_empty[st], (uint32_t)(((nxt / STAGES) - 1) & 1));
        issue(st, nxt);
      }
      cp_async_commit();
      cp_async_wait<STAGES - 1>();
      tc::fence_proxy_async_smem();
      mbar_arrive(&bar_full[kt % STAGES]);
    }
    // ------------------------------------------------------------------ epilogue
    tc::mbar_wait(&bar_acc, 0);
    tc::fence_after_sync();
    const int m = m0 + tid;
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tc::tmem_ld_wait();
      if (m < M) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + (bias ? __ldg(bias + n0 + c0 + j) : 0.f);
        const size_t off = (size_t)m * N + n0 + c0;
        if (EPI == EPI_RESID_F32) {
          float* o = static_cast<float*>(out) + off;
          const float* rr = res + off;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 r4 = *reinterpret_cast<const float4*>(rr + j);
            float4 o4 = make_float4(f[j] + r4.x, f[j + 1] + r4.y, f[j + 2] + r4.z, f[j + 3] + r4.w);
            *reinterpret_cast<float4*>(o + j) = o4;
          }
        } else {
          __align__(16) __nv_bfloat16 h[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x = f[j];
            if (EPI == EPI_GELU_BF16) x = 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
            h[j] = __float2bfloat16_rn(x);
          }
          __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out) + off;
          *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(h);
          *reinterpret_cast<uint4*>(o + 8) = *reinterpret_cast<const uint4*>(h + 8);
        }
      }
    }
  } else if (tid == 128) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = tc::idesc_bf16(BM, BN);
    const uint32_t s0 = tc::smem_u32(smem);
    for (int kt = 0; kt < nk; ++kt) {
      const int st = kt % STAGES;
      tc::mbar_wait(&bar_full[st], (uint32_t)((kt / STAGES) & 1));
      tc::fence_after_sync();
      const uint32_t a0 = s0 + (uint32_t)st * ST_BYTES, b0 = a0 + A_BYTES;
#pragma unroll
      for (int kk = 0; kk < BK / 16; ++kk) {
        const uint64_t ad = tc::smem_desc(a0 + (uint32_t)kk * 2u * BM * 16u, BM * 16u);
        const uint64_t bd = tc::smem_desc(b0 + (uint32_t)kk * 2u * BN * 16u, BN * 16u);
        tc::mma_bf16(tmem, ad, bd, idesc, (kt | kk) ? 1u : 0u);
      }
      tc::mma_commit(&bar_empty[st]);
    }
    tc::mma_commit(&bar_acc);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, BN);
}

template <int BN, int EPI>
static int launch_gemm_t(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, void* out,
                         const float* res, int M, int N, int K, cudaStream_t s) {
  constexpr size_t smem = (size_t)4 * (128 * 64 * 2 + BN * 64 * 2);
  static bool attr = false;
  if (!attr) {
    SVCB_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    attr = true;
  }
  dim3 grid(N / BN, (M + 127) / 128);
  KernelScope ks("whisper_gemm_tc", s, 2.0 * M * (double)N * K,
                 2.0 * ((double)M * K + (double)N * K) + (EPI == EPI_RESID_F32 ? 8.0 : 2.0) * M * (double)N);
  gemm_tc_kernel<BN, EPI><<<grid, 160, smem, s>>>(A, W, bias, out, res, M, N, K);
  SVCB_LAUNCH_CHECK("gemm_tc");
  return SVCB_OK;
}

int launch_gemm_tc(const void* A_bf16, const void* W_bf16, const float* bias, void* out, const float* res,
                   int M, int N, int K, int epi, cudaStream_t s) {
  if (M <= 0) return SVCB_OK;
  if (K % 64 || N % 128) { set_error("gemm_tc: need K % 64 == 0 and N % 128 == 0"); return SVCB_E_BAD_SHAPE; }
  const __nv_bfloat16* A = static_cast<const __nv_bfloat16*>(A_bf16);
  const __nv_bfloat16* W = static_cast<const __nv_bfloat16*>(W_bf16);
  const bool wide = (N % 256 == 0);
  switch (epi) {
    case EPI_BF16:
      return wide ? launch_gemm_t<256, EPI_BF16>(A, W, bias, out, res, M, N, K, s)
                  : launch_gemm_t<128, EPI_BF16>(A, W, bias, out, res, M, N, K, s);
    case EPI_GELU_BF16:
      return wide ? launch_gemm_t<256, EPI_GELU_BF16>(A, W, bias, out, res, M, N, K, s)
                  : launch_gemm_t<128, EPI_GELU_BF16>(A, W, bias, out, res, M, N, K, s);
    case EPI_RESID_F32:
      return wide ? launch_gemm_t<256, EPI_RESID_F32>(A, W, bias, out, res, M, N, K, s)
                  : launch_gemm_t<128, EPI_RESID_F32>(A, W, bias, out, res, M, N, K, s);
  }
  set_error("gemm_tc: unknown epilogue");
  return SVCB_E_BAD_SHAPE;
}

}  // namespace svcb
