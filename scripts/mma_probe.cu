// tcgen05.mma issue-cost probe (round 2): how many SM cycles does ONE MMA cost as a function of N, the
// shared-memory operand layout (K-major SWIZZLE_NONE "panel" layout of csrc/tc.cuh vs SWIZZLE_128B),
// the start-address alignment (the conv-tap trick shifts A by multiples of 16 B), M (128 / 64), the
// operand source of A (shared memory vs tensor memory) and the CTA group (1 / 2)?  Round 1 measured
// ~110 cycles per MMA at N = 48 / 80 / 160 alike inside amp_conv_tc; this isolates the instruction
// from that kernel's producer / epilogue pipeline.
//
// build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/mma_probe scripts/mma_probe.cu
// run:    scripts/mma_probe [sweep]      (prints one line per configuration)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg {
  int M, N;          // MMA shape (M = 256 only with cta_group 2)
  int layout;        // 0 = SWIZZLE_NONE panels (LBO = rows*16, SBO = 128), 1 = SWIZZLE_128B (SBO = 1024)
  int a_off;         // extra start offset of A in bytes (multiple of 16): the conv-tap shift
  int rows_a;        // rows of the A panel (LBO_A = rows_a * 16) for layout 0
  int nk;            // K chunks (of 16) cycled through per "tap"
  int ntaps;         // taps cycled through (A start shifted by tap*dil*16 bytes in layout 0)
  int dil;
  int n_mma;         // MMAs issued per CTA
  int a_tmem;        // 1: A operand from tensor memory (TS form)
  int n_acc;         // accumulators alternated (1 or 2)
  int cg;            // cta_group (1 or 2)
  int b_parts;       // distinct B tiles cycled through (weights of different taps)
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  if (layout == 1) d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t cta_rank = 0;
  if (c.cg == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  // fill shared memory with a harmless bf16 pattern (1.0 / 0.5): content does not change timing much,
  // but keeps the datapath toggling
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = (i & 1) ? 0x3F803F00u : 0x3F003F80u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    if (c.cg == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (c.cg == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  long long t0 = 0, t1 = 0;
  if (warp == 0 && cta_rank == 0) {
    const int Mcta = c.cg == 2 ? c.M / 2 : c.M;
    const int Ncta = c.cg == 2 ? c.N / 2 : c.N;   // B rows held by this CTA
    (void)Mcta;
    const uint32_t a_base = smem_u32(smem) + (uint32_t)c.a_off;
    const uint32_t a_bytes = c.layout == 0 ? (uint32_t)(2 * c.nk) * c.rows_a * 16u : (uint32_t)c.rows_a * 128u * ((c.nk + 3) / 4);
    const uint32_t b_base = (smem_u32(smem) + a_bytes + 4096u + 1023u) & ~1023u;
    const uint32_t lbo_a = c.layout == 0 ? (uint32_t)c.rows_a * 16u : 16u;
    const uint32_t lbo_b = c.layout == 0 ? (uint32_t)Ncta * 16u : 16u;
    const uint32_t sbo = c.layout == 0 ? 128u : 1024u;
    const uint32_t b_tile = c.layout == 0 ? (uint32_t)(2 * c.nk) * Ncta * 16u : (uint32_t)Ncta * 128u * ((c.nk + 3) / 4);
    const uint64_t ad0 = make_desc(a_base, lbo_a, sbo, c.layout);
    const uint64_t bd0 = make_desc(b_base, lbo_b, sbo, c.layout);
    const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
    const uint32_t kstep_a = c.layout == 0 ? (2u * lbo_a) >> 4 : 2u;   // 32 B inside the swizzle row
    const uint32_t kstep_b = c.layout == 0 ? (2u * lbo_b) >> 4 : 2u;
    const uint32_t idesc = idesc_bf16(c.M, c.N);
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    if (pred) {
      t0 = clock64();
      int i = 0;
      while (i < c.n_mma) {
        for (int tap = 0; tap < c.ntaps && i < c.n_mma; ++tap) {
          uint32_t ad = (uint32_t)ad0 + (c.layout == 0 ? (uint32_t)(tap * c.dil) : 0u);
          uint32_t bd = (uint32_t)bd0 + (((uint32_t)(tap % c.b_parts) * b_tile) >> 4);
          for (int kk = 0; kk < c.nk && i < c.n_mma; ++kk, ++i) {
            const uint32_t d = tmem + (uint32_t)((i % c.n_acc) * 256);
            const uint32_t acc = i >= c.n_acc ? 1u : 0u;
            if (c.a_tmem) {
              const uint32_t at = tmem + 480u;   // 8 columns of bf16 pairs = K 16
              if (c.cg == 1)
                asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                             :: "r"(d), "r"(at), "r"(bd), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
              else
                asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                             "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                             :: "r"(d), "r"(at), "r"(bd), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
            } else if (c.cg == 1) {
              asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                           "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                           :: "r"(d), "r"(ad), "r"(a_hi), "r"(bd), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
            } else {
              asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                           "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
                           :: "r"(d), "r"(ad), "r"(a_hi), "r"(bd), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
            }
            ad += kstep_a;
            bd += kstep_b;
          }
        }
      }
      if (c.cg == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
      uint32_t done = 0;
      for (uint32_t spin = 0; spin < (1u << 26) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      t1 = clock64();
      if (!done) t1 = t0 - 1;
      cycles[blockIdx.x] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (c.cg == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (warp == 0) {
    if (c.cg == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ---- probe 2: straight-line issue.  UNR MMAs per loop trip; MODE 0 = all MMAs reuse ONE descriptor pair
// (descriptors live in uniform registers, zero per-MMA integer work: the hardware's own issue/execute
// rate), MODE 1 = descriptor low words advance by compile-time constants (what a K loop needs).
template <int M, int N, int MODE, int ATMEM>
__global__ void __launch_bounds__(128, 1) probe2_kernel(int n_iter, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = (i & 1) ? 0x3F803F00u : 0x3F003F80u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    constexpr int UNR = 32;
    const uint32_t a_base = smem_u32(smem);
    const uint32_t b_base = a_base + 64 * 1024;
    constexpr uint32_t lbo_a = 128 * 16, lbo_b = N * 16;
    const uint64_t ad0 = make_desc(a_base, lbo_a, 128, 0), bd0 = make_desc(b_base, lbo_b, 128, 0);
    const uint32_t a_lo = (uint32_t)ad0, a_hi = (uint32_t)(ad0 >> 32), b_lo = (uint32_t)bd0, b_hi = (uint32_t)(bd0 >> 32);
    constexpr uint32_t idesc = idesc_bf16(M, N);
    constexpr uint32_t ka = (2 * lbo_a) >> 4, kb = (2 * lbo_b) >> 4;
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    if (pred) {
      const long long t0 = clock64();
      for (int it = 0; it < n_iter; ++it) {
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
          const uint32_t al = MODE ? a_lo + (uint32_t)(j & 3) * ka : a_lo;
          const uint32_t bl = MODE ? b_lo + (uint32_t)(j & 3) * kb : b_lo;
          if (ATMEM)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                         :: "r"(tmem), "r"(tmem + 480u), "r"(bl), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                         :: "r"(tmem), "r"(al), "r"(a_hi), "r"(bl), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done = 0;
      for (uint32_t spin = 0; spin < (1u << 26) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      const long long t1 = clock64();
      cycles[blockIdx.x] = done ? t1 - t0 : -1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int M, int N, int MODE, int ATMEM>
static void run2(const char* tag, int grid, long long* d_cycles) {
  const int n_iter = 128;  // x 32 MMAs
  CK(cudaFuncSetAttribute(probe2_kernel<M, N, MODE, ATMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaMemset(d_cycles, 0, sizeof(long long) * 512));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    probe2_kernel<M, N, MODE, ATMEM><<<grid, 128, 200 * 1024>>>(n_iter, d_cycles);
    CK(cudaEventRecord(e1));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-10s FAILED: %s\n", tag, cudaGetErrorString(e)); exit(3); }
  }
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(512);
  CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * 512, cudaMemcpyDeviceToHost));
  std::vector<double> v;
  for (int i = 0; i < grid; ++i) v.push_back((double)h[i] / (n_iter * 32));
  std::sort(v.begin(), v.end());
  printf("%-12s M=%3d N=%3d mode=%d atmem=%d grid=%3d | cyc/mma min %.1f med %.1f max %.1f | kernel %.3f ms | %.0f flop/cyc/SM at med (peak 8192)\n",
         tag, M, N, MODE, ATMEM, grid, v.front(), v[v.size() / 2], v.back(), ms, 2.0 * M * N * 16.0 / v[v.size() / 2]);
  fflush(stdout);
}

template <int MODE, int ATMEM>
static void sweep2(const char* tag, int grid, long long* d) {
  run2<128, 16, MODE, ATMEM>(tag, grid, d);  run2<128, 32, MODE, ATMEM>(tag, grid, d);  run2<128, 48, MODE, ATMEM>(tag, grid, d);
  run2<128, 64, MODE, ATMEM>(tag, grid, d);  run2<128, 80, MODE, ATMEM>(tag, grid, d);  run2<128, 96, MODE, ATMEM>(tag, grid, d);
  run2<128, 128, MODE, ATMEM>(tag, grid, d); run2<128, 160, MODE, ATMEM>(tag, grid, d); run2<128, 192, MODE, ATMEM>(tag, grid, d);
  run2<128, 256, MODE, ATMEM>(tag, grid, d);
}

static void run(const char* tag, Cfg c, int grid, long long* d_cycles) {
  const size_t smem = 200 * 1024;
  CK(cudaMemset(d_cycles, 0, sizeof(long long) * 512));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = c.cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = c.cg > 1 ? 1 : 0;
  for (int rep = 0; rep < 2; ++rep) {   // first = warm-up
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, probe_kernel, c, d_cycles));
    CK(cudaEventRecord(e1));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-10s FAILED: %s\n", tag, cudaGetErrorString(e)); exit(3); }
  }
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(512);
  CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * 512, cudaMemcpyDeviceToHost));
  std::vector<double> v;
  for (int i = 0; i < grid; i += c.cg) v.push_back((double)h[i] / c.n_mma);
  std::sort(v.begin(), v.end());
  const double flop = 2.0 * c.M * c.N * 16.0;
  printf("%-10s M=%3d N=%3d lay=%d aoff=%3d rows=%3d nk=%d taps=%2d dil=%d atmem=%d nacc=%d cg=%d bparts=%d grid=%3d | cyc/mma min %.1f med %.1f max %.1f | kernel %.3f ms | %.0f flop/cyc/SM at med\n",
         tag, c.M, c.N, c.layout, c.a_off, c.rows_a, c.nk, c.ntaps, c.dil, c.a_tmem, c.n_acc, c.cg, c.b_parts, grid,
         v.front(), v[v.size() / 2], v.back(), ms, flop / v[v.size() / 2] / c.cg);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const char* sweep = argc > 1 ? argv[1] : "all";
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, sizeof(long long) * 512));
  int n_sm = 0;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
  const int NM = 4096;
  const int Ns[] = {16, 32, 48, 64, 80, 96, 128, 160, 192, 256};
  auto base = [&](int M, int N) { Cfg c{}; c.M = M; c.N = N; c.layout = 0; c.a_off = 0; c.rows_a = 128; c.nk = 4; c.ntaps = 1; c.dil = 1;
                                  c.n_mma = NM; c.a_tmem = 0; c.n_acc = 1; c.cg = 1; c.b_parts = 1; return c; };
  const bool all = !strcmp(sweep, "all");
  if (all || !strcmp(sweep, "p2")) {
    sweep2<0, 0>("p2_same", n_sm, d_cycles);
    sweep2<1, 0>("p2_kstep", n_sm, d_cycles);
    sweep2<0, 1>("p2_atmem", n_sm, d_cycles);
    run2<64, 48, 0, 0>("p2_m64", n_sm, d_cycles); run2<64, 160, 0, 0>("p2_m64", n_sm, d_cycles); run2<64, 256, 0, 0>("p2_m64", n_sm, d_cycles);
    run2<128, 48, 0, 0>("p2_grid1", 1, d_cycles); run2<128, 160, 0, 0>("p2_grid1", 1, d_cycles); run2<128, 256, 0, 0>("p2_grid1", 1, d_cycles);
  }
  if (all || !strcmp(sweep, "n")) {
    for (int N : Ns) run("none", base(128, N), n_sm, d_cycles);
    for (int N : Ns) { Cfg c = base(128, N); c.layout = 1; run("sw128", c, n_sm, d_cycles); }
  }
  if (all || !strcmp(sweep, "align")) {
    for (int N : {48, 80, 160}) for (int off : {16, 48, 64, 112}) { Cfg c = base(128, N); c.a_off = off; c.rows_a = 160; run("shift", c, n_sm, d_cycles); }
    // the conv pattern: taps shift A rows, weights per tap differ
    for (int N : {48, 80, 96, 160}) { Cfg c = base(128, N); c.rows_a = 178; c.nk = N >= 160 ? 10 : N >= 80 ? 5 : 3; c.ntaps = 11; c.dil = 5; c.b_parts = 2; run("convpat", c, n_sm, d_cycles); }
  }
  if (all || !strcmp(sweep, "m64")) {
    for (int N : {16, 48, 80, 160, 256}) run("m64", base(64, N), n_sm, d_cycles);
  }
  if (all || !strcmp(sweep, "tmem")) {
    for (int N : Ns) { Cfg c = base(128, N); c.a_tmem = 1; run("a_tmem", c, n_sm, d_cycles); }
  }
  if (all || !strcmp(sweep, "acc2")) {
    for (int N : {48, 80, 160}) { Cfg c = base(128, N); c.n_acc = 2; run("acc2", c, n_sm, d_cycles); }
  }
  if (all || !strcmp(sweep, "grid1")) {
    for (int N : {48, 160, 256}) run("grid1", base(128, N), 1, d_cycles);
  }
  if (!strcmp(sweep, "cg2")) {
    for (int N : {32, 48, 64, 96, 160, 256}) { Cfg c = base(256, N); c.cg = 2; run("cg2", c, n_sm & ~1, d_cycles); }
    for (int N : {32, 96, 160, 256}) { Cfg c = base(256, N); c.cg = 2; c.a_tmem = 1; run("cg2_tmem", c, n_sm & ~1, d_cycles); }
  }
  return 0;
}
