// General Conv1d as an implicit GEMM on tcgen05/TMEM (bf16 or bf16x3 operands, fp32 accumulate).
//
// Serves every dense convolution of the prior encoder, the flow and the generator head
// (vits/models.py:44-49, vits/attentions.py:215-223,390-398, vits/modules.py:184-198,296-299,
// vits_decoder/generator.py:177-178) — the F.conv1d call sites whose contraction is wide enough for
// the tensor cores.
//
//   D[t, co] = sum_cc sum_tap  A_cc[t + tap*dil, :] . W[tap, cc][co, :]      M = 128, N = BN, K = KCH
// * Input channels are processed in chunks of KCH (32 or 64).  Eight producer warps gather one
//   chunk of x (fp32, any strides — the time-major PPG input included), apply the optional input
//   mask and the conv's zero padding, split to bf16 hi/lo and write the K-major panel layout of
//   tc.cuh into a 2-deep A ring (generic proxy -> fence.proxy.async -> mbarrier).
// * For every (chunk, tap) the pre-packed weight tiles (hi, lo) arrive by 1-D bulk copy into a
//   3-deep W ring.  One thread issues the MMAs: taps reuse the same A chunk through a row-shifted
//   descriptor; tcgen05.commit releases W slots and A buffers.
// * Epilogue (the producer warps again): tcgen05.ld -> bias -> {none, ReLU, Mish, tanh, WaveNet
//   gate on interleaved channel pairs} -> output mask -> residual -> accumulate -> store [B,C,T].
#include <cstdio>

#include "common.cuh"
#include "tc.cuh"

namespace svcb {

constexpr int CT_M = 128;
constexpr int CT_WST = 3;  // W ring depth (max)
static inline unsigned tc_cols(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u; }

__device__ __forceinline__ float ct_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_MISH: { const float sp = v > 20.f ? v : log1pf(expf(v)); return v * tanhf(sp); }
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__device__ __forceinline__ void ct_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(320, 1)
conv_tc_kernel(const ConvTcParams p, const int wst) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2], w_full[CT_WST], w_empty[CT_WST], bar_acc;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = tc::warp_uniform_idx();
  const int b = blockIdx.z, nt = blockIdx.y;
  const int t0 = blockIdx.x * CT_M;
  const int P = p.pad;
  const int R = CT_M + (p.K - 1) * p.dil;
  const int KC = p.kch / 8;
  const int n0row = t0 - P;                       // sequence position of A row 0
  const uint32_t a_part = (uint32_t)KC * R * 16u; // one of hi / lo
  const int nparts = p.nsplit == 3 ? 2 : 1;
  const uint32_t a_buf = a_part * nparts;
  const uint32_t w_tile = (uint32_t)p.kch * p.bn * 2u;
  uint8_t* A0 = smem;
  uint8_t* W0 = smem + 2 * a_buf;
  const int ncc = p.cin_pad / p.kch;
  const long long len = p.lengths ? p.lengths[b] : (long long)1 << 60;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&a_full[i], 256); tc::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < wst; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    tc::mbar_init(&bar_acc, 1);
    tc::fence_barrier_init();
  }
  const uint32_t ncols = tc::tmem_cols_for(p.bn);
  __syncwarp();
  if (warp == 4) tc::tmem_alloc(&tmem_slot, ncols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < 4 || warp >= 6) {
    // ---------------------------------------------------------------- A producers: the eight warps that run the
    // epilogue later.  A thread fetches TWO items (16 scalar or 4 vector loads in flight) before it converts either:
    // with four warps and one item at a time a chunk cost ~4 global-load latencies and the gather, not the MMAs,
    // set the pace of every small convolution (profiles/r02_notes.md §9).
    const int pid = warp < 4 ? tid : tid - 64;   // 0 .. 255
    const float* xb = p.x + (long long)b * p.sxb;
    const float* x2b = p.x2 ? p.x2 + (long long)b * p.sx2b : nullptr;
    const int n_items = R * KC;
    auto fetch = [&](int item, int cc, float (&v)[8]) {
      const int r = item % R, kc = item / R;
      const int tau = n0row + r;
      const int c0 = cc * p.kch + kc * 8;
      const bool row_ok = tau >= 0 && tau < p.Tin && (!(p.flags & CONV_IN_MASK) || tau < len);
      if (row_ok && x2b && c0 >= p.cin1) {   // second input: channel-contiguous windows (unaligned)
        const float* s2 = x2b + (long long)tau * p.sx2t + (c0 - p.cin1);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (c0 - p.cin1 + e < p.cin2) ? __ldg(s2 + e) : 0.f;
      } else if (row_ok) {
        if (p.sxc == 1 && c0 + 8 <= p.Cin) {  // channel-contiguous input: two 16-byte loads
          const float4* s4 = reinterpret_cast<const float4*>(xb + (long long)tau * p.sxt + c0);
          const float4 u0 = __ldg(s4), u1 = __ldg(s4 + 1);
          v[0] = u0.x; v[1] = u0.y; v[2] = u0.z; v[3] = u0.w; v[4] = u1.x; v[5] = u1.y; v[6] = u1.z; v[7] = u1.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            v[e] = (c0 + e < p.Cin) ? __ldg(xb + (long long)(c0 + e) * p.sxc + (long long)tau * p.sxt) : 0.f;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
    };
    auto put = [&](int item, uint8_t* Ah, uint8_t* Al, const float (&v)[8]) {
      const int r = item % R, kc = item / R;
      __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        hi[e] = __float2bfloat16_rn(v[e]);
        lo[e] = __float2bfloat16_rn(v[e] - __bfloat162float(hi[e]));
      }
      *reinterpret_cast<uint4*>(Ah + ((size_t)kc * R + r) * 16) = *reinterpret_cast<const uint4*>(hi);
      if (nparts == 2)
        *reinterpret_cast<uint4*>(Al + ((size_t)kc * R + r) * 16) = *reinterpret_cast<const uint4*>(lo);
    };
    for (int cc = 0; cc < ncc; ++cc) {
      const int buf = cc & 1;
      if (cc >= 2) tc::mbar_wait(&a_empty[buf], (uint32_t)(((cc >> 1) - 1) & 1));
      uint8_t* Ah = A0 + (size_t)buf * a_buf;
      uint8_t* Al = Ah + a_part;
      for (int it0 = pid; it0 < n_items; it0 += 512) {
        const int it1 = it0 + 256;
        float va[8], vb[8];
        fetch(it0, cc, va);
        if (it1 < n_items) fetch(it1, cc, vb);
        put(it0, Ah, Al, va);
        if (it1 < n_items) put(it1, Ah, Al, vb);
      }
      tc::fence_proxy_async_smem();
      ct_arrive(&a_full[buf]);
    }
  }
  if (warp < 4 || warp >= 6) {
    // ---------------------------------------------------------------- epilogue: two groups of four warps
    // (the A producers, now idle, and warps 6-9) take alternate 16-column strips.  A warp may only
    // read the TMEM lane quarter warp%4, so warps 6..9 own quarters 2,3,0,1.
    const int grp = warp < 4 ? 0 : 1, wq = warp & 3;
    tc::mbar_wait(&bar_acc, 0);
    tc::fence_after_sync();
    const int t = t0 + wq * 32 + lane;
    const bool gate = (p.flags & CONV_GATE) != 0;
    const int cout_real = gate ? p.Cout / 2 : p.Cout;
    const bool keep = !(p.flags & CONV_OUT_MASK) || t < len;
    float* yb = p.y + (long long)b * cout_real * p.Tout;
    const float* rb = p.res ? p.res + (long long)b * cout_real * p.Tout : nullptr;
    if (p.ilv) {   // interleaved store: ilv consecutive samples of one channel per thread (one 8 / 16-byte store)
      float* yi = p.y + (long long)b * p.Cout * p.Tout;   // = [Cout / ilv][Tout * ilv]
      for (int c0 = grp * 16; c0 < p.bn; c0 += 32) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, v);
        tc::tmem_ld_wait();
        if (t < p.Tout) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const int cp = nt * p.bn + c0 + j;
            if (cp < p.Cout) {
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = ct_act(__uint_as_float(v[j + e]) + (p.bias ? __ldg(p.bias + cp + e) : 0.f), p.act);
              if (p.ilv == 4) {
                *reinterpret_cast<float4*>(yi + ((long long)(cp >> 2) * p.Tout + t) * 4) = make_float4(o[0], o[1], o[2], o[3]);
              } else {
                *reinterpret_cast<float2*>(yi + ((long long)(cp >> 1) * p.Tout + t) * 2) = make_float2(o[0], o[1]);
                *reinterpret_cast<float2*>(yi + ((long long)((cp >> 1) + 1) * p.Tout + t) * 2) = make_float2(o[2], o[3]);
              }
            }
          }
        }
      }
    } else
    for (int c0 = grp * 16; c0 < p.bn; c0 += 32) {
      uint32_t v[16];
      tc::tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, v);
      const bool live = t < p.Tout;
      const int step = gate ? 2 : 1;
      // residual / accumulator loads of the strip first (all in flight), stores afterwards
      float add[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = 0.f;
        const int cp = nt * p.bn + c0 + j;
        if (live && (j % step) == 0 && cp + (step - 1) < p.Cout) {
          const long long off = (long long)(gate ? (cp >> 1) : cp) * p.Tout + t;
          if (rb) a = rb[off];
          if (p.flags & CONV_ACCUM) a += yb[off];
        }
        add[j] = a;
      }
      tc::tmem_ld_wait();
      if (live) {
        if (gate) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const int cp = nt * p.bn + c0 + j;
            if (cp + 1 < p.Cout) {
              const float a = __uint_as_float(v[j]) + __ldg(p.bias + cp);
              const float g = __uint_as_float(v[j + 1]) + __ldg(p.bias + cp + 1);
              float o = tanhf(a) * (1.f / (1.f + expf(-g)));
              if (!keep) o = 0.f;
              yb[(long long)(cp >> 1) * p.Tout + t] = o + add[j];
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = nt * p.bn + c0 + j;
            if (co < p.Cout) {
              float o = ct_act(__uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + co) : 0.f), p.act);
              if (!keep) o = 0.f;
              yb[(long long)co * p.Tout + t] = o + add[j];
            }
          }
        }
      }
    }
  } else if (tid == 128) {
    // ---------------------------------------------------------------- W producer
    const int total = ncc * p.K * nparts;
    for (int i = 0; i < total; ++i) {
      const int st = i % wst;
      if (i >= wst) tc::mbar_wait(&w_empty[st], (uint32_t)(((i / wst) - 1) & 1));
      const int part = i % nparts, tap = (i / nparts) % p.K, cc = i / (nparts * p.K);
      const size_t tile = (((size_t)tap * ncc + cc) * 2 + part) * p.ntiles + nt;
      tc::mbar_arrive_expect_tx(&w_full[st], w_tile);
      tc::bulk_g2s(W0 + (size_t)st * w_tile, p.wpk + tile * w_tile, w_tile, &w_full[st]);
    }
  } else if (warp_u == 5) {
    // ---------------------------------------------------------------- MMA issuer (whole warp walks the
    // uniform loop, one elected lane issues: tc::elect_one)
    const uint32_t idesc = tc::idesc_bf16(CT_M, p.bn);
    const uint32_t a_base = tc::smem_u32(A0), w_base = tc::smem_u32(W0);
    const uint32_t lbo_a = (uint32_t)R * 16u, lbo_b = (uint32_t)p.bn * 16u;
    const uint32_t kstep_a = (2u * lbo_a) >> 4, kstep_b = (2u * lbo_b) >> 4;
    const int nk = p.kch / 16;
    uint32_t accumulate = 0;
    int wi = 0;
    for (int cc = 0; cc < ncc; ++cc) {
      const int buf = cc & 1;
      tc::mbar_wait(&a_full[buf], (uint32_t)((cc >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t ah = a_base + (uint32_t)buf * a_buf, al = ah + a_part;
      for (int tap = 0; tap < p.K; ++tap) {
        const uint32_t row_off = (uint32_t)(tap * p.dil) * 16u;
        for (int part = 0; part < nparts; ++part, ++wi) {
          const int st = wi % wst;
          tc::mbar_wait(&w_full[st], (uint32_t)((wi / wst) & 1));
          tc::fence_after_sync();
          const uint64_t bd0 = tc::smem_desc(w_base + (uint32_t)st * w_tile, lbo_b);
          const int n_a = (part == 0 && nparts == 2) ? 2 : 1;
          if (tc::elect_one()) {
            uint32_t acc_flag = accumulate;
            const uint32_t b_hiw = (uint32_t)(bd0 >> 32);
            for (int ap = 0; ap < n_a; ++ap) {
              const uint64_t ad0 = tc::smem_desc((ap == 0 ? ah : al) + row_off, lbo_a);
              const uint32_t a_hiw = (uint32_t)(ad0 >> 32);
              uint32_t ad = (uint32_t)ad0, bd = (uint32_t)bd0;   // low words: only the start-address field moves
              for (int kk = 0; kk < nk; ++kk) {
                tc::mma_bf16_lohi(tmem, ad, a_hiw, bd, b_hiw, idesc, acc_flag);
                acc_flag = 1;
                ad += kstep_a;
                bd += kstep_b;
              }
            }
            tc::mma_commit(&w_empty[st]);
          }
          accumulate = 1;
        }
      }
      if (tc::elect_one()) tc::mma_commit(&a_empty[buf]);
    }
    if (tc::elect_one()) tc::mma_commit(&bar_acc);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, ncols);
}

static size_t conv_tc_smem_bytes(const ConvTcParams& p, int wst) {
  const int R = CT_M + (p.K - 1) * p.dil;
  const size_t a_buf = (size_t)(p.kch / 8) * R * 16 * (p.nsplit == 3 ? 2 : 1);
  return 2 * a_buf + (size_t)wst * p.kch * p.bn * 2 + 128;
}

int launch_conv_tc(const ConvTcParams& p, cudaStream_t s) {
  if (p.B <= 0 || p.Tout <= 0) return SVCB_OK;
  if ((p.kch != 32 && p.kch != 64) || p.cin_pad % p.kch || p.bn % 16 || p.bn < 16 || p.bn > 256 ||
      p.ntiles * p.bn < p.Cout || (p.nsplit != 1 && p.nsplit != 3) || p.Tout > p.Tin) {
    set_error("conv_tc: unsupported tiling (stride-1 'same' convolutions only)");
    return SVCB_E_BAD_SHAPE;
  }
  if (p.x2 && (p.cin1 % 8 || p.cin1 < p.Cin || p.cin1 + p.cin2 > p.cin_pad)) {
    set_error("conv_tc: second input must start at a multiple of 8 channels behind the first");
    return SVCB_E_BAD_SHAPE;
  }
  if (p.ilv && ((p.ilv != 2 && p.ilv != 4) || p.Cout % 4 || p.res || (p.flags & (CONV_ACCUM | CONV_GATE | CONV_OUT_MASK)) ||
                (reinterpret_cast<uintptr_t>(p.y) & 15))) {
    set_error("conv_tc: interleaved output needs ilv in {2, 4}, Cout % 4 == 0, a 16-byte aligned y and a plain epilogue");
    return SVCB_E_BAD_SHAPE;
  }
  // a 2-deep weight ring when that lets two CTAs share an SM (one CTA's epilogue then overlaps the
  // other's gather + MMA; the 1x1 convolutions are epilogue-bound), else the 3-deep ring
  int wst = CT_WST;
  if (conv_tc_smem_bytes(p, 2) <= 113 * 1024 && tc_cols(p.bn) <= 256) wst = 2;
  const size_t smem = conv_tc_smem_bytes(p, wst);
  if (smem > 227 * 1024 - 512) { set_error("conv_tc: tile does not fit shared memory"); return SVCB_E_UNSUPPORTED; }
  static DevSmemCache attr_cache;
  SVCB_CUDA_CHECK(ensure_dyn_smem(conv_tc_kernel, smem, attr_cache));
  dim3 grid((p.Tout + CT_M - 1) / CT_M, p.ntiles, p.B);
  const int cout_real = (p.flags & CONV_GATE) ? p.Cout / 2 : p.Cout;
  char kname[64];
  snprintf(kname, sizeof(kname), "conv_tc_%s_%dto%d_k%d_o1", p.nsplit == 3 ? "bf16x3" : "bf16", p.Cin, p.Cout, p.K);
  KernelScope ks(kname, s,
                 2.0 * p.Cin * p.K * p.Cout * (double)p.Tout * p.B,
                 4.0 * ((double)p.B * p.Cin * p.Tin + (double)p.B * cout_real * p.Tout * (p.res ? 2 : 1)) +
                     2.0 * (double)p.Cin * p.K * p.Cout);
  conv_tc_kernel<<<grid, 320, smem, s>>>(p, wst);
  SVCB_LAUNCH_CHECK("conv_tc");
  return SVCB_OK;
}

}  // namespace svcb
