// attention.cu — fused QK^T . softmax . PV on tcgen05 (flash-style streaming softmax, no N^2 tensor in HBM).
//
// Replaces qkv_attention (reference src/model/attention.rs:5-45 == src/backend.rs:88-128, mask = None):
//   softmax((q d^-1/4)(k d^-1/4)^T) v  ==  softmax(q k^T d^-1/2) v.
//
// One CTA = NG x 128 query rows of one (sample, head); NG = 2 query tiles ping-pong on one K/V stream
// (while the softmax warps of tile 0 run their exponentials, the tensor pipe works for tile 1, and the K/V
// tiles are fetched from L2 once for 256 query rows). Warp roles:
//   warp 0          : TMA producer — Q tiles once, then K tiles [128 keys][d] and V^T tiles [d][128 keys]
//   warp 1          : TMEM allocator + tcgen05.mma issuer: S_g = Q_g K^T (fp32 in TMEM), O_g += P_g V
//   warps 2..2+4*NG : softmax groups — one query row per thread: tcgen05.ld S, running max/sum in fp32
//                     (exp2 with the d^-1/2 scale folded in), lazy O rescale in TMEM, P written as fp16
//                     into 128B-swizzled smem
// S_g(j+1) is issued as soon as the group has copied S_g(j) to registers, so QK^T overlaps the exponentials.
#include "attention.cuh"

#include <type_traits>

namespace sdb {

// VMN = false: V arrives transposed, V^T [d][keys] (K-major B operand of P.V: two boxes of [DPAD rows][64 keys]).
// VMN = true : V arrives as the projection wrote it, V [keys][d] (MN-major B operand: DC boxes of [128 keys][64 channels]) —
//              no transposing GEMM in front of the kernel.
// QK3 = true : q and k arrive as fp16 hi + lo pairs and S = q_hi k_hi^T + q_lo k_hi^T + q_hi k_lo^T (the 3-term split product of
//              the GEMMs): the logits are fp32-class. With single fp16 operands a logit of magnitude ~30 (peaked softmax of a
//              trained checkpoint) carries an absolute error ~1e-2, i.e. ~1 % on the dominant probabilities — the largest
//              single error source of a UNet step on realistic-statistics weights (tests/test_realstats_gpu.py).
template <int DPAD, int NG, bool VMN = false, bool QK3 = false>
struct AttnCfg {
  static constexpr int DC = (DPAD + 63) / 64;                          // 64-wide chunks of the head dim
  static constexpr int QK_PARTS = QK3 ? 2 : 1;                          // hi (+ lo) copies of the Q and K tiles
  static constexpr int Q_HALF = DC * 128 * 128;                         // [128 rows][64] x DC, 128 B rows
  static constexpr int Q_TILE = QK_PARTS * Q_HALF;
  static constexpr int K_HALF = DC * 128 * 128;
  static constexpr int K_BYTES = QK_PARTS * K_HALF;
  static constexpr int V_CHUNK = VMN ? 128 * 128 : ((DPAD * 128 + 1023) / 1024) * 1024;   // VMN: [128 keys][64 ch]; else [DPAD rows][64 keys]
  static constexpr int V_BYTES = (VMN ? DC : 2) * V_CHUNK;
  static constexpr int V_TX = VMN ? DC * 128 * 128 : 2 * DPAD * 128;    // bytes one V stage receives
  // P (the exponentials, fp16) goes to the PV product either through a swizzled shared-memory tile or — when the tensor memory has
  // room for 64 more columns per query tile — through TENSOR MEMORY as the A operand of tcgen05.mma: each softmax thread stores
  // its own row (128 fp16 = 64 columns of its lane) with tcgen05.st, no swizzle, no generic->async proxy fence, and the P tile
  // (32 KB written + 32 KB read per query tile and key tile) leaves shared memory, whose bandwidth bounded the d = 40 kernel
  // (QK^T + PV operands + P + TMA = 344 KB per key tile at 128 B/clk; profiles/r2_attention_timeline_after_elect_2stages.log)
  static constexpr bool PT = NG * (128 + DPAD + 64) <= 512;
  static constexpr int P_TILE = PT ? 0 : 2 * 128 * 128;                 // [128 rows][128 keys] fp16
  static constexpr int FIXED = NG * (Q_TILE + P_TILE) + 512 + 1024;
  // K/V pipeline stages: two when they fit beside the 1 KB of static shared memory (227 KB per CTA = 226 KB dynamic). The split-q/k
  // d = 40 pair-of-query-tiles variant needs 225.5 KB for two: with ONE stage the MMA warp waited ~690 clk per key tile for K(j+1)
  // (clock64 timeline, profiles/r2_attention_timeline_before.log)
  static constexpr int ST = (FIXED + 2 * (K_BYTES + V_BYTES) <= 226 * 1024) ? 2 : 1;
  static constexpr int SMEM = FIXED + ST * (K_BYTES + V_BYTES);
  static constexpr int TMEM_NEED = NG * (128 + DPAD + (PT ? 64 : 0));
  static constexpr int P_COL0 = NG * (128 + DPAD);                      // first P column (PT)
  static constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
  static_assert(TMEM_NEED <= 512, "TMEM budget");
  static_assert(SMEM <= 226 * 1024, "smem budget");
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// RS = true (register split, NG = 2 only): the block is padded to three full warpgroups — warps 0-3 = TMA producer, MMA issuer and
// two idle warps, warps 4-11 = the two softmax groups — so that setmaxnreg can move registers between them: the first warpgroup
// drops to 56 registers per thread, the softmax warpgroups rise to 224. With 10 warps ptxas budgets 65536 / 384 = 168 registers
// per thread (allocation is per 4 warps) and the 128 score registers of a softmax thread left 12 values spilled to local memory,
// ~20 reloads per key tile inside the exponentials loop (LDL in the SASS; ncu: long-scoreboard the top stall).
template <int DPAD, int NG, bool VMN, bool QK3, bool RS>
__global__ void __launch_bounds__((RS ? 128 : 64) + 128 * NG, 1)
attention_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                 const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mq_lo,
                 const __grid_constant__ CUtensorMap mk_lo, const AttnParams p) {
  using Cfg = AttnCfg<DPAD, NG, VMN, QK3>;
  constexpr int DC = Cfg::DC, ST = Cfg::ST;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // NG tiles
  uint8_t* sP = sQ + NG * Cfg::Q_TILE;             // NG tiles
  uint8_t* sK = sP + NG * Cfg::P_TILE;             // ST stages
  uint8_t* sV = sK + ST * Cfg::K_BYTES;            // ST stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * Cfg::V_BYTES);
  uint64_t* q_full = bars;             // 1
  uint64_t* k_full = q_full + 1;       // ST
  uint64_t* k_empty = k_full + ST;     // ST
  uint64_t* v_full = k_empty + ST;     // ST
  uint64_t* v_empty = v_full + ST;     // ST
  uint64_t* s_full = v_empty + ST;     // NG
  uint64_t* s_free = s_full + NG;      // NG (128 arrivals each)
  uint64_t* p_full = s_free + NG;      // NG (128 arrivals each)
  uint64_t* pv_done = p_full + NG;     // NG
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + NG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int SW0 = RS ? 4 : 2;  // first softmax warp
  long long* const dbg = (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.dbg : nullptr;
  constexpr int DJ0 = 8;  // stamped key tiles: DJ0 .. DJ0+3
  if (dbg && threadIdx.x == 0) dbg[255] = clock64();
  pdl_trigger();
  const int q0 = blockIdx.x * (128 * NG);
  const int h = blockIdx.y;
  const int s = blockIdx.z;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int g = 0; g < NG; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 128);
      mbar_init(&p_full[g], 128);
      mbar_init(&pv_done[g], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mq);
    tma_prefetch_desc(&mk);
    tma_prefetch_desc(&mv);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  // a device-side length of 0 (or beyond Nk) would leave T = 0 and the epilogue waiting forever: clamp to [1, Nk]
  const int kvlen = p.kvlen ? max(1, min(p.kvlen[s], p.Nk)) : p.Nk;
  const int T = (kvlen + 127) / 128;
  // columns: S_g at g*128 ; O_g at NG*128 + g*DPAD

  // (setmaxnreg sits INSIDE the role branches: ptxas budgets the code after a join with the smaller of the two limits)
  if (warp < SW0) {
  if (RS) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ======================================================================= TMA producer
    if (elect_one()) {
      mbar_expect_tx(q_full, NG * Cfg::Q_TILE);
#pragma unroll
      for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          tma_load_2d(sQ + g * Cfg::Q_TILE + c * 16384, &mq, q_full, p.q_col0 + h * DPAD + c * 64,
                      s * p.q_rows_per_sample + q0 + g * 128);
          if (QK3)
            tma_load_2d(sQ + g * Cfg::Q_TILE + Cfg::Q_HALF + c * 16384, &mq_lo, q_full, p.q_col0 + h * DPAD + c * 64,
                        s * p.q_rows_per_sample + q0 + g * 128);
        }
      for (int j = 0; j < T; ++j) {
        const int st = j % ST;
        const uint32_t ph = (j / ST) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::K_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          tma_load_2d(sK + st * Cfg::K_BYTES + c * 16384, &mk, &k_full[st], p.k_col0 + h * DPAD + c * 64,
                      s * p.k_rows_per_sample + j * 128);
          if (QK3)
            tma_load_2d(sK + st * Cfg::K_BYTES + Cfg::K_HALF + c * 16384, &mk_lo, &k_full[st], p.k_col0 + h * DPAD + c * 64,
                        s * p.k_rows_per_sample + j * 128);
        }
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], Cfg::V_TX);
        if (VMN) {
#pragma unroll
          for (int c = 0; c < DC; ++c)  // [128 keys][64 channels] boxes; channels past the head (or the matrix) are never multiplied
            tma_load_2d(sV + st * Cfg::V_BYTES + c * Cfg::V_CHUNK, &mv, &v_full[st], p.v_col0 + h * DPAD + c * 64,
                        s * p.k_rows_per_sample + j * 128);
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_2d(sV + st * Cfg::V_BYTES + c * Cfg::V_CHUNK, &mv, &v_full[st],
                        s * p.k_rows_per_sample + j * 128 + c * 64, h * p.d);
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    constexpr uint32_t idesc_s = make_idesc_f16(128, 128);
    constexpr uint32_t idesc_o = make_idesc_f16(128, DPAD, false, /*b_mn_major=*/VMN);
    auto mstamp = [&](int j, int g, int k) {
      if (dbg && lane == 0 && j >= DJ0 && j < DJ0 + 4) dbg[128 + (j - DJ0) * 16 + g * 8 + k] = clock64();
    };
    auto issue_qk = [&](int g, int j) {
      const int st = j % ST;
      mstamp(j - 1, g, 0);
      if (g == 0) mbar_wait(&k_full[st], (j / ST) & 1);
      mstamp(j - 1, g, 4);
      if (j > 0) mbar_wait(&s_free[g], (j - 1) & 1);  // group g has copied S_g(j-1) out of TMEM
      tc_fence_after();
      if (elect_one()) {
        const uint32_t qa = smem_u32(sQ + g * Cfg::Q_TILE), ka = smem_u32(sK + st * Cfg::K_BYTES);
#pragma unroll
        for (int kk = 0; kk < DPAD / 16; ++kk) {
          const uint32_t off = (kk / 4) * 16384 + (kk % 4) * 32;
          umma_f16(tmem_base + g * 128, make_sdesc_sw128(qa + off), make_sdesc_sw128(ka + off), idesc_s, kk > 0 ? 1u : 0u);
          if (QK3) {  // + q_lo k_hi^T + q_hi k_lo^T
            umma_f16(tmem_base + g * 128, make_sdesc_sw128(qa + Cfg::Q_HALF + off), make_sdesc_sw128(ka + off), idesc_s, 1u);
            umma_f16(tmem_base + g * 128, make_sdesc_sw128(qa + off), make_sdesc_sw128(ka + Cfg::K_HALF + off), idesc_s, 1u);
          }
        }
        if (g == NG - 1) umma_commit(&k_empty[st]);  // the K stage is free once the last group's QK retires
        umma_commit(&s_full[g]);
      }
      __syncwarp();
      mstamp(j - 1, g, 1);
    };
    mbar_wait(q_full, 0);
    for (int g = 0; g < NG; ++g) issue_qk(g, 0);
    for (int j = 0; j < T; ++j) {
      const int st = j % ST;
      for (int g = 0; g < NG; ++g) {
        // S_g(j+1) first: it only needs the group to have copied S_g(j) out of TMEM, so it runs while the group is
        // still in its exponentials and the next softmax never waits for the tensor pipe
        if (j + 1 < T) issue_qk(g, j + 1);
        if (g == 0) mbar_wait(&v_full[st], (j / ST) & 1);
        mstamp(j, g, 5);
        mbar_wait(&p_full[g], j & 1);
        mstamp(j, g, 2);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t pa = smem_u32(sP + g * Cfg::P_TILE), va = smem_u32(sV + st * Cfg::V_BYTES);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t poff = (kk / 4) * 16384 + (kk % 4) * 32;
            // K-major V^T: 16 keys = 32 bytes inside a 128-byte row; MN-major V: 16 keys = 16 rows of 128 bytes
            const uint32_t voff = VMN ? kk * 2048 : (kk / 4) * Cfg::V_CHUNK + (kk % 4) * 32;
            const uint64_t vdesc = VMN ? make_sdesc_sw128_mn(va + voff, Cfg::V_CHUNK) : make_sdesc_sw128(va + voff);
            if (Cfg::PT)  // A = P in tensor memory: 16 keys = 8 columns
              umma_f16_ts(tmem_base + NG * 128 + g * DPAD, tmem_base + Cfg::P_COL0 + g * 64 + kk * 8, vdesc, idesc_o,
                          (j > 0 || kk > 0) ? 1u : 0u);
            else
              umma_f16(tmem_base + NG * 128 + g * DPAD, make_sdesc_sw128(pa + poff), vdesc, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          }
          if (g == NG - 1) umma_commit(&v_empty[st]);
          umma_commit(&pv_done[g]);
        }
        __syncwarp();
        mstamp(j, g, 3);
      }
    }
  }
  } else {
    if (RS) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ======================================================================= softmax groups + epilogue
    const int g = (warp - SW0) >> 2;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_sel = uint32_t(qd * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_sel;
    const uint32_t tO = tmem_base + NG * 128 + g * DPAD + lane_sel;
    uint8_t* prow = sP + g * Cfg::P_TILE + r * 128;
    const uint32_t tP = tmem_base + Cfg::P_COL0 + g * 64 + lane_sel;  // this row's 64 P columns (PT)
    const float sl2 = p.scale * 1.4426950408889634f;  // d^-1/2 * log2(e)
    float m_run = -INFINITY, l_run = 0.f;
    auto sstamp = [&](int j, int k) {
      if (dbg && qd == 0 && lane == 0 && j >= DJ0 && j < DJ0 + 4) dbg[g * 64 + (j - DJ0) * 8 + k] = clock64();
    };
    for (int j = 0; j < T; ++j) {
      sstamp(j, 0);
      mbar_wait(&s_full[g], j & 1);
      sstamp(j, 1);
      tc_fence_after();
      uint32_t sv[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[g]);
      sstamp(j, 2);
      int valid = min(128, kvlen - j * 128);
      if (p.causal) valid = max(1, min(valid, q0 + g * 128 + r - j * 128 + 1));  // additive -inf mask above the diagonal
      // 8 independent max chains (a single chain of 128 dependent FMNMX would cost ~500 cycles of pure latency)
      float mxa[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) mxa[a] = -INFINITY;
      if (valid == 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i) mxa[i & 7] = fmaxf(mxa[i & 7], __uint_as_float(sv[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i < valid) mxa[i & 7] = fmaxf(mxa[i & 7], __uint_as_float(sv[i]));
      }
      const float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])), fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
      // Lazy rescale with a threshold: the running maximum is only a reference point (softmax is shift invariant), so it is moved —
      // and O_g / l rescaled — only when a row's maximum grew by more than 2^8; otherwise the tile is exponentiated against the
      // stale reference and P may reach 256 (exact in fp16, fp32 sums). With a plain `m_new > m_run` test one of the 32 rows of
      // a warp has a new maximum in ~97 % of the key tiles, i.e. the "lazy" rescale ran (3 x tcgen05.ld/st of O, ~340 clk) on
      // almost every tile (clock64 timeline, profiles/r2_attention_timeline_p_in_tmem.log).
      const float m_cand = fmaxf(m_run, mx * sl2);
      const bool resc = __any_sync(0xffffffffu, m_cand > m_run + 8.0f);  // first tile: m_run = -inf -> true
      const float m_new = resc ? m_cand : m_run;
      const float alpha = resc ? ex2(m_run - m_new) : 1.0f;  // 0 on the first tile
      sstamp(j, 3);
      if (j > 0) {
        mbar_wait(&pv_done[g], (j - 1) & 1);  // O_g holds PV(j-1); the P buffer is free again
        sstamp(j, 4);
        tc_fence_after();
        if (resc) {
#pragma unroll
          for (int c = 0; c < DPAD; c += 16) {
            uint32_t o[16];
            tmem_ld16(tO + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + c, o);
          }
          tmem_st_wait();
        }
      }
      sstamp(j, 5);
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};  // independent partial row sums (ILP), folded in fixed order below
      const uint32_t prow_s = smem_u32(prow);
      auto emit = [&](auto masked) {
        uint32_t pw[16];  // PT: 32 keys of this row, stored to tensor memory every fourth unit
        (void)pw;
#pragma unroll
        for (int u = 0; u < 16; ++u) {  // 16-byte units of 8 keys
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i0 = u * 8 + 2 * e;
            float p0 = ex2(fmaf(__uint_as_float(sv[i0]), sl2, -m_new));
            float p1 = ex2(fmaf(__uint_as_float(sv[i0 + 1]), sl2, -m_new));
            if (decltype(masked)::value) {
              p0 = (i0 < valid) ? p0 : 0.f;
              p1 = (i0 + 1 < valid) ? p1 : 0.f;
            }
            sum4[e] += p0 + p1;  // fp32 terms; the fp16 rounding of P is unbiased and averages out over the row
            const __half2 hp = __floats2half2_rn(p0, p1);
            w[e] = *reinterpret_cast<const uint32_t*>(&hp);
          }
          if (Cfg::PT) {
#pragma unroll
            for (int e = 0; e < 4; ++e) pw[(u & 3) * 4 + e] = w[e];
            if ((u & 3) == 3) tmem_st16(tP + (u >> 2) * 16, pw);
          } else {
            const int chunk = u >> 3, uu = u & 7;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow_s + chunk * 16384 + ((uu ^ (r & 7)) << 4)),
                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                         : "memory");
          }
        }
      };
      if (valid == 128)
        emit(std::false_type{});
      else
        emit(std::true_type{});
      const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      l_run = l_run * alpha + sum;
      m_run = m_new;
      sstamp(j, 6);
      if (Cfg::PT)
        tmem_st_wait();       // this thread's P columns are in tensor memory
      else
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      sstamp(j, 7);
    }
    // ---- epilogue: O / l -> fp16 hi(/lo)
    mbar_wait(&pv_done[g], (T - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int qrow = q0 + g * 128 + r;
    const bool ok = qrow < p.Nq;
    const size_t orow = (size_t)(s * p.q_rows_per_sample + qrow) * p.ldo + h * p.d;
#pragma unroll
    for (int c = 0; c < DPAD; c += 16) {
      uint32_t o[16];
      tmem_ld16(tO + c, o);
      tmem_ld_wait();
#pragma unroll
      for (int gg = 0; gg < 16; gg += 8) {
        if (ok && c + gg < p.d) {
          __half2 hh[4], hl[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float f0 = __uint_as_float(o[gg + 2 * e]) * inv_l, f1 = __uint_as_float(o[gg + 2 * e + 1]) * inv_l;
            hh[e] = __floats2half2_rn(f0, f1);
            const float2 hf = __half22float2(hh[e]);
            hl[e] = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
          }
          *reinterpret_cast<uint4*>(p.out_hi + orow + c + gg) = *reinterpret_cast<uint4*>(hh);
          if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + orow + c + gg) = *reinterpret_cast<uint4*>(hl);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// option "attn_regsplit": NG = 2 launches use the register-split variant (see attention_kernel). Off by default: it removes every
// spill (STACK 48 -> 0 bytes, no LDL / STL in the SASS) and is bit-identical, but measured neutral (143.38 ms per image with it,
// 143.30 ms without, tools/step_time.py): the spill reloads were not what the softmax warps wait for.
int g_attn_regsplit = 0;

template <int DPAD, int NG, bool VMN, bool QK3, bool RS>
static void launch_attn3(const AttnMaps& m, const AttnParams& p, cudaStream_t st) {
  constexpr int smem = AttnCfg<DPAD, NG, VMN, QK3>::SMEM;
  static DeviceOnce once;
  if (once.first())
    SDB_CUDA(cudaFuncSetAttribute(attention_kernel<DPAD, NG, VMN, QK3, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((p.Nq + 128 * NG - 1) / (128 * NG), p.heads, p.nb);
  launch_k(attention_kernel<DPAD, NG, VMN, QK3, RS>, grid, dim3((RS ? 128 : 64) + 128 * NG), (size_t)smem, st, m.q, m.k, m.v, m.q_lo,
           m.k_lo, p);
}
template <int DPAD, int NG, bool VMN, bool QK3>
static void launch_attn2(const AttnMaps& m, const AttnParams& p, cudaStream_t st) {
  if constexpr (NG == 2) {
    if (g_attn_regsplit) return launch_attn3<DPAD, NG, VMN, QK3, true>(m, p, st);
  }
  launch_attn3<DPAD, NG, VMN, QK3, false>(m, p, st);
}
template <int DPAD, int NG>
static void launch_attn(const AttnMaps& m, const AttnParams& p, cudaStream_t st) {
  if (p.v_mn)
    launch_attn2<DPAD, NG, true, false>(m, p, st);
  else
    launch_attn2<DPAD, NG, false, false>(m, p, st);
}

bool attention_supports_qk3(int dpad) { return dpad == 48 || dpad == 80; }

void attention_launch(const AttnMaps& m, const AttnParams& p, cudaStream_t st) {
  const bool two = p.Nq > 128;  // two ping-pong query tiles per CTA when there are at least two tiles of rows
  if (p.qk3) {
    // split q / k operands (levels 0-1 of the UNet: V always MN-major there). Shared memory: d = 40 keeps two query tiles and two
    // K/V stages (225.5 KB); d = 80 fits one query tile and one stage
    SDB_CHECK(p.v_mn && attention_supports_qk3(p.dpad), "split q/k attention: head dim / V layout");
    if (p.dpad == 48)
      two ? launch_attn2<48, 2, true, true>(m, p, st) : launch_attn2<48, 1, true, true>(m, p, st);
    else
      launch_attn2<80, 1, true, true>(m, p, st);
    return;
  }
  switch (p.dpad) {
    case 48:
      two ? launch_attn<48, 2>(m, p, st) : launch_attn<48, 1>(m, p, st);
      break;
    case 64:
      launch_attn<64, 1>(m, p, st);
      break;
    case 80:
      two ? launch_attn<80, 2>(m, p, st) : launch_attn<80, 1>(m, p, st);
      break;
    case 160:
      launch_attn<160, 1>(m, p, st);
      break;
    default:
      throw Error("attention: unsupported head dim " + std::to_string(p.dpad));
  }
}

}  // namespace sdb
