// gemm_tc.cu — tcgen05 implicit-GEMM for conv3x3 / conv1x1 / Linear on sm_100a.
//
// Replaces burn nn::conv::Conv2d / nn::Linear on the reference hot path
// (call sites: src/model/unet/mod.rs:716,726,729 ResBlock convs; :468,479 proj_in/out;
//  :645-651 q/k/v/out; :580,553 GEGLU/ff; src/model/autoencoder/mod.rs:513-528, 567-606).
//
// One CTA = one 128 x BN output tile (x one K split); with CG = 2 a CTA pair shares one 256 x BN MMA. Warp roles:
//   warp 0   : TMA producer  (cp.async.bulk.tensor 5-D activation boxes + 2-D weight boxes)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (fp16 x fp16 -> fp32 in TMEM)
//   warps 2-9: epilogue (tcgen05.ld -> smem transpose -> bias / time-embedding row / residual / GEGLU -> global)
// Multi-pass products (PASSES = 2, 3) add the low-order fp16 halves of the operands
// (A_lo*B_hi, A_hi*B_lo) into the same accumulator for fp32-class accuracy.
#include "gemm_tc.cuh"

#include <cstring>

namespace sdb {

static constexpr int BM = 128;
static constexpr int BK = 64;  // fp16 elements per k chunk = 128 bytes = one swizzle row
static constexpr int A_TILE_BYTES = BM * BK * 2;

// CG = 1: one CTA per 128 x BN tile. CG = 2: a CTA pair (cluster of 2 along M) runs ONE tcgen05.mma.cta_group::2
// of shape 256 x BN per k-step: each CTA stages its own 128 A rows and only HALF of the weight tile, so the operand
// bytes per FLOP that cross the L2->SM fabric drop by 25-45 % (the bound measured in profiles/r1_gemm_tc_ncu_full.md).
template <int BN, int PASSES, int CG = 1>
struct StageLayout {
  static constexpr int B_TILE_BYTES = (BN / CG) * BK * 2;
  static constexpr int A_TILES = PASSES >= 2 ? 2 : 1;
  static constexpr int B_TILES = PASSES >= 3 ? 2 : 1;
  static constexpr int BYTES = A_TILES * A_TILE_BYTES + B_TILES * B_TILE_BYTES;
};

// Activation + fp32 / fp16(hi,lo) stores of 4 consecutive columns of one output row. Deliberately NOT inlined: the epilogue
// runs once per CTA, so its cost is dominated by cold instruction fetch (ncu: stall_no_inst); one shared copy of this
// body instead of one per unrolled row keeps the epilogue's code footprint small.
__device__ __noinline__ void epilogue_store(float4 f, unsigned int o32, unsigned int o16, float* out_f32, __half* out_f16,
                                            __half* out_f16_lo, int act) {
  if (act == 1) {
    f.x = __fdividef(f.x, 1.0f + __expf(-1.702f * f.x)), f.y = __fdividef(f.y, 1.0f + __expf(-1.702f * f.y));
    f.z = __fdividef(f.z, 1.0f + __expf(-1.702f * f.z)), f.w = __fdividef(f.w, 1.0f + __expf(-1.702f * f.w));
  }
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + o32) = f;
  if (out_f16) {
    __half2 h[2] = {f2h2_sat(f.x, f.y), f2h2_sat(f.z, f.w)};
    *reinterpret_cast<uint2*>(out_f16 + o16) = *reinterpret_cast<uint2*>(h);
    if (out_f16_lo) {
      const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
      __half2 l[2] = {__floats2half2_rn(f.x - f0.x, f.y - f0.y), __floats2half2_rn(f.z - f1.x, f.w - f1.y)};
      *reinterpret_cast<uint2*>(out_f16_lo + o16) = *reinterpret_cast<uint2*>(l);
    }
  }
}

// Bucket reduction of the per-quarter column sums an epilogue left in shared memory (cs[q][col] = (sum, sumsq) of the 32 rows of
// lane quarter q), written as this tile's GroupNorm partials. `nq` quarters per image slice (4 = the whole tile is one image).
// which images the 128 rows of this tile belong to: g_tn images per tile (1, 2 or 4), the first one, the tile's index inside it
struct GnTile {
  int tn, img0, tile, nimg;
};
__device__ __forceinline__ GnTile gn_tile_of(const GemmParams& p, int n0, int th, int tw) {
  GnTile g;
  if (p.gn_rpi) {  // flattened rows: blockIdx.x is the row tile
    g.nimg = p.gn_nimg;
    if (p.gn_rpi >= BM)
      g.tn = 1, g.img0 = (tw * BM) / p.gn_rpi, g.tile = ((tw * BM) % p.gn_rpi) / BM;
    else
      g.tn = BM / p.gn_rpi, g.img0 = tw * g.tn, g.tile = 0;
  } else {
    g.tn = p.TN, g.img0 = n0, g.tile = th * p.tiles_w + tw, g.nimg = p.nimg;
  }
  return g;
}
__device__ __forceinline__ void gn_write_partials(const GemmParams& p, const float2* cs, int BN, int te, int nthreads, const GnTile g,
                                                  int col0, int z) {
  const int nbk_tile = BN / p.gn_bucket, nbk_total = p.N / p.gn_bucket;
  const int qpi = 4 / g.tn;  // lane quarters per image (1, 2 or 4 images per tile: checked by run_gemm)
  const int slot = p.gn_slot0 + g.tile * p.split_k + z;
  // thread = (bucket, quarter): 4 adjacent lanes hold the quarters of one bucket and combine them with shuffles in a fixed
  // order (nbk_tile * 4 <= 256 threads: BN <= 256, bucket >= 4); whole warps take part so that the shuffles are convergent
  (void)nthreads;
  if (te < ((nbk_tile * 4 + 31) & ~31)) {
    const int b = te >> 2, q = te & 3;
    float sm = 0.f, sq = 0.f;
    if (b < nbk_tile) {
      const float2* src = cs + q * BN + b * p.gn_bucket;
      for (int c = 0; c < p.gn_bucket; ++c) sm += src[c].x, sq += src[c].y;
    }
    if (qpi >= 2) sm += __shfl_xor_sync(0xffffffffu, sm, 1), sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    if (qpi == 4) sm += __shfl_xor_sync(0xffffffffu, sm, 2), sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    const int k = q / qpi, img = g.img0 + k, c0 = b * p.gn_bucket;
    if (b < nbk_tile && (q % qpi) == 0 && img < g.nimg && col0 + c0 < p.N) {
      float2* dst = reinterpret_cast<float2*>(p.gn_part) + ((size_t)img * p.gn_cap + slot) * nbk_total + (col0 + c0) / p.gn_bucket;
      *dst = make_float2(sm, sq);
    }
  }
}

// Variants whose pipeline fits twice in an SM's shared memory are launched two CTAs per SM (<= 102 registers per thread);
// the others own the SM and may use the whole register file.
template <int BN, int PASSES, int STAGES, int CG>
__host__ __device__ constexpr int min_ctas_per_sm() {
  return STAGES * StageLayout<BN, PASSES, CG>::BYTES <= 108 * 1024 ? 2 : 1;
}

// Epilogue warps (a multiple of 4: one group per TMEM lane quarter). Measured: 16 warps on the SM-owning variants do not
// shorten the epilogue (6.4k vs 6.7k cycles for 128 x 160) and lengthen the tail, so those use 8. The variants that share an
// SM between two CTAs use 4: with 8 they were capped at 96 registers per thread and spilled 140-950 bytes per thread, and with
// the shared-memory carve-out at its maximum L1 holds nothing, so every spill access is an L2 round trip (ncu on the q|k|v
// projection: 225 K local loads + 152 K local stores, 48 MB of local traffic for 19 MB of output, long-scoreboard the top stall).
// 4 warps (192 threads per CTA) leave 170 registers per thread: no spills.
// The GEGLU epilogue is the exception: it is bound by its own arithmetic (erf-GELU on every output), 16 warps per SM beat 8
// even at 96 registers (measured on the level-0 GEGLU projection: 8 warps 74 us, 4 warps 90 us), and as a compile-time variant
// (EPI_GEGLU*) it no longer carries the other epilogues' registers.
template <int BN, int PASSES, int STAGES, int CG, int EPI>
__host__ __device__ constexpr int epilogue_warps() {
  return (min_ctas_per_sm<BN, PASSES, STAGES, CG>() == 2 && EPI < 4) ? 4 : 8;
}

// EPI selects what the epilogue does beside bias / residual / stores. It is a compile-time choice because the once-per-CTA
// epilogue is instruction-issue bound: statistics code that is merely skipped at run time still cost ~0.5 us per launch.
//   EPI_PLAIN  nothing more          EPI_GN   GroupNorm statistics of the output tensor (column sums per channel bucket)
//   EPI_LNS    LayerNorm row statistics of the output rows (partial sum / sum of squares per N-tile share)
//   EPI_LNC    the A operand is the RAW input of a LayerNorm whose gamma is folded into the weights: the normalisation is applied
//              here as a rank-1 correction, out = rstd_r * (acc - mean_r * u_c) + v_c  (u = column sums of the folded weights,
//              v = beta^T W + bias arrives as `bias`), from the row statistics the producer of A left (EPI_LNS)
//   EPI_GEGLU / EPI_GEGLU_LNC   x * gelu_erf(gate) on column-interleaved (x | gate) tiles (unet/mod.rs:578-592), without / with
//              the LayerNorm-consuming correction
enum : int { EPI_PLAIN = 0, EPI_GN = 1, EPI_LNS = 2, EPI_LNC = 3, EPI_GEGLU = 4, EPI_GEGLU_LNC = 5 };

template <int BN, int PASSES, int STAGES, int CG, int EPI>
__global__ void __launch_bounds__(64 + 32 * epilogue_warps<BN, PASSES, STAGES, CG, EPI>(), min_ctas_per_sm<BN, PASSES, STAGES, CG>())
gemm_tc_kernel(const __grid_constant__ GemmMaps maps, const GemmParams p) {
  constexpr bool kGN = EPI == EPI_GN, kLNS = EPI == EPI_LNS, kLNC = EPI == EPI_LNC || EPI == EPI_GEGLU_LNC;
  constexpr bool kGEGLU = EPI == EPI_GEGLU || EPI == EPI_GEGLU_LNC;
  using L = StageLayout<BN, PASSES, CG>;
  constexpr bool TWO = CG == 2;
  constexpr int EW = epilogue_warps<BN, PASSES, STAGES, CG, EPI>();  // epilogue warps
  constexpr int EG = EW / 4;                                    // warps sharing one TMEM lane quarter
  constexpr int CSTEP = 32 * EG;                                // column stride between the chunks of one warp
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * L::BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* const dbg = (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.dbg : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  if (!p.pdl_late) pdl_trigger();  // the next kernel of the stream may begin its own prologue

  // ---- tile coordinates
  int mt = blockIdx.x;
  const int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th = mt % p.tiles_h;
  const int tn = mt / p.tiles_h;
  const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tn * p.TN;
  const int col0 = blockIdx.y * BN;

  const int main_iters = p.num_taps * p.kc;
  const int total_iters = main_iters + p.xkc;
  const int per_split = (total_iters + p.split_k - 1) / p.split_k;
  // grid.z is the K split — or, for the folded nearest-2x upsample conv (p.up2, never split), the output phase (a, b): the four
  // 2x2-tap phase convolutions of one layer run as ONE launch; phase shifts the taps, the weight rows and the output pixel
  const int up_a = p.up2 ? (int)(blockIdx.z >> 1) : 0, up_b = p.up2 ? (int)(blockIdx.z & 1) : 0;
  const int kz = p.up2 ? 0 : (int)blockIdx.z;
  const int b_row0 = col0 + (p.up2 ? (int)blockIdx.z * p.N : 0);  // first weight row of this tile (phase-major packing)
  const int it_begin = kz * per_split;
  const int it_end = min(total_iters, it_begin + per_split);

  const uint32_t crank = TWO ? cluster_ctarank() : 0;
  const bool leader = crank == 0;  // the CTA that issues the pair's MMAs and owns the "full" barriers
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0][0]);
    tma_prefetch_desc(&maps.b[0]);
    if (p.kc0 < p.kc) tma_prefetch_desc(&maps.a[1][0]);
  }
  constexpr uint32_t TMEM_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);  // power of two >= BN
  if (warp == 1) {
    if (TWO) {
      tmem_alloc2(tmem_slot, TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (TWO)
    cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail. From here on each role waits for the
  // previous kernel (griddepcontrol.wait) only where it first touches data that kernel may have written:
  //   producer : WEIGHTS are immutable, so the weight tiles of the first STAGES k-chunks (and an L2 prefetch of the rest of this
  //              CTA's weight strip when the launch is weight-bound) are issued BEFORE the wait; activation tiles after it
  //   MMA warp : consumes shared memory behind the mbarriers only: no wait
  //   epilogue : waits before its first read of residual / time-embedding rows; all of this kernel's stores follow that wait
  if (warp == 0) {
    // ===================================================== TMA producer (one elected lane: see elect_one in common.cuh)
    if (elect_one()) {
      // iteration -> (activation source, channel chunk, tap shift) and the weight map / K coordinate that go with it
      auto load_b = [&](int it, int s) {
        const CUtensorMap* bm = it < main_iters ? maps.b : maps.bx;
        const int bk = (it < main_iters ? it : it - main_iters) * BK;
        uint8_t* sb = smem + s * L::BYTES + L::A_TILES * A_TILE_BYTES;
        if (!TWO) {
          tma_load_2d(sb, &bm[0], &full_bar[s], bk, b_row0);
          if (PASSES >= 3) tma_load_2d(sb + L::B_TILE_BYTES, &bm[1], &full_bar[s], bk, b_row0);
        } else {
          // rows [crank*BN/2, +BN/2) of the weight tile into this CTA's smem, bytes reported to the leader's barrier
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[s]), 0);
          tma_load_2d_2sm(sb, &bm[0], fb, bk, b_row0 + crank * (BN / 2));
          if (PASSES >= 3) tma_load_2d_2sm(sb + L::B_TILE_BYTES, &bm[1], fb, bk, b_row0 + crank * (BN / 2));
        }
      };
      auto load_a = [&](int it, int s) {
        int src, c0, cw, ch, cp;
        if (it < main_iters) {
          const int tap = it / p.kc;
          const int cc = it - tap * p.kc;
          src = cc >= p.kc0 ? 1 : 0;
          c0 = (cc - (src ? p.kc0 : 0)) * BK;
          cw = w0 + p.tap_dw[tap] + up_b, ch = h0 + p.tap_dh[tap] + up_a, cp = p.tap_ph[tap];
        } else {
          const int e = it - main_iters;
          src = e >= p.xkc0 ? 3 : 2;
          c0 = (e - (src == 3 ? p.xkc0 : 0)) * BK;
          cw = w0, ch = h0, cp = 0;
        }
        uint8_t* st = smem + s * L::BYTES;
        if (!TWO) {
          tma_load_5d(st, &maps.a[src][0], &full_bar[s], c0, cw, ch, cp, n0);
          if (PASSES >= 2) tma_load_5d(st + A_TILE_BYTES, &maps.a[src][1], &full_bar[s], c0, cw, ch, cp, n0);
        } else {
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[s]), 0);
          tma_load_5d_2sm(st, &maps.a[src][0], fb, c0, cw, ch, cp, n0);
          if (PASSES >= 2) tma_load_5d_2sm(st + A_TILE_BYTES, &maps.a[src][1], fb, c0, cw, ch, cp, n0);
        }
      };
      // ---- before the wait: weights only (the pipeline slots are all free: fresh barriers)
      const int npre = min(STAGES, it_end - it_begin);
      for (int i = 0; i < npre; ++i) {
        if (leader) mbar_expect_tx(&full_bar[i], L::BYTES * CG);  // A + B bytes of both CTAs report to the leader's barrier
        load_b(it_begin + i, i);
      }
      if (p.prefetch_w) {
        for (int it = it_begin + npre; it < it_end; ++it) {
          const CUtensorMap* bm = it < main_iters ? maps.b : maps.bx;
          const int bk = (it < main_iters ? it : it - main_iters) * BK;
          const int row = b_row0 + (TWO ? crank * (BN / 2) : 0);
          tma_prefetch_l2_2d(&bm[0], bk, row);
          if (PASSES >= 3) tma_prefetch_l2_2d(&bm[1], bk, row);
        }
      }
      pdl_wait();
      for (int i = 0; i < npre; ++i) load_a(it_begin + i, i);
      if (dbg) dbg[2] = clock64();
      // ---- steady state
      int s = npre == STAGES ? 0 : npre;
      uint32_t ph = npre == STAGES ? 1 : 0;
      for (int it = it_begin + npre; it < it_end; ++it) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (leader) mbar_expect_tx(&full_bar[s], L::BYTES * CG);
        load_a(it, s);
        load_b(it, s);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (pair leader only when CG = 2)
    constexpr uint32_t idesc = make_idesc_f16(BM * CG, BN);
    int s = 0;
    uint32_t ph = 0;
    for (int it = it_begin; leader && it < it_end; ++it) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (dbg && lane == 0 && it == it_begin) dbg[3] = clock64();
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(smem + s * L::BYTES);
        const uint32_t a_lo = a_hi + A_TILE_BYTES;
        const uint32_t b_hi = a_hi + L::A_TILES * A_TILE_BYTES;
        const uint32_t b_lo = b_hi + L::B_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint32_t koff = k * 32;  // 16 fp16 = 32 bytes inside the 128B swizzle row
          const uint64_t da = make_sdesc_sw128(a_hi + koff);
          const uint64_t db = make_sdesc_sw128(b_hi + koff);
          const uint32_t acc = (it > it_begin || k > 0) ? 1u : 0u;
          if (TWO) {
            umma_f16_2sm(tmem_base, da, db, idesc, acc);
            if (PASSES >= 2) umma_f16_2sm(tmem_base, make_sdesc_sw128(a_lo + koff), db, idesc, 1u);
            if (PASSES >= 3) umma_f16_2sm(tmem_base, da, make_sdesc_sw128(b_lo + koff), idesc, 1u);
          } else {
            umma_f16(tmem_base, da, db, idesc, acc);
            if (PASSES >= 2) umma_f16(tmem_base, make_sdesc_sw128(a_lo + koff), db, idesc, 1u);
            if (PASSES >= 3) umma_f16(tmem_base, da, make_sdesc_sw128(b_lo + koff), idesc, 1u);
          }
        }
        if (TWO) {
          umma_commit_2sm(&empty_bar[s], 0x3);                     // release the slot in both CTAs of the pair
          if (it == it_end - 1) umma_commit_2sm(accum_bar, 0x3);   // both epilogues may start
        } else {
          umma_commit(&empty_bar[s]);                    // frees the smem slot when these MMAs retire
          if (it == it_end - 1) umma_commit(accum_bar);  // accumulator complete
        }
        if (dbg && it == it_end - 1) dbg[4] = clock64();
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..9)
    pdl_wait();  // residual / time-embedding rows below may come from the previous kernel; every store of this kernel follows
    // Each thread owns one accumulator row in TMEM (warp w may touch lanes 32*(w%4)..+31); the two warps that
    // share a lane quarter split the 32-column chunks between them. A row-per-thread store pattern would touch 32
    // cache lines per instruction, so every 32x32 block is transposed through shared memory (the pipeline stages
    // are idle by now; rows padded to 144 B keep both the 128-bit writes and reads bank-conflict free) and
    // written/read as 4 rows x 128 contiguous bytes per warp instruction.
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // which share of the column chunks this warp owns (0..EG-1)
    const int r = q * 32 + lane;        // accumulator row
    const int pw = w0 + r % p.TW;
    const int phh = h0 + (r / p.TW) % p.TH;
    const int pn = n0 + r / (p.TW * p.TH);
    const int row_ok = ((pw < p.W) && (phh < p.H) && (pn < p.nimg)) ? 1 : 0;
    // output row index (< 2^31 rows); -1 marks a row outside the tensor
    const int m = row_ok ? ((pn * p.OH + phh * p.os + p.oa + up_a) * p.OW + pw * p.os + p.ob + up_b) : -1;

    // ---- work that needs no accumulator, done while the main loop runs: the element offsets of the 8 rows this lane
    // serves in every column chunk (rr = 4 i + sub), the bias of each chunk, and the first chunk's addends (residual or
    // time-embedding row; run_gemm guarantees at most one of them). The epilogue is a chain of L2 round trips
    // (~650 cycles each, measured with clock64 stamps): everything issued here is off that chain.
    const int sub = lane >> 3;          // row within a group of 4
    const int cq = (lane & 7) * 4;      // 4-column group inside the 32-column chunk
    constexpr int NCHUNK = (BN + CSTEP - 1) / CSTEP;
    int mr8[8], ao[8];
    float4 bvs[NCHUNK], ad[8];
    const float* ad_ptr = p.residual ? p.residual : p.rowbias;
    const bool res_pair = p.res_hi != nullptr;  // the residual lives as an fp16 hi + lo pair (row stride ldc16)
    const bool plain = p.split_k == 1 && !kGEGLU;
    // EPI_LNC: mean / rstd of this lane's own accumulator row, from the partial row sums the producer of A left
    float ln_mu = 0.f, ln_rs = 0.f;
    float mu8[8], rs8[8];
    float4 us[NCHUNK];
    if constexpr (kLNC) {
      if (m >= 0) {
        const float2* sp = reinterpret_cast<const float2*>(p.ln_in) + (size_t)m * p.ln_in_slots;
        float sm = 0.f, sq = 0.f;
        for (int i = 0; i < p.ln_in_slots; ++i) {
          const float2 v = sp[i];
          sm += v.x, sq += v.y;
        }
        const float inv = 1.0f / (float)p.ln_C;
        ln_mu = sm * inv;
        ln_rs = rsqrtf(fmaxf(sq * inv - ln_mu * ln_mu, 0.f) + p.ln_eps);
      }
    }
    {
      const bool use_res = p.residual != nullptr;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + sub;
        const int mr = __shfl_sync(0xffffffffu, m, rr);
        const int pnr = __shfl_sync(0xffffffffu, pn, rr);
        mr8[i] = mr;
        ao[i] = res_pair ? mr * p.ldc16 : (use_res ? mr * p.ldc : pnr * p.N);
        ad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (kLNC) mu8[i] = __shfl_sync(0xffffffffu, ln_mu, rr), rs8[i] = __shfl_sync(0xffffffffu, ln_rs, rr);
      }
      if constexpr (kLNC) {
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) {
          const int col = col0 + half * 32 + j * CSTEP + cq;
          us[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (half * 32 + j * CSTEP < BN && col < p.N) us[j] = *reinterpret_cast<const float4*>(p.ln_u + col);
        }
      }
#pragma unroll
      for (int j = 0; j < NCHUNK; ++j) {
        const int col = col0 + half * 32 + j * CSTEP + cq;
        bvs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (plain && p.bias && half * 32 + j * CSTEP < BN && col < p.N) bvs[j] = *reinterpret_cast<const float4*>(p.bias + col);
      }
    }
    auto issue_addends = [&](int col) {
      if constexpr (kLNC) return;  // a LayerNorm-consuming GEMM has neither residual nor row bias (run_gemm checks): no registers for them
      if (col >= p.N) return;
      if (res_pair) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (mr8[i] >= 0) {
            const uint2 h = *reinterpret_cast<const uint2*>(p.res_hi + (unsigned)(ao[i] + col));
            const uint2 l = *reinterpret_cast<const uint2*>(p.res_lo + (unsigned)(ao[i] + col));
            const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
            const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
            ad[i] = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
          }
        return;
      }
      if (ad_ptr == nullptr) return;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mr8[i] >= 0) ad[i] = *reinterpret_cast<const float4*>(ad_ptr + (unsigned)(ao[i] + col));
    };
    // the residual may be produced by the previous kernel (always complete: stream order) or be this launch's own
    // output buffer written by an EARLIER launch (in-place accumulate): both are safe to read before the MMAs finish
    const bool pre_issued = plain;
    if (pre_issued) issue_addends(col0 + half * 32 + cq);

    mbar_wait(accum_bar, 0);
    tc_fence_after();
    if (p.pdl_late) pdl_trigger();
    if (dbg && threadIdx.x == 64) dbg[5] = clock64();
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    constexpr uint32_t TROW = 144;                                   // padded row pitch of the staging tile (bytes)
    const uint32_t tile_s = smem_u32(smem) + (warp - 2) * (32 * TROW);  // one 32-row tile per warp (4.6 KB each)
    const uint32_t tile2_s = tile_s + EW * 32 * TROW;                    // second bank, GEGLU only (x | gate)

    auto stage = [&](uint32_t t, const uint32_t (&v)[32]) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t + lane * TROW + k * 16), "r"(v[4 * k]),
                     "r"(v[4 * k + 1]), "r"(v[4 * k + 2]), "r"(v[4 * k + 3])
                     : "memory");
    };
    auto unstage = [&](uint32_t t, int rr) {
      float4 f;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w)
                   : "r"(t + rr * TROW + cq * 4)
                   : "memory");
      return f;
    };
    // element offsets fit 32 bits (run_gemm checks rows * ld < 2^31): one IMAD per row instead of 64-bit address chains
    // GroupNorm statistics of the output (p.gn_part): per-quarter column sums, staged in the second transposition bank
    // (which only the GEGLU epilogue uses; run_gemm never combines the two)
    float2* const gn_cs = reinterpret_cast<float2*>(smem + EW * 32 * TROW);
    const GnTile gnt = gn_tile_of(p, n0, th, tw);
    auto store_out = [&](float4 f, int mr, int col) {
      epilogue_store(f, (unsigned)(mr * p.ldc + col), (unsigned)(mr * p.ldc16 + col), p.out_f32, p.out_f16, p.out_f16_lo, p.act);
    };

    if (!kGEGLU && p.split_k > 1) {
      // raw partial sums -> workspace [split][M][N]
      const size_t Mtot = (size_t)p.nimg * p.OH * p.OW;
      float* wsbase = p.ws + (size_t)kz * Mtot * p.N;
#pragma unroll 1
      for (int c = half * 32; c < BN; c += CSTEP) {
        uint32_t v[32];
        tmem_ld32(trow + c, v);
        tmem_ld_wait();
        stage(tile_s, v);
        __syncwarp();
        const int col = col0 + c + cq;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + sub;
          const int mr = __shfl_sync(0xffffffffu, m, rr);
          if (mr >= 0 && col < p.N) __stcg(reinterpret_cast<float4*>(wsbase + (size_t)mr * p.N + col), unstage(tile_s, rr));
        }
        __syncwarp();
      }
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");  // all epilogue warps have published their part of the tile
      // Tile-level rendezvous of the split_k CTAs (all co-resident: run_gemm keeps ctas*split within one wave), then
      // every CTA folds its own slice of the tile rows in z order (deterministic) and runs the epilogue on it.
      unsigned int* tk = p.tickets + 2 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
      if (warp == 2 && lane == 0) {
        atomicAdd(tk, 1u);
        const long long t0 = clock64();
        while (atomicAdd(tk, 0u) < (unsigned)p.split_k) {
          __nanosleep(32);
          if (clock64() - t0 > 4000000000ll) __trap();
        }
        __threadfence();
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
      {
        const int rows_per = (BM + p.split_k - 1) / p.split_k;
        const int r0 = kz * rows_per, r1 = min(BM, r0 + rows_per);
        constexpr int C4 = BN / 4;
        // thread -> (row lane, fixed 4-column group): a thread's GroupNorm column sums stay in registers across its rows
        constexpr int RL = (EW * 32) / C4;  // row lanes
        const int te = threadIdx.x - 64;
        const int rlane = te / C4, cg4 = te - rlane * C4;
        const int rpi = BM / gnt.tn;  // rows per image inside the tile
        float4* const gn_red = reinterpret_cast<float4*>(smem + EW * 32 * TROW);  // [RL][TN][C4][2] float4, <= 32 KB (second bank)
#pragma unroll 1
        for (int k = 0; k < gnt.tn; ++k) {
        float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f), gsq = gsum;
        const int row_a = max(r0, k * rpi), row_b = min(r1, (k + 1) * rpi);
#pragma unroll 1
        for (int rl = row_a + rlane; rlane < RL && rl < row_b; rl += RL) {
          const int col = col0 + cg4 * 4;
          const int qw = w0 + rl % p.TW, qh = h0 + (rl / p.TW) % p.TH, qn = n0 + rl / (p.TW * p.TH);
          if (qw < p.W && qh < p.H && qn < p.nimg && col < p.N) {
            const int mr = (qn * p.OH + qh * p.os + p.oa) * p.OW + qw * p.os + p.ob;
            // the addends first, then the partials four at a time: independent loads in flight together, summed in z order
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), rb = bv, rs = bv;
            if (p.bias) bv = *reinterpret_cast<const float4*>(p.bias + col);
            if (p.rowbias) rb = *reinterpret_cast<const float4*>(p.rowbias + (size_t)qn * p.N + col);
            if (p.residual) rs = *reinterpret_cast<const float4*>(p.residual + (size_t)mr * p.ldc + col);
            if (p.res_hi) {
              const uint2 h = *reinterpret_cast<const uint2*>(p.res_hi + (size_t)mr * p.ldc16 + col);
              const uint2 l = *reinterpret_cast<const uint2*>(p.res_lo + (size_t)mr * p.ldc16 + col);
              const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
              const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
              rs = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
            }
            const float* wp = p.ws + (size_t)mr * p.N + col;
            const size_t zs = Mtot * p.N;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int z = 0;
#pragma unroll 1
            for (; z + 4 <= p.split_k; z += 4) {
              const float4 v0 = __ldcg(reinterpret_cast<const float4*>(wp + (size_t)z * zs));
              const float4 v1 = __ldcg(reinterpret_cast<const float4*>(wp + (size_t)(z + 1) * zs));
              const float4 v2 = __ldcg(reinterpret_cast<const float4*>(wp + (size_t)(z + 2) * zs));
              const float4 v3 = __ldcg(reinterpret_cast<const float4*>(wp + (size_t)(z + 3) * zs));
              acc.x += v0.x, acc.y += v0.y, acc.z += v0.z, acc.w += v0.w;
              acc.x += v1.x, acc.y += v1.y, acc.z += v1.z, acc.w += v1.w;
              acc.x += v2.x, acc.y += v2.y, acc.z += v2.z, acc.w += v2.w;
              acc.x += v3.x, acc.y += v3.y, acc.z += v3.z, acc.w += v3.w;
            }
#pragma unroll 1
            for (; z < p.split_k; ++z) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(wp + (size_t)z * zs));
              acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
            }
            acc.x += bv.x + rb.x + rs.x, acc.y += bv.y + rb.y + rs.y;
            acc.z += bv.z + rb.z + rs.z, acc.w += bv.w + rb.w + rs.w;
            if constexpr (kGN) {
              gsum.x += acc.x, gsum.y += acc.y, gsum.z += acc.z, gsum.w += acc.w;
              gsq.x = fmaf(acc.x, acc.x, gsq.x), gsq.y = fmaf(acc.y, acc.y, gsq.y), gsq.z = fmaf(acc.z, acc.z, gsq.z), gsq.w = fmaf(acc.w, acc.w, gsq.w);
            }
            store_out(acc, mr, col);
          }
        }
        if (kGN && rlane < RL) {
          gn_red[((rlane * gnt.tn + k) * C4 + cg4) * 2] = gsum;
          gn_red[((rlane * gnt.tn + k) * C4 + cg4) * 2 + 1] = gsq;
        }
        }
        if constexpr (kGN) {
          // fold the row lanes in a fixed order, then the channel buckets: this CTA's partial for its slice of the tile rows
          asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
          const int nbk_tile = BN / p.gn_bucket, nbk_total = p.N / p.gn_bucket;
          const int slot = p.gn_slot0 + gnt.tile * p.split_k + kz;
          const float* red = reinterpret_cast<const float*>(gn_red);
          for (int it = te; it < gnt.tn * nbk_tile; it += EW * 32) {
            const int k = it / nbk_tile, b = it - k * nbk_tile;
            const int img = gnt.img0 + k, c0 = b * p.gn_bucket;
            if (img >= gnt.nimg || col0 + c0 >= p.N) continue;
            float sm = 0.f, sq = 0.f;
            for (int l = 0; l < RL; ++l)
              for (int cc = c0; cc < c0 + p.gn_bucket; ++cc) {
                const int base = (((l * gnt.tn + k) * C4 + (cc >> 2)) * 2) * 4 + (cc & 3);
                sm += red[base], sq += red[base + 4];
              }
            float2* dst = reinterpret_cast<float2*>(p.gn_part) + ((size_t)img * p.gn_cap + slot) * nbk_total + (col0 + c0) / p.gn_bucket;
            *dst = make_float2(sm, sq);
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
      if (warp == 2 && lane == 0) {
        // last CTA to finish resets both counters: the buffer is all zero again for the next launch
        if (atomicAdd(tk + 1, 1u) == (unsigned)(p.split_k - 1)) {
          tk[0] = 0u;
          tk[1] = 0u;
        }
      }
    } else if (kGEGLU) {
      // tile columns [0,BN/2) = x, [BN/2,BN) = gate; output columns blockIdx.y*BN/2 + [0,BN/2)
      constexpr int HB = BN / 2;
      const int ocol0 = blockIdx.y * HB;
#pragma unroll 1
      for (int c = half * 32; c < HB; c += CSTEP) {
        uint32_t v[32];
        tmem_ld32(trow + c, v);
        tmem_ld_wait();
        stage(tile_s, v);
        tmem_ld32(trow + HB + c, v);
        tmem_ld_wait();
        stage(tile2_s, v);
        __syncwarp();
        const float4 bx = *reinterpret_cast<const float4*>(p.bias + col0 + c + cq);
        const float4 bg = *reinterpret_cast<const float4*>(p.bias + col0 + HB + c + cq);
        float4 ux = make_float4(0.f, 0.f, 0.f, 0.f), ug = ux;
        if constexpr (kLNC) {
          ux = *reinterpret_cast<const float4*>(p.ln_u + col0 + c + cq);
          ug = *reinterpret_cast<const float4*>(p.ln_u + col0 + HB + c + cq);
        }
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + sub;
          const int mr = __shfl_sync(0xffffffffu, m, rr);
          float rmu = 0.f, rrs = 1.f;
          if constexpr (kLNC) rmu = __shfl_sync(0xffffffffu, ln_mu, rr), rrs = __shfl_sync(0xffffffffu, ln_rs, rr);
          if (mr >= 0) {
            float4 tx = unstage(tile_s, rr), tg = unstage(tile2_s, rr);
            if constexpr (kLNC) {
              tx.x = rrs * (tx.x - rmu * ux.x), tx.y = rrs * (tx.y - rmu * ux.y), tx.z = rrs * (tx.z - rmu * ux.z), tx.w = rrs * (tx.w - rmu * ux.w);
              tg.x = rrs * (tg.x - rmu * ug.x), tg.y = rrs * (tg.y - rmu * ug.y), tg.z = rrs * (tg.z - rmu * ug.z), tg.w = rrs * (tg.w - rmu * ug.w);
            }
            float4 y;
            y.x = (tx.x + bx.x) * gelu_erf_fast(tg.x + bg.x);
            y.y = (tx.y + bx.y) * gelu_erf_fast(tg.y + bg.y);
            y.z = (tx.z + bx.z) * gelu_erf_fast(tg.z + bg.z);
            y.w = (tx.w + bx.w) * gelu_erf_fast(tg.w + bg.w);
            epilogue_store(y, 0u, (unsigned)(mr * p.ldc16 + ocol0 + c + cq), nullptr, p.out_f16, p.out_f16_lo, 0);
          }
        }
        __syncwarp();
      }
    } else {
      constexpr int NCH = NCHUNK;  // column chunks per warp (the warps of a lane quarter interleave them)
      if (!pre_issued) issue_addends(col0 + half * 32 + cq);
      float lrs[8], lrq[8];  // EPI_LNS: this lane's share of the row sums of its 8 rows
#pragma unroll
      for (int k = 0; k < 8; ++k) lrs[k] = 0.f, lrq[k] = 0.f;
      (void)lrs, (void)lrq;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int c = half * 32 + j * CSTEP;
        if (c < BN) {
          uint32_t v[32];
          tmem_ld32(trow + c, v);
          tmem_ld_wait();
          stage(tile_s, v);
          __syncwarp();
          const int col = col0 + c + cq;
          float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f), gsq = gsum;
          (void)gsum, (void)gsq;
          if (col < p.N) {
            const uint32_t tl = tile_s + sub * TROW + cq * 4;
#pragma unroll
            for (int b4 = 0; b4 < 2; ++b4) {
              float4 t[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(t[i].x), "=f"(t[i].y), "=f"(t[i].z), "=f"(t[i].w)
                             : "r"(tl + (b4 * 4 + i) * 4 * TROW)
                             : "memory");
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int k = b4 * 4 + i;
                if (mr8[k] < 0) continue;
                float4 f = t[i];
                if constexpr (kLNC) {  // LayerNorm folded in: rstd * (acc - mean * u) ; beta^T W + bias comes in through bvs
                  f.x = rs8[k] * (f.x - mu8[k] * us[j].x), f.y = rs8[k] * (f.y - mu8[k] * us[j].y);
                  f.z = rs8[k] * (f.z - mu8[k] * us[j].z), f.w = rs8[k] * (f.w - mu8[k] * us[j].w);
                }
                if constexpr (kLNC)
                  f.x += bvs[j].x, f.y += bvs[j].y, f.z += bvs[j].z, f.w += bvs[j].w;
                else
                  f.x += bvs[j].x + ad[k].x, f.y += bvs[j].y + ad[k].y, f.z += bvs[j].z + ad[k].z, f.w += bvs[j].w + ad[k].w;
                if constexpr (kLNS) {
                  lrs[k] += (f.x + f.y) + (f.z + f.w);
                  lrq[k] = fmaf(f.x, f.x, fmaf(f.y, f.y, fmaf(f.z, f.z, fmaf(f.w, f.w, lrq[k]))));
                }
                if constexpr (kGN) {
                  gsum.x += f.x, gsum.y += f.y, gsum.z += f.z, gsum.w += f.w;
                  gsq.x = fmaf(f.x, f.x, gsq.x), gsq.y = fmaf(f.y, f.y, gsq.y), gsq.z = fmaf(f.z, f.z, gsq.z), gsq.w = fmaf(f.w, f.w, gsq.w);
                }
                epilogue_store(f, (unsigned)(mr8[k] * p.ldc + col), (unsigned)(mr8[k] * p.ldc16 + col), p.out_f32, p.out_f16,
                               p.out_f16_lo, p.act);
              }
            }
          }
          if constexpr (kGN) {
            // the 4 lanes that hold the same columns (sub = 0..3) fold in a fixed order; sub 0 publishes the quarter's 32-row sums
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              gsum.x += __shfl_xor_sync(0xffffffffu, gsum.x, o), gsum.y += __shfl_xor_sync(0xffffffffu, gsum.y, o);
              gsum.z += __shfl_xor_sync(0xffffffffu, gsum.z, o), gsum.w += __shfl_xor_sync(0xffffffffu, gsum.w, o);
              gsq.x += __shfl_xor_sync(0xffffffffu, gsq.x, o), gsq.y += __shfl_xor_sync(0xffffffffu, gsq.y, o);
              gsq.z += __shfl_xor_sync(0xffffffffu, gsq.z, o), gsq.w += __shfl_xor_sync(0xffffffffu, gsq.w, o);
            }
            if (sub == 0) {
              float2* d = gn_cs + q * BN + c + cq;
              d[0] = make_float2(gsum.x, gsq.x), d[1] = make_float2(gsum.y, gsq.y);
              d[2] = make_float2(gsum.z, gsq.z), d[3] = make_float2(gsum.w, gsq.w);
            }
          }
          __syncwarp();
          // the next chunk's addends travel while its accumulator columns are read and staged
          if (j + 1 < NCH && c + CSTEP < BN) issue_addends(col0 + c + CSTEP + cq);
        }
      }
      if constexpr (kLNS) {
        // the 8 lanes that share a row (same sub) fold their column groups in a fixed order; one of them publishes this warp's
        // partial for the row: slot = (N tile, share of the chunks). The consumer adds the slots in index order.
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
          for (int o = 1; o <= 4; o <<= 1) {
            lrs[k] += __shfl_xor_sync(0xffffffffu, lrs[k], o);
            lrq[k] += __shfl_xor_sync(0xffffffffu, lrq[k], o);
          }
        }
        if ((lane & 7) == 0) {
          // two slots per N tile whatever the warp count (the consumer's slot count is a function of N alone)
          const int slot = blockIdx.y * 2 + half;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (mr8[k] >= 0) {
              float2* d = reinterpret_cast<float2*>(p.ln_out) + (size_t)mr8[k] * p.ln_slots + slot;
              d[0] = make_float2(lrs[k], lrq[k]);
              if (EG == 1) d[1] = make_float2(0.f, 0.f);
            }
        }
      }
      if constexpr (kGN) {
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");  // every quarter's column sums are in shared memory
        gn_write_partials(p, gn_cs, BN, threadIdx.x - 64, EW * 32, gnt, col0, p.up2 ? (int)blockIdx.z * p.gn_phase_slots : 0);
      }
    }
    tc_fence_before();
    if (dbg && threadIdx.x == 64) dbg[6] = clock64();
  }

  if (TWO)
    cluster_sync_all();  // no CTA may exit while its peer can still signal its barriers / read its smem
  else
    __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (TWO)
      tmem_dealloc2(tmem_base, TMEM_COLS);
    else
      tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (dbg && threadIdx.x == 0) dbg[7] = clock64();
}

// ------------------------------------------------------------------ launcher
template <int BN, int PASSES, int CG>
constexpr int pick_stages() {
  // as many stages as fit in ~200 KB, capped at 8
  constexpr int per = StageLayout<BN, PASSES, CG>::BYTES;
  constexpr int n = (200 * 1024) / per;
  return n > 8 ? 8 : n;
}
template <int BN, int PASSES, int CG>
constexpr int pick_stages_half() {
  // configuration that lets two CTAs share one SM (<= ~110 KB each)
  constexpr int per = StageLayout<BN, PASSES, CG>::BYTES;
  constexpr int n = (104 * 1024) / per;
  return n > 4 ? 4 : (n < 2 ? 2 : n);
}

template <int BN, int PASSES, int STAGES, int CG, int EPI>
static void launch_epi(const GemmMaps& maps, const GemmParams& p, cudaStream_t stream) {
  constexpr int smem = STAGES * StageLayout<BN, PASSES, CG>::BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
  constexpr int EW = epilogue_warps<BN, PASSES, STAGES, CG, EPI>();
  static_assert(STAGES * StageLayout<BN, PASSES, CG>::BYTES >= 2 * EW * 32 * 144, "epilogue staging tiles must fit in the stages");
  static_assert(4 * BN * 8 <= EW * 32 * 144 && ((EW * 32) / (BN / 4)) * 4 * (BN / 4) * 32 <= EW * 32 * 144,
                "GroupNorm column sums must fit in the second staging bank");
  static DeviceOnce once;
  if (once.first())
    SDB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, PASSES, STAGES, CG, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(p.tiles_n * p.tiles_h * p.tiles_w, (p.N + BN - 1) / BN, p.up2 ? 4 : p.split_k);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = dim3(64 + 32 * EW), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = g_pdl_enabled ? 2 : 1;
  SDB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, PASSES, STAGES, CG, EPI>, maps, p));
}

// the statistics-producing epilogues exist for the tile widths their tensors use (run_gemm checks with gemm_tc_supports_epi)
template <int BN, int PASSES, int STAGES, int CG>
static void launch_inst(const GemmMaps& maps, const GemmParams& p, cudaStream_t stream) {
  SDB_CHECK((p.gn_part != nullptr) + (p.ln_out != nullptr) + (p.ln_in != nullptr) <= 1, "one statistics role per launch");
  if constexpr (BN == 128) {
    if (p.geglu) {
      SDB_CHECK(!p.gn_part && !p.ln_out && p.split_k == 1, "GEGLU epilogue: no statistics output, no split-K");
      return p.ln_in ? launch_epi<BN, PASSES, STAGES, CG, EPI_GEGLU_LNC>(maps, p, stream)
                     : launch_epi<BN, PASSES, STAGES, CG, EPI_GEGLU>(maps, p, stream);
    }
  }
  SDB_CHECK(!p.geglu, "the GEGLU epilogue is built for 128-wide tiles");
  if constexpr (BN >= 128) {
    if (p.gn_part) return launch_epi<BN, PASSES, STAGES, CG, EPI_GN>(maps, p, stream);
  }
  if constexpr (BN == 160) {
    if (p.ln_out) return launch_epi<BN, PASSES, STAGES, CG, EPI_LNS>(maps, p, stream);
  }
  if constexpr (BN == 128 || BN == 160) {
    if (p.ln_in) return launch_epi<BN, PASSES, STAGES, CG, EPI_LNC>(maps, p, stream);
  }
  SDB_CHECK(!p.gn_part && !p.ln_out && !p.ln_in, "this statistics epilogue is not built for this tile width");
  launch_epi<BN, PASSES, STAGES, CG, EPI_PLAIN>(maps, p, stream);
}
bool gemm_tc_supports(int BN, int epi) {
  return epi == EPI_PLAIN || (epi == EPI_GN && BN >= 128) || (epi == EPI_LNS && BN == 160) || (epi == EPI_LNC && (BN == 128 || BN == 160));
}

template <int BN, int PASSES, int CG>
static void launch_cg(const GemmMaps& maps, const GemmParams& p, cudaStream_t stream) {
  const long long ctas = (long long)p.tiles_n * p.tiles_h * p.tiles_w * ((p.N + BN - 1) / BN) * (p.up2 ? 4 : p.split_k);
  constexpr int SH = pick_stages_half<BN, PASSES, CG>();
  constexpr bool half_ok = SH * StageLayout<BN, PASSES, CG>::BYTES <= 104 * 1024 && SH * StageLayout<BN, PASSES, CG>::BYTES >= 8 * 32 * 144 * 2;
  // many short tiles: two co-resident CTAs per SM overlap one tile's epilogue with the other's mainloop
  if (half_ok && ctas >= 2 * 148)
    launch_inst<BN, PASSES, half_ok ? SH : pick_stages<BN, PASSES, CG>(), CG>(maps, p, stream);
  else
    launch_inst<BN, PASSES, pick_stages<BN, PASSES, CG>(), CG>(maps, p, stream);
}
template <int BN, int PASSES>
static void launch_bn(const GemmMaps& maps, const GemmParams& p, cudaStream_t stream) {
  if (p.cluster == 2)
    launch_cg<BN, PASSES, 2>(maps, p, stream);
  else
    launch_cg<BN, PASSES, 1>(maps, p, stream);
}

void gemm_tc_launch(const GemmMaps& maps, const GemmParams& p, int BN, int passes, cudaStream_t stream) {
  SDB_CHECK(p.TN * p.TH * p.TW == BM, "M tile must cover 128 rows");
  SDB_CHECK(p.N % 32 == 0, "N must be a multiple of 32");
  SDB_CHECK(p.cluster == 1 || (p.cluster == 2 && (p.tiles_n * p.tiles_h * p.tiles_w) % 2 == 0), "CTA pairs need an even M-tile count");
#define SDB_DISPATCH(bn)                                             \
  case bn:                                                           \
    if (passes == 1) launch_bn<bn, 1>(maps, p, stream);              \
    else if (passes == 2) launch_bn<bn, 2>(maps, p, stream);         \
    else launch_bn<bn, 3>(maps, p, stream);                          \
    break;
  switch (BN) {
    SDB_DISPATCH(64)
    SDB_DISPATCH(128)
    SDB_DISPATCH(160)
    SDB_DISPATCH(256)
    default:
      throw Error("unsupported BN");
  }
#undef SDB_DISPATCH
}

}  // namespace sdb
