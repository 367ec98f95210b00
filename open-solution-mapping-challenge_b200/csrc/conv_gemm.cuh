// conv_gemm.cuh — device kernels for every convolution-shaped op of the U-Net, as tcgen05 implicit GEMMs.
//
//   conv_gemm_kernel<BN, BK, B_MN>   forward convs / transposed convs / data gradients
//       D[pixel, n] = sum over taps t, channels k:  A_t[pixel + off_t, k] * B_t[n, k]
//       A tiles: NHWC bf16 activations fetched by 4-D tiled TMA boxes (channels, bw, bh, bn) placed at the tap
//       offset — out-of-bounds rows/columns are zero-filled by the TMA unit, which IS the conv zero padding.
//       B tiles: bf16 weights [tap][cout][cin]; K-major for forward, MN-major (same storage) for dgrad.
//       Accumulator: 128 x BN fp32 in TMEM.  Epilogue: bias / ReLU / ReLU-mask / BN-statistics, bf16, TMA store
//       (or TMA reduce-add for gradient accumulation).
//
//   wgrad_kernel<BN>                 weight gradients
//       dW_t[co, ci] += sum over pixels: dY[pixel + offA_t, co] * X[pixel + offB_t, ci]
//       both operands MN-major (the pixel axis is GEMM-K), split-K over pixel tiles, fp32 red.add epilogue.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#pragma once
#include "tc.cuh"

namespace mcb {

constexpr int kMaxTaps = 24;
constexpr int kGemmThreads = 192;

struct TapDesc {
  int16_t src;      // index into tmA
  int16_t dx, dy;   // offset in the source view (W, H coordinates)
  int16_t nchunks;  // number of BK-channel chunks of this source
  int32_t wk0;      // first K coordinate in the weight tensor for this source (concat offset)
  int32_t wtap;     // tap coordinate in the weight tensor
};

struct ConvGemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmD[4];  // one per phase (blockIdx.z)
  TapDesc taps[kMaxTaps];
  int tap_start[4];
  int tap_count[4];
  int Wv, Hv, Nimg;  // extents of the output view
  int bw, bh, bn, rows;
  int tiles_x, tiles_y;
  int m_tiles, n_tiles, phases;  // persistent tile space: phases x m_tiles x n_tiles
  int stages;
  int b_resident;  // HALO only, experimental: the 9 per-tap weight tiles are loaded ONCE per CTA into ring slots 0..8
                   // (single N tile, one channel chunk) instead of once per pixel tile
  int n_off;  // first N (weight row / column) coordinate of this launch (concat source slice for dgrad)
  // epilogue:  v = acc * scale[c] + bias[c] + residual[pix][c];  v = relu(v);  v = mask ? v : 0
  const float* scale;              // per-channel multiplier (inference-mode BatchNorm folded into the epilogue), or null
  const __nv_bfloat16* residual;   // NHWC bf16 with the geometry of the output tensor (mask_H/W/C), or null
  const float* bias;
  float* stats;
  int stats_c;  // number of channels in stats (cout)
  int mask_H, mask_W, mask_C, mask_s;  // geometry of `residual`: full dims; mask_s = 1 (plain) or 2 (parity view)
  int relu;
  int accumulate;
  // Auxiliary tile (dgrad only): a tensor with the geometry of the OUTPUT, fetched chunk by chunk with TMA (tmX, same
  // boxes as tmD) into shared memory while the main loop of the tile still runs.
  //   aux_mode 1: the ReLU output y of the producing layer:  g = acc * (y > 0)
  //   aux_mode 2: the BatchNorm input z of the producing conv-BN-ReLU unit: the mask is that unit's own output sign,
  //               (fma(z, gamma*invstd, beta - mean*gamma*invstd) > 0), and the BatchNorm-backward reductions of the
  //               stored gradient ride along:  dbeta += sum g,  dgamma += sum g * (z - mean) * invstd
  CUtensorMap tmX[4];
  int aux_mode;
  const float* bn_mean;
  const float* bn_invstd;
  const float* bn_gamma;
  const float* bn_beta;
  float* bn_dbeta;
  float* bn_dgamma;
};

struct WgradTap {
  int16_t srcA, ax, ay;
  int16_t srcB, bx, by;
  int32_t wtap;
};

struct WgradParams {
  CUtensorMap tmA[4];  // dY views, box (64 | 32 channels, bw, bh, bn)
  CUtensorMap tmB[4];  // X views
  WgradTap taps[16];
  int ntaps;
  int Wv, Hv, Nimg;
  int bw, bh, bn, rows;  // rows % 16 == 0, rows <= 64
  int tiles_x, tiles_y, tiles_total;
  int splits;
  int stages;
  float* dw;  // [tap][cout][cin_total]
  int cout, cin_total, ci_off;
  int a_cw;  // channel width of one A chunk: 64 (SW128) or 32 (SW64)
  int a_chunks;  // chunks actually loaded (1 or 2); missing ones alias chunk 0 (LBO = 0)
};

// (pointer + offset, not an integer round trip: the result provably stays in the shared window, so every access below
// compiles to LDS/STS with 32-bit addressing instead of generic LD/ST with 64-bit address arithmetic)
__device__ __forceinline__ uint8_t* align_up_1024(uint8_t* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  return p + ((1024u - (a & 1023u)) & 1023u);
}

// -------------------------------------------------------------------------------------------------
// Persistent, warp-specialised: one CTA per SM walks tiles t = blockIdx.x, += gridDim.x (n fastest, so CTAs that run
// together share the A pixel tile in L2).  The TMA ring (full/empty mbarriers) runs continuously across tiles; two
// TMEM accumulators (2 x BN columns) let the 8 epilogue warps drain tile i while the MMA warp is already on tile i+1.
constexpr int kConvThreads = 320;  // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue

// HALO (3x3, stride 1): the pixel tile is 8 wide x 16 tall in one image and ONE haloed TMA box (BK channels, 10, 18, 1)
// per channel chunk serves all nine taps: tap (dy, dx) is the same shared-memory tile read through a descriptor whose
// start is shifted by ((1+dy)*10 + (1+dx)) rows and whose 8-row groups are 10 rows apart (the UMMA swizzle is a function
// of the absolute shared-memory address, so unaligned starts and a non-atom SBO are legal — tools/probe/probe.cu,
// tools/probe_umma.py).  A and B then travel in separate mbarrier rings: 1 A load + 9 B loads per channel chunk.
constexpr int kHaloW = 10, kHaloH = 18;  // (8 + 2) x (16 + 2)

template <int BN, int BK, bool B_MN, bool HALO>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  static_assert(BK == 64 || BK == 32, "BK");
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN");
  constexpr int A_ROW_BYTES = BK * 2;              // 128 (SW128) or 64 (SW64)
  constexpr int A_BYTES = HALO ? ((kHaloW * kHaloH * A_ROW_BYTES + 1023) / 1024) * 1024 : 128 * A_ROW_BYTES;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = HALO ? B_BYTES : A_BYTES + B_BYTES;  // HALO: the ring holds B tiles; A has its own 2 slots
  constexpr int A_RING_BYTES = HALO ? 2 * A_BYTES : 0;
  constexpr uint32_t A_LAYOUT = (BK == 64) ? tc::LAYOUT_SW128 : tc::LAYOUT_SW64;
  // MN-major B: rows are K (BK of them), each row holds min(BN,64) n-values
  constexpr int BMN_CW = (BN >= 64) ? 64 : 32;          // n-values per sub-tile row
  constexpr int BMN_ROW_BYTES = BMN_CW * 2;             // 128 or 64
  constexpr int BMN_SUB_BYTES = BK * BMN_ROW_BYTES;     // one sub-tile
  constexpr int BMN_SUBS = BN / BMN_CW;
  constexpr uint32_t BMN_LAYOUT = (BMN_CW == 64) ? tc::LAYOUT_SW128 : tc::LAYOUT_SW64;
  constexpr uint32_t IDESC = tc::make_idesc_bf16(128, BN, 0, B_MN ? 1 : 0);
  // output staging: chunks of OUT_CW channels (one TMA store each)
  constexpr int OUT_CW = (BN >= 64) ? 64 : 32;
  constexpr int OUT_ROW_BYTES = OUT_CW * 2;
  constexpr int OUT_CHUNK_BYTES = 128 * OUT_ROW_BYTES;
  constexpr int OUT_CHUNKS = BN / OUT_CW;
  // small tiles finish faster than a TMA store drains: rotate several staging buffers so the epilogue of tile i+1 never
  // waits for the store of tile i
  constexpr int OUT_BUFS = (BN <= 64) ? 4 : (BN == 128 ? 2 : 1);
  constexpr int OUT_BYTES = OUT_BUFS * OUT_CHUNKS * OUT_CHUNK_BYTES;
  // statistics scratch: RG row groups x BN columns x {sum, sumsq}
  constexpr int STAT_BYTES = 2 * 8192;  // per half: (128 / (OUT_CW/8)) row groups x OUT_CW columns x {a, b} floats

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_up_1024(smem_raw);
  const int stages = p.stages;
  uint8_t* a_ring = smem + (size_t)stages * STAGE_BYTES;               // HALO only: 2 haloed A tiles
  uint8_t* out_stage = a_ring + A_RING_BYTES;                          // 1024-aligned (all pieces are)
  float* stat_scratch = reinterpret_cast<float*>(out_stage + OUT_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + OUT_BYTES + STAT_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* a_full_bar = tmem_empty_bar + 2;      // [2] (HALO)
  uint64_t* a_empty_bar = a_full_bar + 2;         // [2] (HALO)
  uint64_t* aux_full_bar = a_empty_bar + 2;       // [2] one per epilogue half
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full_bar + 2);
  int* row_pix = reinterpret_cast<int*>(tmem_slot + 2);  // [2][128] pixel index in the full-resolution tensor, -1 = invalid

  // the shuffle makes the warp index provably warp-uniform, so the role loops below compile onto the uniform datapath
  // (descriptors, barrier addresses and loop counters in uniform registers; no ELECT/R2UR round trip per tcgen05.mma)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tiles_per_phase = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_per_phase * p.phases;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(&tmem_full_bar[0], 1);
    tc::mbar_init(&tmem_full_bar[1], 1);
    tc::mbar_init(&tmem_empty_bar[0], 8);  // one arrival per epilogue warp
    tc::mbar_init(&tmem_empty_bar[1], 8);
    tc::mbar_init(&a_full_bar[0], 1);
    tc::mbar_init(&a_full_bar[1], 1);
    tc::mbar_init(&a_empty_bar[0], 1);
    tc::mbar_init(&a_empty_bar[1], 1);
    tc::mbar_init(&aux_full_bar[0], 1);
    tc::mbar_init(&aux_full_bar[1], 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&p.tmB);
    tc::prefetch_tmap(&p.tmA[0]);
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 2 * BN);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  pdl_prologue();  // everything above touched only shared memory / TMEM / kernel parameters
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp walks the loop, one elected lane issues)
    {
      const uint32_t a_bytes = (uint32_t)p.rows * A_ROW_BYTES;
      uint32_t kb = 0;
      uint32_t ag = 0;  // HALO: haloed A tiles issued so far
      int ring_s = 0;            // pipeline slot / phase parity, advanced without integer division: the single-thread
      uint32_t ring_ph = 0;      // producer and MMA loops are latency chains, every instruction in them is exposed
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        // tile order: phase fastest (the phases of one pixel tile share the A tile in L2), then pixel tile, N tile
        // slowest (CTAs running together share the weight tile; a CTA keeps its N tile for many tiles in a row)
        const int phase_id = t % p.phases;
        const int mt = (t / p.phases) % p.m_tiles;
        const int nt = t / (p.phases * p.m_tiles);
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tn = mt / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tn * p.bn;
        const int ncol0 = nt * BN;
        const int tap_begin = p.tap_start[phase_id], tap_end = tap_begin + p.tap_count[phase_id];
        if (HALO) {
          // taps are listed tap-major, source-minor: entry (t, s) at tap_begin + t * nsrc + s
          const int nsrc = (tap_end - tap_begin) / 9;
          for (int sidx = 0; sidx < nsrc; ++sidx) {
            const TapDesc t0 = p.taps[tap_begin + sidx];
            for (int ch = 0; ch < t0.nchunks; ++ch, ++ag) {
              const int as = ag & 1;
              tc::mbar_wait(&a_empty_bar[as], ((ag >> 1) & 1) ^ 1);
              if (tc::elect_one()) {
                tc::mbar_expect_tx(&a_full_bar[as], (uint32_t)(kHaloW * kHaloH * A_ROW_BYTES));
                tc::tma_load_4d(a_ring + (size_t)as * A_BYTES, &p.tmA[t0.src], &a_full_bar[as], ch * BK, x0 - 1, y0 - 1, n0);
              }
              for (int t = 0; t < 9; ++t, ++kb) {
                const TapDesc tap = p.taps[tap_begin + t * nsrc + sidx];
                int s = ring_s;
                const uint32_t ph = ring_ph;
                if (++ring_s == stages) { ring_s = 0; ring_ph ^= 1; }
                if (p.b_resident) {
                  if (ag != 0) continue;   // weights of tap t already sit in slot t
                  s = t;
                } else {
                  tc::mbar_wait(&empty_bar[s], ph ^ 1);
                }
                uint8_t* sb = smem + (size_t)s * STAGE_BYTES;
                if (tc::elect_one()) {
                  tc::mbar_expect_tx(&full_bar[s], B_BYTES);
                  if (!B_MN) {
                    tc::tma_load_3d(sb, &p.tmB, &full_bar[s], tap.wk0 + ch * BK, p.n_off + ncol0, tap.wtap);
                  } else {
#pragma unroll
                    for (int j = 0; j < BMN_SUBS; ++j)
                      tc::tma_load_3d(sb + j * BMN_SUB_BYTES, &p.tmB, &full_bar[s], p.n_off + ncol0 + j * BMN_CW,
                                      tap.wk0 + ch * BK, tap.wtap);
                  }
                }
              }
            }
          }
          continue;
        }
        for (int tp = tap_begin; tp < tap_end; ++tp) {
          const TapDesc tap = p.taps[tp];
          const CUtensorMap* mA = &p.tmA[tap.src];
          for (int ch = 0; ch < tap.nchunks; ++ch, ++kb) {
            const int s = ring_s;
            const uint32_t ph = ring_ph;
            if (++ring_s == stages) { ring_s = 0; ring_ph ^= 1; }
            tc::mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + (size_t)s * STAGE_BYTES;
            uint8_t* sb = sa + A_BYTES;
            if (tc::elect_one()) {
              tc::mbar_expect_tx(&full_bar[s], a_bytes + B_BYTES);
              tc::tma_load_4d(sa, mA, &full_bar[s], ch * BK, x0 + tap.dx, y0 + tap.dy, n0);
              if (!B_MN) {
                tc::tma_load_3d(sb, &p.tmB, &full_bar[s], tap.wk0 + ch * BK, p.n_off + ncol0, tap.wtap);
              } else {
#pragma unroll
                for (int j = 0; j < BMN_SUBS; ++j)
                  tc::tma_load_3d(sb + j * BMN_SUB_BYTES, &p.tmB, &full_bar[s], p.n_off + ncol0 + j * BMN_CW,
                                  tap.wk0 + ch * BK, tap.wtap);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    uint32_t kb = 0;
    uint32_t ag = 0;
    int ring_s = 0;
    uint32_t ring_ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int phase_id = t % p.phases;
      const int tap_begin = p.tap_start[phase_id], tap_end = tap_begin + p.tap_count[phase_id];
      int num_kb = 0;
      for (int tp = tap_begin; tp < tap_end; ++tp) num_kb += p.taps[tp].nchunks;
      const int acc = it & 1;
      tc::mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);  // epilogue drained this accumulator
      tc::tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BN);
      if (HALO) {
        const int nsrc = (tap_end - tap_begin) / 9;
        int i = 0;
        for (int sidx = 0; sidx < nsrc; ++sidx) {
          const int nch = p.taps[tap_begin + sidx].nchunks;
          for (int ch = 0; ch < nch; ++ch, ++ag) {
            const int as = ag & 1;
            tc::mbar_wait(&a_full_bar[as], (ag >> 1) & 1);
            const uint32_t a_base = tc::smem_u32(a_ring + (size_t)as * A_BYTES);
            for (int t = 0; t < 9; ++t, ++kb, ++i) {
              const TapDesc tap = p.taps[tap_begin + t * nsrc + sidx];
              int s = ring_s;
              const uint32_t ph = ring_ph;
              if (++ring_s == stages) { ring_s = 0; ring_ph ^= 1; }
              if (p.b_resident) {
                s = t;
                if (ag == 0) tc::mbar_wait(&full_bar[s], 0);   // first tile of this CTA only
              } else {
                tc::mbar_wait(&full_bar[s], ph);
              }
              tc::tc_fence_after();
              if (tc::elect_one()) {
                const uint32_t sa = a_base + (uint32_t)(((1 + tap.dy) * kHaloW + (1 + tap.dx)) * A_ROW_BYTES);
                const uint32_t sb = tc::smem_u32(smem + (size_t)s * STAGE_BYTES);
                const uint64_t da0 = tc::make_smem_desc(sa, 16, kHaloW * A_ROW_BYTES, A_LAYOUT);
                uint64_t db0;
                if (!B_MN) db0 = tc::make_smem_desc(sb, 16, 8 * A_ROW_BYTES, A_LAYOUT);
                else db0 = tc::make_smem_desc(sb, BMN_SUB_BYTES, 8 * BMN_ROW_BYTES, BMN_LAYOUT);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint64_t da = da0 + (uint64_t)((k * 32) >> 4);
                  const uint64_t db = B_MN ? db0 + (uint64_t)((k * 16 * BMN_ROW_BYTES) >> 4) : db0 + (uint64_t)((k * 32) >> 4);
                  tc::umma_bf16(tmem_acc, da, db, IDESC, (i > 0 || k > 0) ? 1u : 0u);
                }
                if (!p.b_resident) tc::umma_commit(&empty_bar[s]);
                if (t == 8) tc::umma_commit(&a_empty_bar[as]);
                if (i == num_kb - 1) tc::umma_commit(&tmem_full_bar[acc]);
              }
              __syncwarp();
            }
          }
        }
        continue;
      }
      for (int i = 0; i < num_kb; ++i, ++kb) {
        const int s = ring_s;
        const uint32_t ph = ring_ph;
        if (++ring_s == stages) { ring_s = 0; ring_ph ^= 1; }
        tc::mbar_wait(&full_bar[s], ph);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t sa = tc::smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da0 = tc::make_smem_desc(sa, 16, 8 * A_ROW_BYTES, A_LAYOUT);
          uint64_t db0;
          if (!B_MN) db0 = tc::make_smem_desc(sb, 16, 8 * A_ROW_BYTES, A_LAYOUT);
          else db0 = tc::make_smem_desc(sb, BMN_SUB_BYTES, 8 * BMN_ROW_BYTES, BMN_LAYOUT);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = da0 + (uint64_t)((k * 32) >> 4);
            const uint64_t db = B_MN ? db0 + (uint64_t)((k * 16 * BMN_ROW_BYTES) >> 4) : db0 + (uint64_t)((k * 32) >> 4);
            tc::umma_bf16(tmem_acc, da, db, IDESC, (i > 0 || k > 0) ? 1u : 0u);
          }
          tc::umma_commit(&empty_bar[s]);
          if (i == num_kb - 1) tc::umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..9 = two halves of 128 threads)
    // Each half (4 warps = the 4 TMEM lane quadrants = all 128 rows) owns whole OUT_CW-column chunks: half h takes
    // chunks h, h+2, ...  Per chunk: wait until the TMA store that last used this chunk's staging buffer has read it,
    // TMEM -> registers -> bias/ReLU/mask -> bf16 -> swizzled smem, per-channel reductions from smem, TMA store.
    // The two halves never synchronise with each other; stores of one chunk overlap the math of the next.
    const int ew = warp - 2;           // 0..7
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int half = ew >> 2;
    const int eth = (threadIdx.x - 64) & 127;  // thread index inside the half
    const int bar_id = 2 + half;
    constexpr int CH = (OUT_CHUNKS + 1) / 2;   // chunks per half (half 1 may own fewer)
    constexpr int CGc = OUT_CW / 8;            // 8-channel groups per chunk
    constexpr int RGc = 128 / CGc;             // row groups
    constexpr int ROWSc = 128 / RGc;           // rows per thread in the reduction
    int* my_row_pix = row_pix + half * 128;
    float* my_scratch = stat_scratch + half * (RGc * OUT_CW * 2);
    const int row = q * 32 + lane;
    const int wi = row % p.bw;
    const int hi = (row / p.bw) % p.bh;
    const int ni = row / (p.bw * p.bh);
    // single-chunk tiles (BN <= 64): the halves alternate TILES instead (half h drains tiles with it % 2 == h, i.e. its
    // own TMEM accumulator), so two epilogues are in flight at once where the per-tile main loop is shortest
    constexpr bool ALT = (OUT_CHUNKS == 1);
    const int my_chunks = ALT ? 1 : (OUT_CHUNKS - half + 1) / 2;
    // per-channel reductions are kept in registers (threads eth < OUT_CW, one channel per chunk) across all tiles of
    // this CTA that share an N tile, and flushed with one atomic per channel when the N tile changes / at the end
    float racc1[CH], racc2[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) racc1[j] = racc2[j] = 0.f;
    int racc_nt = -1;
    const int aux_mode = p.aux_mode;
    // aux 1 with bn_dbeta set: per-channel sum of the masked gradient (bias gradient of the producing layer)
    const bool do_red = (p.stats != nullptr || aux_mode == 2 || (aux_mode == 1 && p.bn_dbeta != nullptr));
    // aux modes use the four chunk buffers as {out[half], aux[half]}
    uint8_t* abuf = out_stage + (size_t)(2 + half) * OUT_CHUNK_BYTES;
    uint32_t aux_n = 0;
    auto issue_aux = [&](int t2, int chunk2) {   // one thread: fetch the aux chunk of tile t2 into abuf
      const int ph2 = t2 % p.phases;
      const int mt2 = (t2 / p.phases) % p.m_tiles;
      const int nt2 = t2 / (p.phases * p.m_tiles);
      const int tx2 = mt2 % p.tiles_x, ty2 = (mt2 / p.tiles_x) % p.tiles_y, tn2 = mt2 / (p.tiles_x * p.tiles_y);
      tc::mbar_expect_tx(&aux_full_bar[half], (uint32_t)p.rows * OUT_ROW_BYTES);
      tc::tma_load_4d(abuf, &p.tmX[ph2], &aux_full_bar[half], p.n_off + nt2 * BN + chunk2 * OUT_CW, tx2 * p.bw,
                      ty2 * p.bh, tn2 * p.bn);
    };
    if (aux_mode != 0 && eth == 0) {
      const int t0 = blockIdx.x + (ALT ? half * gridDim.x : 0);
      if (t0 < total_tiles) issue_aux(t0, ALT ? 0 : half);
    }
    auto flush_reductions = [&]() {
      if (racc_nt >= 0 && eth < OUT_CW) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (j < my_chunks) {
            const int ch = p.n_off + racc_nt * BN + (ALT ? 0 : half + 2 * j) * OUT_CW + eth;
            if (aux_mode == 0) {
              atomicAdd(p.stats + ch, racc1[j]);
              atomicAdd(p.stats + p.stats_c + ch, racc2[j]);
            } else {
              atomicAdd(p.bn_dbeta + ch, racc1[j]);
              if (aux_mode == 2) atomicAdd(p.bn_dgamma + ch, racc2[j]);
            }
            racc1[j] = racc2[j] = 0.f;
          }
        }
      }
    };
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int phase_id = t % p.phases;
      const int mt = (t / p.phases) % p.m_tiles;
      const int nt = t / (p.phases * p.m_tiles);
      const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tn = mt / (p.tiles_x * p.tiles_y);
      const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tn * p.bn;
      const int ncol0 = nt * BN;
      const int ox = x0 + wi, oy = y0 + hi, on = n0 + ni;
      const bool valid = row < p.rows && ox < p.Wv && oy < p.Hv && on < p.Nimg;
      // every row of the tile is a real output pixel (the common case): the second pass skips the per-row checks
      const bool tile_full = p.rows == 128 && x0 + p.bw <= p.Wv && y0 + p.bh <= p.Hv && n0 + p.bn <= p.Nimg;
      const int acc = it & 1;
      if (ALT && acc != half) {
        // not this half's tile: only keep the accumulator hand-off in lock-step (8 arrivals per tile)
        tc::mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
        tc::tc_fence_before();
        if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[acc]);
        continue;
      }
      if (do_red && nt != racc_nt) {
        flush_reductions();
        racc_nt = nt;
      }
      int pix = -1;
      if (valid) {
        pix = 0;
        if (p.mask_H > 0) {
          const int fy = oy * p.mask_s + (p.mask_s == 2 ? (phase_id >> 1) : 0);
          const int fx = ox * p.mask_s + (p.mask_s == 2 ? (phase_id & 1) : 0);
          pix = (on * p.mask_H + fy) * p.mask_W + fx;
        }
      }
      const __nv_bfloat16* rrow = nullptr;
      if (p.residual != nullptr && valid) rrow = p.residual + (size_t)pix * p.mask_C + p.n_off + ncol0;

      tc::mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);

      for (int cj = 0; cj < my_chunks; ++cj) {
        const int chunk = ALT ? 0 : half + 2 * cj;
        uint8_t* cbuf = aux_mode != 0 ? out_stage + (size_t)half * OUT_CHUNK_BYTES
                                      : out_stage + (size_t)((it % OUT_BUFS) * OUT_CHUNKS + chunk) * OUT_CHUNK_BYTES;
        // Staging-buffer reuse.  Plain / statistics mode: this half rotates over two buffers, and the wait for the
        // buffer of the NEXT chunk sits just before this chunk's hand-off barrier below (one barrier less per chunk).
        // Aux modes have a single output buffer per half: wait for its previous store here.
        if (cj == 0) my_row_pix[row] = pix;
        if (aux_mode != 0) {
          if (eth == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        }

        // all TMEM loads of the chunk are in flight before the first one is consumed
        uint32_t vv[OUT_CW / 32][32];
#pragma unroll
        for (int sub = 0; sub < OUT_CW / 32; ++sub) tc::tmem_ld_32x32(tmem_acc + (uint32_t)(chunk * OUT_CW + sub * 32), vv[sub]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int sub = 0; sub < OUT_CW / 32; ++sub) {
          const int c0 = chunk * OUT_CW + sub * 32;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(vv[sub][i]);
          if (p.scale != nullptr) {
            const float4* sp = reinterpret_cast<const float4*>(p.scale + p.n_off + ncol0 + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 sv = __ldg(sp + i);
              f[4 * i] *= sv.x; f[4 * i + 1] *= sv.y; f[4 * i + 2] *= sv.z; f[4 * i + 3] *= sv.w;
            }
          }
          if (p.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + p.n_off + ncol0 + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bv = __ldg(bp + i);
              f[4 * i] += bv.x; f[4 * i + 1] += bv.y; f[4 * i + 2] += bv.z; f[4 * i + 3] += bv.w;
            }
          }
          if (rrow != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(rrow + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 rv = __ldg(rp + j);
              const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 xy = __bfloat1622float2(r2[e]);
                f[j * 8 + e * 2] += xy.x;
                f[j * 8 + e * 2 + 1] += xy.y;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          uint8_t* rowp = cbuf + (size_t)row * OUT_ROW_BYTES;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = tc::pack_bf16x2(f[j * 8 + 0], f[j * 8 + 1]);
            o.y = tc::pack_bf16x2(f[j * 8 + 2], f[j * 8 + 3]);
            o.z = tc::pack_bf16x2(f[j * 8 + 4], f[j * 8 + 5]);
            o.w = tc::pack_bf16x2(f[j * 8 + 6], f[j * 8 + 7]);
            int unit = sub * 4 + j;
            if (OUT_CW == 64) unit ^= (row & 7);          // SWIZZLE_128B
            else unit ^= ((row >> 1) & 3);                // SWIZZLE_64B
            *reinterpret_cast<uint4*>(rowp + unit * 16) = o;
          }
        }
        if (cj == my_chunks - 1) {
          // this warp is done with the accumulator
          tc::tc_fence_before();
          if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[acc]);
        }
        tc::fence_proxy_async_smem();
        // (all previous stores of this thread retired their smem reads -> the other rotating buffer is free for the
        // next chunk; the newest of them was issued a whole chunk ago, so this does not stall in practice)
        if (aux_mode == 0 && eth == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");

        if (do_red || aux_mode != 0) {
          // Second pass over the STORED (bf16) chunk in a (row group, 8-channel group) layout: 16-byte smem accesses,
          // per-channel coefficients in registers.
          //   forward:  (sum v, sum v^2) of the valid rows -> BatchNorm batch statistics of the following layer
          //   aux 1:    g = v * (y > 0), written back in place
          //   aux 2:    g = v * (bn(z) > 0) written back, (sum g, sum g * xhat) -> dbeta / dgamma of that BatchNorm
          if (aux_mode != 0) {
            tc::mbar_wait(&aux_full_bar[half], aux_n & 1);
            ++aux_n;
          }
          const bool bnred = aux_mode == 2;
          const int cg = eth % CGc, rg = eth / CGc;
          const int ch0 = p.n_off + ncol0 + chunk * OUT_CW + cg * 8;
          float s1[8], s2[8], mu[8], is[8], sc[8], sh[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) s1[j] = s2[j] = mu[j] = is[j] = sc[j] = sh[j] = 0.f;
          if (bnred) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              mu[j] = __ldg(p.bn_mean + ch0 + j);
              is[j] = __ldg(p.bn_invstd + ch0 + j);
              sc[j] = __ldg(p.bn_gamma + ch0 + j) * is[j];       // same expressions as the forward's
              sh[j] = __ldg(p.bn_beta + ch0 + j) - mu[j] * sc[j];  // bn_train_coef (elementwise.cu)
            }
          }
          // rows rg*ROWSc .. +ROWSc-1; the swizzle term of row r depends only on rr (OUT_CW 64: r & 7 == rr; 32:
          // (r >> 1) & 3 == ((rg & 1) * 2 + (rr >> 1))), so every offset is base + compile-time pieces
          const uint32_t row0_off = (uint32_t)(rg * ROWSc) * OUT_ROW_BYTES;
          const int swz_rg = (OUT_CW == 64) ? 0 : (rg & 1) * 2;
#pragma unroll
          for (int rr = 0; rr < ROWSc; ++rr) {
            if (!tile_full && my_row_pix[rg * ROWSc + rr] < 0) continue;   // ragged tiles only
            const int unit = (OUT_CW == 64) ? (cg ^ rr) : (cg ^ (swz_rg + (rr >> 1)));
            const uint32_t off = row0_off + (uint32_t)rr * OUT_ROW_BYTES + (uint32_t)unit * 16;
            const uint4 pk = *reinterpret_cast<const uint4*>(cbuf + off);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
            if (aux_mode == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 xy = __bfloat1622float2(h2[j]);
                s1[2 * j] += xy.x; s2[2 * j] += xy.x * xy.x;
                s1[2 * j + 1] += xy.y; s2[2 * j + 1] += xy.y * xy.y;
              }
            } else if (aux_mode == 1) {
              const uint4 ak = *reinterpret_cast<const uint4*>(abuf + off);
              const uint32_t w[4] = {ak.x, ak.y, ak.z, ak.w};
              uint32_t o[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // bf16 > 0  <=>  sign bit clear and magnitude nonzero
                const uint32_t lo = w[e] & 0xFFFFu, hi2 = w[e] >> 16;
                if (!(lo != 0 && lo < 0x8000u)) o[e] &= 0xFFFF0000u;
                if (!(hi2 != 0 && hi2 < 0x8000u)) o[e] &= 0x0000FFFFu;
              }
              *reinterpret_cast<uint4*>(cbuf + off) = make_uint4(o[0], o[1], o[2], o[3]);
              if (do_red) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  s1[2 * e] += __uint_as_float(o[e] << 16);
                  s1[2 * e + 1] += __uint_as_float(o[e] & 0xFFFF0000u);
                }
              }
            } else {
              const uint4 zk = *reinterpret_cast<const uint4*>(abuf + off);
              const __nv_bfloat162* z2 = reinterpret_cast<const __nv_bfloat162*>(&zk);
              uint32_t o[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 g = __bfloat1622float2(h2[j]);
                const float2 zz = __bfloat1622float2(z2[j]);
                if (!(fmaf(zz.x, sc[2 * j], sh[2 * j]) > 0.f)) { g.x = 0.f; o[j] &= 0xFFFF0000u; }
                if (!(fmaf(zz.y, sc[2 * j + 1], sh[2 * j + 1]) > 0.f)) { g.y = 0.f; o[j] &= 0x0000FFFFu; }
                s1[2 * j] += g.x; s2[2 * j] += g.x * ((zz.x - mu[2 * j]) * is[2 * j]);
                s1[2 * j + 1] += g.y; s2[2 * j + 1] += g.y * ((zz.y - mu[2 * j + 1]) * is[2 * j + 1]);
              }
              *reinterpret_cast<uint4*>(cbuf + off) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
          if (aux_mode != 0) tc::fence_proxy_async_smem();  // the masked chunk is read by the TMA store below
          const int wq = eth >> 5;  // warp inside the half
          if (do_red) {
            // combine the row groups of this warp with shuffles (lanes l, l+CGc, ... share a channel group), then the
            // four warps through a small shared-memory table
#pragma unroll
            for (int off = CGc; off < 32; off <<= 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
              }
            }
            if (lane < CGc) {
              float4* scq = reinterpret_cast<float4*>(my_scratch + ((size_t)wq * CGc + lane) * 16);
              scq[0] = make_float4(s1[0], s2[0], s1[1], s2[1]);
              scq[1] = make_float4(s1[2], s2[2], s1[3], s2[3]);
              scq[2] = make_float4(s1[4], s2[4], s1[5], s2[5]);
              scq[3] = make_float4(s1[6], s2[6], s1[7], s2[7]);
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (do_red && eth < OUT_CW) {
            float a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) {
              const float2 v2 = *reinterpret_cast<const float2*>(my_scratch + ((size_t)w4 * CGc + (eth >> 3)) * 16 + (eth & 7) * 2);
              a1 += v2.x;
              a2 += v2.y;
            }
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (j == cj) { racc1[j] += a1; racc2[j] += a2; }
          }
        }

        if (eth == 0) {
          const CUtensorMap* mD = &p.tmD[phase_id];
          if (p.accumulate) tc::tma_reduce_add_4d(mD, cbuf, p.n_off + ncol0 + chunk * OUT_CW, x0, y0, n0);
          else tc::tma_store_4d(mD, cbuf, p.n_off + ncol0 + chunk * OUT_CW, x0, y0, n0);
          tc::tma_store_commit();
          if (aux_mode != 0) {
            // abuf is free (every thread of the half passed the barrier after its last read): fetch the next chunk
            if (cj + 1 < my_chunks) issue_aux(t, chunk + 2);
            else {
              const int t2 = t + (int)gridDim.x * (ALT ? 2 : 1);
              if (t2 < total_tiles) issue_aux(t2, ALT ? 0 : half);
            }
          }
        }
      }
    }
    if (do_red) flush_reductions();
    if (eth == 0) tc::tma_store_wait_read0();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 2 * BN);
  }
}

// -------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kGemmThreads) wgrad_kernel(const __grid_constant__ WgradParams p) {
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN");
  constexpr int KROWS = 64;                       // max pixel rows per K block
  constexpr int B_CW = (BN >= 64) ? 64 : 32;      // channels per B sub-tile row
  constexpr int B_ROW_BYTES = B_CW * 2;
  constexpr int B_SUB_BYTES = KROWS * B_ROW_BYTES;
  constexpr int B_SUBS = BN / B_CW;
  constexpr int B_BYTES = B_SUBS * B_SUB_BYTES;
  constexpr uint32_t B_LAYOUT = (B_CW == 64) ? tc::LAYOUT_SW128 : tc::LAYOUT_SW64;
  constexpr int A_SUB_BYTES = KROWS * 128;        // sized for the 64-channel case
  constexpr int A_BYTES = 2 * A_SUB_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t IDESC = tc::make_idesc_bf16(128, BN, 1, 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_up_1024(smem_raw);
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int ncol0 = blockIdx.x * BN;       // cin tile
  const int m0 = blockIdx.y * 128;         // cout tile
  const int tap_id = blockIdx.z / p.splits;
  const int split = blockIdx.z % p.splits;
  const WgradTap tap = p.taps[tap_id];
  // K blocks (pixel tiles) of this split
  const int per = (p.tiles_total + p.splits - 1) / p.splits;
  const int kb_begin = split * per;
  const int kb_end = min(p.tiles_total, kb_begin + per);
  const int num_kb = max(0, kb_end - kb_begin);
  const int a_row_bytes = p.a_cw * 2;
  const uint32_t a_layout = (p.a_cw == 64) ? tc::LAYOUT_SW128 : tc::LAYOUT_SW64;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(tmem_full_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, BN);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  pdl_prologue();  // everything above touched only shared memory / TMEM / kernel parameters
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (num_kb > 0) {
    if (warp == 0) {
      {
        const CUtensorMap* mA = &p.tmA[tap.srcA];
        const CUtensorMap* mB = &p.tmB[tap.srcB];
        const uint32_t tx_bytes = (uint32_t)p.rows * (uint32_t)(a_row_bytes * p.a_chunks + B_ROW_BYTES * B_SUBS);
        int ring_s = 0;
        uint32_t ring_ph = 0;
        // pixel-tile coordinates advance incrementally (no integer division inside the single-thread issue loop)
        int tx = kb_begin % p.tiles_x;
        int ty = (kb_begin / p.tiles_x) % p.tiles_y;
        int tn = kb_begin / (p.tiles_x * p.tiles_y);
        for (int i = 0; i < num_kb; ++i) {
          const int s = ring_s;
          const uint32_t ph = ring_ph;
          if (++ring_s == stages) { ring_s = 0; ring_ph ^= 1; }
          const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tn * p.bn;
          if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++tn; } }
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + (size_t)s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&full_bar[s], tx_bytes);
            for (int j = 0; j < p.a_chunks; ++j)
              tc::tma_load_4d(sa + j * A_SUB_BYTES, mA, &full_bar[s], m0 + j * p.a_cw, x0 + tap.ax, y0 + tap.ay, n0);
#pragma unroll
            for (int j = 0; j < B_SUBS; ++j)
              tc::tma_load_4d(sb + j * B_SUB_BYTES, mB, &full_bar[s], ncol0 + j * B_CW, x0 + tap.bx, y0 + tap.by, n0);
          }
        }
      }
    } else if (warp == 1) {
      const int ksteps = p.rows / 16;
      // A: M = 128 = (128 / a_cw) chunks of a_cw channels; chunks beyond a_chunks alias chunk 0 via LBO = 0
      const uint32_t a_lbo = (p.a_chunks * p.a_cw >= 128) ? (uint32_t)A_SUB_BYTES : 0u;
      int ring_s = 0;
      uint32_t ring_ph = 0;
      for (int i = 0; i < num_kb; ++i) {
        const int s = ring_s;
        const uint32_t ph = ring_ph;
        if (++ring_s == stages) { ring_s = 0; ring_ph ^= 1; }
        tc::mbar_wait(&full_bar[s], ph);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t sa = tc::smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da0 = tc::make_smem_desc(sa, a_lbo, 8 * a_row_bytes, a_layout);
          const uint64_t db0 = tc::make_smem_desc(sb, B_SUB_BYTES, 8 * B_ROW_BYTES, B_LAYOUT);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = da0 + (uint64_t)((k * 16 * a_row_bytes) >> 4);
            const uint64_t db = db0 + (uint64_t)((k * 16 * B_ROW_BYTES) >> 4);
            tc::umma_bf16(tmem_base, da, db, IDESC, (i > 0 || k > 0) ? 1u : 0u);
          }
          tc::umma_commit(&empty_bar[s]);
          if (i == num_kb - 1) tc::umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
    } else {
      const int q = warp & 3;
      const int co = m0 + q * 32 + lane;
      const bool valid = (q * 32 + lane) < p.a_chunks * p.a_cw && co < p.cout;
      tc::mbar_wait(tmem_full_bar, 0);
      tc::tc_fence_after();
      float* drow = p.dw + ((size_t)tap.wtap * p.cout + (valid ? co : 0)) * p.cin_total + p.ci_off + ncol0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tc::tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + i),
                         "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])),
                         "f"(__uint_as_float(v[i + 3]))
                         : "memory");
          }
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, BN);
  }
}

}  // namespace mcb
