// spmm_tcgen05.cuh -- the sm_100a tensor-core SpMM kernel.
//
// What it computes (reference semantics; reference kernels:
// voltrix/include/voltrix/spmm_kernels.cuh:1458-2001):
//   C[16w + r, n] = sum over TC blocks b of window w, columns c < 8 of  bit(b, r, c) * B[hind[8b + c], n]
// with fp32 accumulation, consuming the reference's own tile format (blk_offsets, hspa_packed, hind).
//
// How (nothing here is the reference's mma.sync m16n8k8 pipeline):
//   * operands are swapped: one tcgen05.mma computes  D^T[128 features x 16 rows] +=
//     Bg^T[128 features x 16 gathered rows] * A^T[16 gathered rows x 16 window rows], so the dense
//     width fills the MMA M dimension and the 16-row window is the MMA N dimension;
//   * Bg^T is the "A" operand, MN-major, SWIZZLE_128B: exactly the image TMA tile::gather4 leaves in
//     shared memory (4 B rows x 128 B per instruction; two of them fill one 8-row swizzle atom);
//   * A^T is the "B" operand, K-major, no swizzle: 512 B expanded from two 16-byte bitmaps;
//   * D^T lives in TMEM (128 lanes x 16 fp32 columns, double buffered across work items) and is read
//     back with tcgen05.ld 32x32b: a thread holds the 16 window rows of one feature, so every store
//     instruction of a warp writes 32 consecutive floats of one C row;
//   * a K-step is 16 gathered rows (two TC blocks); a pipeline stage is NPW K-steps (one per producer warp: 7 x (4 KB of
//     B rows + 512 B of A^T) in the 14/7 variant), so the MMA warp pays one mbarrier wait and one tcgen05.commit per stage;
//   * warp roles: NPW producers -- warp w owns K-step w of every stage: it expands the two bitmaps into the A^T tile with all
//     32 lanes (nibble table in shared memory, one STS.128 per lane), then ONE elected lane issues the 8 gather4 copies
//     (elect.sync keeps the TMA operands in uniform registers: no per-lane serialisation loop); 4 epilogue warps (TMEM lane
//     quarter = warp % 4); the MMA issuer + TMEM owner; the loader, which is the CTA's scheduler and streams each item's
//     hind / bitmap arrays into a shared-memory ring with 1-D bulk copies so the producers never wait on a global load;
//   * full/empty mbarriers per stage (tcgen05.commit frees a stage when the MMAs that read it retire), full/empty per
//     metadata chunk, per TMEM accumulator and per slot of the unit ring;
//   * persistent CTAs claim units (feature tile, item) from the LPT-sorted work list (schedule.cuh) with an atomic
//     ticket, feature-tile-major; the loader lane publishes each unit to the other roles through a small shared-memory ring
//     (static striding when the caller passes no ticket counter);
//   * TWO OR THREE CTAs PER SM (tc_ctas_per_sm): small rings, several MMA-issuing warps per SM -- the ~68 clk a back-to-back
//     tcgen05.mma costs is a per-issuer cost, and co-resident CTAs fill one another's barrier waits;
//   * TERMS = 2 (fp32 input as two bf16 terms) doubles the gathered tile and the MMAs of a K-step, same accumulator; fp32
//     input as ONE fp16 term runs the TERMS = 1 kernel behind a device-side range gate (spmm_kernels.cuh, model 4);
//   * WEIGHTED: per-edge values arrive as ready-made A^T tiles by bulk copy, nothing is expanded;
//   * the optional Epilogue (row scale / bias / ReLU) is applied to whole-window items here and to K-split windows in
//     vx_spmm_fixup_kernel.
// What paces it (profiles/r1c_bottleneck_isolation.md, profiles/r2n_multi_cta_variants.txt): TMA writes 32 and the tensor core
// reads 36 shared-memory wavefronts per K-step (68 clk at one per clock) and the L2 slices deliver the gather at 86 % of their
// peak at best (61 clk); the 14/7 variant runs at ~70 clk per K-step on the L2-resident Reddit-shaped graph.
#ifndef VOLTRIX_B200_SPMM_TCGEN05_CUH_
#define VOLTRIX_B200_SPMM_TCGEN05_CUH_

#include "voltrix/common.cuh"
#include "voltrix/ptx.cuh"

// Bottleneck-isolation builds (scripts/isolate.py; results are garbage, timing only):
//   VX_TC_DBG=1  producers + TMA gather run, the MMA issuer only commits      -> gather path alone
//   VX_TC_DBG=2  bitmap expansion + MMAs run on stale shared memory, no TMA   -> MMA operand reads alone
#ifndef VX_TC_DBG
#define VX_TC_DBG 0
#endif
// Timing-only probes: 5 = producers skip the bitmap expansion (gather issue + loop bookkeeping only)
#ifndef VX_TC_EXP
#define VX_TC_EXP 0
#endif
// Probe: force the number of resident CTAs per SM the persistent kernel is compiled and launched for (0 = what the
// variant's shared memory, threads and registers allow, see tc_ctas_per_sm).
#ifndef VX_TC_CTAS_PER_SM
#define VX_TC_CTAS_PER_SM 0
#endif

namespace voltrix {

template <typename T> struct TcFmt;
template <> struct TcFmt<__half> {
  static constexpr uint32_t kFmt = 0, kOne = 0x3C00u;
  static constexpr CUtensorMapDataType kTmapType = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
};
template <> struct TcFmt<__nv_bfloat16> {
  static constexpr uint32_t kFmt = 1, kOne = 0x3F80u;
  static constexpr CUtensorMapDataType kTmapType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};

// TERMS: the dense operand is stored as TERMS 16-bit terms per value (column blocks of `term_stride` elements in the
// tensor map) whose products accumulate into the same TMEM tile.  TERMS = 1 for fp16 / bf16 input; TERMS = 2 is the fp32
// path: x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits, full fp32 exponent range; the reference
// rounds fp32 to TF32 = 10 mantissa bits, spmm_kernels.cuh:1631-1678).
// FT: features per work unit = MMA M.  128 fills the tensor core's M; 64 is for dense operands of at most 64 columns: the M = 64
// instruction reads one 128-byte swizzle atom per gathered row instead of two, which is what such a launch is bound by
// (the second atom would be padding: Reddit-shaped N = 64 ran as fast as N = 128 with the 128-wide tile).
template <int NPW, int TERMS = 1, int FT = 128>
struct TcGeom {
  static_assert(FT == 128 || FT == 64, "feature tile = MMA M: 128 or 64");
  static constexpr int kFeatTile = FT;                  // MMA M: features per work unit
  static constexpr int kAtomCols = 64;                  // 128-byte swizzle span in 16-bit elements
  static constexpr int kTermB = kFeatTile * 16 * 2;     // one term of one K-step: 16 gathered rows x FT features x 2 B (4096 / 2048)
  static constexpr int kKGroupB = kFeatTile * 8 * 2;    // 8 gathered rows (one k-group): FT / 64 swizzle atoms of 1024 B
  static constexpr int kKsB = kTermB * TERMS;           // one K-step, all terms
  static constexpr int kKsA = 16 * 16 * 2;              // one K-step: densified 16 x 16 tile = 512
  // K-steps per stage.  Up to 16 producer warps: one stage = one K-step per warp.  More warps: stages of 8
  // K-steps owned round-robin by NPW/8 warp groups (more warps in flight without growing the stage).
  static constexpr int kKsPerStage = NPW > 16 ? 8 : NPW;
  static constexpr int kGroups = NPW / kKsPerStage;
  static constexpr int kStageB = kKsPerStage * kKsB;
  static constexpr int kStageA = kKsPerStage * kKsA;
  // metadata ring: 4 slots of 4-stage chunks for the one-CTA-per-SM variants; the small rings that share an SM keep
  // 3 slots of 2-stage chunks (every byte of shared memory they do not spend is ring depth for the co-resident CTA)
  static constexpr int kStagesPerChunk = NPW <= 11 ? 2 : 4;
  static constexpr int kChunkBlks = 2 * kKsPerStage * kStagesPerChunk;   // TC blocks per metadata chunk
  static constexpr int kMetaSlots = NPW <= 11 ? 3 : 4;
  static constexpr int kMetaH = kChunkBlks * 32;        // hind bytes per chunk
  static constexpr int kMetaP = kChunkBlks * 16;        // bitmap bytes per chunk
  // Warp layout: producers 0..NPW-1; the four epilogue warps start at a multiple of 4 (epilogue warp e may only read TMEM
  // lanes 32 * (warp % 4) ...); the MMA and loader warps go wherever that costs fewer warps -- right behind the producers
  // (filling the gap up to the next multiple of 4) or behind the epilogue warps.  Fewer threads = more co-resident CTAs.
  static constexpr int kWarpsTail = (NPW + 3) / 4 * 4 + 6;        // [producers][pad][epilogue x4][mma][loader]
  static constexpr int kWarpsGap = (NPW + 2 + 3) / 4 * 4 + 4;     // [producers][mma][loader][pad][epilogue x4]
  static constexpr bool kServiceInGap = kWarpsGap < kWarpsTail;
  static constexpr int kEpilogueWarp0 = kServiceInGap ? (NPW + 2 + 3) / 4 * 4 : (NPW + 3) / 4 * 4;
  static constexpr int kMmaWarp = kServiceInGap ? NPW : kEpilogueWarp0 + 4;
  static constexpr int kLoaderWarp = kMmaWarp + 1;
  static constexpr int kThreads = (kServiceInGap ? kWarpsGap : kWarpsTail) * 32;
  static constexpr uint32_t kTmemCols = 32;             // two 16-column accumulators
  static constexpr int kUnitSlots = 4;                  // work-unit ring (loader -> every other role), 32 B per slot
  static constexpr int kUnitConsumers = NPW + 5;        // producer warps + MMA warp + 4 epilogue warps
};

// KSTEPS = K-steps in flight (the autotuned "stages" knob): ring depth = KSTEPS / NPW stages.
template <int KSTEPS, int NPW, int TERMS = 1, int FT = 128>
constexpr size_t tc_smem_bytes() {
  using G = TcGeom<NPW, TERMS, FT>;
  constexpr int S = KSTEPS / G::kKsPerStage;
  return size_t(S) * (G::kStageB + G::kStageA) + size_t(G::kMetaSlots) * (G::kMetaH + G::kMetaP) +
         (2 * S + 2 * G::kMetaSlots + 4 + 2 * G::kUnitSlots) * 8 + 16 + 128 /*nibble table*/ + G::kUnitSlots * 32 +
         1024 /*align slack*/;
}

// WEIGHTED: A carries per-edge values.  `packed` then points at VALUE TILES instead of bitmaps: 256 B per TC block,
// element (row r, column slot c) of type T at (r >> 3) * 128 + (r & 7) * 16 + c * 2 -- exactly the K-major shared-memory image
// of one 8-column chunk of the A^T operand (core matrices 128 B apart along the window rows, 256 B apart along K), so a
// K-step's A^T tile is ONE 512-byte bulk copy from global memory into the stage and the producers expand nothing.
// Resident CTAs per SM.  Two or three small rings on one SM mean two or three MMA-issuing warps: the ~68 clk a 128x16x16
// tcgen05.mma costs back to back is a per-issuer cost, and with co-resident CTAs the tensor core, the TMA unit
// and the shared-memory port are kept busy by one CTA while another waits on a barrier (Reddit-shaped C2, N=128 fp16: 42/14 with
// one CTA per SM 1.88 ms, 20/10 with two 1.71 ms).  Bounded by shared memory (228 KB per SM, 1 KB reserved per CTA), threads
// (2048) and registers (the kernel needs 48; __launch_bounds__ holds the compiler to what the count allows).
template <int KSTEPS, int NPW, int TERMS = 1, int FT = 128>
constexpr int tc_ctas_per_sm() {
  if (VX_TC_CTAS_PER_SM > 0) return VX_TC_CTAS_PER_SM;
  constexpr int by_smem = int((228 * 1024) / (tc_smem_bytes<KSTEPS, NPW, TERMS, FT>() + 1024));
  constexpr int by_threads = 2048 / TcGeom<NPW, TERMS, FT>::kThreads;
  constexpr int by_regs = 65536 / (TcGeom<NPW, TERMS, FT>::kThreads * 40);   // the kernel compiles to 40-48 registers
  constexpr int m = by_smem < by_threads ? (by_smem < by_regs ? by_smem : by_regs) : (by_threads < by_regs ? by_threads : by_regs);
  return m < 1 ? 1 : (m > 4 ? 4 : m);
}

template <typename T, int KSTEPS, int NPW, int TERMS = 1, bool WEIGHTED = false, int FT = 128>
__global__ void __launch_bounds__(TcGeom<NPW, TERMS, FT>::kThreads, tc_ctas_per_sm<KSTEPS, NPW, TERMS, FT>())
vx_spmm_tc_kernel(const __grid_constant__ CUtensorMap tmap, const WorkItem *__restrict__ items, int32_t num_items,
                  int32_t n_feat_tiles, const int32_t *__restrict__ blk_offsets, const uint4 *__restrict__ packed,
                  const int4 *__restrict__ hind4, int32_t num_nodes, int32_t N, float *__restrict__ C,
                  float *__restrict__ scratch, int32_t term_stride, Epilogue epi, int32_t *__restrict__ ticket,
                  const int32_t *__restrict__ gate, int32_t gate_want) {
  using G = TcGeom<NPW, TERMS, FT>;
  // Optional launch gate (fp32 operands, model 4): two alternative pipelines are enqueued and a flag written by an earlier
  // kernel on the stream decides which one runs; the other returns here, before it touches a barrier or TMEM.
  if (gate != nullptr && *gate != gate_want) return;
  static_assert(NPW % G::kKsPerStage == 0, "producer warps must form whole groups");
  static_assert(KSTEPS % G::kKsPerStage == 0 && KSTEPS / G::kKsPerStage > G::kGroups,
                "the ring must hold more stages than there are producer groups");
  static_assert(!WEIGHTED || TERMS == 1, "per-edge values ride on the 16-bit path only");
  constexpr uint32_t S = KSTEPS / G::kKsPerStage;
  constexpr uint32_t MR = G::kMetaSlots;
  constexpr uint32_t UR = G::kUnitSlots;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB = sbase;                               // [S][4 K-steps][4096]  1024-aligned atoms
  const uint32_t sA = sB + S * G::kStageB;                 // [S][4 K-steps][512]
  const uint32_t sMetaH = sA + S * G::kStageA;             // [MR][64 blocks][8 int32]
  const uint32_t sMetaP = sMetaH + MR * G::kMetaH;         // [MR][64 blocks][4 uint32]
  const uint32_t sBar = sMetaP + MR * G::kMetaP;           // full[S] empty[S] mfull[MR] mempty[MR] tfull[2] tempty[2] ufull[UR] uempty[UR]
  const uint32_t sTmem = sBar + (2 * S + 2 * MR + 4 + 2 * UR) * 8;
  const uint32_t sLut = sTmem + 16;                        // [16] nibble -> four 16-bit {0, 1.0} values (8 B each)
  const uint32_t sUnit = sLut + 128;                       // [UR] {WorkItem, feature-tile base, valid}
  auto full_bar = [&](uint32_t s) { return sBar + s * 8; };
  auto empty_bar = [&](uint32_t s) { return sBar + (S + s) * 8; };
  auto mfull_bar = [&](uint32_t m) { return sBar + (2 * S + m) * 8; };
  auto mempty_bar = [&](uint32_t m) { return sBar + (2 * S + MR + m) * 8; };
  auto tfull_bar = [&](uint32_t a) { return sBar + (2 * S + 2 * MR + a) * 8; };
  auto tempty_bar = [&](uint32_t a) { return sBar + (2 * S + 2 * MR + 2 + a) * 8; };
  auto ufull_bar = [&](uint32_t q) { return sBar + (2 * S + 2 * MR + 4 + q) * 8; };
  auto uempty_bar = [&](uint32_t q) { return sBar + (2 * S + 2 * MR + 4 + UR + q) * 8; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t total_units = num_items * n_feat_tiles;
  // Without a schedule (items == nullptr, kernel-level API) item i is simply window i, whole.
  auto load_item = [&](int32_t i) -> WorkItem {
    if (items != nullptr) return items[i];
    WorkItem it;
    it.window = i;
    it.blk_begin = blk_offsets[i];
    it.blk_count = blk_offsets[i + 1] - it.blk_begin;
    it.slot = -1;
    return it;
  };
  // Every role but the loader learns its next unit from the unit ring (whole warp; lane 0 frees the slot).
  auto next_unit = [&](uint32_t &uc, WorkItem &it, int32_t &c_base) -> bool {
    const uint32_t q = uc % UR;
    ptx::mbar_wait(ufull_bar(q), (uc / UR) & 1u);
    const int4 a = ptx::lds128(sUnit + q * 32), b = ptx::lds128(sUnit + q * 32 + 16);
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(uempty_bar(q));
    ++uc;
    it.window = a.x; it.blk_begin = a.y; it.blk_count = a.z; it.slot = a.w;
    c_base = b.x;
    return b.y != 0;
  };

  if (warp == G::kMmaWarp) {
    if (lane == 0) {
      for (uint32_t s = 0; s < S; ++s) {
        ptx::mbar_init(full_bar(s), G::kKsPerStage);   // one arrive(.expect_tx) per producer warp
        ptx::mbar_init(empty_bar(s), 1);               // tcgen05.commit
      }
      for (uint32_t m = 0; m < MR; ++m) {
        ptx::mbar_init(mfull_bar(m), 1);               // loader's arrive.expect_tx
        ptx::mbar_init(mempty_bar(m), NPW);            // one arrive per producer warp
      }
      for (uint32_t a = 0; a < 2; ++a) {
        ptx::mbar_init(tfull_bar(a), 1);               // tcgen05.commit after the item's last MMA
        ptx::mbar_init(tempty_bar(a), 4);              // one arrive per epilogue warp
      }
      for (uint32_t q = 0; q < UR; ++q) {
        ptx::mbar_init(ufull_bar(q), 1);               // loader publishes a unit
        ptx::mbar_init(uempty_bar(q), G::kUnitConsumers);   // every consumer warp has read it
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc<G::kTmemCols>(sTmem);
  }
  if (warp == 0) {
    if (lane == 0) ptx::prefetch_tensormap(&tmap);
    if (lane < 16) {
      constexpr uint32_t one = TcFmt<T>::kOne;
      const uint32_t w0 = ((lane & 1) ? one : 0u) | ((lane & 2) ? (one << 16) : 0u);
      const uint32_t w1 = ((lane & 4) ? one : 0u) | ((lane & 8) ? (one << 16) : 0u);
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sLut + lane * 8), "r"(w0), "r"(w1) : "memory");
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));

  if (warp == G::kLoaderWarp) {
    // ------------------------------------------------------------------ scheduler + metadata loader (one lane)
    // Units are (feature tile, item) pairs, feature-tile-major: all items of one 128-feature column slice of B before
    // the next slice, so that only one slice has to stay L2-resident at a time.  Within a slice the items come in LPT
    // order.  With a ticket counter every CTA claims its next unit with one atomicAdd (claimed one unit ahead, so the
    // round trip to L2 hides behind the current unit's metadata stream); without one, CTAs stride over the list.
    if (ptx::elect_one()) {
      const uint64_t pol = ptx::policy_evict_first();   // hind / bitmaps are streamed once: do not displace B in L2
      uint32_t gc = 0, uc = 0;
      int32_t u = blockIdx.x;
      int32_t ahead = ticket != nullptr ? int32_t(gridDim.x) + atomicAdd(ticket, 1) : u + int32_t(gridDim.x);
      for (;;) {
        const uint32_t q = uc % UR;
        ptx::mbar_wait(uempty_bar(q), ((uc / UR) & 1u) ^ 1u);
        const bool valid = u < total_units;
        WorkItem it;
        it.window = 0; it.blk_begin = 0; it.blk_count = 0; it.slot = -1;
        int32_t c_base = 0;
        if (valid) {
          const int32_t tile = u / num_items;
          it = load_item(u - tile * num_items);
          c_base = tile * G::kFeatTile;
        }
        ptx::sts128(sUnit + q * 32, uint32_t(it.window), uint32_t(it.blk_begin), uint32_t(it.blk_count), uint32_t(it.slot));
        ptx::sts128(sUnit + q * 32 + 16, uint32_t(c_base), valid ? 1u : 0u, 0u, 0u);
        ptx::mbar_arrive(ufull_bar(q));   // release: the two stores above are visible to whoever sees the phase flip
        ++uc;
        if (!valid) break;
        for (int32_t b0 = 0; b0 < it.blk_count; b0 += G::kChunkBlks, ++gc) {
          const uint32_t m = gc % MR;
          const uint32_t nb = uint32_t(min(G::kChunkBlks, it.blk_count - b0));
          ptx::mbar_wait(mempty_bar(m), ((gc / MR) & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(mfull_bar(m), WEIGHTED ? nb * 32u : nb * 48u);
          ptx::bulk_g2s_hint(sMetaH + m * G::kMetaH, hind4 + int64_t(it.blk_begin + b0) * 2, nb * 32u, mfull_bar(m), pol);
          if (!WEIGHTED)   // value tiles do not go through the metadata ring: they land in the stage itself
            ptx::bulk_g2s_hint(sMetaP + m * G::kMetaP, packed + int64_t(it.blk_begin + b0), nb * 16u, mfull_bar(m), pol);
        }
        u = ahead;
        ahead = ticket != nullptr ? int32_t(gridDim.x) + atomicAdd(ticket, 1) : u + int32_t(gridDim.x);
      }
    }
  } else if (warp < NPW) {
    // ------------------------------------------------------------------ producers: bitmap expansion + TMA gather
    // Ring / chunk positions are carried incrementally (no div / mod in the loop); the common K-step (both TC
    // blocks present, both feature halves) runs a select-free path.
    const int n = lane & 15;     // window row
    const int kc = lane >> 4;    // which TC block of the K-step (= which 8-column chunk of K)
    const int kw = warp % G::kKsPerStage;   // this warp's K-step within a stage
    const int grp = warp / G::kKsPerStage;  // stages st with st % kGroups == grp are this warp's
    const uint32_t a_off = uint32_t(kw) * G::kKsA + (n >> 3) * 256 + kc * 128 + (n & 7) * 16;
    const uint32_t p_off = uint32_t(2 * kw + kc) * 16 + uint32_t(n >> 3) * 4, shift = uint32_t(n & 7) << 2;
    uint32_t s = 0, par = 0, m = 0, mpar = 0, uc = 0;
    int32_t turn = 0;                       // global stage counter mod kGroups
#ifdef VX_TC_HUB_POPC
    // L2-policy probe (R-MAT ids: a column's expected degree falls with the number of set bits of its id): gathers whose
    // four rows are all "hub" rows are tagged evict_last, the rest evict_first.
    const uint64_t pol_hot = ptx::policy_evict_last(), pol_cold = ptx::policy_evict_first();
#endif
    auto g4 = [&](uint32_t d, uint32_t bar, int32_t c, const int4 &r) {
#ifdef VX_TC_HUB_POPC
      const bool hot = max(max(__popc(r.x), __popc(r.y)), max(__popc(r.z), __popc(r.w))) <= VX_TC_HUB_POPC;
      ptx::tma_gather4_hint(d, &tmap, bar, c, r.x, r.y, r.z, r.w, hot ? pol_hot : pol_cold);
#else
      ptx::tma_gather4(d, &tmap, bar, c, r.x, r.y, r.z, r.w);
#endif
    };
    constexpr uint32_t KG = G::kKGroupB;    // byte distance of the second 8-row k-group of a K-step
    WorkItem it;
    int32_t c_base;
    while (next_unit(uc, it, c_base)) {
      const int32_t c1 = c_base + G::kAtomCols;
      const bool two_halves = FT == 128 && c1 < N;     // the 64-wide tile has one swizzle atom per gathered row
      const int32_t nks = (it.blk_count + 1) >> 1;
      const int32_t nst = (nks + G::kKsPerStage - 1) / G::kKsPerStage;
      const int32_t full_ks = it.blk_count >> 1;    // K-steps with both TC blocks present
      int32_t cst = 0, ks = kw;
      for (int32_t st = 0; st < nst; ++st, ks += G::kKsPerStage) {
        if (cst == 0) ptx::mbar_wait(mfull_bar(m), mpar);
        const bool mine = G::kGroups == 1 || turn == grp;
        if (G::kGroups > 1 && ++turn == G::kGroups) turn = 0;
        const uint32_t bar = full_bar(s);
        if (mine) ptx::mbar_wait(empty_bar(s), par ^ 1u);
        if (!mine) {
          // another group's stage: only the ring / chunk bookkeeping below
        } else if (ks < nks) {
          const uint32_t moff = uint32_t(cst) * (G::kKsPerStage * 2);   // first block of this stage, chunk-local
          const uint32_t pa = sMetaP + m * G::kMetaP + moff * 16 + p_off;
          const uint32_t ha = sMetaH + m * G::kMetaH + (moff + 2 * kw) * 32;
          const uint32_t dst = sB + s * G::kStageB + uint32_t(kw) * G::kKsB;
          const bool has_b1 = ks < full_ks;                 // odd block count: the item's last K-step is half empty
          // A^T fragment of this lane: 8 K values (one TC block's 8 columns) of window row n, via the nibble table
#if VX_TC_EXP != 5
          if constexpr (!WEIGHTED) {
            uint32_t lo = 0, hi = 0;
            if (has_b1 || kc == 0) {
              lo = (ptx::lds32(pa) >> shift) & 0xfu;          // columns 0..3 of row n
              hi = (ptx::lds32(pa + 8) >> shift) & 0xfu;      // columns 4..7
            }
            const uint2 e0 = ptx::lds64(sLut + lo * 8), e1 = ptx::lds64(sLut + hi * 8);
            ptx::sts128(sA + s * G::kStageA + a_off, e0.x, e0.y, e1.x, e1.y);
            ptx::fence_proxy_async_smem();
            __syncwarp();
          } else if (!has_b1) {
            // odd block count: the K-step's second 8-column chunk does not exist -- zero its 256 bytes (the gathered
            // padding rows must meet zeros), the first chunk arrives by bulk copy below
            if (lane < 16) ptx::sts128(sA + s * G::kStageA + uint32_t(kw) * G::kKsA + 256 + lane * 16, 0u, 0u, 0u, 0u);
            ptx::fence_proxy_async_smem();
            __syncwarp();
          }
#endif
#if VX_TC_DBG == 2
          if (lane == 0) ptx::mbar_arrive(bar);
          if (false) {
#else
          if (ptx::elect_one()) {
#endif
            // atom (k-group kg, feature half j) sits at (kg*2 + j) * 1024; 4 rows = half an atom (512 B);
            // issued row-group-major so a group's 4 row coordinates are converted to uniform registers once
            if (has_b1 && two_halves) {
              const int4 r0 = ptx::lds128(ha), r1 = ptx::lds128(ha + 16), r2 = ptx::lds128(ha + 32),
                         r3 = ptx::lds128(ha + 48);
              ptx::mbar_arrive_expect_tx(bar, uint32_t(G::kKsB) + (WEIGHTED ? 512u : 0u));
              if constexpr (WEIGHTED)
                ptx::bulk_g2s(sA + s * G::kStageA + uint32_t(kw) * G::kKsA, packed + (int64_t(it.blk_begin) + 2 * ks) * 16,
                              512u, bar);
#pragma unroll
              for (int t = 0; t < TERMS; ++t) {   // term t: columns shifted by t * term_stride, tile t of the K-step
                const uint32_t d = dst + t * G::kTermB;
                const int32_t ca = c_base + t * term_stride, cb = c1 + t * term_stride;
                g4(d, bar, ca, r0);
                g4(d + 1024, bar, cb, r0);
                g4(d + 512, bar, ca, r1);
                g4(d + 1536, bar, cb, r1);
                g4(d + KG, bar, ca, r2);
                g4(d + KG + 1024, bar, cb, r2);
                g4(d + KG + 512, bar, ca, r3);
                g4(d + KG + 1536, bar, cb, r3);
              }
            } else {
              const int4 r0 = ptx::lds128(ha), r1 = ptx::lds128(ha + 16);
              int4 r2 = make_int4(0, 0, 0, 0), r3 = make_int4(0, 0, 0, 0);   // tail: row 0, bitmap bits are 0
              if (has_b1) { r2 = ptx::lds128(ha + 32); r3 = ptx::lds128(ha + 48); }
              const uint32_t vbytes = WEIGHTED ? (has_b1 ? 512u : 256u) : 0u;
              ptx::mbar_arrive_expect_tx(bar, uint32_t(two_halves || FT == 64 ? G::kKsB : G::kKsB / 2) + vbytes);
              if constexpr (WEIGHTED)
                ptx::bulk_g2s(sA + s * G::kStageA + uint32_t(kw) * G::kKsA, packed + (int64_t(it.blk_begin) + 2 * ks) * 16,
                              vbytes, bar);
#pragma unroll
              for (int t = 0; t < TERMS; ++t) {
                const uint32_t d = dst + t * G::kTermB;
                const int32_t ca = c_base + t * term_stride, cb = c1 + t * term_stride;
                g4(d, bar, ca, r0);
                g4(d + 512, bar, ca, r1);
                g4(d + KG, bar, ca, r2);
                g4(d + KG + 512, bar, ca, r3);
                if (two_halves) {
                  g4(d + 1024, bar, cb, r0);
                  g4(d + 1536, bar, cb, r1);
                  g4(d + KG + 1024, bar, cb, r2);
                  g4(d + KG + 1536, bar, cb, r3);
                }
              }
            }
          }
        } else if (lane == 0) {
          ptx::mbar_arrive(bar);   // K-step past the item's end: nothing to load, keep the arrival count
        }
        if (++s == S) { s = 0; par ^= 1u; }
        if (cst == G::kStagesPerChunk - 1 || st == nst - 1) {   // done with this metadata chunk
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(mempty_bar(m));
          cst = 0;
          if (++m == MR) { m = 0; mpar ^= 1u; }
        } else {
          ++cst;
        }
      }
    }
  } else if (warp == G::kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::make_idesc(TcFmt<T>::kFmt, /*A MN-major*/ true, /*B K-major*/ false, FT, 16);
    uint32_t s = 0, par = 0, unit = 0, uc = 0;
    WorkItem it;
    int32_t c_base;
    while (next_unit(uc, it, c_base)) {
      const int32_t nks = (it.blk_count + 1) >> 1;
      if (nks == 0) continue;
      const int32_t nst = (nks + G::kKsPerStage - 1) / G::kKsPerStage;
      const uint32_t acc = unit & 1u;
      ptx::mbar_wait(tempty_bar(acc), ((unit >> 1) & 1u) ^ 1u);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * 16;
      for (int32_t st = 0; st < nst; ++st) {
        ptx::mbar_wait(full_bar(s), par);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
          const int32_t kn = min(G::kKsPerStage, nks - st * G::kKsPerStage);
          // A = gathered rows: MN-major SW128, LBO = feature-atom stride (1024), SBO = k-group stride (2048)
          // B = densified tile: K-major, no swizzle, LBO = k-chunk stride (128), SBO = 8-row group stride (256)
          // Only the 14-bit start-address field changes from K-step to K-step (+4096 B / +512 B, no carry out of
          // the field: shared memory is < 256 KB), so the low words are advanced by constants.
          const uint64_t a_desc = ptx::smem_desc(sB + s * G::kStageB, 1024, G::kKGroupB, ptx::kLayoutSw128);
          // (value tiles: one TC block = 256 contiguous bytes, so k-chunk stride 256 and 8-row group stride 128)
          const uint64_t b_desc = WEIGHTED ? ptx::smem_desc(sA + s * G::kStageA, 256, 128, ptx::kLayoutNone)
                                           : ptx::smem_desc(sA + s * G::kStageA, 128, 256, ptx::kLayoutNone);
          const uint32_t a_lo = uint32_t(a_desc), a_hi = uint32_t(a_desc >> 32);
          const uint32_t b_lo = uint32_t(b_desc), b_hi = uint32_t(b_desc >> 32);
#if VX_TC_DBG == 1
          if (false) {
#else
          if (kn == G::kKsPerStage) {
#endif
#pragma unroll
            for (int32_t k = 0; k < G::kKsPerStage; ++k)
#pragma unroll
              for (int32_t t = 0; t < TERMS; ++t)   // every term of the K-step multiplies the same densified tile
                ptx::umma_f16_split(d_tmem, a_lo + ((k * G::kKsB + t * G::kTermB) >> 4), a_hi, b_lo + k * (G::kKsA >> 4),
                                    b_hi, idesc, (k | t) > 0 ? 1u : (st > 0 ? 1u : 0u));
          } else if (VX_TC_DBG != 1) {
            for (int32_t k = 0; k < kn; ++k)
#pragma unroll
              for (int32_t t = 0; t < TERMS; ++t)
                ptx::umma_f16_split(d_tmem, a_lo + ((k * G::kKsB + t * G::kTermB) >> 4), a_hi, b_lo + k * (G::kKsA >> 4),
                                    b_hi, idesc, (st | k | t) > 0 ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(s));
          if (st == nst - 1) ptx::umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++s == S) { s = 0; par ^= 1u; }
      }
      ++unit;
    }
  } else if (warp >= G::kEpilogueWarp0 && warp < G::kEpilogueWarp0 + 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> C
    const int ew = warp - G::kEpilogueWarp0;   // == warp % 4: the TMEM lane quarter this warp may read
    uint32_t unit = 0, uc = 0;
    WorkItem it;
    int32_t c_base;
    while (next_unit(uc, it, c_base)) {
      // M = 128: TMEM lane = feature, a warp reads its 32-lane quarter.  M = 64: the accumulator's 64 rows sit in lanes
      // 0-15 of every quarter (rows 16 e ... 16 e + 15 in quarter e; scripts/probes/tmem_m64_probe.cu), the other lanes idle.
      const bool holds_row = FT == 128 || lane < 16;
      const int32_t f = FT == 128 ? c_base + ew * 32 + lane : c_base + ew * 16 + lane;
      uint32_t v[16];
      if (it.blk_count > 0) {
        const uint32_t acc = unit & 1u;
        ptx::mbar_wait(tfull_bar(acc), (unit >> 1) & 1u);
        ptx::tc_fence_after_sync();
        ptx::tmem_ld_32x32b_x16(tmem_base + (uint32_t(ew * 32) << 16) + acc * 16, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
        ++unit;
      } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = 0u;
      }
      if (holds_row && f < N) {
        float *dst;
        int32_t nrows = BLK_H;
        if (it.slot < 0) {
          dst = C + int64_t(it.window) * BLK_H * N + f;
          nrows = min(BLK_H, num_nodes - it.window * BLK_H);   // partial tail window: rows >= M do not exist
          if (epi.any()) {   // fused epilogue; K-split partial tiles (slot >= 0) get it in the fix-up pass instead
            const float bf = epi.bias_of(f), pre = epi.pre_scale();
#pragma unroll
            for (int r = 0; r < BLK_H; ++r)
              if (r < nrows)
                v[r] = __float_as_uint(epi.apply(__uint_as_float(v[r]), epi.scale_of(int64_t(it.window) * BLK_H + r) * pre, bf));
          }
        } else {
          dst = scratch + int64_t(it.slot) * BLK_H * N + f;
        }
#pragma unroll
        for (int r = 0; r < BLK_H; ++r)
          if (r < nrows) __stcs(dst + int64_t(r) * N, __uint_as_float(v[r]));
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == G::kMmaWarp) ptx::tmem_dealloc<G::kTmemCols>(tmem_base);
}

// Sums the partial tiles of K-split windows in slot order (fixed order => deterministic).
__global__ void vx_spmm_fixup_kernel(const FixupItem *__restrict__ fixups, int32_t num_fixups,
                                const float *__restrict__ scratch, int32_t num_nodes, int32_t N,
                                float *__restrict__ C, Epilogue epi, const int32_t *__restrict__ gate = nullptr,
                                int32_t gate_want = 0) {
  if (gate != nullptr && *gate != gate_want) return;
  const int32_t i = blockIdx.x;
  if (i >= num_fixups) return;
  const float pre = epi.pre_scale();
  const FixupItem fx = fixups[i];
  const int32_t nrows = min(BLK_H, num_nodes - fx.window * BLK_H);
  const int32_t elems = nrows * N;   // rows of a slot are laid out [16][N]
  for (int32_t e = threadIdx.x; e < elems; e += blockDim.x) {
    float s = 0.f;
    for (int32_t k = 0; k < fx.slot_count; ++k) s += scratch[int64_t(fx.slot_begin + k) * BLK_H * N + e];
    if (epi.any()) s = epi.apply(s, epi.scale_of(int64_t(fx.window) * BLK_H + e / N) * pre, epi.bias_of(e % N));
    C[int64_t(fx.window) * BLK_H * N + e] = s;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_vxTensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                               const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_vxTensorMapEncodeTiled get_tensor_map_encoder() {
  static PFN_vxTensorMapEncodeTiled fn = nullptr;
  if (fn) return fn;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_vxTensorMapEncodeTiled>(p);
  return fn;
}

// 2-D map over B[rows, N] whose box is one 128-byte row segment: the shape tile::gather4 needs.
inline int make_gather_tensor_map(CUtensorMap *out, const void *B, CUtensorMapDataType dt, int elem_bytes,
                                  int64_t rows, int64_t N) {
  PFN_vxTensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return VX_ERR_CUDA;
  cuuint64_t gdim[2] = {cuuint64_t(N), cuuint64_t(rows)};
  cuuint64_t gstride[1] = {cuuint64_t(N) * elem_bytes};
  cuuint32_t box[2] = {cuuint32_t(128 / elem_bytes), 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void *>(B), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[voltrix] cuTensorMapEncodeTiled failed: %d\n", int(r));
    return VX_ERR_CUDA;
  }
  return VX_OK;
}

inline int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// Launch the tensor-core kernel over a prepared work list.  B must be 16-byte aligned with N % 8 == 0
// (TMA global-stride rule); hind / hspa_packed must be 16-byte aligned.
// `B` holds TERMS column blocks of N elements per row (row stride TERMS * N): plain fp16 / bf16 input has TERMS = 1.
template <typename T, int STAGES, int NPW, int TERMS = 1, bool WEIGHTED = false, int FT = 128>
inline int launch_spmm_tc(const WorkItem *items, int32_t num_items, const FixupItem *fixups, int32_t num_fixups,
                          const int32_t *blk_offsets, const uint32_t *hspa_packed, const int32_t *hind,
                          int32_t num_nodes, int64_t b_rows,
                          int32_t N, const T *B, float *C, float *scratch, cudaStream_t stream,
                          const Epilogue &epi = Epilogue(), int32_t *ticket = nullptr, const int32_t *gate = nullptr,
                          int32_t gate_want = 0, bool reset_ticket = true) {
  if (num_items <= 0) return VX_OK;
  if (N % 8 != 0 || (reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(hind) & 15) ||
      (reinterpret_cast<uintptr_t>(hspa_packed) & 15))
    return VX_ERR_UNSUPPORTED;
  CUtensorMap tmap;
  int rc = make_gather_tensor_map(&tmap, B, TcFmt<T>::kTmapType, 2, b_rows, int64_t(N) * TERMS);
  if (rc != VX_OK) return rc;
  using G = TcGeom<NPW, TERMS, FT>;
  auto kern = vx_spmm_tc_kernel<T, STAGES, NPW, TERMS, WEIGHTED, FT>;
  constexpr size_t smem = tc_smem_bytes<STAGES, NPW, TERMS, FT>();
  static_assert(smem <= 227 * 1024, "stage ring does not fit in shared memory");
  // Set on every launch (~1 us): a function-local `static bool` would be a GNU_UNIQUE symbol shared by every
  // JIT artefact / library that instantiates this template, while each of them owns a distinct kernel copy.
  VX_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int32_t n_feat_tiles = ceil_div(N, G::kFeatTile);
  const int64_t total_units = int64_t(num_items) * n_feat_tiles;
  const int64_t resident = int64_t(device_sm_count()) * tc_ctas_per_sm<STAGES, NPW, TERMS, FT>();
  const int grid = int(total_units < resident ? total_units : resident);
  // `ticket` (4 bytes of device memory owned by the caller, one per stream in flight): dynamic unit claiming.
  // Zeroed here, on the stream, so a launch never depends on how the previous one ended; nullptr = static striding.
  // reset_ticket = false: the caller has zeroed it already on this stream (the gated pipelines of model 4 share one memset).
  if (ticket != nullptr && reset_ticket) VX_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(int32_t), stream));
  kern<<<grid, G::kThreads, smem, stream>>>(tmap, items, num_items, n_feat_tiles, blk_offsets,
                                                 reinterpret_cast<const uint4 *>(hspa_packed),
                                                 reinterpret_cast<const int4 *>(hind), num_nodes, N, C, scratch, N, epi,
                                                 ticket, gate, gate_want);
  VX_LAUNCH_CHECK();
  if (num_fixups > 0) {
    vx_spmm_fixup_kernel<<<num_fixups, 256, 0, stream>>>(fixups, num_fixups, scratch, num_nodes, N, C, epi, gate, gate_want);
    VX_LAUNCH_CHECK();
  }
  return VX_OK;
}

// fp32 -> [hi | lo] bf16 terms: out[r, c] = bf16(x), out[r, N + c] = bf16(x - hi).  One thread per 4 values.
__global__ void vx_spmm_split_bf16x2_kernel(const float4 *__restrict__ in, __nv_bfloat16 *__restrict__ out, int64_t rows,
                                       int32_t N, const int32_t *__restrict__ gate = nullptr, int32_t gate_want = 0) {
  if (gate != nullptr && *gate != gate_want) return;
  const int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;   // quad index
  const int32_t qpr = N >> 2;
  if (q >= rows * qpr) return;
  const int64_t r = q / qpr;
  const int32_t c = int32_t(q - r * qpr) << 2;
  const float4 v = in[q];
  const float x[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hi[i] = __float2bfloat16_rn(x[i]);
    // Inf, or a finite value that rounds to bf16 Inf: the high term alone carries it (x - hi would be Inf - Inf = NaN)
    const float h = __bfloat162float(hi[i]);
    lo[i] = __float2bfloat16_rn(isfinite(h) ? x[i] - h : 0.f);
  }
  __nv_bfloat16 *o = out + r * (2 * int64_t(N)) + c;
  *reinterpret_cast<uint2 *>(o) = *reinterpret_cast<const uint2 *>(hi);
  *reinterpret_cast<uint2 *>(o + N) = *reinterpret_cast<const uint2 *>(lo);
}

inline int launch_split_bf16x2(const float *in, __nv_bfloat16 *out, int64_t rows, int32_t N, cudaStream_t stream,
                               const int32_t *gate = nullptr, int32_t gate_want = 0) {
  if (N % 4 != 0 || (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 7))
    return VX_ERR_UNSUPPORTED;
  const int64_t quads = rows * (N >> 2);
  if (quads <= 0) return VX_OK;
  vx_spmm_split_bf16x2_kernel<<<unsigned((quads + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4 *>(in), out,
                                                                            rows, N, gate, gate_want);
  VX_LAUNCH_CHECK();
  return VX_OK;
}

// ---- model 4: an fp32 operand carried as ONE fp16 term -------------------------------------------------------------------
// fp16 keeps 11 significant bits -- one more than the TF32 the reference rounds its operand to (spmm_kernels.cuh:1631-1678)
// -- for values inside its normal range, 2^-14 .. 65504.  The operand is therefore scaled by a power of two (exact) that
// puts its largest magnitude just under fp16's maximum, which leaves 2^29 of dynamic range below it; the accumulator is
// scaled back in the epilogue.  Three launches on the stream, no host round trip:
//   range pass    max |x| (atomicMax on the bit pattern), non-finite values
//   convert pass  x * 2^s -> fp16, and counts the groups of 128 consecutive values in which a non-zero value lands below
//                 fp16's normal range (it keeps fewer than 11 bits there, none below 2^-39 of the largest magnitude)
//   decide        flag = 1 if the operand holds Inf / NaN, if 2^s is not a normal float, or if more than one group in
//                 kF16TailGroups (and more than kF16TailFloor groups) has such a value -- then the gated two-term bf16
//                 pipeline runs instead.
// The count, not the minimum, decides: one Gaussian sample in 10^8 is that close to zero, and a gate on the minimum sends a
// large operand down the slow pipeline at random (1.8x the time); an operand with a structurally wide range -- rows or
// columns scaled apart by more than 2^29 -- fails the count by orders of magnitude.  A value under the gate's radar carries
// an absolute error below 2^-40 of the operand's largest magnitude.
// state[0] = flag (after `decide`), state[1] = bits of max |x|, state[2] = tail-group count / not-representable bit.
constexpr int32_t kF16TailGroups = 8192;          // tolerated: one 128-value group in 8192 with a sub-normal-range value
constexpr int32_t kF16TailFloor = 3;              // small operands: a handful of stray values is still not structure
constexpr int32_t kF16NotRepresentable = 1 << 30;

__global__ void vx_spmm_absrange_kernel(const float4 *__restrict__ in, int64_t quads, int32_t *__restrict__ state) {
  uint32_t mx = 0u;
  bool nonfinite = false;
  for (int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < quads; q += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = in[q];
    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t a = __float_as_uint(x[i]) & 0x7fffffffu;
      nonfinite |= a >= 0x7f800000u;
      mx = max(mx, a);
    }
  }
  for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  nonfinite = __any_sync(0xffffffffu, nonfinite);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<unsigned int *>(state + 1), mx);
    if (nonfinite) atomicOr(state + 2, kF16NotRepresentable);
  }
}

__global__ void vx_spmm_cvt_f16_kernel(const float4 *__restrict__ in, __half *__restrict__ out, int64_t quads,
                                       int32_t *__restrict__ state) {
  const int32_t max_bits = state[1];
  const int s = Epilogue::carrier_shift(max_bits);
  const int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (!(s > -100 && s < 100)) {                 // 2^s must be a normal float
    if (q == 0) atomicOr(state + 2, kF16NotRepresentable);
    return;
  }
  bool tail = false;
  if (q < quads) {
    const float k = __int_as_float((127 + s) << 23);   // 2^s
    const float4 v = in[q];
    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t a = __float_as_uint(x[i]) & 0x7fffffffu;
      tail |= a != 0u && int(a >> 23) - 127 + s < -14;
    }
    __half2 h[2] = {__floats2half2_rn(v.x * k, v.y * k), __floats2half2_rn(v.z * k, v.w * k)};
    *reinterpret_cast<uint2 *>(out + q * 4) = *reinterpret_cast<const uint2 *>(h);
  }
  if (__any_sync(0xffffffffu, tail) && (threadIdx.x & 31) == 0) atomicAdd(state + 2, 1);
}

__global__ void vx_spmm_f16_gate_kernel(int32_t *__restrict__ state, int32_t limit) {
  state[0] = state[2] > limit ? 1 : 0;          // the not-representable bit is far above any limit
}

// state: 3 ints of device memory (flag, max bits, tail count), initialised here
// state_zeroed: the caller has already zeroed the three words on this stream
inline int launch_cvt_f16(const float *in, __half *out, int64_t rows, int32_t N, int32_t *state, cudaStream_t stream,
                          bool state_zeroed = false) {
  if (N % 4 != 0 || (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 7) || state == nullptr)
    return VX_ERR_UNSUPPORTED;
  if (!state_zeroed) VX_CUDA_TRY(cudaMemsetAsync(state, 0, 3 * sizeof(int32_t), stream));   // flag, max |x|, tail groups
  const int64_t quads = rows * (N >> 2);
  if (quads <= 0) return VX_OK;
  const int64_t groups = (quads + 31) / 32;                                 // 128-value groups = warps of the convert pass
  if (groups >= kF16NotRepresentable) return VX_ERR_UNSUPPORTED;            // the count shares a word with that bit
  const int64_t want_ctas = (quads + 255) / 256, cap_ctas = int64_t(device_sm_count()) * 16;
  const int grid = int(want_ctas < cap_ctas ? want_ctas : cap_ctas);
  vx_spmm_absrange_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4 *>(in), quads, state);
  VX_LAUNCH_CHECK();
  vx_spmm_cvt_f16_kernel<<<unsigned((quads + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4 *>(in), out, quads,
                                                                            state);
  VX_LAUNCH_CHECK();
  const int64_t tolerated = groups / kF16TailGroups;                          // ... and never fewer than kF16TailFloor groups
  vx_spmm_f16_gate_kernel<<<1, 1, 0, stream>>>(state, int32_t(tolerated > kF16TailFloor ? tolerated : kF16TailFloor));
  VX_LAUNCH_CHECK();
  return VX_OK;
}

}  // namespace voltrix

#endif  // VOLTRIX_B200_SPMM_TCGEN05_CUH_
