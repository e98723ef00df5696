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
//   * A^T is the "B" operand, K-major, no swizzle: 512 B expanded from two 16-byte bitmaps by one warp;
//   * D^T lives in TMEM (128 lanes x 16 fp32 columns, double buffered across work items) and is read
//     back with tcgen05.ld 32x32b: a thread holds the 16 window rows of one feature, so every store
//     instruction of a warp writes 32 consecutive floats of one C row;
//   * warp roles: 0 = TMA gather producer (4 lanes per K-step, 8 K-steps in flight per pass),
//     1 = MMA issuer (lane 0), 2 = bitmap expander, 3 = TMEM allocator, 4-7 = epilogue;
//   * a ring of STAGES (B tile + A tile) slots guarded by full/empty mbarriers; tcgen05.commit frees a
//     slot when the MMA that read it retires;
//   * persistent CTAs stride over the LPT-sorted work list (schedule.cuh), every role derives the same
//     item sequence independently, so no intra-CTA work broadcast is needed.
#ifndef VOLTRIX_B200_SPMM_TCGEN05_CUH_
#define VOLTRIX_B200_SPMM_TCGEN05_CUH_

#include "voltrix/common.cuh"
#include "voltrix/ptx.cuh"

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

struct TcGeom {
  static constexpr int kFeatTile = 128;                 // MMA M: features per work unit
  static constexpr int kAtomCols = 64;                  // 128-byte swizzle span in 16-bit elements
  static constexpr int kStageB = kFeatTile * 16 * 2;    // 16 gathered rows x 128 features x 2 B = 4096
  static constexpr int kStageA = 16 * 16 * 2;           // densified 16 x 16 tile = 512
  static constexpr int kThreads = 256;
  static constexpr uint32_t kTmemCols = 32;             // two 16-column accumulators
};

template <int STAGES>
constexpr size_t tc_smem_bytes() {
  return size_t(STAGES) * (TcGeom::kStageB + TcGeom::kStageA) + (2 * STAGES + 4) * 8 + 16 + 1024 /*align slack*/;
}

template <typename T, int STAGES>
__global__ void __launch_bounds__(TcGeom::kThreads, 1)
vx_spmm_tc_kernel(const __grid_constant__ CUtensorMap tmap, const WorkItem *__restrict__ items, int32_t num_items,
                  int32_t n_feat_tiles, const int32_t *__restrict__ blk_offsets, const uint4 *__restrict__ packed,
                  const int4 *__restrict__ hind4, int32_t num_nodes, int32_t N, float *__restrict__ C,
                  float *__restrict__ scratch) {
  static_assert(STAGES >= 8 && (STAGES & (STAGES - 1)) == 0, "STAGES must be a power of two >= 8");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB = sbase;                                   // [STAGES][4096]  1024-aligned atoms
  const uint32_t sA = sB + STAGES * TcGeom::kStageB;           // [STAGES][512]
  const uint32_t sBar = sA + STAGES * TcGeom::kStageA;         // full[STAGES], empty[STAGES], tfull[2], tempty[2]
  const uint32_t sTmem = sBar + (2 * STAGES + 4) * 8;
  auto full_bar = [&](uint32_t s) { return sBar + s * 8; };
  auto empty_bar = [&](uint32_t s) { return sBar + (STAGES + s) * 8; };
  auto tfull_bar = [&](uint32_t a) { return sBar + (2 * STAGES + a) * 8; };
  auto tempty_bar = [&](uint32_t a) { return sBar + (2 * STAGES + 2 + a) * 8; };

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

  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 4 + 1);   // 4 producer lanes (arrive.expect_tx) + expander
      ptx::mbar_init(empty_bar(s), 1);      // tcgen05.commit
    }
    for (uint32_t a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);      // tcgen05.commit after the item's last MMA
      ptx::mbar_init(tempty_bar(a), 4);     // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) ptx::prefetch_tensormap(&tmap);
  if (warp == 3) ptx::tmem_alloc<TcGeom::kTmemCols>(sTmem);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA gather producer
    const int q = lane & 3;      // which 4 of the 16 gathered rows of the K-step
    const int sub = lane >> 2;   // which of the 8 K-steps of this pass
    uint32_t g = 0;
    for (int32_t u = blockIdx.x; u < total_units; u += gridDim.x) {
      const WorkItem it = load_item(u / n_feat_tiles);
      const int32_t c_base = (u % n_feat_tiles) * TcGeom::kFeatTile;
      const int32_t nj = min(2, (N - c_base + TcGeom::kAtomCols - 1) / TcGeom::kAtomCols);
      const int32_t nks = (it.blk_count + 1) >> 1;
      for (int32_t ks0 = 0; ks0 < nks; ks0 += 8) {
        const int32_t ks = ks0 + sub;
        if (ks < nks) {
          const uint32_t gg = g + ks, stage = gg % STAGES, par = ((gg / STAGES) & 1u) ^ 1u;
          const int32_t blk = 2 * ks + (q >> 1);
          int4 rows = make_int4(0, 0, 0, 0);   // K-step tail past an odd block count: row 0, bitmap bits are 0
          if (blk < it.blk_count) rows = __ldg(hind4 + (int64_t(it.blk_begin + blk) * 2 + (q & 1)));
          ptx::mbar_wait(empty_bar(stage), par);
          // atom (k-group kg = q>>1, feature half j) sits at (kg*2 + j) * 1024; 4 rows = half an atom
          const uint32_t dst = sB + stage * TcGeom::kStageB + (q >> 1) * 2048 + (q & 1) * 512;
          ptx::mbar_arrive_expect_tx(full_bar(stage), nj * 512);
          for (int32_t j = 0; j < nj; ++j)
            ptx::tma_gather4(dst + j * 1024, &tmap, full_bar(stage), c_base + j * TcGeom::kAtomCols, rows.x, rows.y,
                             rows.z, rows.w);
        }
      }
      g += nks;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::make_idesc(TcFmt<T>::kFmt, /*A MN-major*/ true, /*B K-major*/ false, 128, 16);
    uint32_t g = 0, unit = 0;
    for (int32_t u = blockIdx.x; u < total_units; u += gridDim.x, ++unit) {
      const WorkItem it = load_item(u / n_feat_tiles);
      const int32_t nks = (it.blk_count + 1) >> 1;
      const uint32_t acc = unit & 1u;
      ptx::mbar_wait(tempty_bar(acc), ((unit >> 1) & 1u) ^ 1u);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * 16;
      for (int32_t ks = 0; ks < nks; ++ks) {
        const uint32_t gg = g + ks, stage = gg % STAGES;
        ptx::mbar_wait(full_bar(stage), (gg / STAGES) & 1u);
        ptx::tc_fence_after_sync();
        if (lane == 0) {
          // A = gathered rows: MN-major SW128, LBO = feature-atom stride (1024), SBO = k-group stride (2048)
          const uint64_t a_desc = ptx::smem_desc(sB + stage * TcGeom::kStageB, 1024, 2048, ptx::kLayoutSw128);
          // B = densified tile: K-major, no swizzle, LBO = k-chunk stride (128), SBO = 8-row group stride (256)
          const uint64_t b_desc = ptx::smem_desc(sA + stage * TcGeom::kStageA, 128, 256, ptx::kLayoutNone);
          ptx::umma_f16(d_tmem, a_desc, b_desc, idesc, ks > 0 ? 1u : 0u);
          ptx::umma_commit(empty_bar(stage));
          if (ks == nks - 1) ptx::umma_commit(tfull_bar(acc));
        }
        __syncwarp();
      }
      g += nks;
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ bitmap -> dense A^T tile
    const int n = lane & 15;     // window row
    const int kc = lane >> 4;    // which TC block of the K-step (= which 8-column chunk of K)
    const uint32_t a_off = (n >> 3) * 256 + kc * 128 + (n & 7) * 16;
    const int word = n >> 3, shift = (n & 7) << 2;
    uint32_t g = 0;
    for (int32_t u = blockIdx.x; u < total_units; u += gridDim.x) {
      const WorkItem it = load_item(u / n_feat_tiles);
      const int32_t nks = (it.blk_count + 1) >> 1;
      // one coalesced load covers 32 blocks = 16 K-steps; words are redistributed by shuffle
      for (int32_t ks0 = 0; ks0 < nks; ks0 += 16) {
        const int32_t myblk = 2 * ks0 + lane;
        uint4 bits = make_uint4(0, 0, 0, 0);
        if (myblk < it.blk_count) bits = __ldg(packed + it.blk_begin + myblk);
        const int32_t kend = min(16, nks - ks0);
        for (int32_t i = 0; i < kend; ++i) {
          const int src = 2 * i + kc;
          const uint32_t w0 = __shfl_sync(0xffffffffu, bits.x, src), w1 = __shfl_sync(0xffffffffu, bits.y, src);
          const uint32_t w2 = __shfl_sync(0xffffffffu, bits.z, src), w3 = __shfl_sync(0xffffffffu, bits.w, src);
          const uint32_t lo = ((word ? w1 : w0) >> shift) & 0xfu;   // columns 0..3 of row n
          const uint32_t hi = ((word ? w3 : w2) >> shift) & 0xfu;   // columns 4..7
          constexpr uint32_t one = TcFmt<T>::kOne;
          uint4 v;
          v.x = ((lo & 1u) ? one : 0u) | ((lo & 2u) ? (one << 16) : 0u);
          v.y = ((lo & 4u) ? one : 0u) | ((lo & 8u) ? (one << 16) : 0u);
          v.z = ((hi & 1u) ? one : 0u) | ((hi & 2u) ? (one << 16) : 0u);
          v.w = ((hi & 4u) ? one : 0u) | ((hi & 8u) ? (one << 16) : 0u);
          const uint32_t gg = g + ks0 + i, stage = gg % STAGES;
          ptx::mbar_wait(empty_bar(stage), ((gg / STAGES) & 1u) ^ 1u);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sA + stage * TcGeom::kStageA + a_off),
                       "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                       : "memory");
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(full_bar(stage));
        }
      }
      g += nks;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> C
    const int ew = warp - 4;   // == warp % 4: the TMEM lane quarter this warp may read
    uint32_t unit = 0;
    for (int32_t u = blockIdx.x; u < total_units; u += gridDim.x, ++unit) {
      const WorkItem it = load_item(u / n_feat_tiles);
      const int32_t f = (u % n_feat_tiles) * TcGeom::kFeatTile + ew * 32 + lane;
      const uint32_t acc = unit & 1u;
      ptx::mbar_wait(tfull_bar(acc), (unit >> 1) & 1u);
      ptx::tc_fence_after_sync();
      uint32_t v[16];
      ptx::tmem_ld_32x32b_x16(tmem_base + (uint32_t(ew * 32) << 16) + acc * 16, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
      if (f < N) {
        float *dst;
        int32_t nrows = BLK_H;
        if (it.slot < 0) {
          dst = C + int64_t(it.window) * BLK_H * N + f;
          nrows = min(BLK_H, num_nodes - it.window * BLK_H);   // partial tail window: rows >= M do not exist
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
  if (warp == 3) ptx::tmem_dealloc<TcGeom::kTmemCols>(tmem_base);
}

// Sums the partial tiles of K-split windows in slot order (fixed order => deterministic).
__global__ void vx_fixup_kernel(const FixupItem *__restrict__ fixups, int32_t num_fixups,
                                const float *__restrict__ scratch, int32_t num_nodes, int32_t N,
                                float *__restrict__ C) {
  const int32_t i = blockIdx.x;
  if (i >= num_fixups) return;
  const FixupItem fx = fixups[i];
  const int32_t nrows = min(BLK_H, num_nodes - fx.window * BLK_H);
  const int32_t elems = nrows * N;   // rows of a slot are laid out [16][N]
  for (int32_t e = threadIdx.x; e < elems; e += blockDim.x) {
    float s = 0.f;
    for (int32_t k = 0; k < fx.slot_count; ++k) s += scratch[int64_t(fx.slot_begin + k) * BLK_H * N + e];
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
                                  int64_t rows, int32_t N) {
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
template <typename T, int STAGES>
inline int launch_spmm_tc(const WorkItem *items, int32_t num_items, const FixupItem *fixups, int32_t num_fixups,
                          const int32_t *blk_offsets, const uint32_t *hspa_packed, const int32_t *hind,
                          int32_t num_nodes, int64_t b_rows,
                          int32_t N, const T *B, float *C, float *scratch, cudaStream_t stream) {
  if (num_items <= 0) return VX_OK;
  if (N % 8 != 0 || (reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(hind) & 15) ||
      (reinterpret_cast<uintptr_t>(hspa_packed) & 15))
    return VX_ERR_UNSUPPORTED;
  CUtensorMap tmap;
  int rc = make_gather_tensor_map(&tmap, B, TcFmt<T>::kTmapType, 2, b_rows, N);
  if (rc != VX_OK) return rc;
  auto kern = vx_spmm_tc_kernel<T, STAGES>;
  constexpr size_t smem = tc_smem_bytes<STAGES>();
  // Set on every launch (~1 us): a function-local `static bool` would be a GNU_UNIQUE symbol shared by every
  // JIT artefact / library that instantiates this template, while each of them owns a distinct kernel copy.
  VX_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int32_t n_feat_tiles = ceil_div(N, TcGeom::kFeatTile);
  const int64_t total_units = int64_t(num_items) * n_feat_tiles;
  const int grid = int(total_units < device_sm_count() ? total_units : device_sm_count());
  kern<<<grid, TcGeom::kThreads, smem, stream>>>(tmap, items, num_items, n_feat_tiles, blk_offsets,
                                                 reinterpret_cast<const uint4 *>(hspa_packed),
                                                 reinterpret_cast<const int4 *>(hind), num_nodes, N, C, scratch);
  VX_LAUNCH_CHECK();
  if (num_fixups > 0) {
    vx_fixup_kernel<<<num_fixups, 256, 0, stream>>>(fixups, num_fixups, scratch, num_nodes, N, C);
    VX_LAUNCH_CHECK();
  }
  return VX_OK;
}

}  // namespace voltrix

#endif  // VOLTRIX_B200_SPMM_TCGEN05_CUH_
