// Where do the rows of an M=64 tcgen05.mma accumulator land in TMEM?  D[m, n] = (m + 1) * (n + 1) from rank-1 operands,
// read back with tcgen05.ld 32x32b.x16 by four warps (lane quarter = warp % 4) and printed.
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -Ivoltrix-spmm_b200/csrc scripts/probes/tmem_m64_probe.cu -o /tmp/m64 && /tmp/m64
#include <cstdio>
#include <cuda_fp16.h>
#include "voltrix/ptx.cuh"
using namespace voltrix;

__global__ void probe(float *out, int M) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base;            // A: MN-major SW128, 16 k x 64 (or 128) m, one 8-row k-group = 1024 B per 64-m atom
  const uint32_t sB = base + 8192;     // B: K-major no swizzle 16 k x 16 n: (n>>3)*256 + (k>>3)*128 + (n&7)*16 + (k&7)*2
  const uint32_t sBar = base + 8192 + 1024, sTm = sBar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A[m, k]: k = 0 row holds m + 1, other k rows zero.  Atom j (64 m) of k-group g at (g * atoms + j) * 1024; inside an atom row
  // r (k & 7) is 128 B, 16-byte chunk c sits at chunk (c ^ r).
  for (int i = threadIdx.x; i < 8192 / 2; i += blockDim.x) reinterpret_cast<__half *>(raw + (base - ptx::smem_u32(raw)))[i] = __float2half(0.f);
  __syncthreads();
  const int atoms = M / 64;
  __half *A = reinterpret_cast<__half *>(raw + (sA - ptx::smem_u32(raw)));
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const int j = m / 64, mm = m % 64, chunk = mm / 8, r = 0;       // k = 0 -> k-group 0, row 0
    A[((0 * atoms + j) * 1024 + r * 128 + ((chunk ^ r) * 16)) / 2 + (mm % 8)] = __float2half(float(m + 1));
  }
  __half *B = reinterpret_cast<__half *>(raw + (sB - ptx::smem_u32(raw)));
  for (int i = threadIdx.x; i < 256; i += blockDim.x) B[i] = __float2half(0.f);
  __syncthreads();
  if (threadIdx.x < 16) { const int n = threadIdx.x; B[((n >> 3) * 256 + 0 * 128 + (n & 7) * 16 + 0) / 2] = __float2half(float(n + 1)); }
  ptx::fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) { ptx::mbar_init(sBar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc<32>(sTm);
  }
  ptx::tc_fence_before_sync(); __syncthreads(); ptx::tc_fence_after_sync();
  uint32_t tm; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tm) : "r"(sTm));
  if (warp == 0 && ptx::elect_one()) {
    const uint32_t idesc = ptx::make_idesc(0, true, false, M, 16);
    const uint64_t ad = ptx::smem_desc(sA, 1024, uint32_t(atoms) * 1024, ptx::kLayoutSw128);
    const uint64_t bd = ptx::smem_desc(sB, 128, 256, ptx::kLayoutNone);
    ptx::umma_f16(tm, ad, bd, idesc, 0);
    ptx::umma_commit(sBar);
  }
  __syncthreads();
  ptx::mbar_wait(sBar, 0);
  ptx::tc_fence_after_sync();
  uint32_t v[16];
  ptx::tmem_ld_32x32b_x16(tm + (uint32_t(warp * 32) << 16), v);
  ptx::tmem_ld_wait();
  for (int c = 0; c < 16; ++c) out[(warp * 32 + lane) * 16 + c] = __uint_as_float(v[c]);
  ptx::tc_fence_before_sync(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<32>(tm);
}

int main() {
  float *d; cudaMalloc(&d, 128 * 16 * 4);
  for (int M : {128, 64}) {
    cudaMemset(d, 0xff, 128 * 16 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    probe<<<1, 128, 16384>>>(d, M);
    cudaError_t e = cudaDeviceSynchronize();
    float h[128 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=%d (%s): TMEM lane -> row (from column 0 = m + 1), check column 3 = 4 (m + 1)\n", M, cudaGetErrorString(e));
    for (int l = 0; l < 128; ++l) {
      const float r = h[l * 16];
      printf("%s%3d:%s", l % 16 == 0 ? "\n  " : " ", l, (r == r && r > 0 && r < 200 && h[l * 16 + 3] == 4 * r) ? "" : "");
      if (r == r && r > 0 && r < 200 && h[l * 16 + 3] == 4 * r) printf("%3d", int(r) - 1); else printf("  .");
    }
    printf("\n");
  }
  return 0;
}
