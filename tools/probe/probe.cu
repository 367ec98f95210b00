// probe.cu — hardware probe (debug entry point, not on the product path): does a K-major UMMA shared-memory descriptor
// accept a start address that is NOT aligned to the swizzle atom (arbitrary row shift) and a stride-byte-offset that is
// not a multiple of the atom?  This decides whether one haloed activation tile can serve all nine taps of a 3x3 conv.
// D[m][n] = A[m][n] with B = identity, A rows taken from a TMA-written (swizzled) R x (ROWB bytes) tile:
//   row(m) = shift + (m / 8) * (sbo / ROWB) + m % 8.
#include "../../open-solution-mapping-challenge_b200/csrc/host_common.h"
#include "../../open-solution-mapping-challenge_b200/csrc/tc.cuh"


namespace mcb {

struct ProbeParams {
  CUtensorMap tmA;  // (rowb/2 elements, R rows), box = whole tile
  CUtensorMap tmB;  // identity (K x K), K-major
  float* out;       // [128][N]
  int rows, rowb, shift, sbo, base_off_mode, kdim;
};

__global__ void __launch_bounds__(128) umma_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                       // up to 256 rows x 128 B = 32 KB
  uint8_t* sb = smem + 32768;               // identity 64 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_init(bar2, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) { tc::tmem_alloc(slot, 64); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  const int N = p.kdim;  // B is kdim x kdim identity
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(bar, (uint32_t)p.rows * p.rowb + (uint32_t)p.kdim * p.rowb);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(tc::smem_u32(sa)), "l"(reinterpret_cast<uint64_t>(&p.tmA)), "r"(tc::smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(tc::smem_u32(sb)), "l"(reinterpret_cast<uint64_t>(&p.tmB)), "r"(tc::smem_u32(bar)), "r"(0), "r"(0) : "memory");
    tc::mbar_wait(bar, 0);
    tc::tc_fence_after();
    const uint32_t layout = p.rowb == 128 ? tc::LAYOUT_SW128 : tc::LAYOUT_SW64;
    const uint32_t a_addr = tc::smem_u32(sa) + (uint32_t)p.shift * p.rowb;
    uint64_t da = tc::make_smem_desc(a_addr, 16, (uint32_t)p.sbo, layout);
    if (p.base_off_mode == 1) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t db = tc::make_smem_desc(tc::smem_u32(sb), 16, 8 * p.rowb, layout);
    const uint32_t idesc = tc::make_idesc_bf16(128, (uint32_t)N, 0, 0);
    for (int k = 0; k < p.kdim / 16; ++k)
      tc::umma_bf16(tmem, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc, k > 0 ? 1u : 0u);
    tc::umma_commit(bar2);
  }
  tc::mbar_wait(bar2, 0);
  tc::tc_fence_after();
  uint32_t v[32];
  for (int c0 = 0; c0 < N; c0 += 32) {
    tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) p.out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 64); }
}

}  // namespace mcb

using namespace mcb;

// a: bf16 [rows][rowb/2]; ident: bf16 [kdim][kdim] (kdim = rowb/2); out: fp32 [128][kdim]
extern "C" int mcb_debug_umma_probe(const void* a, const void* ident, float* out, int rows, int rowb, int shift, int sbo,
                                    int base_off_mode, void* stream) {
  MCB_REQUIRE(a && ident && out, "probe: null pointer");
  MCB_REQUIRE((rowb == 128 || rowb == 64) && rows <= 256, "probe: bad shape");
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  const int kdim = rowb / 2;
  uint64_t dims[2] = {(uint64_t)kdim, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)rowb};
  uint32_t box[2] = {(uint32_t)kdim, (uint32_t)rows};
  if (int r = encode_tmap(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, dims, str, box, rowb)) return r;
  uint64_t dimsb[2] = {(uint64_t)kdim, (uint64_t)kdim};
  uint32_t boxb[2] = {(uint32_t)kdim, (uint32_t)kdim};
  if (int r = encode_tmap(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ident, dimsb, str, boxb, rowb)) return r;
  p.out = out; p.rows = rows; p.rowb = rowb; p.shift = shift; p.sbo = sbo; p.base_off_mode = base_off_mode; p.kdim = kdim;
  static bool attr = false;
  if (!attr) {
    MCB_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  umma_probe_kernel<<<1, 128, 48 * 1024, static_cast<cudaStream_t>(stream)>>>(p);
  MCB_LAUNCH_CHECK();
  return MCB_OK;
}
