// sync.cu — one-shot all-reduce of small per-layer vectors over NVLink peer memory, for synchronised BatchNorm
// (SURVEY.md 8e collective (2), 5.8: "all-reduce [sum x, sum x^2] per BN in forward and [sum dy, sum dy xhat] in
// backward, fp32, low-latency").  A BatchNorm needs 2C floats (512 B .. 16 KB) from every rank between the conv that
// produces them and the normalisation pass that consumes them: pure latency.  NCCL spends 10-20 us per such call.  Here
// every rank owns a RECEIVE buffer [world][sum 2C] in symmetric memory (mapped into every peer by torch's
// symmetric-memory rendezvous -- plumbing), and one small CTA per rank
//   1. PUSHES its partial sums into slot [my rank] of every peer's receive buffer (posted NVLink stores, one-way
//      latency; no request / response round trip),
//   2. fences, then stamps a step-numbered flag in every peer's flag table (st.release.sys),
//   3. spins on its OWN flag table until every peer has stamped this exchange (ld.acquire.sys, local memory),
//   4. adds the world contributions IN RANK ORDER from local memory (its own partial sums + the received slots), so
//      every rank computes bit-identical totals and the replicas stay identical.
// The partial sums are complete when the kernel starts (stream order behind their producer).  Reuse across steps is
// safe because a peer can push the next step's values of an exchange only after the gradient all-reduce of this step,
// which every rank joins after its last exchange.  The step stamp lives in device memory (bumped by a tiny kernel at
// the start of every forward), so the launches replay unchanged inside CUDA graphs.
#include "host_common.h"
#include "../../include/mcb200.h"

namespace mcb {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_volatile_f32(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__global__ void sync_step_bump_kernel(unsigned* step) {
  if (threadIdx.x == 0) *step = *step + 1u;
}

// recv / flags: device arrays of `world` pointers (entry r = rank r's symmetric buffer as mapped in THIS process).
// recv layout per rank: [world][stride] floats (slot s = contribution of rank s); flags per rank: [exchange][world]
// unsigned (rank s stamps slot [exchange][s] of every peer).
constexpr int kSyncThreads = 1024;
__global__ void __launch_bounds__(kSyncThreads) sync_exchange_kernel(const float* __restrict__ partial,
                                                                     float* const* __restrict__ recv,
                                                                     unsigned* const* __restrict__ flags, int rank,
                                                                     int world, long stride, long offset, int count,
                                                                     int exchange, const unsigned* __restrict__ step_ptr,
                                                                     float* __restrict__ out, float* __restrict__ out2a,
                                                                     float* __restrict__ out2b, int split, float scale2) {
  const unsigned step = *step_ptr;
  const int t = threadIdx.x;
  // 1. push (count is a multiple of 4 and every pointer 16-byte aligned: float4 stores)
  const int n4 = count >> 2;
  const float4* src4 = reinterpret_cast<const float4*>(partial + offset);
  for (int i = t; i < n4; i += kSyncThreads) {
    const float4 v = src4[i];
    for (int p = 0; p < world; ++p) {
      if (p == rank) continue;
      reinterpret_cast<float4*>(recv[p] + (long)rank * stride + offset)[i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  // 2. stamp, 3. wait
  if (t < world && t != rank) {
    st_release_sys(flags[t] + (long)exchange * world + rank, step);
    const unsigned* mine = flags[rank] + (long)exchange * world + t;
    // bounded spin (~10 s): a peer that never arrives (crashed rank, mismatched plans) must surface as a CUDA error on
    // this rank, not as a silent hang of the whole job
    long spins = 0;
    while (ld_acquire_sys(mine) != step) {
      if (++spins > (1L << 26)) __trap();
    }
  }
  __syncthreads();
  // 4. reduce in rank order from local memory
  const float* mine_recv = recv[rank];
  for (int c = t; c < count; c += kSyncThreads) {
    float s = 0.f;
    for (int r = 0; r < world; ++r)
      s += (r == rank) ? partial[offset + c] : ld_volatile_f32(mine_recv + (long)r * stride + offset + c);
    out[c] = s;
    if (out2a != nullptr) {
      if (c < split) out2a[c] = s * scale2;
      else out2b[c - split] = s * scale2;
    }
  }
}

}  // namespace mcb

using namespace mcb;

extern "C" int mcb_sync_step_bump(unsigned* step, void* stream) {
  MCB_REQUIRE(step, "sync_step_bump: null pointer");
  sync_step_bump_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(step);
  MCB_LAUNCH_CHECK();
  return MCB_OK;
}

extern "C" int mcb_sync_exchange(const float* partial, float* const* peer_recv, unsigned* const* peer_flags, int rank,
                                 int world, long stride, long offset, int count, int exchange, const unsigned* step,
                                 float* out, float* out2_first, float* out2_second, int split, float scale2, void* stream) {
  MCB_REQUIRE(partial && peer_recv && peer_flags && step && out, "sync_exchange: null pointer");
  MCB_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && count > 0, "sync_exchange: bad rank / world / count");
  MCB_REQUIRE(count % 4 == 0 && offset % 4 == 0 && stride % 4 == 0, "sync_exchange: count / offset / stride must be multiples of 4");
  MCB_REQUIRE((out2_first == nullptr) == (out2_second == nullptr) && split >= 0 && split <= count, "sync_exchange: bad split");
  sync_exchange_kernel<<<1, kSyncThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      partial, peer_recv, peer_flags, rank, world, stride, offset, count, exchange, step, out, out2_first, out2_second, split,
      scale2);
  MCB_LAUNCH_CHECK();
  return MCB_OK;
}
