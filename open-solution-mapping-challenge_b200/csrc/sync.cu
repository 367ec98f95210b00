// sync.cu — one-shot all-reduce of small per-layer vectors over NVLink peer memory, for synchronised BatchNorm
// (SURVEY.md 8e collective (2), 5.8: "all-reduce [sum x, sum x^2] per BN in forward and [sum dy, sum dy xhat] in
// backward, fp32, low-latency").  A BatchNorm needs 2C floats (512 B .. 16 KB) from every rank between the conv that
// produces them and the normalisation pass that consumes them: pure latency.  NCCL spends 10-20 us per such call; here
// every rank keeps its partial sums in SYMMETRIC memory (mapped into every peer by torch's symmetric-memory
// rendezvous -- plumbing), and ONE small CTA per rank
//   1. pushes a step-stamped flag into every peer's flag table (st.release.sys over NVLink),
//   2. spins on its OWN flag table until every peer has stamped this exchange (ld.acquire.sys, local memory),
//   3. reads the peers' partial sums through their mapped pointers and adds them IN RANK ORDER (so every rank
//      computes bit-identical totals and the replicas stay identical), writing the totals to local memory.
// The partial sums of an exchange are complete when the kernel starts (stream order behind their producer), so the flag
// only has to say "my producer has finished".  Reuse across steps is safe because no rank can run more than one exchange
// ahead of its slowest peer, and the gradient all-reduce at the end of a step orders everything before the next step's
// zero-fill.  The step stamp lives in device memory (bumped by a tiny kernel at the start of every forward), so the
// launches replay unchanged inside CUDA graphs.
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

// bufs / flags: device arrays of `world` pointers (entry r = rank r's symmetric buffer as mapped in THIS process).
// flags layout per rank: [exchange][world] unsigned; rank s writes slot [exchange][s] of every peer.
__global__ void __launch_bounds__(256) sync_exchange_kernel(const float* const* __restrict__ bufs,
                                                            unsigned* const* __restrict__ flags, int rank, int world,
                                                            long offset, int count, int exchange,
                                                            const unsigned* __restrict__ step_ptr,
                                                            float* __restrict__ out, float* __restrict__ out2a,
                                                            float* __restrict__ out2b, int split, float scale2) {
  const unsigned step = *step_ptr;
  const int t = threadIdx.x;
  if (t < world && t != rank) {
    __threadfence_system();
    st_release_sys(flags[t] + (long)exchange * world + rank, step);
  }
  if (t < world && t != rank) {
    const unsigned* mine = flags[rank] + (long)exchange * world + t;
    // bounded spin (~10 s): a peer that never arrives (crashed rank, mismatched plans) must surface as a CUDA error on
    // this rank, not as a silent hang of the whole job
    long spins = 0;
    while (ld_acquire_sys(mine) != step) {
      if (++spins > (1L << 26)) __trap();
    }
  }
  __syncthreads();
  for (int c = t; c < count; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += ld_volatile_f32(bufs[r] + offset + c);   // rank order: identical on every rank
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

extern "C" int mcb_sync_exchange(const float* const* peer_bufs, unsigned* const* peer_flags, int rank, int world,
                                 long offset, int count, int exchange, const unsigned* step, float* out, float* out2_first,
                                 float* out2_second, int split, float scale2, void* stream) {
  MCB_REQUIRE(peer_bufs && peer_flags && step && out, "sync_exchange: null pointer");
  MCB_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && count > 0, "sync_exchange: bad rank / world / count");
  MCB_REQUIRE((out2_first == nullptr) == (out2_second == nullptr) && split >= 0 && split <= count, "sync_exchange: bad split");
  sync_exchange_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(peer_bufs, peer_flags, rank, world, offset, count,
                                                                        exchange, step, out, out2_first, out2_second,
                                                                        split, scale2);
  MCB_LAUNCH_CHECK();
  return MCB_OK;
}
