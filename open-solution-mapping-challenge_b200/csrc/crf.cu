// crf.cu — dense-CRF mean-field (src/postprocessing.py:183-225 -> pydensecrf DenseCRF2D) as shared-memory-tiled
// kernels.  PARITY UNPINNED: pydensecrf is absent; semantics follow oracle/post_oracle.py::dense_crf, which restates the
// published algorithm (Kraehenbuehl & Koltun 2011) with EXACT Gaussian filtering inside a (2R+1)^2 window instead of
// the library's permutohedral-lattice approximation (with sxy = 1 the tail beyond R = 6 is < 1.5e-8).
//
//   U = -log(max(p, 1e-5)); Q0 = softmax(-U)
//   K_g(i,j) = exp(-|pi-pj|^2 / (2 sxy_g^2)),  K_b(i,j) = exp(-|pi-pj|^2/(2 sxy_b^2) - |Ii-Ij|^2/(2 srgb^2))
//   n_k = 1/sqrt(K_k 1 + 1e-20)  (symmetric normalisation);  msg_k = n_k (.) K_k (n_k (.) Q)
//   Q <- softmax(-U + compat_g msg_g + compat_b msg_b), `iterations` times.
//
// One launch per iteration (every pixel needs its 13x13 neighbourhood of the previous iterate); Q ping-pongs through
// L2 (2 x 300 x 300 x 4 B per image).  Tiles of 32x32 outputs stage (Q*n_g, Q*n_b, RGB) with a 6-pixel halo in shared
// memory; each thread owns one output pixel and walks the 169 taps.  HBM/L2 + MUFU bound, no tensor cores.
#include "host_common.h"
#include "../../include/mcb200.h"
#include <algorithm>
#include <math.h>

namespace mcb {

constexpr int CRF_R = 6;
constexpr int CRF_D = 2 * CRF_R + 1;
constexpr int CRF_T = 32;
constexpr int CRF_S = CRF_T + 2 * CRF_R;  // staged tile edge

__constant__ float c_sp_g[CRF_D * CRF_D];
__constant__ float c_sp_b[CRF_D * CRF_D];

// de-normalise (x*std+mean)*255 and cast like numpy's float64 -> uint8 C cast (truncate, wrap modulo 256)
__global__ void crf_rgb_kernel(const float* __restrict__ img, uint8_t* __restrict__ rgb, long hw, long total) {
  const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = i % 3;
    const long p = (i / 3) % hw, n = i / (3 * hw);
    const double v = ((double)img[(n * 3 + c) * hw + p] * stdv[c] + mean[c]) * 255.0;
    rgb[i] = (uint8_t)(long long)v;
  }
}

// MODE 0: norms (writes n_g, n_b).  MODE 1: one mean-field iteration.
template <int MODE>
__global__ void __launch_bounds__(CRF_T* CRF_T) crf_kernel(const float* __restrict__ probs, const float* __restrict__ q_in,
                                                           const uint8_t* __restrict__ rgb, float* __restrict__ norms,
                                                           float* __restrict__ q_out, int H, int W, float inv_2srgb2,
                                                           float compat_g, float compat_b, int first_iter) {
  __shared__ float2 s_qg[CRF_S][CRF_S + 1];
  __shared__ float2 s_qb[CRF_S][CRF_S + 1];
  __shared__ uchar4 s_rgb[CRF_S][CRF_S + 1];
  const int img = blockIdx.z;
  const long hw = (long)H * W;
  const int x0 = blockIdx.x * CRF_T, y0 = blockIdx.y * CRF_T;
  const float* ng = norms + (long)img * 2 * hw;
  const float* nb = ng + hw;
  // stage the halo tile
  for (int i = threadIdx.x; i < CRF_S * CRF_S; i += blockDim.x) {
    const int sy = i / CRF_S, sx = i % CRF_S;
    const int y = y0 + sy - CRF_R, x = x0 + sx - CRF_R;
    float2 qg = make_float2(0.f, 0.f), qb = make_float2(0.f, 0.f);
    uchar4 c = make_uchar4(0, 0, 0, 0);  // .w = 1 marks an in-image pixel
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const long p = (long)y * W + x;
      const uint8_t* cp = rgb + ((long)img * hw + p) * 3;
      c = make_uchar4(cp[0], cp[1], cp[2], 1);
      if (MODE == 1) {
        float q0, q1;
        if (first_iter) {
          // Q0 = softmax(-U) = clipped probabilities renormalised
          const float p0 = fmaxf(probs[((long)img * 2) * hw + p], 1e-5f), p1 = fmaxf(probs[((long)img * 2 + 1) * hw + p], 1e-5f);
          const float u0 = -logf(p0), u1 = -logf(p1);
          const float m = fmaxf(-u0, -u1);
          const float e0 = expf(-u0 - m), e1 = expf(-u1 - m);
          q0 = e0 / (e0 + e1);
          q1 = e1 / (e0 + e1);
        } else {
          q0 = q_in[((long)img * 2) * hw + p];
          q1 = q_in[((long)img * 2 + 1) * hw + p];
        }
        const float a = ng[p], b = nb[p];
        qg = make_float2(q0 * a, q1 * a);
        qb = make_float2(q0 * b, q1 * b);
      }
    }
    s_qg[sy][sx] = qg;
    s_qb[sy][sx] = qb;
    s_rgb[sy][sx] = c;
  }
  __syncthreads();
  const int tx = threadIdx.x % CRF_T, ty = threadIdx.x / CRF_T;
  const int x = x0 + tx, y = y0 + ty;
  if (x >= W || y >= H) return;
  const uchar4 me = s_rgb[ty + CRF_R][tx + CRF_R];
  float sum_g = 0.f, sum_b = 0.f;
  float mg0 = 0.f, mg1 = 0.f, mb0 = 0.f, mb1 = 0.f;
#pragma unroll 1
  for (int dy = 0; dy < CRF_D; ++dy) {
#pragma unroll
    for (int dx = 0; dx < CRF_D; ++dx) {
      const uchar4 o = s_rgb[ty + dy][tx + dx];
      if (o.w == 0) continue;
      const float dr = (float)me.x - (float)o.x, dg = (float)me.y - (float)o.y, db = (float)me.z - (float)o.z;
      const float col = expf(-0.5f * (dr * dr + dg * dg + db * db) * (2.f * inv_2srgb2));
      const float kg = c_sp_g[dy * CRF_D + dx];
      const float kb = c_sp_b[dy * CRF_D + dx] * col;
      if (MODE == 0) {
        sum_g += kg;
        sum_b += kb;
      } else {
        const float2 qg = s_qg[ty + dy][tx + dx], qb = s_qb[ty + dy][tx + dx];
        mg0 += kg * qg.x; mg1 += kg * qg.y;
        mb0 += kb * qb.x; mb1 += kb * qb.y;
      }
    }
  }
  const long p = (long)y * W + x;
  if (MODE == 0) {
    norms[(long)img * 2 * hw + p] = 1.f / sqrtf(sum_g + 1e-20f);
    norms[(long)img * 2 * hw + hw + p] = 1.f / sqrtf(sum_b + 1e-20f);
  } else {
    const float a = ng[p], b = nb[p];
    const float p0 = fmaxf(probs[((long)img * 2) * hw + p], 1e-5f), p1 = fmaxf(probs[((long)img * 2 + 1) * hw + p], 1e-5f);
    const float e0 = logf(p0) + compat_g * (mg0 * a) + compat_b * (mb0 * b);   // -U = log p
    const float e1 = logf(p1) + compat_g * (mg1 * a) + compat_b * (mb1 * b);
    const float m = fmaxf(e0, e1);
    const float x0e = expf(e0 - m), x1e = expf(e1 - m);
    q_out[((long)img * 2) * hw + p] = x0e / (x0e + x1e);
    q_out[((long)img * 2 + 1) * hw + p] = x1e / (x0e + x1e);
  }
}

}  // namespace mcb

using namespace mcb;

extern "C" int mcb_crf_rgb_from_normalized(const float* img, uint8_t* rgb, int n, int h, int w, void* stream) {
  MCB_REQUIRE(img && rgb, "crf_rgb: null pointer");
  const long hw = (long)h * w, total = hw * 3 * n;
  const int grid = (int)std::max(1L, std::min((total + 255) / 256, (long)num_sms() * 8));
  crf_rgb_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, rgb, hw, total);
  MCB_LAUNCH_CHECK();
  return MCB_OK;
}

extern "C" int mcb_dense_crf(const float* probs, const uint8_t* rgb, float* out, float* workspace, int n, int h, int w,
                             float compat_gaussian, float sxy_gaussian, float compat_bilateral, float sxy_bilateral,
                             float srgb, int iterations, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCB_REQUIRE(probs && rgb && out && workspace, "dense_crf: null pointer");
  MCB_REQUIRE(iterations >= 1, "dense_crf: iterations %d", iterations);
  float spg[CRF_D * CRF_D], spb[CRF_D * CRF_D];
  for (int dy = -CRF_R; dy <= CRF_R; ++dy)
    for (int dx = -CRF_R; dx <= CRF_R; ++dx) {
      const double d2 = (double)(dy * dy + dx * dx);
      spg[(dy + CRF_R) * CRF_D + dx + CRF_R] = (float)exp(-0.5 * d2 / ((double)sxy_gaussian * sxy_gaussian));
      spb[(dy + CRF_R) * CRF_D + dx + CRF_R] = (float)exp(-0.5 * d2 / ((double)sxy_bilateral * sxy_bilateral));
    }
  MCB_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_sp_g, spg, sizeof(spg), 0, cudaMemcpyHostToDevice, st));
  MCB_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_sp_b, spb, sizeof(spb), 0, cudaMemcpyHostToDevice, st));
  const long plane = (long)n * 2 * h * w;
  float* norms = workspace;            // [n][2][h][w]
  float* qa = workspace + plane;       // ping
  float* qb = workspace + 2 * plane;   // pong
  dim3 grid((w + CRF_T - 1) / CRF_T, (h + CRF_T - 1) / CRF_T, n);
  const float inv_2srgb2 = 0.5f / (srgb * srgb);
  crf_kernel<0><<<grid, CRF_T * CRF_T, 0, st>>>(probs, nullptr, rgb, norms, nullptr, h, w, inv_2srgb2, compat_gaussian,
                                               compat_bilateral, 0);
  MCB_LAUNCH_CHECK();
  const float* qin = nullptr;
  for (int it = 0; it < iterations; ++it) {
    float* qout = (it == iterations - 1) ? out : ((it & 1) ? qb : qa);
    crf_kernel<1><<<grid, CRF_T * CRF_T, 0, st>>>(probs, qin, rgb, norms, qout, h, w, inv_2srgb2, compat_gaussian,
                                                 compat_bilateral, it == 0 ? 1 : 0);
    MCB_LAUNCH_CHECK();
    qin = qout;
  }
  return MCB_OK;
}
