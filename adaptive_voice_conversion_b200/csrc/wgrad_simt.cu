// Weight gradient of pad_layer + Conv1d (autograd of model.py:21-32 under solver.py:90):
//   dW[co][ci][j] += sum_{b, t} dc[b][co][t] * xpad[b][ci][t*stride + j]
// A [co x ci] GEMM per tap with the reduction over (sample, time).  One CTA: 128 co x
// 128 ci for one tap and one slice of the batch; partial results are accumulated into the
// (pre-zeroed) canonical nn.Conv1d-layout gradient with fp32 atomics.
#include "common.cuh"

namespace avc {

constexpr int WG_TK = 16;
constexpr int WG_LD = 132;  // 128 + 4: conflict-free float4 staging stores

struct WgradArgs {
  avc_wgrad_desc d;
  int nsl, bps;  // batch slices, samples per slice
};

__global__ void __launch_bounds__(256, 2) conv_wgrad_kernel(const WgradArgs a) {
  __shared__ __align__(16) float As[WG_TK * WG_LD];  // dc  [t][co]
  __shared__ __align__(16) float Bs[WG_TK * WG_LD];  // x   [t][ci] (tap-shifted, reflect-padded)
  const avc_wgrad_desc& d = a.d;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int ci0 = blockIdx.x * 128, co0 = blockIdx.y * 128;
  const int j = blockIdx.z / a.nsl;
  const int sl = blockIdx.z - j * a.nsl;
  const int bbeg = sl * a.bps;
  const int bend = min(d.B, bbeg + a.bps);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;

  for (int b = bbeg; b < bend; ++b) {
    for (int tc0 = 0; tc0 < d.Tout; tc0 += WG_TK) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int idx = tid + r * 256;  // 32 chunks x 16 time steps
        const int q = idx >> 4, tt = idx & 15;
        const int t = tc0 + tt;
        float4 va = zero4(), vb = zero4();
        if (t < d.Tout) {
          const int co = co0 + 4 * q;
          if (co < d.Cout) va = ldg4(d.dc + (int64_t)b * d.dc_bstride + ((int64_t)(co >> 2) * d.Tout + t) * 4);
          const int ci = ci0 + 4 * q;
          if (ci < d.Cin) {
            const int p = src_pos(t * d.stride + j - d.pad_left, d.Tin, AVC_PAD_REFLECT, 1);
            if (p >= 0) vb = ldg4(d.x + (int64_t)b * d.x_bstride + ((int64_t)(ci >> 2) * d.Tin + p) * 4);
          }
        }
        st4(As + tt * WG_LD + 4 * q, va);
        st4(Bs + tt * WG_LD + 4 * q, vb);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < WG_TK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(As + kk * WG_LD + ty * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(As + kk * WG_LD + ty * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + kk * WG_LD + tx * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(Bs + kk * WG_LD + tx * 8 + 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(av[i], bv[k], acc[i][k]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = co0 + ty * 8 + i;
    if (co >= d.Cout) continue;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ci = ci0 + tx * 8 + k;
      if (ci < d.Cin) atomicAdd(d.dw + ((int64_t)co * d.Cin + ci) * d.K + j, acc[i][k]);
    }
  }
}

}  // namespace avc

using namespace avc;

extern "C" int avc_conv_wgrad(const avc_wgrad_desc* d, void* stream) {
  AVC_REQUIRE(d && d->x && d->dc && d->dw, AVC_ERR_INVALID, "avc_conv_wgrad: null argument");
  AVC_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0 && d->K >= 1 && d->Tin > 0 && d->Tout > 0, AVC_ERR_INVALID,
              "avc_conv_wgrad: bad shape");
  AVC_REQUIRE(d->Cin % 4 == 0 && d->Cout % 4 == 0, AVC_ERR_INVALID, "avc_conv_wgrad: channels must be multiples of 4");
  WgradArgs a;
  a.d = *d;
  const int tiles = cdiv(d->Cin, 128) * cdiv(d->Cout, 128) * d->K;
  // enough CTAs for ~2 waves of 148 SMs, but keep >= 256 reduction steps per CTA so the
  // atomic epilogue stays a minor cost
  int64_t kdepth = (int64_t)d->B * d->Tout;
  int nsl = (int)cdiv64(2 * 148 * 2, tiles);
  int max_by_depth = (int)(kdepth / 256);
  if (max_by_depth < 1) max_by_depth = 1;
  if (nsl > max_by_depth) nsl = max_by_depth;
  if (nsl > d->B) nsl = d->B;
  if (nsl < 1) nsl = 1;
  a.bps = cdiv(d->B, nsl);
  a.nsl = cdiv(d->B, a.bps);
  dim3 grid(cdiv(d->Cin, 128), cdiv(d->Cout, 128), d->K * a.nsl);
  AVC_LAUNCH(conv_wgrad_kernel, grid, 256, 0, (cudaStream_t)stream, a);
  AVC_CHECK_LAUNCH("conv_wgrad");
  return AVC_OK;
}
