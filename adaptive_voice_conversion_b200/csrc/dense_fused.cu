// Fused small dense layers (launch-latency work):
//  * avc_dense_stack_fwd/bwd : the SpeakerEncoder tail (model.py:252-263, :273-276) -- n residual
//    blocks of two Linear(C,C)+ReLU and the output Linear -- as ONE kernel per direction.  A CTA
//    carries DS_R batch rows through all 2n+1 layers; a thread owns one output feature and holds its
//    whole weight row (forward) / column (backward) in registers, activations sit in shared memory.
//  * avc_linear_batch_fwd / _dx / _dw : L same-shape Linear layers in one launch each -- the 12
//    AdaIN affine layers of the decoder (model.py:342-343) and the 2n+1 weight gradients of the stack.
// fp32 FFMA throughout (these layers are ~0.1 % of the step's FLOPs).
#include "common.cuh"

namespace avc {

constexpr int DS_C = 128;  // hidden width = threads per CTA
constexpr int DS_R = 4;    // batch rows per CTA

__device__ __forceinline__ void ds_load_row(const float* __restrict__ W, int n, float4 (&w)[DS_C / 4]) {
  const float4* p = reinterpret_cast<const float4*>(W + (size_t)n * DS_C);
#pragma unroll
  for (int q = 0; q < DS_C / 4; ++q) w[q] = __ldg(p + q);
}
__device__ __forceinline__ void ds_load_col(const float* __restrict__ W, int k, float (&w)[DS_C]) {
#pragma unroll
  for (int n = 0; n < DS_C; ++n) w[n] = __ldg(W + (size_t)n * DS_C + k);
}
// acc[r] += sum_k w[k] * v[r][k]   (v in shared memory, read as broadcast float4)
__device__ __forceinline__ void ds_dot_rows(const float4 (&w)[DS_C / 4], const float (*v)[DS_C], float (&acc)[DS_R]) {
#pragma unroll
  for (int q = 0; q < DS_C / 4; ++q) {
#pragma unroll
    for (int r = 0; r < DS_R; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(&v[r][4 * q]);
      acc[r] = fmaf(w[q].x, x.x, acc[r]);
      acc[r] = fmaf(w[q].y, x.y, acc[r]);
      acc[r] = fmaf(w[q].z, x.z, acc[r]);
      acc[r] = fmaf(w[q].w, x.w, acc[r]);
    }
  }
}
__device__ __forceinline__ void ds_dot_cols(const float (&w)[DS_C], const float (*v)[DS_C], float (&acc)[DS_R]) {
#pragma unroll
  for (int q = 0; q < DS_C / 4; ++q) {
#pragma unroll
    for (int r = 0; r < DS_R; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(&v[r][4 * q]);
      acc[r] = fmaf(w[4 * q + 0], x.x, acc[r]);
      acc[r] = fmaf(w[4 * q + 1], x.y, acc[r]);
      acc[r] = fmaf(w[4 * q + 2], x.z, acc[r]);
      acc[r] = fmaf(w[4 * q + 3], x.w, acc[r]);
    }
  }
}

// params table (device): [W1_l, b1_l] l<n | [W2_l, b2_l] l<n | Wo, bo
// save planes  [B][C]: 0..n: h_0..h_n (block inputs, h_n feeds the output layer) | n+1..2n: y_l =
// relu(W1 h + b1) | 2n+1..3n: a_l = relu(W2 y + b2)
__global__ void __launch_bounds__(DS_C) dense_stack_fwd_kernel(const avc_dense_stack_desc d) {
  __shared__ __align__(16) float h[DS_R][DS_C];
  __shared__ __align__(16) float y[DS_R][DS_C];
  const int n = threadIdx.x, r0 = blockIdx.x * DS_R;
  const int nr = min(DS_R, d.B - r0), nb = d.n_blocks;
  const size_t plane = (size_t)d.B * DS_C;
#pragma unroll
  for (int r = 0; r < DS_R; ++r) h[r][n] = r < nr ? __ldg(d.x + (size_t)(r0 + r) * DS_C + n) : 0.f;
  __syncthreads();
  float4 w[DS_C / 4];
  float acc[DS_R];
  for (int l = 0; l < nb; ++l) {
    const float* W1 = d.params[2 * l];
    const float* b1 = d.params[2 * l + 1];
    const float* W2 = d.params[2 * nb + 2 * l];
    const float* b2 = d.params[2 * nb + 2 * l + 1];
    ds_load_row(W1, n, w);
    float bias = __ldg(b1 + n);
#pragma unroll
    for (int r = 0; r < DS_R; ++r) acc[r] = bias;
    ds_dot_rows(w, h, acc);
#pragma unroll
    for (int r = 0; r < DS_R; ++r) {
      acc[r] = fmaxf(acc[r], 0.f);
      if (d.save && r < nr) {
        d.save[(size_t)l * plane + (size_t)(r0 + r) * DS_C + n] = h[r][n];
        d.save[(size_t)(nb + 1 + l) * plane + (size_t)(r0 + r) * DS_C + n] = acc[r];
      }
      y[r][n] = acc[r];
    }
    __syncthreads();  // y complete; every thread is done reading h
    ds_load_row(W2, n, w);
    bias = __ldg(b2 + n);
#pragma unroll
    for (int r = 0; r < DS_R; ++r) acc[r] = bias;
    ds_dot_rows(w, y, acc);
#pragma unroll
    for (int r = 0; r < DS_R; ++r) {
      acc[r] = fmaxf(acc[r], 0.f);
      if (d.save && r < nr) d.save[(size_t)(2 * nb + 1 + l) * plane + (size_t)(r0 + r) * DS_C + n] = acc[r];
      h[r][n] += acc[r];  // residual: only this thread touches h[.][n] in this phase
    }
    __syncthreads();  // h complete; every thread is done reading y
  }
  const float* Wo = d.params[4 * nb];
  const float* bo = d.params[4 * nb + 1];
  ds_load_row(Wo, n, w);
  const float bias = __ldg(bo + n);
#pragma unroll
  for (int r = 0; r < DS_R; ++r) acc[r] = bias;
  ds_dot_rows(w, h, acc);
#pragma unroll
  for (int r = 0; r < DS_R; ++r)
    if (r < nr) {
      if (d.save) d.save[(size_t)nb * plane + (size_t)(r0 + r) * DS_C + n] = h[r][n];
      d.out[(size_t)(r0 + r) * DS_C + n] = acc[r];
    }
}

// gsave planes [B][C]: 0..n-1: g1_l (into W1_l) | n..2n-1: g2_l (into W2_l) | 2n: dout (into Wo);
// each is the upstream gradient AFTER the layer's ReLU mask = the left operand of its weight gradient
__global__ void __launch_bounds__(DS_C) dense_stack_bwd_kernel(const avc_dense_stack_desc d) {
  __shared__ __align__(16) float g[DS_R][DS_C];
  const int k = threadIdx.x, r0 = blockIdx.x * DS_R;
  const int nr = min(DS_R, d.B - r0), nb = d.n_blocks;
  const size_t plane = (size_t)d.B * DS_C;
  float wc[DS_C];
  float dh[DS_R], dy[DS_R];
#pragma unroll
  for (int r = 0; r < DS_R; ++r) {
    const float gv = r < nr ? __ldg(d.dout + (size_t)(r0 + r) * DS_C + k) : 0.f;
    g[r][k] = gv;
    if (r < nr) d.gsave[(size_t)(2 * nb) * plane + (size_t)(r0 + r) * DS_C + k] = gv;
    dh[r] = 0.f;
  }
  __syncthreads();
  ds_load_col(d.params[4 * nb], k, wc);
  ds_dot_cols(wc, g, dh);
  for (int l = nb - 1; l >= 0; --l) {
    const float* W1 = d.params[2 * l];
    const float* W2 = d.params[2 * nb + 2 * l];
    __syncthreads();  // every thread is done reading g
#pragma unroll
    for (int r = 0; r < DS_R; ++r) {
      float gv = 0.f;
      if (r < nr) {
        const float a = __ldg(d.save + (size_t)(2 * nb + 1 + l) * plane + (size_t)(r0 + r) * DS_C + k);
        gv = a > 0.f ? dh[r] : 0.f;
        d.gsave[(size_t)(nb + l) * plane + (size_t)(r0 + r) * DS_C + k] = gv;
      }
      g[r][k] = gv;
      dy[r] = 0.f;
    }
    __syncthreads();
    ds_load_col(W2, k, wc);
    ds_dot_cols(wc, g, dy);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < DS_R; ++r) {
      float gv = 0.f;
      if (r < nr) {
        const float a = __ldg(d.save + (size_t)(nb + 1 + l) * plane + (size_t)(r0 + r) * DS_C + k);
        gv = a > 0.f ? dy[r] : 0.f;
        d.gsave[(size_t)l * plane + (size_t)(r0 + r) * DS_C + k] = gv;
      }
      g[r][k] = gv;
    }
    __syncthreads();
    ds_load_col(W1, k, wc);
    ds_dot_cols(wc, g, dh);  // + identity branch: dh already holds the gradient of the block output
  }
#pragma unroll
  for (int r = 0; r < DS_R; ++r)
    if (r < nr) d.dx[(size_t)(r0 + r) * DS_C + k] = dh[r];
}

// ------------------------------------------------------------------ L same-shape linears per launch
// out_l[b][n] = W_l[n] . x_l[b] + bias_l[n];  block = 32 (n) x 8 (rows), grid.z = layer
__global__ void __launch_bounds__(256) linear_batch_fwd_kernel(const avc_linear_batch_desc d) {
  __shared__ float xs[8][33];
  __shared__ float ws[32][33];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 32, b0 = blockIdx.y * 8, l = blockIdx.z;
  const float* W = d.params[2 * l];
  const float* bias = d.params[2 * l + 1];
  const float* x = d.x + d.x_off[l];
  float acc = 0.f;
  for (int k0 = 0; k0 < d.K; k0 += 32) {
    {
      const int b = b0 + ly, k = k0 + lx;
      xs[ly][lx] = (b < d.B && k < d.K) ? __ldg(x + (int64_t)b * d.x_bstride + k) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = n0 + ly + 8 * r, k = k0 + lx;
      ws[ly + 8 * r][lx] = (n < d.N && k < d.K) ? __ldg(W + (int64_t)n * d.K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) acc = fmaf(xs[ly][kk], ws[lx][kk], acc);
    __syncthreads();
  }
  const int b = b0 + ly, n = n0 + lx;
  if (b < d.B && n < d.N) d.out[d.y_off[l] + (int64_t)b * d.y_bstride + n] = acc + (bias ? __ldg(bias + n) : 0.f);
}

// part[l][b][k] = sum_n g_l[b][n] W_l[n][k];  grid.z = layer
__global__ void __launch_bounds__(256) linear_batch_dx_kernel(const avc_linear_batch_desc d) {
  __shared__ float gs[8][33];
  __shared__ float ws[32][33];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int k0 = blockIdx.x * 32, b0 = blockIdx.y * 8, l = blockIdx.z;
  const float* W = d.params[2 * l];
  const float* g = d.y + d.y_off[l];
  float acc = 0.f;
  for (int n0 = 0; n0 < d.N; n0 += 32) {
    {
      const int b = b0 + ly, n = n0 + lx;
      gs[ly][lx] = (b < d.B && n < d.N) ? __ldg(g + (int64_t)b * d.y_bstride + n) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = n0 + ly + 8 * r, k = k0 + lx;
      ws[ly + 8 * r][lx] = (n < d.N && k < d.K) ? __ldg(W + (int64_t)n * d.K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int nn = 0; nn < 32; ++nn) acc = fmaf(gs[ly][nn], ws[nn][lx], acc);
    __syncthreads();
  }
  const int b = b0 + ly, k = k0 + lx;
  if (b < d.B && k < d.K) d.part[((int64_t)l * d.B + b) * d.K + k] = acc;
}

// out[i] = sum_l part[l][i] (+ add[i])
__global__ void __launch_bounds__(256) sum_slices_kernel(const float* __restrict__ part, int L, int64_t n, const float* __restrict__ add,
                                                         float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = add ? __ldg(add + i) : 0.f;
    for (int l = 0; l < L; ++l) s += __ldg(part + (int64_t)l * n + i);
    out[i] = s;
  }
}

// dW_l[n][k] += sum_b g_l[b][n] x_l[b][k];  db_l[n] += sum_b g_l[b][n];  grid.z = layer
__global__ void __launch_bounds__(256) linear_batch_dw_kernel(const avc_linear_batch_desc d) {
  __shared__ float gs[32][9];   // [b][n]
  __shared__ float xs[32][33];  // [b][k]
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 8, l = blockIdx.z;
  const float* g = d.y + d.y_off[l];
  const float* x = d.x + d.x_off[l];
  float* dW = d.grads[2 * l];
  float* db = d.grads[2 * l + 1];
  float acc = 0.f, bacc = 0.f;
  for (int b0 = 0; b0 < d.B; b0 += 32) {
    {
      const int b = b0 + lx, n = n0 + ly;  // lx walks the batch here
      gs[lx][ly] = (b < d.B && n < d.N) ? __ldg(g + (int64_t)b * d.y_bstride + n) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int b = b0 + ly + 8 * r, k = k0 + lx;
      xs[ly + 8 * r][lx] = (b < d.B && k < d.K) ? __ldg(x + (int64_t)b * d.x_bstride + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int bb = 0; bb < 32; ++bb) {
      acc = fmaf(gs[bb][ly], xs[bb][lx], acc);
      bacc += gs[bb][ly];
    }
    __syncthreads();
  }
  const int n = n0 + ly, k = k0 + lx;
  if (n < d.N && k < d.K) dW[(int64_t)n * d.K + k] += acc;
  if (db && blockIdx.x == 0 && lx == 0 && n < d.N) db[n] += bacc;
}

static int check_stack(const avc_dense_stack_desc* d, const char* who) {
  AVC_REQUIRE(d && d->params && d->B > 0 && d->n_blocks >= 0, AVC_ERR_INVALID, "%s: bad argument", who);
  AVC_REQUIRE(d->C == DS_C && d->c_out == DS_C, AVC_ERR_UNSUPPORTED, "%s: the fused dense stack is built for C = c_out = %d (got %d, %d)", who,
              DS_C, d->C, d->c_out);
  return AVC_OK;
}
static int check_batch(const avc_linear_batch_desc* d, const char* who) {
  AVC_REQUIRE(d && d->L > 0 && d->L <= AVC_LINEAR_BATCH_MAX && d->B > 0 && d->N > 0 && d->K > 0, AVC_ERR_INVALID, "%s: bad argument", who);
  return AVC_OK;
}

}  // namespace avc

using namespace avc;

extern "C" int avc_dense_stack_fwd(const avc_dense_stack_desc* d, void* stream) {
  int rc = check_stack(d, "avc_dense_stack_fwd");
  if (rc != AVC_OK) return rc;
  AVC_REQUIRE(d->x && d->out, AVC_ERR_INVALID, "avc_dense_stack_fwd: null x/out");
  AVC_LAUNCH(dense_stack_fwd_kernel, cdiv(d->B, DS_R), DS_C, 0, (cudaStream_t)stream, *d);
  AVC_CHECK_LAUNCH("dense_stack_fwd");
  return AVC_OK;
}
extern "C" int avc_dense_stack_bwd(const avc_dense_stack_desc* d, void* stream) {
  int rc = check_stack(d, "avc_dense_stack_bwd");
  if (rc != AVC_OK) return rc;
  AVC_REQUIRE(d->save && d->dout && d->gsave && d->dx, AVC_ERR_INVALID, "avc_dense_stack_bwd: null save/dout/gsave/dx");
  AVC_LAUNCH(dense_stack_bwd_kernel, cdiv(d->B, DS_R), DS_C, 0, (cudaStream_t)stream, *d);
  AVC_CHECK_LAUNCH("dense_stack_bwd");
  return AVC_OK;
}
extern "C" int avc_linear_batch_fwd(const avc_linear_batch_desc* d, void* stream) {
  int rc = check_batch(d, "avc_linear_batch_fwd");
  if (rc != AVC_OK) return rc;
  AVC_REQUIRE(d->params && d->x && d->out, AVC_ERR_INVALID, "avc_linear_batch_fwd: null params/x/out");
  dim3 grid(cdiv(d->N, 32), cdiv(d->B, 8), d->L);
  AVC_LAUNCH(linear_batch_fwd_kernel, grid, 256, 0, (cudaStream_t)stream, *d);
  AVC_CHECK_LAUNCH("linear_batch_fwd");
  return AVC_OK;
}
extern "C" int avc_linear_batch_dx(const avc_linear_batch_desc* d, void* stream) {
  int rc = check_batch(d, "avc_linear_batch_dx");
  if (rc != AVC_OK) return rc;
  AVC_REQUIRE(d->params && d->y && d->part && d->dx, AVC_ERR_INVALID, "avc_linear_batch_dx: null params/y/part/dx");
  dim3 grid(cdiv(d->K, 32), cdiv(d->B, 8), d->L);
  AVC_LAUNCH(linear_batch_dx_kernel, grid, 256, 0, (cudaStream_t)stream, *d);
  AVC_CHECK_LAUNCH("linear_batch_dx");
  const int64_t n = (int64_t)d->B * d->K;
  int blocks = (int)cdiv64(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  AVC_LAUNCH(sum_slices_kernel, blocks, 256, 0, (cudaStream_t)stream, (const float*)d->part, d->L, n, d->dx_add, d->dx);
  AVC_CHECK_LAUNCH("sum_slices");
  return AVC_OK;
}
extern "C" int avc_linear_batch_dw(const avc_linear_batch_desc* d, void* stream) {
  int rc = check_batch(d, "avc_linear_batch_dw");
  if (rc != AVC_OK) return rc;
  AVC_REQUIRE(d->grads && d->y && d->x, AVC_ERR_INVALID, "avc_linear_batch_dw: null grads/y/x");
  dim3 grid(cdiv(d->K, 32), cdiv(d->N, 8), d->L);
  AVC_LAUNCH(linear_batch_dw_kernel, grid, 256, 0, (cudaStream_t)stream, *d);
  AVC_CHECK_LAUNCH("linear_batch_dw");
  return AVC_OK;
}
