// tcgen05 self-test: D[128][N] = sum over `nk` K=8 steps of A_k (128x8) * B_k (Nx8)^T with
// tf32 inputs / fp32 accumulation, operands given as raw shared-memory images plus the
// descriptor strides.  tests/test_gpu_tc.py uses it to pin the descriptor conventions the
// conv kernels rely on (no-swizzle core-matrix layout, row-offset tap shifts, MN-major).
#include "common.cuh"
#include "tc_common.cuh"

namespace avc {

struct ProbeArgs {
  const float* a_img;
  const float* b_img;
  int a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep, a_off, b_off, a_layout, b_layout;
  int nk, N, reps, ld_shift;
  uint32_t idesc;
  float* D;
  int* status;
};

__global__ void __launch_bounds__(128) tc_probe_kernel(const ProbeArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a.a_bytes + 1023) / 1024) * 1024;
  for (int i = tid; i < a.a_bytes / 16; i += 128) reinterpret_cast<float4*>(sa)[i] = reinterpret_cast<const float4*>(a.a_img)[i];
  for (int i = tid; i < a.b_bytes / 16; i += 128) reinterpret_cast<float4*>(sb)[i] = reinterpret_cast<const float4*>(a.b_img)[i];
  tc::fence_proxy_async_smem();
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  uint32_t ncols = 32;
  while ((int)ncols < a.N) ncols <<= 1;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = tmem_slot;
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int r = 0; r < a.reps; ++r)
      for (int k = 0; k < a.nk; ++k) {
        const uint64_t ad = tc::make_sdesc(tc::smem_u32(sa) + a.a_off + k * a.a_kstep, a.a_lbo, a.a_sbo, a.a_layout);
        const uint64_t bd = tc::make_sdesc(tc::smem_u32(sb) + a.b_off + k * a.b_kstep, a.b_lbo, a.b_sbo, a.b_layout);
        tc::mma_tf32(tbase, ad, bd, a.idesc, (k > 0 || r > 0) ? 1u : 0u);
      }
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0, a.status, 1);
  if (tid == 0) a.status[1] = (int)(clock64() - t0);  // cycles: issue of the first MMA -> all complete
  tc::tc_fence_after();
  if (ok) {
    // ld_shift != 0: the 16-column loads start at column ld_shift + 16 j (NOT a multiple of 16) -- pins that
    // tcgen05.ld accepts any start column (the conv epilogue reads samples stacked at an arbitrary pitch)
    const int sh = a.ld_shift;
    for (int c0 = 0; c0 + sh + 16 <= (int)ncols || c0 == 0; c0 += 16) {
      if (c0 + sh >= a.N) break;
      float v[16];
      tc::tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c0 + sh), v);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + sh + i < a.N) a.D[(size_t)tid * a.N + c0 + sh + i] = v[i];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, ncols);
}

// Store-path probe: every CTA writes `bytes` to its own region of `dst`; mode 0 = coalesced STG.128 from registers,
// mode 1 = 2 KB bulk (TMA) stores from shared memory, mode 2 = both at once (half the bytes each).  cycles[cta] =
// clock64 span of the CTA.  Answers "how many bytes per clock can one SM push to L2".
__global__ void __launch_bounds__(256) store_probe_kernel(float* dst, long long bytes, int mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  float4* sm4 = reinterpret_cast<float4*>(smem);
  for (int i = tid; i < 4096; i += 256) sm4[i] = make_float4((float)i, 1.f, 2.f, 3.f);   // 64 KB
  tc::fence_proxy_async_smem();
  __syncthreads();
  float* base = dst + (size_t)blockIdx.x * (bytes / 4);
  const long long n16 = bytes / 16;
  const long long t0 = clock64();
  if (mode == 0 || mode == 2) {
    const long long lim = mode == 2 ? n16 / 2 : n16;
    const float4 v = make_float4((float)tid, 1.f, 2.f, 3.f);
    for (long long i = tid; i < lim; i += 256) st4(base + i * 4, v);
  }
  if (mode == 1 || mode == 2) {
    const long long first = mode == 2 ? n16 / 2 : 0;
    const long long rows = (n16 - first) / 128;   // 2 KB rows
    if ((tid & 31) == 0) {
      for (long long r = tid >> 5; r < rows; r += 8) tc::bulk_s2g(base + (first + r * 128) * 4, smem + (size_t)(r & 31) * 2048, 2048u);
      tc::bulk_commit();
      tc::bulk_wait_all();
    }
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

}  // namespace avc

using namespace avc;

extern "C" int avc_probe_store(float* dst, long long bytes_per_cta, int ctas, int mode, long long* cycles, void* stream) {
  AVC_REQUIRE(dst && cycles && bytes_per_cta > 0 && bytes_per_cta % 4096 == 0 && ctas > 0 && mode >= 0 && mode <= 2, AVC_ERR_INVALID,
              "avc_probe_store: bad argument");
  cudaError_t e = cudaFuncSetAttribute(store_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  if (e != cudaSuccess) { set_error("avc_probe_store: %s", cudaGetErrorString(e)); return AVC_ERR_CUDA; }
  store_probe_kernel<<<ctas, 256, 64 * 1024, (cudaStream_t)stream>>>(dst, bytes_per_cta, mode, cycles);
  AVC_CHECK_LAUNCH("store_probe");
  return AVC_OK;
}

static int g_probe_ld_shift = 0;
extern "C" void avc_tc_probe_set_ld_shift(int shift) { g_probe_ld_shift = shift < 0 ? 0 : shift; }

extern "C" int avc_tc_probe_gemm(const float* a_img, int a_bytes, const float* b_img, int b_bytes, const uint32_t* strides /*[10]*/,
                                 int nk, int N, int a_mn, int b_mn, int reps, float* D, int* status, void* stream) {
  AVC_REQUIRE(a_img && b_img && strides && D && status, AVC_ERR_INVALID, "avc_tc_probe_gemm: null argument");
  AVC_REQUIRE(a_bytes % 16 == 0 && b_bytes % 16 == 0 && N % 16 == 0 && N >= 16 && N <= 256 && nk >= 1, AVC_ERR_INVALID,
              "avc_tc_probe_gemm: bad sizes");
  ProbeArgs a;
  a.a_img = a_img; a.b_img = b_img; a.a_bytes = a_bytes; a.b_bytes = b_bytes;
  a.a_lbo = strides[0]; a.a_sbo = strides[1]; a.b_lbo = strides[2]; a.b_sbo = strides[3];
  a.a_kstep = strides[4]; a.b_kstep = strides[5]; a.a_off = strides[6]; a.b_off = strides[7];
  a.a_layout = strides[8]; a.b_layout = strides[9];
  a.nk = nk; a.N = N; a.reps = reps < 1 ? 1 : reps; a.ld_shift = g_probe_ld_shift; a.idesc = tc::make_idesc_tf32(128, N, a_mn, b_mn); a.D = D; a.status = status;
  const int smem = ((a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
  AVC_REQUIRE(smem <= 200 * 1024, AVC_ERR_INVALID, "avc_tc_probe_gemm: images too large");
  cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("avc_tc_probe_gemm: %s", cudaGetErrorString(e)); return AVC_ERR_CUDA; }
  tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a);
  AVC_CHECK_LAUNCH("tc_probe");
  return AVC_OK;
}
