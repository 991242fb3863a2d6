// Conv weight gradient on the tcgen05 tensor cores (TF32 in, fp32 accumulate in TMEM):
//   dW[co][ci][j] += sum_{b,t} dc[b][co][t] * xpad[b][ci][t + j]        (stride 1)
// (autograd of pad_layer + nn.Conv1d w.r.t. the weight, model.py:21-32 under solver.py:90).
//
// GEMM view: M = co (128 per CTA), N = ci (64 per CTA), reduction K = time rows of a batch
// slice.  Both operands are MN-major: the A4 activation layout [c/4][t][4] already keeps 4
// channels of one time step in a 16-byte unit, so staging is a plain LDG.128 -> STS.128 of
// units into the tensor core's MN-major TF32 layout (UMMA layout code 1, "SW128_32B"):
//     byte(c, row) = (c/32)*LBO + row*128 + ((((c%32)/8) ^ (row%4)) * 32) + (c%8)*4
// measured on B200 with tools/diag_mn4.py / diag_mn5.py (the XOR is keyed on the absolute
// shared-memory row, so a descriptor start shifted by whole rows addresses shifted rows).
// The K taps are therefore K descriptor starts (row shifts) into ONE staged input tile whose
// reflect padding is resolved while staging; tap j accumulates into TMEM columns [j*N, (j+1)*N).
// Each CTA owns (co tile, ci tile, batch slice); partial sums go to a scratch buffer
// [slice][tap][ci/4][co][4] with coalesced 16-byte stores and are reduced into the canonical
// nn.Conv1d gradient layout by wgrad_tc_reduce_kernel (deterministic, no atomics).
#include "common.cuh"
#include "tc_common.cuh"

namespace avc {

constexpr int WT_NT = 64;  // ci columns per CTA

struct WgTcArgs {
  avc_wgrad_desc d;
  float* scratch;
  long long* dbg;   // optional per-CTA phase cycle counters (avc_wgrad_tc_set_debug)
  int nslices, tiles_per_slice, G, RA, RX, ntpad, ncols_tmem, coutp, TX, H;  // H: rows of one parity block (stride 2)
  uint32_t buf_bytes, x_off;
  int* status;
};

__device__ __forceinline__ float rtf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 rtf32_4(float4 v) { return make_float4(rtf32(v.x), rtf32(v.y), rtf32(v.z), rtf32(v.w)); }

// 16-byte async global->shared copy; !valid writes zeros (src-size 0, nothing is read)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}

// byte offset of the 16-byte unit (channel chunk q = c/4, row) inside an operand buffer
__device__ __forceinline__ uint32_t mn_unit_off(int q, int row, uint32_t atom_bytes) {
  return (uint32_t)(q >> 3) * atom_bytes + (uint32_t)row * 128u + (uint32_t)((((q & 7) >> 1) ^ (row & 3)) << 5) + (uint32_t)((q & 1) << 4);
}

// UI (uniform issue): see conv_tc.cu; UI = false is the round-1 issue loop.
// ATOMIC: the epilogue adds the CTA's partial straight into ONE accumulation buffer per layer
// ([tap][ci/4][co][4], vector red.global.add.v4.f32) instead of writing a per-slice partial: no
// scratch round trip and no per-layer reduction launch (avc_wgrad_acc_flush unpacks every layer at
// the end of the backward pass); the price is a run-to-run summation order.
template <bool UI, bool ATOMIC>
__global__ void __launch_bounds__(128, 1) conv_wgrad_tc_kernel(const WgTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_free[2], bar_done;
  __shared__ uint32_t tmem_slot;
  const avc_wgrad_desc& d = a.d;
  const int tid = threadIdx.x, warp = UI ? tc::warp_idx_sync() : (tid >> 5);
  const int ci0 = blockIdx.x * WT_NT, co0 = blockIdx.y * 128, sl = blockIdx.z;
  const int K = d.K, T = d.Tout, TX = a.TX;  // padded input positions per sample
  const int S = d.stride, H = a.H;
  const int tile0 = sl * a.tiles_per_slice;
  const int ntiles_all = cdiv(d.B, a.G);
  const int tile1 = min(ntiles_all, tile0 + a.tiles_per_slice);

  if (tid == 0) {
    tc::mbar_init(&bar_free[0], 1);
    tc::mbar_init(&bar_free[1], 1);
    tc::mbar_init(&bar_done, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, (uint32_t)a.ncols_tmem);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_slot, 0);
  const uint32_t atomA = (uint32_t)a.RA * 128u, atomX = (uint32_t)a.RX * 128u;
  const uint32_t idesc = tc::make_idesc_tf32(128, a.ntpad, 1, 1);
  const int nq_x = a.ntpad >> 2;  // 16-byte units per row of the x operand
  bool ok = true;
  const uint32_t smem_base = tc::smem_u32(smem);
  long long t_begin = 0, c_stage = 0, c_free = 0, c_issue = 0;
  if (a.dbg) t_begin = clock64();

  for (int tile = tile0; tile < tile1; ++tile) {
    const int it = tile - tile0;
    const int buf = it & 1;
    const long long q0 = a.dbg ? clock64() : 0;
    if (it >= 2) ok = tc::mbar_wait(&bar_free[buf], (uint32_t)((it >> 1) - 1) & 1u, a.status, 6) && ok;
    const long long q1 = a.dbg ? clock64() : 0;
    uint8_t* sA = smem + (size_t)buf * a.buf_bytes;
    uint8_t* sX = sA + a.x_off;
    const int b0 = tile * a.G;
    const int nsamp = min(a.G, d.B - b0);
    // ---- stage dc: [4 atoms of 32 co][G*T rows][128 B]; cp.async keeps ~50 16-byte copies per
    // thread in flight (the tensor core truncates fp32 -> tf32; a uniform ~1e-3 shrink of dW).
    // Row -> (sample, time) is decomposed once per row and reused for all channel chunks.
    for (int r = tid; r < nsamp * T; r += 128) {
      const int g = r / T, t = r - g * T;
      const float* src = d.dc + (size_t)(b0 + g) * d.dc_bstride + (size_t)t * 4;
#pragma unroll 8
      for (int q = 0; q < 32; ++q) {
        const int co = co0 + 4 * q;
        const bool cv = co < d.Cout;
        cp_async16(sA + mn_unit_off(q, r, atomA), src + (size_t)((cv ? co : 0) >> 2) * T * 4, cv);
      }
    }
    // ---- stage x with the reflect padding resolved: [ntpad/32 atoms][G*(T+K-1) rows][128 B]
    for (int r0 = tid; r0 < nsamp * TX; r0 += 128) {
      const int g = r0 / TX, u = r0 - g * TX;
      // stride 2: even and odd padded positions go to separate row blocks, so tap j addresses
      // rows (j&1)*H + t + (j>>1) -- contiguous in t
      const int r = S == 1 ? r0 : g * 2 * H + (u & 1) * H + (u >> 1);
      const int p = src_pos(u - d.pad_left, d.Tin, AVC_PAD_REFLECT, 1);
      const float* src = d.x + (size_t)(b0 + g) * d.x_bstride + (size_t)(p >= 0 ? p : 0) * 4;
#pragma unroll 8
      for (int q = 0; q < nq_x; ++q) {
        const int ci = ci0 + 4 * q;
        const bool v = ci < d.Cin && p >= 0;
        cp_async16(sX + mn_unit_off(q, r, atomX), src + (size_t)((v ? ci : 0) >> 2) * d.Tin * 4, v);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    tc::fence_proxy_async_smem();
    ok = __syncthreads_and(ok) != 0;  // also makes `ok` block-uniform for the issue loop below
    const long long q2 = a.dbg ? clock64() : 0;
    c_free += q1 - q0;
    c_stage += q2 - q1;
    if (warp == 0) {  // warp-converged issue loop, one elected lane per instruction (uniform descriptors)
      if (__all_sync(0xffffffffu, ok)) {
        if constexpr (UI) {
          tc::tc_fence_after();
          // everything the MMAs consume is re-derived from lane-0 broadcasts: ptxas then proves the
          // values warp-uniform and keeps the whole issue loop on the uniform datapath
          const int it_u = __shfl_sync(0xffffffffu, it, 0), nsamp_u = __shfl_sync(0xffffffffu, nsamp, 0);
          const uint32_t tb_u = __shfl_sync(0xffffffffu, tbase, 0);
          const uint32_t sbase = __shfl_sync(0xffffffffu, smem_base, 0) + (uint32_t)(it_u & 1) * a.buf_bytes;
          const uint32_t a_lo0 = tc::sdesc_lo(sbase, atomA), b_lo0 = tc::sdesc_lo(sbase + a.x_off, atomX);
          const uint32_t hi = tc::sdesc_hi(512, 1);
          const int nks = T >> 3;
          for (int g = 0; g < nsamp_u; ++g) {
            // loop-carried 64-bit descriptors: one k-step = 8 rows = 64 descriptor units on both operands
            uint64_t a_desc = tc::sdesc64(a_lo0 + (uint32_t)(g * T) * 8u, hi);
            uint64_t b_ks = tc::sdesc64(b_lo0 + (uint32_t)(S == 1 ? g * TX : g * 2 * H) * 8u, hi);
            for (int ks = 0; ks < nks; ++ks) {
              const uint32_t acc = (it_u | g | ks) ? 1u : 0u;
              uint32_t dcol = tb_u;
              if (S == 1) {
                uint64_t b_desc = b_ks;
                if (K == 5) {
  #pragma unroll
                  for (int j = 0; j < 5; ++j) {
                    tc::mma_tf32_elect(dcol, a_desc, b_desc, idesc, acc);
                    b_desc += 8u;
                    dcol += (uint32_t)a.ntpad;
                  }
                } else {
                  for (int j = 0; j < K; ++j) {
                    tc::mma_tf32_elect(dcol, a_desc, b_desc, idesc, acc);
                    b_desc += 8u;
                    dcol += (uint32_t)a.ntpad;
                  }
                }
              } else {
                for (int j = 0; j < K; ++j) {
                  tc::mma_tf32_elect(dcol, a_desc, b_ks + (uint32_t)((j & 1) * H + (j >> 1)) * 8u, idesc, acc);
                  dcol += (uint32_t)a.ntpad;
                }
              }
              a_desc += 64u;
              b_ks += 64u;
            }
          }
        } else {
          tc::tc_fence_after();
          const uint32_t sbase = smem_base + (uint32_t)buf * a.buf_bytes;
          const uint32_t a_lo0 = tc::sdesc_lo(sbase, atomA), b_lo0 = tc::sdesc_lo(sbase + a.x_off, atomX);
          const uint32_t hi = tc::sdesc_hi(512, 1);
          for (int g = 0; g < nsamp; ++g)
            for (int ks = 0; ks < T / 8; ++ks) {
              const uint32_t a_lo = a_lo0 + (uint32_t)(g * T + 8 * ks) * 8u;         // rows * 128 B >> 4
              const uint32_t b_base = b_lo0 + (uint32_t)((S == 1 ? g * TX : g * 2 * H) + 8 * ks) * 8u;
              uint32_t dcol = tbase;
              const uint32_t acc = (it == 0 && g == 0 && ks == 0) ? 0u : 1u;
              if (S == 1) {
                uint32_t b_lo = b_base;
                for (int j = 0; j < K; ++j) {
                  if (tc::elect_one()) tc::mma_tf32_lohi(dcol, a_lo, hi, b_lo, hi, idesc, acc);
                  b_lo += 8u;
                  dcol += (uint32_t)a.ntpad;
                }
              } else {
                for (int j = 0; j < K; ++j) {
                  const uint32_t b_lo = b_base + (uint32_t)((j & 1) * H + (j >> 1)) * 8u;
                  if (tc::elect_one()) tc::mma_tf32_lohi(dcol, a_lo, hi, b_lo, hi, idesc, acc);
                  dcol += (uint32_t)a.ntpad;
                }
              }
            }
        }
        __syncwarp();
        if (tc::elect_one()) tc::mma_commit(&bar_free[buf]);
      }
      if (a.dbg) c_issue += clock64() - q2;
    }
  }
  if (warp == 0) {
    __syncwarp();
    if (tc::elect_one()) tc::mma_commit(&bar_done);
  }
  const long long q3 = a.dbg ? clock64() : 0;
  ok = tc::mbar_wait(&bar_done, 0, a.status, 7) && ok;
  ok = __syncthreads_and(ok) != 0;
  tc::tc_fence_after();
  const long long q4 = a.dbg ? clock64() : 0;
  if (ok && tile1 > tile0) {
    // partial dW of this slice: scratch[sl][tap][ci/4][co][4]
    const int co = co0 + tid;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int j = 0; j < K; ++j) {
      float* sbase = a.scratch + (((size_t)(ATOMIC ? 0 : sl) * K + j) * (size_t)(d.Cin >> 2)) * (size_t)a.coutp * 4;
      for (int c0 = 0; c0 < a.ntpad; c0 += 16) {
        float v[16];
        tc::tmem_ld16(lane_addr + (uint32_t)(j * a.ntpad + c0), v);
#pragma unroll
        for (int i4 = 0; i4 < 16; i4 += 4) {
          const int ci = ci0 + c0 + i4;
          if (ci < d.Cin) {
            float* p = sbase + ((size_t)(ci >> 2) * a.coutp + co) * 4;
            if constexpr (ATOMIC)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[i4]), "f"(v[i4 + 1]), "f"(v[i4 + 2]), "f"(v[i4 + 3]) : "memory");
            else
              st4(p, make_float4(v[i4], v[i4 + 1], v[i4 + 2], v[i4 + 3]));
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (a.dbg && tid == 0) {
    long long* o = a.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
    const long long q5 = clock64();
    o[0] = t_begin; o[1] = q5; o[2] = c_stage; o[3] = c_free; o[4] = c_issue; o[5] = q4 - q3; o[6] = q5 - q4; o[7] = tile1 - tile0;
  }
  if (warp == 0) tc::tmem_dealloc(tbase, (uint32_t)a.ncols_tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// Round-2 variant ("wgrad_split"): 8 staging / epilogue warps + a dedicated MMA warp, so that staging tile i+1 runs
// under the MMAs of tile i (the kernel above issues from staging warp 0, whose blocking issue loop serialises the two:
// 17.3 K staging + 16.6 K issue of a 43 K-cycle CTA, tools/diag_wgrad.py).  Pipelines: full[2] (256 staging arrivals),
// free[2] (tcgen05.commit), done.  (Staging the operands with swizzled tensor-map TMA copies instead was tried and
// dropped: CU_TENSOR_MAP_SWIZZLE_128B* needs 128-byte inner rows, the A4 layout has 16-byte ones.)
constexpr int WS_STAGE_THREADS = 256;
constexpr int WS_THREADS = WS_STAGE_THREADS + 32;

template <bool ATOMIC>
__global__ void __launch_bounds__(WS_THREADS, 1) conv_wgrad_split_kernel(const WgTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[2], bar_free[2], bar_done;
  __shared__ uint32_t tmem_slot;
  const avc_wgrad_desc& d = a.d;
  const int tid = threadIdx.x, warp = tc::warp_idx_sync(), lane = tid & 31;
  const int ci0 = blockIdx.x * WT_NT, co0 = blockIdx.y * 128, sl = blockIdx.z;
  const int K = d.K, T = d.Tout, TX = a.TX;
  const int S = d.stride, H = a.H;
  const int tile0 = sl * a.tiles_per_slice;
  const int tile1 = min(cdiv(d.B, a.G), tile0 + a.tiles_per_slice);
  if (tid == 0) {
    tc::mbar_init(&bar_full[0], WS_STAGE_THREADS);
    tc::mbar_init(&bar_full[1], WS_STAGE_THREADS);
    tc::mbar_init(&bar_free[0], 1);
    tc::mbar_init(&bar_free[1], 1);
    tc::mbar_init(&bar_done, 1);
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, (uint32_t)a.ncols_tmem);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t atomA = (uint32_t)a.RA * 128u, atomX = (uint32_t)a.RX * 128u;
  const int nq_x = a.ntpad >> 2;
  long long t_begin = 0, c_stage = 0, c_free = 0, c_issue = 0, q3 = 0, q4 = 0;
  if (a.dbg) t_begin = clock64();
  bool ok = true;

  if (warp < 8) {
    // ============================================================ staging warps: one tile of copies at a time; the
    // tile is handed to the MMA warp as soon as its copies have landed, BEFORE waiting for the next free buffer
    for (int tile = tile0; tile < tile1 && ok; ++tile) {
      const int it = tile - tile0;
      const int buf = it & 1;
      const long long q0 = a.dbg ? clock64() : 0;
      if (it >= 2) ok = tc::mbar_wait(&bar_free[buf], (uint32_t)((it >> 1) - 1) & 1u, a.status, 6);
      const long long q1 = a.dbg ? clock64() : 0;
      if (!ok) break;
      uint8_t* sA = smem + (size_t)buf * a.buf_bytes;
      uint8_t* sX = sA + a.x_off;
      const int b0 = tile * a.G;
      const int nsamp = min(a.G, d.B - b0);
      // Unit -> lane mapping: a warp writes 4 rows x the 8 units of one 128-byte atom row = four whole bank sweeps.
      // (One row per lane, as in the round-1 kernel, puts 32 lanes on 4 bank groups: an 8-way conflict that made the
      // shared-memory side of every cp.async 8x slower -- staging was 6.5 K cycles per tile against 1.5 K of issue.)
      // Each thread walks rows r_lo, r_lo + 16 (dc: + 8), ... with (g, t) kept incrementally: no division per unit.
      // Unit -> lane mapping: a warp writes 4 rows x the 8 units of one 128-byte atom row = four whole bank sweeps, and
      // each thread walks rows r_lo, r_lo + 8 (x: + 16), ... with (g, t) kept incrementally: no division per unit.
      // (Measured alternatives, tools/diag_wgrad.py: one row per lane as in round 1 -- same time; LDG.128 -> registers
      // -> STS.128 with eight loads in flight -- 40 % slower.  The ~6.5 K cycles per 97 KB tile are not the bank mapping.)
      {
        // dc: 4 atoms x 8 units per row; 256 threads = 8 warps = (2 row groups of 4 rows) x 4 atoms
        const int q = ((tid >> 5) & 3) * 8 + (tid & 7), r_lo = ((tid >> 7) << 2) + ((tid >> 3) & 3);
        const int co = co0 + 4 * q;
        const bool cv = co < d.Cout;
        const float* colsrc = d.dc + (size_t)((cv ? co : 0) >> 2) * T * 4;
        int g = r_lo / T, t = r_lo - g * T;
        for (int r = r_lo; r < nsamp * T; r += 8) {
          cp_async16(sA + mn_unit_off(q, r, atomA), colsrc + (size_t)(b0 + g) * d.dc_bstride + (size_t)t * 4, cv);
          t += 8;
          while (t >= T) { t -= T; ++g; }
        }
      }
      {
        // x: ntpad/32 atoms x 8 units per row; a warp = 4 rows of one atom, the 8 warps cover natom atoms x (8/natom) row groups
        const int natom = nq_x >> 3, wpa = 8 / natom;             // warps per atom
        const int w = tid >> 5, atom = w % natom, rg = w / natom;  // row group of this warp
        const int q = atom * 8 + (tid & 7), r_lo = (rg << 2) + ((tid >> 3) & 3), rstep = wpa << 2;
        const int ci = ci0 + 4 * q;
        const bool civ = ci < d.Cin;
        const float* colsrc = d.x + (size_t)((civ ? ci : 0) >> 2) * d.Tin * 4;
        int g = r_lo / TX, u = r_lo - g * TX;
        for (int r0 = r_lo; r0 < nsamp * TX; r0 += rstep) {
          const int r = S == 1 ? r0 : g * 2 * H + (u & 1) * H + (u >> 1);
          const int p = src_pos(u - d.pad_left, d.Tin, AVC_PAD_REFLECT, 1);
          const bool v = civ && p >= 0;
          cp_async16(sX + mn_unit_off(q, r, atomX), colsrc + (size_t)(b0 + g) * d.x_bstride + (size_t)(p >= 0 ? p : 0) * 4, v);
          u += rstep;
          while (u >= TX) { u -= TX; ++g; }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&bar_full[buf]);
      if (a.dbg) {
        const long long q2 = clock64();
        c_free += q1 - q0;
        c_stage += q2 - q1;
      }
    }
  } else {
    // ============================================================ MMA issuer (warp 8), uniform datapath.
    // ALL TAPS IN ONE MMA: for a 32-channel block of x, tap j is the same swizzled atom one row (128 B) further, so
    // a B descriptor with LBO = 128 B addresses the K taps as K consecutive 32-column atoms: one MMA of N = 32 K
    // columns per k-step and channel block (stride 2: the even and the odd taps form two such groups) instead of K
    // MMAs of N = 64 -- the dc operand is read once per k-step instead of K times (the MMAs here are bound by
    // shared-memory operand traffic), 2 instead of 5 instructions per k-step.
    const int NT = K * 32;                 // TMEM columns of one 32-channel block
    const int natom = a.ntpad >> 5;
    const int ne = (K + 1) >> 1, no = K >> 1;
    const uint32_t idesc_all = tc::make_idesc_tf32(128, NT, 1, 1);
    const uint32_t idesc_e = tc::make_idesc_tf32(128, 32 * ne, 1, 1), idesc_o = tc::make_idesc_tf32(128, no ? 32 * no : 32, 1, 1);
    const uint32_t tb_u = __shfl_sync(0xffffffffu, tbase, 0);
    const uint32_t smem_base = __shfl_sync(0xffffffffu, tc::smem_u32(smem), 0);
    const uint32_t hi = tc::sdesc_hi(512, 1);
    const int nks = T >> 3;
    for (int tile = tile0; tile < tile1 && ok; ++tile) {
      const int it_u = tile - tile0;
      const int nsamp_u = min(a.G, d.B - tile * a.G);
      ok = __all_sync(0xffffffffu, tc::mbar_wait(&bar_full[it_u & 1], (uint32_t)(it_u >> 1) & 1u, a.status, 8));
      if (!ok) break;
      tc::tc_fence_after();
      const long long q2 = a.dbg ? clock64() : 0;
      const uint32_t sbase = smem_base + (uint32_t)(it_u & 1) * a.buf_bytes;
      const uint32_t a_lo0 = tc::sdesc_lo(sbase, atomA);
      for (int g = 0; g < nsamp_u; ++g) {
        uint64_t a_desc = tc::sdesc64(a_lo0 + (uint32_t)(g * T) * 8u, hi);
        const uint32_t xrow0 = (uint32_t)(S == 1 ? g * TX : g * 2 * H) * 8u;   // in 16-byte descriptor units
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t acc = (it_u | g | ks) ? 1u : 0u;
          for (int cb = 0; cb < natom; ++cb) {
            const uint32_t b_lo = tc::sdesc_lo(sbase + a.x_off + (uint32_t)cb * atomX, 128u) + xrow0 + (uint32_t)ks * 64u;
            const uint32_t dcol = tb_u + (uint32_t)(cb * NT);
            if (S == 1) {
              tc::mma_tf32_elect(dcol, a_desc, tc::sdesc64(b_lo, hi), idesc_all, acc);
            } else {
              tc::mma_tf32_elect(dcol, a_desc, tc::sdesc64(b_lo, hi), idesc_e, acc);
              if (no) tc::mma_tf32_elect(dcol + (uint32_t)(32 * ne), a_desc, tc::sdesc64(b_lo + (uint32_t)H * 8u, hi), idesc_o, acc);
            }
          }
          a_desc += 64u;
        }
      }
      __syncwarp();
      if (tc::elect_one()) tc::mma_commit(&bar_free[it_u & 1]);
      __syncwarp();
      if (a.dbg) c_issue += clock64() - q2;
    }
    __syncwarp();
    if (tc::elect_one()) tc::mma_commit(&bar_done);
  }
  if (a.dbg) q3 = clock64();
  ok = tc::mbar_wait(&bar_done, 0, a.status, 7) && ok;
  ok = __syncthreads_and(ok) != 0;
  tc::tc_fence_after();
  if (a.dbg) q4 = clock64();
  if (ok && tile1 > tile0 && warp < 8) {
    // TMEM column of (32-channel block cb, tap j, channel c) = cb * 32 K + slot(j) * 32 + c, slot(j) = j (stride 1) or the
    // even taps first (stride 2); 8 warps: lane quarter = warp & 3, the 16-column chunks are split between its two warps
    const int quarter = warp & 3, half = warp >> 2;
    const int co = co0 + quarter * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(quarter * 32) << 16);
    const int nch = a.ntpad >> 4;
    const int NT = K * 32, ne = (K + 1) >> 1;
    for (int j = 0; j < K; ++j) {
      const int slot = S == 1 ? j : ((j & 1) ? ne + (j >> 1) : (j >> 1));
      float* sbase = a.scratch + (((size_t)(ATOMIC ? 0 : sl) * K + j) * (size_t)(d.Cin >> 2)) * (size_t)a.coutp * 4;
      for (int ch = half; ch < nch; ch += 2) {
        const int c0 = ch << 4;
        float v[16];
        tc::tmem_ld16(lane_addr + (uint32_t)((c0 >> 5) * NT + slot * 32 + (c0 & 31)), v);
#pragma unroll
        for (int i4 = 0; i4 < 16; i4 += 4) {
          const int ci = ci0 + c0 + i4;
          if (ci < d.Cin) {
            float* p = sbase + ((size_t)(ci >> 2) * a.coutp + co) * 4;
            if constexpr (ATOMIC)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[i4]), "f"(v[i4 + 1]), "f"(v[i4 + 2]), "f"(v[i4 + 3]) : "memory");
            else
              st4(p, make_float4(v[i4], v[i4 + 1], v[i4 + 2], v[i4 + 3]));
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (a.dbg && (tid == 0 || tid == WS_STAGE_THREADS)) {
    long long* o = a.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
    const long long q5 = clock64();
    if (tid == 0) { o[0] = t_begin; o[1] = q5; o[2] = c_stage; o[3] = c_free; o[5] = q4 - q3; o[6] = q5 - q4; o[7] = tile1 - tile0; }
    else o[4] = c_issue;
  }
  if (warp == 8) tc::tmem_dealloc(tbase, (uint32_t)a.ncols_tmem);
}

// dW[co][ci][j] += sum over slices of scratch[sl][j][ci/4][co][ci%4]
// block (32, 8): x = output float4 (coalesced 512 B per warp), y = slice group
template <bool V2>
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int Cout, int Cin,
                                                              int K, int coutp, int nslices) {
  __shared__ float4 part[8][32];
  const int64_t n = (int64_t)K * (Cin >> 2) * coutp;
  const int64_t slice_stride = n * 4;
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  float4 s = zero4();
  if (i < n) {
    int sl = threadIdx.y;
    if (V2) {  // four independent 16-byte loads in flight per thread (the partials sit in L2)
      float4 s1 = zero4();
      const float* p = scratch + i * 4;
      for (; sl + 24 < nslices; sl += 32) {
        const float4 v0 = ldg4(p + (sl + 0) * slice_stride), v1 = ldg4(p + (sl + 8) * slice_stride);
        const float4 v2 = ldg4(p + (sl + 16) * slice_stride), v3 = ldg4(p + (sl + 24) * slice_stride);
        s.x += v0.x + v2.x; s.y += v0.y + v2.y; s.z += v0.z + v2.z; s.w += v0.w + v2.w;
        s1.x += v1.x + v3.x; s1.y += v1.y + v3.y; s1.z += v1.z + v3.z; s1.w += v1.w + v3.w;
      }
      s.x += s1.x; s.y += s1.y; s.z += s1.z; s.w += s1.w;
    }
    for (; sl < nslices; sl += 8) {
      const float4 v = ldg4(scratch + sl * slice_stride + i * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
#pragma unroll
    for (int y = 1; y < 8; ++y) { const float4 v = part[y][threadIdx.x]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    const int co = (int)(i % coutp);
    if (co < Cout) {
      const int64_t r = i / coutp;
      const int c4 = (int)(r % (Cin >> 2)), j = (int)(r / (Cin >> 2));
      float* o = dw + ((int64_t)co * Cin + c4 * 4) * K + j;
      o[0] += s.x; o[K] += s.y; o[2 * K] += s.z; o[3 * K] += s.w;
    }
  }
}

// dW[co][ci][j] += acc[j][ci/4][co][ci%4]; acc = 0.   grid.y = layer (device item table), grid.x covers the largest
// layer (surplus blocks exit).  A block moves a 32 co x 32 ci x K tile through shared memory so that BOTH sides are
// coalesced: 512-byte runs of the accumulation buffer in, (32 ci x K)-float runs of the nn.Conv1d gradient out.  (Round
// 1 wrote the gradient with a 4-byte scatter at stride K: 209 us per step, now the kernel is copy-bound.)
__global__ void __launch_bounds__(256) wgrad_acc_flush_kernel(const avc_wgrad_acc_item* __restrict__ items) {
  __shared__ float tile[32][32 * 8 + 1];
  const avc_wgrad_acc_item it = items[blockIdx.y];
  const int K = it.K, C4 = it.Cin >> 2;
  const int coutp = cdiv(it.Cout, 128) * 128;
  const int ncg = cdiv(C4, 8);                      // groups of 8 four-channel chunks (32 input channels)
  const int nblk = (coutp >> 5) * ncg;
  if ((int)blockIdx.x >= nblk) return;
  const int cot = blockIdx.x / ncg, cg = blockIdx.x - cot * ncg;
  const int co0 = cot << 5, c40 = cg << 3;
  const int tid = threadIdx.x, col = tid & 31, c4l = tid >> 5;
  if (c40 + c4l < C4) {
    for (int j = 0; j < K; ++j) {
      float4* a4 = reinterpret_cast<float4*>(it.acc) + ((size_t)j * C4 + c40 + c4l) * coutp + co0 + col;
      const float4 s = *a4;
      *a4 = zero4();
      float* t = &tile[col][(c4l * 4) * K + j];
      t[0] = s.x; t[K] = s.y; t[2 * K] = s.z; t[3 * K] = s.w;
    }
  }
  __syncthreads();
  const int nci = min(32, it.Cin - c40 * 4);         // input channels of this tile
  const int run = nci * K;                            // contiguous floats of one output-channel row
  for (int r = tid >> 5; r < 32; r += 8) {
    const int co = co0 + r;
    if (co >= it.Cout) break;
    float* o = it.dw + ((size_t)co * it.Cin + c40 * 4) * K;
    for (int e = tid & 31; e < run; e += 32) o[e] += tile[r][e];
  }
}

static int wgrad_tc_plan(const avc_wgrad_desc* d, WgTcArgs& a) {
  const int T = d->Tout, K = d->K;
  a.d = *d;
  a.G = T >= 128 ? 1 : 128 / T;
  if (d->stride == 2) a.G = T >= 64 ? 1 : 64 / T;  // the parity-split input tile is twice as tall
  a.RA = a.G * T;
  // padded input positions one sample contributes: stride 1: T+K-1; stride 2: 2(T-1)+K
  a.TX = d->stride == 1 ? T + K - 1 : 2 * (T - 1) + K;
  a.H = ((a.TX + 1) / 2 + 3) / 4 * 4;
  // atom stride must keep every atom base 512 B aligned: the swizzle XOR is keyed on absolute
  // shared-memory address bits [7,9)
  // (and 1024 B aligned for the tensor-map copies of the TMA-staged kernel: a swizzled destination must sit on the
  // swizzle pattern's 8-row period)
  a.RX = d->stride == 1 ? (a.G * a.TX + 7) / 8 * 8 : a.G * 2 * a.H;
  a.ntpad = WT_NT;
  a.coutp = cdiv(d->Cout, 128) * 128;
  int ncols = 32;
  while (ncols < K * a.ntpad) ncols <<= 1;
  a.ncols_tmem = ncols;
  a.x_off = (uint32_t)(4 * a.RA * 128 + 1023) / 1024 * 1024;
  a.buf_bytes = (a.x_off + (uint32_t)((a.ntpad / 32) * a.RX * 128) + 1023) / 1024 * 1024;
  const int ntiles = cdiv(d->B, a.G);
  const int cta_per_slice = cdiv(d->Cin, WT_NT) * cdiv(d->Cout, 128);
  int nsl = 148 / cta_per_slice;
  if (nsl < 1) nsl = 1;
  if (nsl > ntiles) nsl = ntiles;
  a.tiles_per_slice = cdiv(ntiles, nsl);
  a.nslices = cdiv(ntiles, a.tiles_per_slice);
  return AVC_OK;
}

static bool wgrad_tc_supported(const avc_wgrad_desc* d) {
  return (d->stride == 1 || d->stride == 2) && d->Tout % 8 == 0 && d->Tout <= 128 && d->K >= 1 && d->K <= 8 && d->Cin % 4 == 0 &&
         d->Cout % 4 == 0 && d->Tin + d->K - 1 >= (d->Tout - 1) * d->stride + 1 && (d->stride == 1 || d->Tout <= 64);
}

}  // namespace avc

using namespace avc;

extern "C" int64_t avc_wgrad_tc_scratch_floats(const avc_wgrad_desc* d) {
  if (!d || !wgrad_tc_supported(d)) return -1;
  WgTcArgs a;
  wgrad_tc_plan(d, a);
  return (int64_t)a.nslices * d->K * d->Cin * a.coutp;
}

static long long* g_wg_dbg = nullptr;
static int wgrad_tc_launch(const avc_wgrad_desc* d, float* scratch, int* status, void* stream, bool accumulate, const char* who) {
  AVC_REQUIRE(d && d->x && d->dc && scratch && status && (accumulate || d->dw), AVC_ERR_INVALID, "%s: null argument", who);
  AVC_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0 && d->Tin > 0 && d->Tout > 0, AVC_ERR_INVALID, "%s: bad shape", who);
  AVC_REQUIRE(wgrad_tc_supported(d), AVC_ERR_UNSUPPORTED, "%s: needs stride 1 (Tout <= 128) or 2 (Tout <= 64), Tout %% 8 == 0, K <= 8", who);
  WgTcArgs a;
  wgrad_tc_plan(d, a);
  a.scratch = scratch;
  a.status = status;
  a.dbg = g_wg_dbg;
  const int smem = 2 * (int)a.buf_bytes;
  AVC_REQUIRE(smem <= 224 * 1024, AVC_ERR_UNSUPPORTED, "%s: tile does not fit shared memory", who);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
      return AVC_ERR_CUDA;
    }
    attr_done = true;
  }
  dim3 grid(cdiv(d->Cin, WT_NT), cdiv(d->Cout, 128), a.nslices);
  if (opt_wgrad_split()) {   // "wgrad_split": dedicated MMA warp
    static bool attr2 = false;
    if (!attr2) {
      cudaError_t e = cudaFuncSetAttribute(conv_wgrad_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
      if (e != cudaSuccess) {
        set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
        return AVC_ERR_CUDA;
      }
      attr2 = true;
    }
    if (accumulate) AVC_LAUNCH(conv_wgrad_split_kernel<true>, grid, WS_THREADS, smem, (cudaStream_t)stream, a);
    else AVC_LAUNCH(conv_wgrad_split_kernel<false>, grid, WS_THREADS, smem, (cudaStream_t)stream, a);
  } else {
    void (*kern)(const WgTcArgs) = opt_tc_uniform_issue() ? conv_wgrad_tc_kernel<true, false> : conv_wgrad_tc_kernel<false, false>;
    if (accumulate) kern = conv_wgrad_tc_kernel<true, true>;
    AVC_LAUNCH(kern, grid, 128, smem, (cudaStream_t)stream, a);
  }
  AVC_CHECK_LAUNCH(who);
  if (accumulate) return AVC_OK;
  const int64_t n = (int64_t)d->K * (d->Cin / 4) * a.coutp;
  if (opt_wgrad_reduce_v2())
    AVC_LAUNCH(wgrad_tc_reduce_kernel<true>, (int)cdiv64(n, 32), dim3(32, 8), 0, (cudaStream_t)stream, scratch, d->dw, d->Cout, d->Cin, d->K, a.coutp, a.nslices);
  else
    AVC_LAUNCH(wgrad_tc_reduce_kernel<false>, (int)cdiv64(n, 32), dim3(32, 8), 0, (cudaStream_t)stream, scratch, d->dw, d->Cout, d->Cin, d->K, a.coutp, a.nslices);
  AVC_CHECK_LAUNCH("wgrad_tc_reduce");
  return AVC_OK;
}

extern "C" int avc_conv_wgrad_tc(const avc_wgrad_desc* d, float* scratch, int* status, void* stream) {
  return wgrad_tc_launch(d, scratch, status, stream, false, "avc_conv_wgrad_tc");
}

// ---- accumulate-in-place variant: every layer owns a zeroed [K][Cin/4][coutp][4] accumulation
// buffer; avc_conv_wgrad_tc_acc adds into it (vector atomics), avc_wgrad_acc_flush folds every
// layer's buffer into its nn.Conv1d gradient and zeroes it again -- ONE launch per backward pass.
extern "C" int64_t avc_wgrad_acc_floats(int Cout, int Cin, int K) {
  if (Cout <= 0 || Cin <= 0 || K <= 0 || Cin % 4 != 0) return -1;
  return (int64_t)K * Cin * (cdiv(Cout, 128) * 128);
}
extern "C" int avc_conv_wgrad_tc_acc(const avc_wgrad_desc* d, float* acc, int* status, void* stream) {
  return wgrad_tc_launch(d, acc, status, stream, true, "avc_conv_wgrad_tc_acc");
}
extern "C" int avc_wgrad_acc_flush(const avc_wgrad_acc_item* items_dev, int n_items, int64_t max_units, void* stream) {
  AVC_REQUIRE(items_dev && n_items > 0 && max_units > 0, AVC_ERR_INVALID, "avc_wgrad_acc_flush: bad argument");
  // a layer needs (coutp / 32) * ceil(Cin / 32) blocks <= units / (256 K) + coutp / 32: max_units / 256 plus a margin
  // covers every layer (surplus blocks exit at once)
  dim3 grid((unsigned)cdiv64(max_units, 256) + 64u, (unsigned)n_items);
  AVC_LAUNCH(wgrad_acc_flush_kernel, grid, 256, 0, (cudaStream_t)stream, items_dev);
  AVC_CHECK_LAUNCH("wgrad_acc_flush");
  return AVC_OK;
}

extern "C" void avc_wgrad_tc_set_debug(void* dev_buffer) { g_wg_dbg = (long long*)dev_buffer; }
