// gate_tc.cuh -- fused 4- and 5-qubit fp32 gates on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), 3xTF32 split for fp32-level accuracy.
//
// Why.  A G=4 fp32 pass on the CUDA cores needs 32 FFMA2 per amplitude: it sits at the FMA/HBM
// ridge (ncu: FMA pipe ~70 % busy while DRAM is at ~88 %), and under the board power cap the
// FMA work is what keeps it off the copy roofline; G=5 is FMA-bound outright.  On the tensor
// cores the same mat-vec costs 12 (G=4) or 24 (G=5) MMA instructions per 128 groups, ~15-30 %
// of the tile's HBM time, and the CUDA cores only move data.
//
// Formulation.  One CTA tile = 128 groups (M = 128).  With the state in normal order a group's
// 2^G amplitudes ARE a row of K = 2*2^G floats (re, im interleaved), so
//     D[128 x 2N] = A[128 x 2N] * W^T,   W[2r][2c] = Re U[r][c]   W[2r][2c+1] = -Im U[r][c]
//                                        W[2r+1][2c] = Im U[r][c] W[2r+1][2c+1] = Re U[r][c]
// maps the interleaved input row to the interleaved output row: no (re, im) de-interleaving
// anywhere.  A and W are split x = hi + lo with hi = rn_tf32(x) (exact in fp32); the tile is
// D = A_lo W_hi + A_hi W_lo + A_hi W_hi (the 2^-22 lo*lo term is dropped), fp32 accumulation
// in TMEM.
//
// Data path per tile (256 threads, two per row: thread t owns half t>>7 of row t&127):
// LDG of the thread's half row into registers, issued one tile ahead -> hi/lo split in
// registers -> two conflict-free STS.128 per 16-byte chunk straight into the canonical K-major
// SWIZZLE_128B operand tiles (shared memory is written once and read only by the tensor core:
// the pass is smem-bandwidth sensitive, a cp.async staging ring + in-place split costs twice
// the wavefronts) -> fence.proxy.async + barrier -> one thread issues the MMAs and commits to
// an mbarrier -> every thread tcgen05.ld's its half of its TMEM lane (= its group) and stores
// the results straight from registers.  With NBUF = 2 operand/accumulator buffers the MMAs of
// tile i run while tile i-1 is read out and stored; co-resident CTAs cover each other's barriers.
#pragma once

#include "gate_pipe.cuh"

namespace qb200 {

namespace tc {

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  // UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start address >> 4 in [0,14),
  // leading byte offset (unused for swizzled K-major) = 1 in [16,30), stride byte offset =
  // 1024 B (8 rows x 128 B) >> 4 in [32,46), descriptor version 1 in [46,48), layout type 2 in [61,64)
  return (uint64_t) ((saddr & 0x3FFFFu) >> 4) | (uint64_t{1} << 16) | (uint64_t{64} << 32) |
         (uint64_t{1} << 46) | (uint64_t{2} << 61);
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t instr_desc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(mbar), "r"(parity)
      : "memory");
}
// one lane of a converged warp (the compiler keeps the guarded code on the uniform datapath)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 %%rx;\n\t"
      ".reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t"
      "}\n"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 / 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// tcgen05.st: 16 / 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]: A is M x K with rows on TMEM lanes and K along 32-bit columns
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// hi = x rounded to nearest tf32 (10-bit mantissa), exact as an fp32 bit pattern
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

}  // namespace tc

template <int G>
struct TcShape {
  static constexpr int N = 1 << G;
  static constexpr int KF = 2 * N;                  // floats per row (K) = output floats (MMA N)
  static constexpr int ATOMS = KF / 32;             // 128-byte swizzle atoms along K
  static constexpr int A_ATOM = 128 * 128;          // bytes: 128 rows x 128 B
  static constexpr int A_TILE = ATOMS * A_ATOM;
  static constexpr int B_ATOM = KF * 128;           // bytes: KF rows x 128 B
  static constexpr int B_TILE = ATOMS * B_ATOM;
  
};

// [W_hi][W_lo][NBUF x (A_hi, A_lo)] + alignment slack
template <int G, int NBUF>
constexpr size_t tc_smem_bytes() {
  return 1024 + 2 * TcShape<G>::B_TILE + (size_t) (2 * NBUF) * TcShape<G>::A_TILE;
}

// byte offset of (row, 16-byte chunk c16 of the row's 128-byte atom slice) in a K-major SW128 atom
__device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t c16) {
  return (row >> 3) * 1024 + (row & 7) * 128 + ((c16 ^ (row & 7)) << 4);
}

constexpr int kTcThreads = 256;

// PF = tiles of register prefetch (1 or 2): PF * 64 B (G=4) per thread in flight
template <int G, bool PAIR, int NBUF, int PF, int MINB>
__global__ void __launch_bounds__(kTcThreads, MINB)
k_gate_tc(float* __restrict__ st, const __grid_constant__ Geom g,
          const __grid_constant__ MatParam<float, G> mat) {
  using S = TcShape<G>;
  constexpr int N = S::N, KF = S::KF, HN = N / 2;  // HN amplitudes per thread
  constexpr int CHUNKS = HN / 2;                   // 16-byte chunks (2 amplitudes) per thread
  static_assert(NBUF == 1 || NBUF == 2, "one or two operand/accumulator buffers");
  extern __shared__ unsigned char tc_raw[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t t = threadIdx.x;
  const uint32_t warp = __shfl_sync(0xffffffffu, t >> 5, 0);  // warp-uniform for the compiler too
  const uint32_t row = t & 127, half = t >> 7;
  const uint32_t raw_s = smem_u32(tc_raw);
  const uint32_t base_s = (raw_s + 1023u) & ~1023u;
  unsigned char* const base_p = tc_raw + (base_s - raw_s);
  const uint32_t bhi_s = base_s, blo_s = base_s + S::B_TILE;
  const uint32_t a_s = base_s + 2 * S::B_TILE;  // buffer b: hi at a_s + 2b*A_TILE, lo right behind
  unsigned char* const bhi_p = base_p;
  unsigned char* const blo_p = base_p + S::B_TILE;
  unsigned char* const a_p = base_p + 2 * S::B_TILE;

  // ---- one-time setup: W = real form of the gate, split hi/lo, swizzled K-major ----
  for (uint32_t idx = t; idx < (uint32_t) (KF * KF); idx += kTcThreads) {
    const uint32_t n = idx / KF, k = idx % KF;
    const uint32_t r = n >> 1, c = k >> 1;
    const float ur = mat.m[2 * (r * N + c)], ui = mat.m[2 * (r * N + c) + 1];
    const float w = (n & 1) ? ((k & 1) ? ur : ui) : ((k & 1) ? -ui : ur);
    const float hi = tc::tf32_hi(w);
    const uint32_t off = (k >> 5) * S::B_ATOM + sw128_off(n, (k & 31) >> 2) + (k & 3) * 4;
    *reinterpret_cast<float*>(bhi_p + off) = hi;
    *reinterpret_cast<float*>(blo_p + off) = w - hi;
  }
  constexpr int TCOLS = NBUF * KF;  // 32, 64 or 128: powers of two >= 32
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "n"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0) {
    tc::mbar_init(smem_u32(&mbar[0]), 1);
    tc::mbar_init(smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  // a warp reaches TMEM lanes 32*(warp%4)..+31; this thread's columns: its half of the row
  const uint32_t tmem_mine = tmem_base + (((warp & 3) * 32u) << 16) + half * N;
  constexpr uint32_t idesc = tc::instr_desc_tf32(KF);

  const uint64_t ntiles = g.work >> 7;
  const uint64_t stride = gridDim.x;

  const uint32_t row_off = (row >> 3) * 1024 + (row & 7) * 128;
  const uint32_t row_x = row & 7;
  const int k0 = (int) half * HN;  // first amplitude of this thread's half row
  // Addressing, hoisted out of the tile loop (it was ~2/3 of all issued instructions):
  //   amplitude index of (tile, row, element k0 + j) = dep(tile << 7) | dep(row) | cbits + half * xs[G-1] + eo(j)
  // dep() deposits index bits at the free positions; the 7 row bits take the 7 lowest free ones, so
  // the two deposits are disjoint and the row part is a per-thread constant.
  const uint64_t thread_off = 8 * (expand_index(row, g) + (half ? g.xs[G - 1] : 0));  // bytes, includes cbits
  unsigned char* const st_b = reinterpret_cast<unsigned char*>(st);
  auto tile_ptr = [&](uint64_t tile) {  // warp-uniform part, evaluated on the uniform datapath
    uint64_t i = tile << 7;
    for (uint32_t k = 0; k < g.npos; ++k) {
      const uint64_t lo = i & ((uint64_t{1} << g.pos[k]) - 1);
      i = ((i - lo) << 1) | lo;
    }
    return st_b + 8 * i + thread_off;
  };
  auto eo = [&](int j) { return 8 * elem_offset<G>(j, g); };  // byte offset of element j, uniform

  auto load_mine = [&](uint64_t tile, uint4 (&x)[CHUNKS]) {
    const unsigned char* const p = tile_ptr(tile);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      if constexpr (PAIR) {
        x[c] = *reinterpret_cast<const uint4*>(p + eo(2 * c));
      } else {
        const uint2 a = *reinterpret_cast<const uint2*>(p + eo(2 * c));
        const uint2 b = *reinterpret_cast<const uint2*>(p + eo(2 * c + 1));
        x[c] = make_uint4(a.x, a.y, b.x, b.y);
      }
    }
  };

  // hi/lo split in registers, one STS.128 each into the operand tiles of buffer b
  auto split_store = [&](const uint4 (&x)[CHUNKS], int b) {
    unsigned char* const hrow = a_p + (2 * b) * S::A_TILE + row_off;
    unsigned char* const lrow = hrow + S::A_TILE;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int q = (k0 >> 1) + c;  // chunk within the row (over all atoms)
      // physical chunk (q ^ row&7): the 8 lanes of a quarter-warp hit 8 different chunks
      const uint32_t off = (q >> 3) * S::A_ATOM + ((((uint32_t) q & 7) ^ row_x) << 4);
      const float x0 = __uint_as_float(x[c].x), x1 = __uint_as_float(x[c].y);
      const float x2 = __uint_as_float(x[c].z), x3 = __uint_as_float(x[c].w);
      float4 h, l;
      h.x = tc::tf32_hi(x0); h.y = tc::tf32_hi(x1); h.z = tc::tf32_hi(x2); h.w = tc::tf32_hi(x3);
      l.x = x0 - h.x; l.y = x1 - h.y; l.z = x2 - h.z; l.w = x3 - h.w;
      *reinterpret_cast<float4*>(hrow + off) = h;
      *reinterpret_cast<float4*>(lrow + off) = l;
    }
  };

  auto issue_mmas = [&](int b, uint32_t mb) {
    tc::fence_after();
    const uint32_t d = tmem_base + b * KF;
    const uint32_t ahi = a_s + (2 * b) * S::A_TILE, alo = ahi + S::A_TILE;
    uint32_t acc = 0;
#pragma unroll
    for (int term = 0; term < 3; ++term) {
      const uint32_t ab = term == 0 ? alo : ahi;
      const uint32_t bb = term == 1 ? blo_s : bhi_s;
#pragma unroll
      for (int k = 0; k < KF / 8; ++k) {
        const uint64_t ad = tc::smem_desc_sw128(ab + (k >> 2) * S::A_ATOM + (k & 3) * 32);
        const uint64_t bd = tc::smem_desc_sw128(bb + (k >> 2) * S::B_ATOM + (k & 3) * 32);
        tc::mma_tf32(d, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    tc::mma_commit(mb);
  };

  auto wait_tile = [&](uint32_t i) {  // MMAs of this CTA's i-th tile are complete
    tc::mbar_wait(smem_u32(&mbar[i % NBUF]), (i / NBUF) & 1);
    tc::fence_after();
  };

  auto store_mine = [&](uint64_t tile, const uint32_t (&v)[N]) {
    unsigned char* const p = tile_ptr(tile);
#pragma unroll
    for (int j = 0; j < HN; j += (PAIR ? 2 : 1)) {
      if constexpr (PAIR) {
        *reinterpret_cast<uint4*>(p + eo(j)) =
            make_uint4(v[2 * j], v[2 * j + 1], v[2 * j + 2], v[2 * j + 3]);
      } else {
        *reinterpret_cast<uint2*>(p + eo(j)) = make_uint2(v[2 * j], v[2 * j + 1]);
      }
    }
  };

  static_assert(PF == 1 || PF == 2, "prefetch one or two tiles ahead");
  uint4 cur[CHUNKS], nxt[CHUNKS], nx2[PF == 2 ? CHUNKS : 1];
  uint64_t tile = blockIdx.x, prev_tile = 0;
  if (tile < ntiles) load_mine(tile, cur);
  if constexpr (PF == 2) {
    if (tile + stride < ntiles) load_mine(tile + stride, nxt);
  }
  uint32_t it = 0;
  for (; tile < ntiles; tile += stride, ++it) {
    const int b = it % NBUF;
    uint32_t v[N];
    if constexpr (NBUF == 1) {
      // one buffer: tile it-1 must be out of the tensor core (operands) and TMEM before it is reused
      if (it > 0) {
        wait_tile(it - 1);
        tc::tmem_ld(tmem_mine, v);
      }
    }
    split_store(cur, b);  // NBUF == 2: buffer b was released by the wait of iteration it-1
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    if (warp == 0) {
      // single-thread MMA issue from converged, warp-uniform control flow: descriptors stay in
      // uniform registers (a divergent `if (threadIdx.x == 0)` costs ~100 cycles per MMA in
      // R2UR waterfall loops)
      if (tc::elect_one()) {
        issue_mmas(b, smem_u32(&mbar[b]));
      }
      __syncwarp();
    }
    // Prefetch for the following tiles, issued AFTER the proxy fence and the barrier: a fence
    // completes only when the thread's earlier memory operations have, so loads issued in front
    // of it would be waited for right there (measured: no prefetch effect at all).
    if constexpr (PF == 2) {
      if (tile + 2 * stride < ntiles) load_mine(tile + 2 * stride, nx2);
    } else {
      if (tile + stride < ntiles) load_mine(tile + stride, nxt);
    }
    if (it > 0) {
      if constexpr (NBUF == 2) {
        wait_tile(it - 1);
            tc::tmem_ld(tmem_mine + ((it - 1) & 1) * KF, v);
          }
      store_mine(prev_tile, v);
      }
    prev_tile = tile;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      cur[c] = nxt[c];
      if constexpr (PF == 2) nxt[c] = nx2[c];
    }
  }
  if (it > 0) {
    uint32_t v[N];
    wait_tile(it - 1);
    tc::tmem_ld(tmem_mine + ((it - 1) % NBUF) * KF, v);
    store_mine(prev_tile, v);
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// Variant with the A operand in TENSOR MEMORY.  The split halves of the state tile never touch
// shared memory: each thread tcgen05.st's the hi and lo halves of its half row into its own TMEM
// lane, the MMAs read A from TMEM (tcgen05.mma [d], [a], b-desc) and only W (a few KB) comes from
// shared memory.  No generic->async proxy fence per tile, no operand traffic on the shared-memory
// port, no operand buffers in shared memory (residency is bounded by TMEM columns alone).
// TMEM columns of buffer b: MT x [A_hi KF][A_lo KF][D KF].
// ---------------------------------------------------------------------------------------------
template <int G, int NBUF>
__host__ __device__ constexpr int tca_tmem_cols() {
  const int need = NBUF * 3 * TcShape<G>::KF;
  return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}
// W_hi, W_lo, W_c (+ a RING-deep cp.async staging ring of raw tiles, 64 B (G=4) / 128 B (G=5) per thread and stage)
template <int G, int RING, int MT>
constexpr size_t tca_smem_bytes() {
  return 1024 + 3 * TcShape<G>::B_TILE + (size_t) RING * kTcThreads * MT * (TcShape<G>::N / 2) * 8;
}

// Mean relative shrink of one pass caused by the tensor core's truncating fp32 accumulation,
// measured as norm drift over 64 random passes (tools/tc_check.py): -1.7355e-7 (G = 4) and
// -3.0130e-7 (G = 5) on the squared norm.  It is cancelled by a fourth, tiny MMA term
// A_hi * (c W_hi) with c = half the drift.
template <int G> __host__ __device__ constexpr float tc_bias() { return G == 4 ? 0.86777e-7f : 1.50648e-7f; }

// MT = 128-row MMA tiles per CTA iteration (256 threads each): MT = 2 halves the barriers per byte for G = 4
// COMP = add the accumulation-bias compensation term
// RING = 0: register prefetch PF tiles ahead; RING >= 2: cp.async staging ring, RING - 1 tiles ahead
// (cp.async groups are waited for exactly; LDG results share counting scoreboards, so a register
// prefetch deeper than one tile does not buy latency)
template <int G, bool PAIR, int NBUF, int PF, int MT, bool COMP, int RING, int MINB>
__global__ void __launch_bounds__(kTcThreads * MT, MINB)
k_gate_tca(float* __restrict__ st, const __grid_constant__ Geom g,
           const __grid_constant__ MatParam<float, G> mat) {
  using S = TcShape<G>;
  constexpr int N = S::N, KF = S::KF, HN = N / 2;
  constexpr int CHUNKS = HN / 2;
  constexpr int TCOLS = tca_tmem_cols<G, NBUF * MT>();
  constexpr int BUFC = MT * 3 * KF;  // TMEM columns of one buffer: MT x [A_hi | A_lo | D]
  static_assert(NBUF == 1 || NBUF == 2, "one or two operand/accumulator buffers");
  static_assert(NBUF * MT * 3 * KF <= 512, "tensor memory has 512 columns");
  extern __shared__ unsigned char tc_raw[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ __align__(8) uint64_t full_bar;  // all rows of the tile are in TMEM (count = threads)
  __shared__ uint32_t tmem_slot;

  const uint32_t t = threadIdx.x;
  const uint32_t warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const uint32_t row = t & 127, half = (t >> 7) & 1, mt = MT == 1 ? 0 : t >> 8;
  const uint32_t raw_s = smem_u32(tc_raw);
  const uint32_t base_s = (raw_s + 1023u) & ~1023u;
  unsigned char* const base_p = tc_raw + (base_s - raw_s);
  const uint32_t bhi_s = base_s, blo_s = base_s + S::B_TILE, bc_s = base_s + 2 * S::B_TILE;
  unsigned char* const bhi_p = base_p;
  unsigned char* const blo_p = base_p + S::B_TILE;
  unsigned char* const bc_p = base_p + 2 * S::B_TILE;
  // staging ring: stage s, 16-byte slot q, thread t at ((s * CHUNKS + q) * NT + t) * 16 (thread-private, conflict free)
  constexpr int NT = kTcThreads * MT;
  unsigned char* const ring_p = base_p + 3 * S::B_TILE + t * 16;
  const uint32_t ring_s = base_s + 3 * S::B_TILE + t * 16;

  for (uint32_t idx = t; idx < (uint32_t) (KF * KF); idx += kTcThreads * MT) {
    const uint32_t n = idx / KF, k = idx % KF;
    const uint32_t r = n >> 1, c = k >> 1;
    const float ur = mat.m[2 * (r * N + c)], ui = mat.m[2 * (r * N + c) + 1];
    const float w = (n & 1) ? ((k & 1) ? ur : ui) : ((k & 1) ? -ui : ur);
    const float hi = tc::tf32_hi(w);
    const uint32_t off = (k >> 5) * S::B_ATOM + sw128_off(n, (k & 31) >> 2) + (k & 3) * 4;
    *reinterpret_cast<float*>(bhi_p + off) = hi;
    *reinterpret_cast<float*>(blo_p + off) = w - hi;
    *reinterpret_cast<float*>(bc_p + off) = COMP ? tc_bias<G>() * hi : 0.f;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "n"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0) {
    tc::mbar_init(smem_u32(&mbar[0]), 1);
    tc::mbar_init(smem_u32(&mbar[1]), 1);
    tc::mbar_init(smem_u32(&full_bar), kTcThreads * MT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc::fence_async_smem();  // W was written through the generic proxy, the MMAs read it through the async proxy
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  // this thread's lane (its row) and its half of the K / N columns inside a [KF]-wide block
  const uint32_t tmem_mine = tmem_base + (((warp & 3) * 32u) << 16) + mt * 3 * KF + half * N;
  constexpr uint32_t idesc = tc::instr_desc_tf32(KF);

  const uint64_t ntiles = g.work >> 7 >> (MT - 1);  // CTA tiles of 128 * MT groups
  const uint64_t stride = gridDim.x;

  const uint64_t thread_off = 8 * (expand_index(row, g) + (half ? g.xs[G - 1] : 0));
  unsigned char* const st_b = reinterpret_cast<unsigned char*>(st);
  auto tile_ptr = [&](uint64_t tile) {
    uint64_t i = (tile * MT + mt) << 7;
    for (uint32_t k = 0; k < g.npos; ++k) {
      const uint64_t lo = i & ((uint64_t{1} << g.pos[k]) - 1);
      i = ((i - lo) << 1) | lo;
    }
    return st_b + 8 * i + thread_off;
  };
  auto eo = [&](int j) { return 8 * elem_offset<G>(j, g); };

  auto load_mine = [&](const unsigned char* p, uint4 (&x)[CHUNKS]) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      if constexpr (PAIR) {
        x[c] = *reinterpret_cast<const uint4*>(p + eo(2 * c));
      } else {
        const uint2 a = *reinterpret_cast<const uint2*>(p + eo(2 * c));
        const uint2 b = *reinterpret_cast<const uint2*>(p + eo(2 * c + 1));
        x[c] = make_uint4(a.x, a.y, b.x, b.y);
      }
    }
  };

  auto ring_issue = [&](uint64_t tile, int stg) {
    if (tile < ntiles) {
      const unsigned char* const p = tile_ptr(tile);
      const uint32_t dst = ring_s + (uint32_t) (stg * CHUNKS) * NT * 16;
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        if constexpr (PAIR) {
          cp_async16(dst + c * NT * 16, p + eo(2 * c));
        } else {
          cp_async8(dst + c * NT * 16, p + eo(2 * c));
          cp_async8(dst + c * NT * 16 + 8, p + eo(2 * c + 1));
        }
      }
    }
    cp_async_commit();
  };
  auto ring_read = [&](int stg, uint4 (&x)[CHUNKS]) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
      x[c] = *reinterpret_cast<const uint4*>(ring_p + (size_t) (stg * CHUNKS + c) * NT * 16);
  };

  // hi/lo split in registers, written to this thread's TMEM lane: columns [A_hi | A_lo] of buffer b
  auto split_to_tmem = [&](const uint4 (&x)[CHUNKS], int b) {
    uint32_t h[N], l[N];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const uint32_t w[4] = {x[c].x, x[c].y, x[c].z, x[c].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xf = __uint_as_float(w[e]);
        const float hf = tc::tf32_hi(xf);
        h[4 * c + e] = __float_as_uint(hf);
        l[4 * c + e] = __float_as_uint(xf - hf);
      }
    }
    tc::tmem_st(tmem_mine + b * BUFC, h);
    tc::tmem_st(tmem_mine + b * BUFC + KF, l);
    tc::tmem_st_wait();
  };

  auto issue_mmas = [&](int b, uint32_t mb) {
    tc::fence_after();
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const uint32_t ahi = tmem_base + b * BUFC + m * 3 * KF, alo = ahi + KF, d = ahi + 2 * KF;
      uint32_t acc = 0;
      // smallest terms first: A_lo W_hi, A_hi W_lo, A_hi W_hi (the accumulation-bias compensation is an FMA in
      // store_mine, not a fourth MMA term: a quarter less tensor work per tile)
#pragma unroll
      for (int term = 1; term < 4; ++term) {
        const uint32_t ab = term == 1 ? alo : ahi;
        const uint32_t bb = term == 2 ? blo_s : bhi_s;
#pragma unroll
        for (int k = 0; k < KF / 8; ++k) {
          const uint64_t bd = tc::smem_desc_sw128(bb + (k >> 2) * S::B_ATOM + (k & 3) * 32);
          tc::mma_tf32_ts(d, ab + 8 * k, bd, idesc, acc);
          acc = 1;
        }
      }
    }
    tc::mma_commit(mb);
  };

  auto wait_tile = [&](uint32_t i) {
    tc::mbar_wait(smem_u32(&mbar[i % NBUF]), (i / NBUF) & 1);
    tc::fence_after();
  };

  auto store_mine = [&](unsigned char* p, uint32_t (&v)[N]) {
    if constexpr (COMP) {
      // the tensor core truncates its fp32 accumulation: a pass shrinks the state by ~(1 - tc_bias); v + c v with one
      // rounding puts the mean back
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), tc_bias<G>(), __uint_as_float(v[j])));
    }
#pragma unroll
    for (int j = 0; j < HN; j += (PAIR ? 2 : 1)) {
      if constexpr (PAIR) {
        *reinterpret_cast<uint4*>(p + eo(j)) = make_uint4(v[2 * j], v[2 * j + 1], v[2 * j + 2], v[2 * j + 3]);
      } else {
        *reinterpret_cast<uint2*>(p + eo(j)) = make_uint2(v[2 * j], v[2 * j + 1]);
      }
    }
  };

  static_assert(PF == 1 || PF == 2, "prefetch one or two tiles ahead");
  if constexpr (RING >= 2) {
    const uint64_t first = blockIdx.x;
#pragma unroll
    for (int sg = 0; sg < RING - 1; ++sg) ring_issue(first + sg * stride, sg);
    int sg = 0;
    uint32_t it = 0;
    uint64_t prev_tile = 0;
    for (uint64_t tile = first; tile < ntiles; tile += stride, ++it) {
      const int b = it % NBUF;
      uint32_t v[N];
      if constexpr (NBUF == 1) {
        if (it > 0) {
          wait_tile(it - 1);
          tc::tmem_ld(tmem_mine + 2 * KF, v);
        }
      }
      cp_async_wait<RING - 2>();  // this thread's pieces of tile `tile` (stage sg) have landed
      uint4 x[CHUNKS];
      ring_read(sg, x);
      split_to_tmem(x, b);
      // refill the stage read one iteration ago (by this same thread: no cross-thread hazard)
      int sp = sg + RING - 1;
      if (sp >= RING) sp -= RING;
      ring_issue(tile + (RING - 1) * stride, sp);
      // split barrier: everybody arrives, only the issuing warp waits -- the other warps go straight
      // on to storing the previous tile instead of idling until the slowest warp has written its rows
      tc::fence_before();
      tc::mbar_arrive(smem_u32(&full_bar));
      if (warp == 0) {
        tc::mbar_wait(smem_u32(&full_bar), it & 1);
        if (tc::elect_one()) issue_mmas(b, smem_u32(&mbar[b]));
        __syncwarp();
      }
      if (it > 0) {
        if constexpr (NBUF == 2) {
          wait_tile(it - 1);
          tc::tmem_ld(tmem_mine + ((it - 1) & 1) * BUFC + 2 * KF, v);
        }
        store_mine(tile_ptr(prev_tile), v);
      }
      prev_tile = tile;
      if (++sg == RING) sg = 0;
    }
    if (it > 0) {
      uint32_t v[N];
      wait_tile(it - 1);
      tc::tmem_ld(tmem_mine + ((it - 1) % NBUF) * BUFC + 2 * KF, v);
      store_mine(tile_ptr(prev_tile), v);
    }
    cp_async_wait<0>();
  } else {
    // tile pointers are computed once per tile (when its loads are issued) and carried along
    uint4 cur[CHUNKS], nxt[CHUNKS], nx2[PF == 2 ? CHUNKS : 1];
    unsigned char *p_cur = nullptr, *p_nxt = nullptr, *p_nx2 = nullptr, *p_prev = nullptr;
    uint64_t tile = blockIdx.x;
    if (tile < ntiles) { p_cur = tile_ptr(tile); load_mine(p_cur, cur); }
    if constexpr (PF == 2) {
      if (tile + stride < ntiles) { p_nxt = tile_ptr(tile + stride); load_mine(p_nxt, nxt); }
    }
    uint32_t it = 0;
    for (; tile < ntiles; tile += stride, ++it) {
      const int b = it % NBUF;
      uint32_t v[N];
      if constexpr (NBUF == 1) {
        if (it > 0) {
          wait_tile(it - 1);
          tc::tmem_ld(tmem_mine + 2 * KF, v);
        }
      }
      split_to_tmem(cur, b);
      if constexpr (PF == 2) {
        if (tile + 2 * stride < ntiles) { p_nx2 = tile_ptr(tile + 2 * stride); load_mine(p_nx2, nx2); }
      } else {
        if (tile + stride < ntiles) { p_nxt = tile_ptr(tile + stride); load_mine(p_nxt, nxt); }
      }
      // split barrier: everybody arrives, only the issuing warp waits -- the other warps go straight
      // on to storing the previous tile instead of idling until the slowest warp has written its rows
      tc::fence_before();
      tc::mbar_arrive(smem_u32(&full_bar));
      if (warp == 0) {
        tc::mbar_wait(smem_u32(&full_bar), it & 1);
        if (tc::elect_one()) issue_mmas(b, smem_u32(&mbar[b]));
        __syncwarp();
      }
      if (it > 0) {
        if constexpr (NBUF == 2) {
          wait_tile(it - 1);
          tc::tmem_ld(tmem_mine + ((it - 1) & 1) * BUFC + 2 * KF, v);
        }
        store_mine(p_prev, v);
      }
      p_prev = p_cur;
      p_cur = p_nxt;
      if constexpr (PF == 2) p_nxt = p_nx2;
  #pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        cur[c] = nxt[c];
        if constexpr (PF == 2) nxt[c] = nx2[c];
      }
    }
    if (it > 0) {
      uint32_t v[N];
      wait_tile(it - 1);
      tc::tmem_ld(tmem_mine + ((it - 1) % NBUF) * BUFC + 2 * KF, v);
      store_mine(p_prev, v);
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// k_gate_tcx: the same tensor-memory data path, written for the cases k_gate_tca does not cover:
// 6-qubit gates (K = N = 128: 384 TMEM columns, W tiles of 64 KB, one CTA per SM, the half row
// processed in 32-float pieces to stay inside the register file) and EXPECTATION VALUES for
// G = 4, 5, 6 (read-only: D = U x through the MMAs, then <x|D> per row in the epilogue with x
// re-read from tensor memory as A_hi + A_lo, which is exact; products in fp32, accumulation in
// double like lib/simulator_basic.h:323-324).  Matrix from device memory (a 6-qubit matrix does
// not fit the kernel parameter space).  One operand/accumulator buffer, one tile of prefetch.
// ---------------------------------------------------------------------------------------------
template <int G>
constexpr size_t tcx_smem_bytes() { return 1024 + 2 * TcShape<G>::B_TILE; }   // W_hi, W_lo

template <int G, bool PAIR, bool EXPECT>
__global__ void __launch_bounds__(kTcThreads, 1)
k_gate_tcx(float* __restrict__ st, const __grid_constant__ Geom g, const float* __restrict__ umat,
           const float comp, double* __restrict__ partials) {
  using S = TcShape<G>;
  constexpr int N = S::N, KF = S::KF, HN = N / 2;  // this thread: HN amplitudes = N floats
  constexpr int CHUNKS = HN / 2;
  constexpr int PIECE = N < 32 ? N : 32;           // floats per TMEM transfer
  constexpr int NPIECE = N / PIECE;
  constexpr int TCOLS = tca_tmem_cols<G, 1>();
  extern __shared__ unsigned char tc_raw[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ __align__(8) uint64_t full_bar;
  __shared__ uint32_t tmem_slot;

  const uint32_t t = threadIdx.x;
  const uint32_t warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const uint32_t row = t & 127, half = t >> 7;
  const uint32_t raw_s = smem_u32(tc_raw);
  const uint32_t base_s = (raw_s + 1023u) & ~1023u;
  unsigned char* const base_p = tc_raw + (base_s - raw_s);
  const uint32_t bhi_s = base_s, blo_s = base_s + S::B_TILE;

  for (uint32_t idx = t; idx < (uint32_t) (KF * KF); idx += kTcThreads) {
    const uint32_t n = idx / KF, k = idx % KF;
    const uint32_t r = n >> 1, c = k >> 1;
    const float ur = umat[2 * (r * N + c)], ui = umat[2 * (r * N + c) + 1];
    const float w = (n & 1) ? ((k & 1) ? ur : ui) : ((k & 1) ? -ui : ur);
    const float hi = tc::tf32_hi(w);
    const uint32_t off = (k >> 5) * S::B_ATOM + sw128_off(n, (k & 31) >> 2) + (k & 3) * 4;
    *reinterpret_cast<float*>(base_p + off) = hi;
    *reinterpret_cast<float*>(base_p + S::B_TILE + off) = w - hi;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "n"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0) {
    tc::mbar_init(smem_u32(&mbar), 1);
    tc::mbar_init(smem_u32(&full_bar), kTcThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  // TMEM columns: [A_hi KF][A_lo KF][D KF]; this thread: its lane, its half of each block
  const uint32_t tmem_mine = tmem_base + (((warp & 3) * 32u) << 16) + half * N;
  constexpr uint32_t idesc = tc::instr_desc_tf32(KF);

  const uint64_t ntiles = g.work >> 7;
  const uint64_t stride = gridDim.x;
  const uint64_t thread_off = 8 * (expand_index(row, g) + (half ? g.xs[G - 1] : 0));
  unsigned char* const st_b = reinterpret_cast<unsigned char*>(st);
  auto tile_ptr = [&](uint64_t tile) {
    uint64_t i = tile << 7;
    for (uint32_t k = 0; k < g.npos; ++k) {
      const uint64_t lo = i & ((uint64_t{1} << g.pos[k]) - 1);
      i = ((i - lo) << 1) | lo;
    }
    return st_b + 8 * i + thread_off;
  };
  auto eo = [&](int j) { return 8 * elem_offset<G>(j, g); };

  auto load_mine = [&](const unsigned char* p, uint4 (&x)[CHUNKS]) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      if constexpr (PAIR) {
        x[c] = *reinterpret_cast<const uint4*>(p + eo(2 * c));
      } else {
        const uint2 a = *reinterpret_cast<const uint2*>(p + eo(2 * c));
        const uint2 b = *reinterpret_cast<const uint2*>(p + eo(2 * c + 1));
        x[c] = make_uint4(a.x, a.y, b.x, b.y);
      }
    }
  };

  auto split_to_tmem = [&](const uint4 (&x)[CHUNKS]) {
#pragma unroll
    for (int pc = 0; pc < NPIECE; ++pc) {
      uint32_t h[PIECE], l[PIECE];
#pragma unroll
      for (int c = 0; c < PIECE / 4; ++c) {
        const uint4 xv = x[pc * (PIECE / 4) + c];
        const uint32_t w[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xf = __uint_as_float(w[e]);
          const float hf = tc::tf32_hi(xf);
          h[4 * c + e] = __float_as_uint(hf);
          l[4 * c + e] = __float_as_uint(xf - hf);
        }
      }
      tc::tmem_st(tmem_mine + pc * PIECE, h);
      tc::tmem_st(tmem_mine + KF + pc * PIECE, l);
    }
    tc::tmem_st_wait();
  };

  // accumulator buffer b (gates: two of them, so that tile i's MMAs run while tile i-1 is read out and stored; the
  // allocation is the power of two above 3 KF = 4 KF columns anyway)
  auto issue_mmas = [&](uint32_t b) {
    tc::fence_after();
    const uint32_t ahi = tmem_base, alo = ahi + KF, d = ahi + 2 * KF + b * KF;
    uint32_t acc = 0;
#pragma unroll
    for (int term = 1; term < 4; ++term) {   // A_lo W_hi + A_hi W_lo + A_hi W_hi
      const uint32_t ab = term == 1 ? alo : ahi;
      const uint32_t bb = term == 2 ? blo_s : bhi_s;
#pragma unroll 4
      for (int k = 0; k < KF / 8; ++k) {
        const uint64_t bd = tc::smem_desc_sw128(bb + (k >> 2) * S::B_ATOM + (k & 3) * 32);
        tc::mma_tf32_ts(d, ab + 8 * k, bd, idesc, acc);
        acc = 1;
      }
    }
    tc::mma_commit(smem_u32(&mbar));
  };

  double ere = 0, eim = 0;
  // tile it-1 is complete in TMEM: write it back (gate) or fold <x|D> into the accumulators (expectation)
  auto epilogue = [&](unsigned char* p, uint32_t b) {
#pragma unroll
    for (int pc = 0; pc < NPIECE; ++pc) {
      uint32_t v[PIECE];
      tc::tmem_ld(tmem_mine + 2 * KF + b * KF + pc * PIECE, v);
      if constexpr (EXPECT) {
        uint32_t xh[PIECE], xl[PIECE];
        tc::tmem_ld(tmem_mine + pc * PIECE, xh);
        tc::tmem_ld(tmem_mine + KF + pc * PIECE, xl);
#pragma unroll
        for (int j = 0; j < PIECE / 2; ++j) {
          const float xr = __uint_as_float(xh[2 * j]) + __uint_as_float(xl[2 * j]);
          const float xi = __uint_as_float(xh[2 * j + 1]) + __uint_as_float(xl[2 * j + 1]);
          const float re = __uint_as_float(v[2 * j]), im = __uint_as_float(v[2 * j + 1]);
          ere += xr * re + xi * im;
          eim += xr * im - xi * re;
        }
      } else {
        // accumulation-bias compensation (tc_bias): the tensor core truncates its fp32 accumulation, a pass
        // shrinks the state by a factor ~(1 - comp); v + comp * v, one rounding, puts the mean back.  (A fourth
        // MMA A_hi (comp W_hi) did the same at 25 % more tensor time and a third 64 KB W tile.)
#pragma unroll
        for (int j = 0; j < PIECE; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), comp, __uint_as_float(v[j])));
#pragma unroll
        for (int j = 0; j < PIECE / 2; j += (PAIR ? 2 : 1)) {
          const int a = pc * (PIECE / 2) + j;  // amplitude within the half row
          if constexpr (PAIR) {
            *reinterpret_cast<uint4*>(p + eo(a)) = make_uint4(v[2 * j], v[2 * j + 1], v[2 * j + 2], v[2 * j + 3]);
          } else {
            *reinterpret_cast<uint2*>(p + eo(a)) = make_uint2(v[2 * j], v[2 * j + 1]);
          }
        }
      }
    }
  };

  uint4 cur[CHUNKS], nxt[CHUNKS];
  unsigned char *p_cur = nullptr, *p_nxt = nullptr, *p_prev = nullptr;
  uint64_t tile = blockIdx.x;
  if (tile < ntiles) { p_cur = tile_ptr(tile); load_mine(p_cur, cur); }
  uint32_t it = 0;
  for (; tile < ntiles; tile += stride, ++it) {
    if constexpr (EXPECT) {
      // one accumulator buffer: the epilogue re-reads x from the A columns, which the next split overwrites
      if (it > 0) {
        tc::mbar_wait(smem_u32(&mbar), (it - 1) & 1);
        tc::fence_after();
        epilogue(p_prev, 0);
      }
    } else if (it > 0) {
      // tile it-1's MMAs are complete: the A columns are free, D[(it-1) & 1] is ready -- read out BELOW, after this
      // tile's MMAs have been issued into the other accumulator buffer
      tc::mbar_wait(smem_u32(&mbar), (it - 1) & 1);
      tc::fence_after();
    }
    split_to_tmem(cur);
    if (tile + stride < ntiles) { p_nxt = tile_ptr(tile + stride); load_mine(p_nxt, nxt); }
    tc::fence_before();
    tc::mbar_arrive(smem_u32(&full_bar));
    if (warp == 0) {
      tc::mbar_wait(smem_u32(&full_bar), it & 1);
      if (tc::elect_one()) issue_mmas(EXPECT ? 0u : (it & 1u));
      __syncwarp();
    }
    if constexpr (!EXPECT) {
      if (it > 0) epilogue(p_prev, (it - 1) & 1);   // overlaps the MMAs just issued
    }
    p_prev = p_cur;
    p_cur = p_nxt;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) cur[c] = nxt[c];
  }
  if (it > 0) {
    tc::mbar_wait(smem_u32(&mbar), (it - 1) & 1);
    tc::fence_after();
    epilogue(p_prev, EXPECT ? 0u : ((it - 1) & 1u));
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
  }
  if constexpr (EXPECT) {
    block_sum2<kTcThreads>(ere, eim);
    if (t == 0) {
      partials[2 * blockIdx.x] = ere;
      partials[2 * blockIdx.x + 1] = eim;
    }
  }
}

}  // namespace qb200
