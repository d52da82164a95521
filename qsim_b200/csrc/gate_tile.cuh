// gate_tile.cuh -- warp-cooperative fused-gate kernel (fp32, G = 4; G = 5, 6 in gate_big.cuh) for ANY
// target layout, including the lowest index bits.
//
// Tile.  A warp owns 32 groups whose 2^(G+5) amplitudes form the index set
//   base(tile) + deposit(j, B),  j in [0, 2^(G+5)),  B = targets U {5 lowest free bits}
// Because B always contains the lowest free bits, the tile is made of contiguous runs
// of >= 256 bytes (512 bytes or more whenever a target sits below bit 5), so the warp
// moves it with fully coalesced 16-byte-per-lane cp.async / STG.128 requests no matter
// where the targets are -- the job the reference does with its separate "L" kernels and
// per-call lane tables (lib/simulator_cuda_kernels.h:113-202).
//
// Shared memory.  The tile lives in a per-warp D-deep ring (deep prefetch, see
// gate_pipe.cuh) at swz(j) = j ^ X(j): X XORs free tile bits >= 4 into the target bits
// among {1,2,3}, chosen on the host so that (a) the lane-per-group reads (LDS.64, or
// LDS.128 pairs when bit 0 is a target) and (b) the cooperative 16-byte chunk accesses are
// both bank-conflict free.  X is linear over GF(2): every address is
// (per-lane constant) ^ (uniform constant), one LOP3 per access.
//
// Only warp-level synchronisation is used (__syncwarp + cp.async.wait_group).
#pragma once

#include "gate_pipe.cuh"

namespace qb200 {

struct TileGeom {
  uint64_t work;        // number of warp tiles
  uint64_t cbits;       // control values at control positions
  uint64_t goff_m[32];  // global amplitude offset of chunk-row m (lane-independent part)
  uint32_t sm_m[32];    // swizzled BYTE offset of chunk-row m inside the tile
  uint32_t skb[64];     // swizzled BYTE offset of group element k
  uint32_t npos;        // zero bits inserted to form the tile base ...
  uint8_t pos[44];      // ... at these global positions (B and controls), ascending
  uint8_t bpos[12];     // global bit of tile-local bit i
  uint8_t fl[5];        // tile-local positions of the 5 free (lane) bits
  uint8_t nsw;          // swizzle terms
  uint8_t sw_src[3], sw_dst[3];
  uint8_t pair;         // 1: tile-local bit 0 is a target (elements k, k+1 adjacent)
};

__device__ __forceinline__ uint32_t tile_swz(uint32_t j, const TileGeom& t) {
  uint32_t x = j;
  for (int i = 0; i < t.nsw; ++i) x ^= ((j >> t.sw_src[i]) & 1u) << t.sw_dst[i];
  return x;
}

template <int G, bool PAIR, int NT, int D, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_gate_tile(float* __restrict__ st, const __grid_constant__ TileGeom t,
            const __grid_constant__ MatParam<float, G> mat) {
  constexpr int N = 1 << G;
  constexpr int TILE_BYTES = 8 << (G + 5);
  constexpr int ROWS = N / 2;  // 16-byte chunk rows per lane
  constexpr int WARPS = NT / 32;
  using C = CT<float>::type;
  extern __shared__ __align__(1024) unsigned char ring_raw[];  // [WARPS][D][TILE_BYTES] + pad
  __shared__ uint64_t base_ring[WARPS][D];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // tiles must be TILE_BYTES-aligned in the shared window: addresses are composed with | and ^
  const uint32_t pad = (TILE_BYTES - (smem_u32(ring_raw) & (TILE_BYTES - 1))) & (TILE_BYTES - 1);
  unsigned char* const wring = ring_raw + pad + (size_t) warp * D * TILE_BYTES;
  const uint32_t wring_s = smem_u32(wring);

  // per-lane constants
  uint64_t goff_l = 0;
  uint32_t jl = (uint32_t) lane << 1, jw = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    goff_l |= (uint64_t) ((lane >> i) & 1) << t.bpos[i + 1];
    jw |= (uint32_t) ((lane >> i) & 1) << t.fl[i];
  }
  const uint32_t chunk_l = tile_swz(jl, t) << 3;  // byte offset of this lane's chunk column
  const uint32_t group_l = tile_swz(jw, t) << 3;  // byte offset of this lane's group

  const uint64_t stride = uint64_t{gridDim.x} * WARPS;
  const uint64_t first = blockIdx.x * uint64_t{WARPS} + warp;

  auto tile_base = [&](uint64_t i) {
    for (uint32_t k = 0; k < t.npos; ++k) {
      const uint64_t lo = i & ((uint64_t{1} << t.pos[k]) - 1);
      i = ((i - lo) << 1) | lo;
    }
    return i | t.cbits;
  };

  auto issue = [&](uint64_t i, int s) {
    if (i < t.work) {
      const uint64_t base = tile_base(i);
      if (lane == 0) base_ring[warp][s] = base;
      const float* p = st + 2 * (base + goff_l);
      const uint32_t dst = (wring_s + s * TILE_BYTES) | chunk_l;
#pragma unroll
      for (int m = 0; m < ROWS; ++m) cp_async16(dst ^ t.sm_m[m], p + 2 * t.goff_m[m]);
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < D - 1; ++s) issue(first + s * stride, s);

  int s = 0;
  for (uint64_t i = first; i < t.work; i += stride) {
    int sp = s + D - 1;
    if (sp >= D) sp -= D;
    __syncwarp();  // every lane is done reading stage sp (consumed last iteration)
    issue(i + (D - 1) * stride, sp);
    cp_async_wait<D - 1>();
    __syncwarp();  // all lanes' copies of stage s have landed

    unsigned char* const tile = wring + s * TILE_BYTES;
    const uint32_t ga = group_l;
    C x[N], ix[N];
    if constexpr (PAIR) {
#pragma unroll
      for (int k = 0; k < N; k += 2) {
        const uint4 w = *reinterpret_cast<const uint4*>(tile + (ga ^ t.skb[k]));
        x[k] = (uint64_t) w.x | ((uint64_t) w.y << 32);
        x[k + 1] = (uint64_t) w.z | ((uint64_t) w.w << 32);
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) x[k] = *reinterpret_cast<const uint64_t*>(tile + (ga ^ t.skb[k]));
    }
#pragma unroll
    for (int k = 0; k < N; ++k) ix[k] = CT<float>::rot(x[k]);

    // in-place mat-vec: every lane owns its group's slots, results overwrite them
    if constexpr (PAIR) {
#pragma unroll
      for (int r = 0; r < N; r += 2) {
        if (r % kRowBatch == 0 && r > 0) CT<float>::fence(x[0]);
        const C a = row_dot<float, G>(x, ix, mat, r);
        const C b = row_dot<float, G>(x, ix, mat, r + 1);
        *reinterpret_cast<uint4*>(tile + (ga ^ t.skb[r])) =
            make_uint4((uint32_t) a, (uint32_t) (a >> 32), (uint32_t) b, (uint32_t) (b >> 32));
      }
    } else {
#pragma unroll
      for (int r = 0; r < N; ++r) {
        if (r % kRowBatch == 0 && r > 0) CT<float>::fence(x[0]);
        *reinterpret_cast<uint64_t*>(tile + (ga ^ t.skb[r])) = row_dot<float, G>(x, ix, mat, r);
      }
    }
    __syncwarp();

    // coalesced write-back: 16 bytes per lane, 512 contiguous bytes per request
    float* const p = st + 2 * (base_ring[warp][s] + goff_l);
#pragma unroll
    for (int m = 0; m < ROWS; ++m) {
      const uint4 v = *reinterpret_cast<const uint4*>(tile + (chunk_l ^ t.sm_m[m]));
      *reinterpret_cast<uint4*>(p + 2 * t.goff_m[m]) = v;
    }
    if (++s == D) s = 0;
  }
  cp_async_wait<0>();
}

template <int G, int NT, int D>
constexpr size_t tile_smem_bytes() { return (size_t) (NT / 32) * D * (8 << (G + 5)) + (8 << (G + 5)); }

// Host: builds the tile geometry.  Returns QB200_ERR_UNSUPPORTED when the layout cannot
// use the tile kernel (too few free bits, or bit 0 is a control).
inline int make_tile_geom(unsigned n, const unsigned* qs, unsigned nq, const unsigned* cqs,
                          unsigned nc, uint64_t cvals, TileGeom* t) {
  if (n > kMaxQubits || nq > 6 || nq < 1 || nq + nc + 5 > n) return QB200_ERR_UNSUPPORTED;
  uint64_t tmask = 0, cmask = 0;
  for (unsigned j = 0; j < nq; ++j) {
    if (qs[j] >= n || ((tmask >> qs[j]) & 1) || (j > 0 && qs[j] < qs[j - 1])) return QB200_ERR_INVALID;
    tmask |= uint64_t{1} << qs[j];
  }
  for (unsigned j = 0; j < nc; ++j) {
    if (cqs[j] >= n || (((tmask | cmask) >> cqs[j]) & 1)) return QB200_ERR_INVALID;
    cmask |= uint64_t{1} << cqs[j];
  }
  if (cmask & 1) return QB200_ERR_UNSUPPORTED;
  // the 5 lowest free bits
  uint64_t fmask = 0;
  unsigned nf = 0;
  for (unsigned b = 0; b < n && nf < 5; ++b)
    if (!(((tmask | cmask) >> b) & 1)) { fmask |= uint64_t{1} << b; ++nf; }
  const uint64_t bmask = tmask | fmask;
  const unsigned nb = nq + 5;
  unsigned li = 0, tl[6], ntl = 0, nfl = 0;
  for (unsigned b = 0; b < n; ++b) {
    if (!((bmask >> b) & 1)) continue;
    t->bpos[li] = (uint8_t) b;
    if ((tmask >> b) & 1) tl[ntl++] = li; else t->fl[nfl++] = (uint8_t) li;
    ++li;
  }
  if (t->bpos[0] != 0) return QB200_ERR_UNSUPPORTED;  // bit 0 must be inside the tile
  t->pair = (tmask & 1) ? 1 : 0;
  // swizzle: target tile bits among {1,2,3} <- free tile bits >= 4 (lowest first)
  t->nsw = 0;
  unsigned src_i = 0;
  for (unsigned d = 1; d <= 3; ++d) {
    bool is_target = false;
    for (unsigned k = 0; k < ntl; ++k) is_target |= tl[k] == d;
    if (!is_target) continue;
    while (src_i < 5 && t->fl[src_i] < 4) ++src_i;
    if (src_i >= 5) break;
    t->sw_src[t->nsw] = t->fl[src_i++];
    t->sw_dst[t->nsw] = (uint8_t) d;
    ++t->nsw;
  }
  auto swz = [&](uint32_t j) {
    uint32_t x = j;
    for (int i = 0; i < t->nsw; ++i) x ^= ((j >> t->sw_src[i]) & 1u) << t->sw_dst[i];
    return x;
  };
  auto deposit = [&](uint32_t j) {  // tile-local index -> global amplitude offset
    uint64_t o = 0;
    for (unsigned i = 0; i < nb; ++i) o |= (uint64_t) ((j >> i) & 1) << t->bpos[i];
    return o;
  };
  const unsigned rows = 1u << (nq - 1);
  for (unsigned m = 0; m < 32; ++m) {
    const uint32_t j = m << 6;
    t->goff_m[m] = m < rows ? deposit(j) : 0;
    t->sm_m[m] = m < rows ? swz(j) << 3 : 0;
  }
  for (unsigned k = 0; k < 64; ++k) {
    uint32_t j = 0;
    for (unsigned b = 0; b < nq; ++b) j |= ((k >> b) & 1u) << tl[b];
    t->skb[k] = k < (1u << nq) ? swz(j) << 3 : 0;
  }
  // tile base: zero bits at B and control positions, control values ORed in
  uint64_t cbits = 0;
  unsigned kk = 0;
  for (unsigned b = 0; b < n; ++b)
    if ((cmask >> b) & 1) { cbits |= ((cvals >> kk) & 1) << b; ++kk; }
  t->cbits = cbits;
  t->npos = 0;
  for (unsigned b = 0; b < n; ++b)
    if (((bmask | cmask) >> b) & 1) t->pos[t->npos++] = (uint8_t) b;
  t->work = uint64_t{1} << (n - t->npos);
  return QB200_OK;
}

}  // namespace qb200
