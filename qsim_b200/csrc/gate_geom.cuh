// gate_geom.cuh -- addressing of one fused-gate pass.
//
// A pass over an n-qubit state with target set T (|T| = G) and control set C
// visits 2^(n-G-|C|) *groups*; group i owns the 2^G amplitudes
//     base(i) + sum_j bit_j(k) << T[j],   k in [0, 2^G)
// where base(i) spreads the bits of i over the positions outside T u C and
// carries the required control values at the positions in C.  This is the same
// index space as the reference's ms[]/xss[] tables (lib/simulator_cuda.h:
// 640-740, lib/simulator.h:40-66) but is evaluated arithmetically from a
// 64-byte kernel parameter instead of per-call uploaded tables.
#pragma once

#include "common.cuh"

namespace qb200 {

struct Geom {
  uint64_t work;     // number of work items (one per thread-iteration)
  uint64_t cbits;    // control values deposited at the control positions
  uint64_t xs[6];    // amplitude stride of target j (1 << T[j])
  uint32_t npos;     // zero bits to insert into the work index ...
  uint8_t pos[43];   // ... at these bit positions, ascending
  uint8_t nt;        // number of targets (runtime-generic kernel only)
};

__device__ __forceinline__ uint64_t expand_index(uint64_t i, const Geom& g) {
  for (uint32_t k = 0; k < g.npos; ++k) {
    const uint64_t lo = i & ((uint64_t{1} << g.pos[k]) - 1);
    i = ((i - lo) << 1) | lo;
  }
  return i | g.cbits;
}

// offset of group element k (compile-time k in the unrolled kernels)
template <int G>
__device__ __forceinline__ uint64_t elem_offset(int k, const Geom& g) {
  uint64_t o = 0;
#pragma unroll
  for (int j = 0; j < G; ++j)
    if ((k >> j) & 1) o += g.xs[j];
  return o;
}

// Host side: validates a (targets, controls) pair and fills the geometry.
// `extra_low` = number of lowest index bits kept *inside* a work item
// (1 when a thread moves two adjacent amplitudes with one 128-bit access and
// bit 0 is neither a target nor a control).
// Returns QB200_OK / QB200_ERR_INVALID.
inline int make_geom(unsigned n, const unsigned* qs, unsigned nq, const unsigned* cqs,
                     unsigned nc, uint64_t cvals, bool reserve_bit0, Geom* g) {
  if (n > kMaxQubits || nq > kMaxTargets || nq + nc > n) return QB200_ERR_INVALID;
  uint64_t tmask = 0, cmask = 0;
  for (unsigned j = 0; j < nq; ++j) {
    if (qs[j] >= n || ((tmask >> qs[j]) & 1)) return QB200_ERR_INVALID;
    if (j > 0 && qs[j] < qs[j - 1]) return QB200_ERR_INVALID;  // must be sorted
    tmask |= uint64_t{1} << qs[j];
  }
  for (unsigned j = 0; j < nc; ++j) {
    if (cqs[j] >= n || (((tmask | cmask) >> cqs[j]) & 1)) return QB200_ERR_INVALID;
    cmask |= uint64_t{1} << cqs[j];
  }
  uint64_t special = tmask | cmask;
  if (reserve_bit0) {
    if (special & 1) return QB200_ERR_INVALID;
    special |= 1;
  }
  // control values: bit i of cvals goes to the i-th lowest control position
  // (bits::ExpandBits, lib/bits.h:53-66; lib/simulator.h:364-375)
  uint64_t cbits = 0;
  unsigned k = 0;
  for (unsigned b = 0; b < n; ++b) {
    if ((cmask >> b) & 1) {
      cbits |= ((cvals >> k) & 1) << b;
      ++k;
    }
  }
  g->cbits = cbits;
  g->npos = 0;
  for (unsigned b = 0; b < n; ++b)
    if ((special >> b) & 1) g->pos[g->npos++] = (uint8_t) b;
  for (unsigned j = 0; j < 6; ++j) g->xs[j] = j < nq ? uint64_t{1} << qs[j] : 0;
  g->nt = (uint8_t) nq;
  g->work = uint64_t{1} << (n - g->npos);
  return QB200_OK;
}

}  // namespace qb200
