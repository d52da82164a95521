// p2p_swap.cu -- local<->global qubit swap as ONE kernel over NVLink peer memory.
// See the block comment in include/qsim_b200.h.  HBM/NVLink-bound byte work:
// 128-bit accesses, 4 independent remote loads in flight per thread.
#include <cstring>

#include "common.cuh"

namespace qb200 {

constexpr int kMaxSwapBits = 3;  // 8 GPUs

struct SwapGeom {
  void* peers[1 << kMaxSwapBits];
  uint64_t items_per_peer;   // vector items this rank swaps with each peer (half a slice)
  uint64_t slice_items;      // vector items per slice
  uint32_t k;
  uint32_t my;
  uint32_t vshift;           // log2(amplitudes per vector item)
  uint32_t cshift;           // log2(items per peer-interleave chunk)
  uint32_t lbits[kMaxSwapBits];
};

template <typename V>
__global__ void __launch_bounds__(256)
k_p2p_swap(V* __restrict__ local, const __grid_constant__ SwapGeom g) {
  constexpr int U = 4;
  const uint64_t per_peer = g.items_per_peer;
  const uint32_t npeers = (1u << g.k) - 1;
  const uint64_t total = per_peer * npeers;
  const uint64_t stride = uint64_t{gridDim.x} * blockDim.x;
  for (uint64_t t0 = blockIdx.x * uint64_t{blockDim.x} + threadIdx.x; t0 < total; t0 += stride * U) {
    V x[U], y[U];
    uint64_t il[U], ir[U];
    V* rp[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t t = t0 + u * stride;
      ok[u] = t < total;
      if (!ok[u]) continue;
      // Peers are interleaved at chunk granularity (64 KB): at any moment every GPU exchanges with
      // ALL its 2^k - 1 peers, so ingress and egress of every GPU stay balanced.  (Walking the
      // peers one after the other makes several ranks hit the same peer at once: 216 GB/s per
      // direction at k = 2 against 658 GB/s at k = 1.)
      const uint64_t q = t >> g.cshift;
      const uint32_t pi = (uint32_t) (q % npeers);
      const uint32_t b = pi < g.my ? pi : pi + 1;
      // the rank with the smaller value swaps the first half of the pair's slice
      uint64_t e = ((q / npeers) << g.cshift) + (t & ((uint64_t{1} << g.cshift) - 1)) + (g.my < b ? 0 : per_peer);
      // vector item -> amplitude index with zero bits inserted at the swapped local bits
      uint64_t idx = e << g.vshift;
      uint64_t lb = 0, rb = 0;
      for (uint32_t j = 0; j < g.k; ++j) {
        const uint64_t lo = idx & ((uint64_t{1} << g.lbits[j]) - 1);
        idx = ((idx - lo) << 1) | lo;
        lb |= (uint64_t) ((b >> j) & 1) << g.lbits[j];
        rb |= (uint64_t) ((g.my >> j) & 1) << g.lbits[j];
      }
      il[u] = (idx | lb) >> g.vshift;
      ir[u] = (idx | rb) >> g.vshift;
      rp[u] = reinterpret_cast<V*>(g.peers[b]);
      y[u] = rp[u][ir[u]];   // remote read over NVLink
      x[u] = local[il[u]];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      rp[u][ir[u]] = x[u];   // remote write over NVLink
      local[il[u]] = y[u];
    }
  }
}

}  // namespace qb200

using namespace qb200;

extern "C" {

int qb200_ipc_export(const void* state, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!state || !handle) return QB200_ERR_INVALID;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, const_cast<void*>(state)) != cudaSuccess) {
    (void) cudaGetLastError();
    return QB200_ERR_CUDA;
  }
  memcpy(handle, &h, 64);
  return QB200_OK;
}

int qb200_ipc_import(const unsigned char handle[64], void** peer_state) {
  if (!handle || !peer_state) return QB200_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  if (cudaIpcOpenMemHandle(peer_state, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    (void) cudaGetLastError();
    *peer_state = nullptr;
    return QB200_ERR_CUDA;
  }
  return QB200_OK;
}

int qb200_ipc_close(void* peer_state) {
  if (!peer_state) return QB200_OK;
  if (cudaIpcCloseMemHandle(peer_state) != cudaSuccess) {
    (void) cudaGetLastError();
    return QB200_ERR_CUDA;
  }
  return QB200_OK;
}

int qb200_swap_global_local(qb200_ctx* ctx, int dtype, void* state, unsigned n_local,
                            void* const* peer_states, unsigned k, const unsigned* local_bits,
                            unsigned my_value) {
  if (!ctx || !state || !peer_states || !local_bits || k < 1 || k > kMaxSwapBits) return QB200_ERR_INVALID;
  note_state_written();
  if (dtype != QB200_F32 && dtype != QB200_F64) return QB200_ERR_INVALID;
  if (n_local < k + 2 || n_local > kMaxQubits || my_value >= (1u << k)) return QB200_ERR_INVALID;
  SwapGeom g{};
  for (unsigned j = 0; j < k; ++j) {
    if (local_bits[j] >= n_local || (j > 0 && local_bits[j] <= local_bits[j - 1])) return QB200_ERR_INVALID;
    g.lbits[j] = local_bits[j];
  }
  for (unsigned b = 0; b < (1u << k); ++b) {
    if (b != my_value && !peer_states[b]) return QB200_ERR_INVALID;
    g.peers[b] = peer_states[b];
  }
  g.k = k;
  g.my = my_value;
  // 16-byte vector items: two fp32 amplitudes (needs local bit 0 untouched) or one fp64 amplitude
  const bool vec2 = dtype == QB200_F32 && local_bits[0] >= 1;
  g.vshift = vec2 ? 1 : 0;
  const uint64_t slice_amps = uint64_t{1} << (n_local - k);
  g.slice_items = slice_amps >> g.vshift;
  g.items_per_peer = g.slice_items / 2;
  g.cshift = 0;
  while (g.cshift < 12 && (uint64_t{2} << g.cshift) <= g.items_per_peer) ++g.cshift;  // <= 4096 items, divides items_per_peer
  DeviceGuard guard(ctx);
  const uint64_t total = g.items_per_peer * ((1u << k) - 1);
  uint64_t blocks = (total + 256 * 4 - 1) / (256 * 4);
  if (blocks > uint64_t{kNumSMs} * 16) blocks = uint64_t{kNumSMs} * 16;
  if (blocks < 1) blocks = 1;
  if (dtype == QB200_F32 && vec2)
    k_p2p_swap<float4><<<(uint32_t) blocks, 256, 0, ctx->stream>>>((float4*) state, g);
  else if (dtype == QB200_F32)
    k_p2p_swap<float2><<<(uint32_t) blocks, 256, 0, ctx->stream>>>((float2*) state, g);
  else
    k_p2p_swap<double2><<<(uint32_t) blocks, 256, 0, ctx->stream>>>((double2*) state, g);
  QB_LAUNCHED(ctx);
  return QB200_OK;
}

}  // extern "C"
