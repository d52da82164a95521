// sv_plan.h -- look-ahead remap pass for a state sharded by global qubits (host-only C++).
//
// Input: the fused-gate list of a circuit (per op: the set of qubits it touches, targets and controls alike),
// the number of global qubits g and the current global set.  Output: a schedule = the ops, possibly REORDERED,
// with local<->global swaps in between, such that every op runs while all its qubits are local.
//
// The reference reaches this step only through the closed cuStateVecEx scheduler (custatevecExSVUpdaterEnqueue* /
// Apply, lib/run_custatevecex.h:243-305); the policy here is ours:
//  * ops on disjoint qubits commute, so an epoch (the stretch between two swaps) executes EVERY pending op that
//    is not blocked -- an op is blocked when it touches a global qubit or a qubit of an earlier blocked op
//    (per-qubit program order is preserved, nothing else is);
//  * when nothing is executable the next global set is chosen by a small depth-first search over "which g qubits
//    are global next" minimising the bytes exchanged, sum over swaps of (1 - 2^-k) shards, k = qubits that change
//    sides; every candidate set is tried when C(n, g) is small (8 GPUs, 37 qubits: 7770 sets), otherwise the
//    Belady choice (furthest next use);
//  * reorder = false keeps program order (an epoch ends at the first blocked op): the round-1 planner.
// On the depth-20 RQCs of BASELINE configs 2-4 this needs 1-2 swaps (0.5 / 0.75 / 1.375 shards at 2 / 4 / 8 GPUs)
// where the in-order Belady planner needs 3-4 (1.5 / 3.0 / 3.5 shards).
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

namespace qb200 {

struct PlanStep {
  bool is_swap = false;
  uint32_t op = 0;                       // gate step: index into the input list
  std::vector<unsigned> victims;         // swap step: local qubits that become global ...
  std::vector<unsigned> incoming;        // ... and global qubits that become local (same length)
};

class SwapPlanner {
 public:
  // touch[i]: every qubit op i acts on (targets and controls) -- what orders it against other ops;
  // need[i]: the qubits that must be LOCAL when it runs (its targets: a control on a global qubit only decides
  // per shard whether the shard takes part).
  SwapPlanner(unsigned n, unsigned g, const std::vector<uint64_t>& touch, const std::vector<uint64_t>& need,
              uint64_t glob, bool reorder)
      : n_(n), g_(g), ops_(touch), need_(need), reorder_(reorder), glob_(glob), done_(touch.size(), 0) {}

  // Exchanged shards of a schedule: sum over its swaps of (1 - 2^-k).
  static double Cost(const std::vector<PlanStep>& plan) {
    double c = 0;
    for (const auto& s : plan)
      if (s.is_swap) c += 1.0 - 1.0 / (double) (uint64_t{1} << s.victims.size());
    return c;
  }

  // The best INITIAL global set when the caller may choose it -- the state is |0...0> (or uniform), which looks the
  // same under every qubit map, so the map can be picked for the circuit before the first gate at no cost (the "wire
  // ordering" freedom of lib/vectorspace_custatevecex.h:163-177).  Every g-subset is rated by the ops its first epoch
  // executes; the best few (and the default) are planned in full and the cheapest schedule wins, the default on ties.
  // 33 qubits on 8 GPUs (rqc depth 20): {30, 31, 32} needs exchanges of 3 + 1 qubits (1.375 shards), the best set one
  // exchange of 3 (0.875).
  static uint64_t BestInitial(unsigned n, unsigned g, const std::vector<uint64_t>& touch, const std::vector<uint64_t>& need,
                              uint64_t default_glob, bool reorder) {
    if (g == 0 || Binomial(n, g) > kMaxSets) return default_glob;
    SwapPlanner probe(n, g, touch, need, default_glob, reorder);
    std::vector<std::pair<size_t, uint64_t>> rated;
    std::vector<uint32_t> ex;
    uint64_t s = (uint64_t{1} << g) - 1;
    const uint64_t limit = uint64_t{1} << n;
    while (s < limit) {
      ex.clear();
      probe.Closure(probe.done_, 0, s, &ex);
      rated.emplace_back(ex.size(), s);
      const uint64_t c = s & (~s + 1), r = s + c;
      s = (((r ^ s) >> 2) / c) | r;
    }
    std::stable_sort(rated.begin(), rated.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    if (rated.size() > kInitialCandidates) rated.resize(kInitialCandidates);
    uint64_t best = default_glob;
    double best_cost = Cost(SwapPlanner(n, g, touch, need, default_glob, reorder).Run());
    for (const auto& cand : rated) {
      const double c = Cost(SwapPlanner(n, g, touch, need, cand.second, reorder).Run());
      if (c < best_cost - 1e-12) {
        best_cost = c;
        best = cand.second;
      }
    }
    return best;
  }

  std::vector<PlanStep> Run() {
    std::vector<PlanStep> out;
    std::vector<uint32_t> ex;
    while (true) {
      ex.clear();
      Closure(done_, lo_, glob_, &ex);
      for (uint32_t i : ex) {
        done_[i] = 1;
        PlanStep s;
        s.op = i;
        out.push_back(std::move(s));
      }
      while (lo_ < ops_.size() && done_[lo_]) ++lo_;
      if (lo_ >= ops_.size()) break;
      if (!ex.empty()) continue;  // the window moved: look again before paying for a swap
      const uint64_t next = Choose();
      PlanStep s;
      s.is_swap = true;
      for (unsigned q = 0; q < n_; ++q) {
        const bool was = (glob_ >> q) & 1, is = (next >> q) & 1;
        if (is && !was) s.victims.push_back(q);
        if (was && !is) s.incoming.push_back(q);
      }
      glob_ = next;
      out.push_back(std::move(s));
    }
    return out;
  }

 private:
  static constexpr size_t kWindow = 512;       // ops considered per epoch (keeps the search O(window))
  static constexpr uint64_t kMaxSets = 200000;  // exhaustive candidate enumeration up to this many global sets

  // Ops of [lo, lo + window) executable with global set `glob`, in index order.
  void Closure(const std::vector<char>& done, size_t lo, uint64_t glob, std::vector<uint32_t>* ex) const {
    uint64_t blocked = 0;
    const size_t hi = std::min(ops_.size(), lo + kWindow);
    for (size_t i = lo; i < hi; ++i) {
      if (done[i]) continue;
      if ((need_[i] & glob) | (ops_[i] & blocked)) {
        if (!reorder_) return;
        blocked |= ops_[i];
      } else {
        ex->push_back((uint32_t) i);
      }
    }
  }

  static uint64_t Binomial(unsigned n, unsigned k) {
    uint64_t r = 1;
    for (unsigned i = 1; i <= k; ++i) {
      r = r * (n - k + i) / i;
      if (r > (uint64_t{1} << 40)) return r;
    }
    return r;
  }

  // g qubits with the furthest next use among the pending ops (never a qubit of the first pending op).
  uint64_t Belady(const std::vector<char>& done, size_t lo, uint64_t glob) const {
    std::vector<size_t> next(n_, ops_.size() + 1);
    size_t first = ops_.size();
    for (size_t i = ops_.size(); i-- > lo;) {
      if (done[i]) continue;
      first = i;
      for (unsigned q = 0; q < n_; ++q)
        if ((ops_[i] >> q) & 1) next[q] = i;
    }
    std::vector<unsigned> cand;
    for (unsigned q = 0; q < n_; ++q)
      if (first >= ops_.size() || !((need_[first] >> q) & 1)) cand.push_back(q);
    std::stable_sort(cand.begin(), cand.end(), [&](unsigned a, unsigned b) {
      if (next[a] != next[b]) return next[a] > next[b];
      return ((glob >> a) & 1) > ((glob >> b) & 1);  // ties: keep what is global already
    });
    uint64_t m = 0;
    for (unsigned i = 0; i < g_ && i < cand.size(); ++i) m |= uint64_t{1} << cand[i];
    return m;
  }

  struct Cand {
    uint64_t set;
    unsigned k;
    std::vector<uint32_t> ex;
  };

  void Candidates(const std::vector<char>& done, size_t lo, uint64_t glob, size_t beam, std::vector<Cand>* out) const {
    out->clear();
    auto consider = [&](uint64_t set) {
      const unsigned k = (unsigned) __builtin_popcountll(set & ~glob);
      if (k == 0) return;
      Cand c{set, k, {}};
      Closure(done, lo, set, &c.ex);
      if (c.ex.empty()) return;
      out->push_back(std::move(c));
      if (out->size() > 4 * beam + 16) Prune(out, beam);
    };
    if (Binomial(n_, g_) <= kMaxSets) {
      // Gosper's hack over all g-subsets of n qubits
      uint64_t s = (uint64_t{1} << g_) - 1;
      const uint64_t limit = uint64_t{1} << n_;
      while (s < limit) {
        consider(s);
        const uint64_t c = s & (~s + 1), r = s + c;
        s = (((r ^ s) >> 2) / c) | r;
      }
    } else {
      consider(Belady(done, lo, glob));
    }
    Prune(out, beam);
  }

  static void Prune(std::vector<Cand>* v, size_t beam) {
    std::stable_sort(v->begin(), v->end(), [](const Cand& a, const Cand& b) {
      // most ops per exchanged byte first; then more ops; then fewer moved qubits
      const double ra = a.ex.size() / (1.0 - 1.0 / (1u << a.k)), rb = b.ex.size() / (1.0 - 1.0 / (1u << b.k));
      if (ra != rb) return ra > rb;
      if (a.ex.size() != b.ex.size()) return a.ex.size() > b.ex.size();
      if (a.k != b.k) return a.k < b.k;
      // equal otherwise: prefer HIGH qubits as the new global ones -- victims on high index bits travel as long
      // contiguous runs (no sorting inside the push kernel's tiles; the copy engines can take them)
      return a.set > b.set;
    });
    if (v->size() > beam) v->resize(beam);
  }

  // cheapest continuation from (done, glob): returns cost, writes the first global set to *first
  double Search(std::vector<char>& done, size_t lo, uint64_t glob, int depth, uint64_t* first) const {
    while (lo < ops_.size() && done[lo]) ++lo;
    size_t pending = 0;
    for (size_t i = lo; i < std::min(ops_.size(), lo + kWindow); ++i) pending += !done[i];
    if (pending == 0) return 0.0;
    if (depth == 0) return 0.5 + 0.01 * (double) pending;
    std::vector<Cand> cands;
    Candidates(done, lo, glob, depth >= 2 ? kBeam : 1, &cands);
    double best = 1e30;
    for (const Cand& c : cands) {
      for (uint32_t i : c.ex) done[i] = 1;
      uint64_t dummy = 0;
      const double cost = (1.0 - 1.0 / (1u << c.k)) + Search(done, lo, c.set, depth - 1, &dummy);
      for (uint32_t i : c.ex) done[i] = 0;
      if (cost < best - 1e-12) {
        best = cost;
        *first = c.set;
      }
    }
    if (cands.empty()) {  // cannot happen for a gate that fits a shard; keep the planner total anyway
      *first = Belady(done, lo, glob);
      return 1e6;
    }
    return best;
  }

  uint64_t Choose() {
    uint64_t first = 0;
    std::vector<char> scratch = done_;
    Search(scratch, lo_, glob_, kDepth, &first);
    return first;
  }

  static constexpr size_t kBeam = 6;
  static constexpr int kDepth = 3;
  static constexpr size_t kInitialCandidates = 12;

  unsigned n_, g_;
  std::vector<uint64_t> ops_, need_;
  bool reorder_;
  uint64_t glob_;
  std::vector<char> done_;
  size_t lo_ = 0;
};

}  // namespace qb200
