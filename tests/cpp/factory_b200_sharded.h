// factory_b200_sharded.h -- Factory for the reference fixtures on a SHARDED state.  QB200_TEST_SHARDS (default 4)
// shards; when the box has fewer GPUs than that, several shards share a device (the exchange kernels then move
// data between allocations of one GPU), so a one-GPU box runs the multi-shard path too.
#ifndef TESTS_CPP_FACTORY_B200_SHARDED_H_
#define TESTS_CPP_FACTORY_B200_SHARDED_H_

#include <cstdlib>

#include "qsim_b200/simulator_b200_sharded.h"

namespace qsim {

inline b200::ShardedParameter TestShardedParameter() {
  b200::ShardedParameter p;
  const char* e = std::getenv("QB200_TEST_SHARDS");
  const int shards = e ? std::atoi(e) : 4;
  int count = 1;
  qb200_device_count(&count);
  if (count < 1) count = 1;
  for (int r = 0; r < shards; ++r) p.devices.push_back(r % count);
  if (const char* m = std::getenv("QB200_TEST_SWAP_MODE")) p.swap_mode = std::atoi(m);
  return p;
}

template <typename FP>
struct Factory {
  using fp_type = FP;
  using Simulator = qsim::SimulatorB200Sharded<fp_type>;
  using StateSpace = typename Simulator::StateSpace;

  Factory() : param(TestShardedParameter()) {}
  explicit Factory(const typename StateSpace::Parameter& param) : param(param) {}

  StateSpace CreateStateSpace() const { return StateSpace(param); }
  Simulator CreateSimulator() const { return Simulator(); }

  typename StateSpace::Parameter param;
};

}  // namespace qsim

#endif  // TESTS_CPP_FACTORY_B200_SHARDED_H_
