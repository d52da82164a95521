// factory_b200.h -- the Factory every reference test fixture consumes
// (same shape as tests/simulator_cuda_test.cu:32-48 in the reference).
#ifndef TESTS_CPP_FACTORY_B200_H_
#define TESTS_CPP_FACTORY_B200_H_

#include "qsim_b200/simulator_b200.h"

namespace qsim {

template <typename FP>
struct Factory {
  using fp_type = FP;
  using Simulator = qsim::SimulatorB200<fp_type>;
  using StateSpace = typename Simulator::StateSpace;

  Factory() {}
  explicit Factory(const typename StateSpace::Parameter& param) : param(param) {}

  StateSpace CreateStateSpace() const { return StateSpace(param); }
  Simulator CreateSimulator() const { return Simulator(); }

  typename StateSpace::Parameter param;
};

}  // namespace qsim

#endif  // TESTS_CPP_FACTORY_B200_H_
