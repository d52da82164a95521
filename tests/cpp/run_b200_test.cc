// B200Runner (include/qsim_b200/run_b200.h) against the known answers of the reference's runner tests
// (tests/run_qsim_test.cc:80-245: the 4-qubit circuit with entropy 2.2192848, the measurement-consistency
// circuit, CirqCircuit1 from tests/gates_cirq_testfixture.h), on the single-GPU backend and on a sharded state.
#include <cmath>
#include <complex>
#include <cstdint>
#include <sstream>
#include <vector>

#include "gates_cirq_testfixture.h"
#include "gtest/gtest.h"

#include "circuit_qsim_parser.h"
#include "fuser_basic.h"
#include "fuser_mqubit.h"
#include "io.h"
#include "operation.h"

#include "qsim_b200/run_b200.h"
#include "qsim_b200/simulator_b200.h"
#include "factory_b200_sharded.h"

namespace qsim {

template <typename FP>
struct SingleFactory {
  using fp_type = FP;
  using Simulator = qsim::SimulatorB200<FP>;
  using StateSpace = typename Simulator::StateSpace;
  StateSpace CreateStateSpace() const { return StateSpace(); }
  Simulator CreateSimulator() const { return Simulator(); }
};

// the circuit of tests/run_qsim_test.cc:36-64 (known answer: entropy of the final distribution)
constexpr char kEntropyCircuit[] =
    "4\n0 h 0\n0 h 1\n0 h 2\n0 h 3\n1 cz 0 1\n1 cz 2 3\n2 t 0\n2 x 1\n2 y 2\n2 t 3\n3 y 0\n3 cz 1 2\n3 x 3\n"
    "4 t 1\n4 t 2\n5 cz 1 2\n6 x 1\n6 y 2\n7 cz 1 2\n8 t 1\n8 t 2\n9 cz 0 1\n9 cz 2 3\n10 h 0\n10 h 1\n10 h 2\n10 h 3\n";
constexpr double kEntropy = 2.2192848;
// tests/run_qsim_test.cc:159-171
constexpr char kSampleCircuit[] = "2\n0 h 0\n0 x 1\n1 m 1\n2 cx 0 1\n3 m 0 1\n4 m 0\n5 cx 1 0\n6 m 0\n7 x 0\n7 h 1\n8 m 0 1\n";

template <typename StateSpace, typename State>
double Entropy(const StateSpace& ss, const State& state) {
  double e = 0;
  for (uint64_t i = 0; i < (uint64_t{1} << state.num_qubits()); ++i) {
    const double p = std::norm(ss.GetAmpl(state, i));
    if (p > 0) e -= p * std::log(p);
  }
  return e;
}

template <class F>
class RunB200Test : public testing::Test {};
using Factories = ::testing::Types<SingleFactory<float>, Factory<float>>;
TYPED_TEST_SUITE(RunB200Test, Factories);

TYPED_TEST(RunB200Test, MeasureCallbackAtTheEnd) {
  using StateSpace = typename TypeParam::StateSpace;
  using State = typename StateSpace::State;
  std::stringstream ss(kEntropyCircuit);
  Circuit<Operation<float>> circuit;
  ASSERT_TRUE(CircuitQsimParser<IO>::FromStream(99, "run_b200_test", ss, circuit));
  ASSERT_EQ(circuit.ops.size(), 27u);
  for (int fuser = 0; fuser < 2; ++fuser) {
    double entropy = 0;
    unsigned calls = 0;
    auto measure = [&](unsigned, const StateSpace& space, const State& state) { entropy = Entropy(space, state); ++calls; };
    bool ok;
    if (fuser == 0) {
      using Runner = B200Runner<IO, BasicGateFuser<IO>, TypeParam>;
      typename Runner::Parameter param;
      param.seed = 1; param.verbosity = 0;
      ok = Runner::Run(param, TypeParam(), circuit, measure);
    } else {
      using Runner = B200Runner<IO, MultiQubitGateFuser<IO>, TypeParam>;
      typename Runner::Parameter param;
      param.max_fused_size = 3; param.seed = 1; param.verbosity = 0;
      ok = Runner::Run(param, TypeParam(), {3, 7, 10}, circuit, measure);
      EXPECT_EQ(calls, 3u);
    }
    EXPECT_TRUE(ok);
    EXPECT_NEAR(entropy, kEntropy, 1e-6);
  }
}

TYPED_TEST(RunB200Test, FinalState) {
  using StateSpace = typename TypeParam::StateSpace;
  using Runner = B200Runner<IO, BasicGateFuser<IO>, TypeParam>;
  std::stringstream ss(kEntropyCircuit);
  Circuit<Operation<float>> circuit;
  ASSERT_TRUE(CircuitQsimParser<IO>::FromStream(99, "run_b200_test", ss, circuit));
  TypeParam factory;
  StateSpace space = factory.CreateStateSpace();
  auto state = space.Create(circuit.num_qubits);
  ASSERT_FALSE(space.IsNull(state));
  space.SetStateZero(state);
  typename Runner::Parameter param;
  param.seed = 1; param.verbosity = 0;
  EXPECT_TRUE(Runner::Run(param, factory, circuit, state));
  EXPECT_NEAR(Entropy(space, state), kEntropy, 1e-6);
}

TYPED_TEST(RunB200Test, MeasurementGatesInsideTheCircuit) {
  using StateSpace = typename TypeParam::StateSpace;
  using Result = typename StateSpace::MeasurementResult;
  using Runner = B200Runner<IO, BasicGateFuser<IO>, TypeParam>;
  std::stringstream ss(kSampleCircuit);
  Circuit<Operation<float>> circuit;
  ASSERT_TRUE(CircuitQsimParser<IO>::FromStream(99, "run_b200_test", ss, circuit));
  TypeParam factory;
  StateSpace space = factory.CreateStateSpace();
  auto state = space.Create(circuit.num_qubits);
  ASSERT_FALSE(space.IsNull(state));
  space.SetStateZero(state);
  std::vector<Result> results;
  typename Runner::Parameter param;
  param.seed = 1; param.verbosity = 0;
  ASSERT_TRUE(Runner::Run(param, factory, circuit, state, results));
  ASSERT_EQ(results.size(), 5u);
  EXPECT_TRUE(results[0].bitstring[0]);                                 // q1 @ 1: |01)
  EXPECT_EQ(results[1].bitstring[0], !results[1].bitstring[1]);        // |01) or |10)
  EXPECT_EQ(results[1].bitstring[0], results[2].bitstring[0]);         // repeated measurement agrees
  EXPECT_TRUE(results[3].bitstring[0]);
  EXPECT_FALSE(results[4].bitstring[0]);
  EXPECT_FALSE(results[4].bitstring[1]);
}

TYPED_TEST(RunB200Test, CirqGatesKnownAmplitudes) {
  using StateSpace = typename TypeParam::StateSpace;
  using Runner = B200Runner<IO, BasicGateFuser<IO>, TypeParam>;
  auto circuit = CirqCircuit1::GetCircuit<float>(true);
  const auto& expected = CirqCircuit1::expected_results1;
  TypeParam factory;
  StateSpace space = factory.CreateStateSpace();
  auto state = space.Create(circuit.num_qubits);
  ASSERT_FALSE(space.IsNull(state));
  ASSERT_EQ(uint64_t{1} << circuit.num_qubits, expected.size());
  space.SetStateZero(state);
  typename Runner::Parameter param;
  param.seed = 1; param.verbosity = 0;
  EXPECT_TRUE(Runner::Run(param, factory, circuit, state));
  for (uint64_t i = 0; i < expected.size(); ++i) {
    const auto a = space.GetAmpl(state, i);
    EXPECT_NEAR(std::real(a), std::real(expected[i]), 2e-6);
    EXPECT_NEAR(std::imag(a), std::imag(expected[i]), 2e-6);
  }
}

}  // namespace qsim

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
