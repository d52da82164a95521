// Runs the reference's own simulator known-answer suite
// (tests/simulator_testfixture.h, consumed in place) on a state SHARDED over several B200s (or several shards on one).
#include <type_traits>

#include "simulator_testfixture.h"
#include "gtest/gtest.h"

#include "factory_b200_sharded.h"

namespace qsim {

template <class T>
class SimulatorB200ShardedTest : public testing::Test {};

using fp_impl = ::testing::Types<float, double>;
TYPED_TEST_SUITE(SimulatorB200ShardedTest, fp_impl);

#define B200_SIM_TEST(Name, ...)                       \
  TYPED_TEST(SimulatorB200ShardedTest, Name) {                \
    qsim::Factory<TypeParam> factory;                  \
    Test##Name(factory, ##__VA_ARGS__);                \
  }

B200_SIM_TEST(ApplyGate1)
B200_SIM_TEST(ApplyGate2)
B200_SIM_TEST(ApplyGate3)
B200_SIM_TEST(ApplyGate5)
B200_SIM_TEST(CircuitWithControlledGates)
B200_SIM_TEST(CircuitWithControlledGatesDagger)
B200_SIM_TEST(MultiQubitGates)
B200_SIM_TEST(ControlledGates, (std::is_same<TypeParam, double>::value))
B200_SIM_TEST(GlobalPhaseGate)
B200_SIM_TEST(ExpectationValue1)
B200_SIM_TEST(ExpectationValue2)

}  // namespace qsim

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
