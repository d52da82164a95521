// Runs the reference's quantum-trajectory suite (tests/qtrajectory_testfixture.h) on a SHARDED state through
// B200Runner (include/qsim_b200/run_b200.h): the drop-in check for lib/qtrajectory.h on the multi-device backend
// and for the runner's Run overloads (mirror of lib/run_qsim.h:198-315).
#include "qtrajectory_testfixture.h"
#include "gtest/gtest.h"

#include "fuser_mqubit.h"
#include "gates_cirq.h"
#include "io.h"

#include "factory_b200_sharded.h"
#include "qsim_b200/run_b200.h"

namespace qsim {

#define B200_QT_TEST(Name)                                      \
  TEST(QTrajectoryB200ShardedTest, Name) {                      \
    using Factory = qsim::Factory<float>;                       \
    using Fuser = MultiQubitGateFuser<IO>;                      \
    using Runner = B200Runner<IO, Fuser, Factory>;              \
    Factory factory;                                            \
    Test##Name<Runner>(factory);                                \
  }

B200_QT_TEST(BitFlip)
B200_QT_TEST(GenDump)
B200_QT_TEST(ReusingResults)
B200_QT_TEST(CollectKopStat)
B200_QT_TEST(CleanCircuit)
B200_QT_TEST(InitialState)
B200_QT_TEST(UncomputeFinalState)

}  // namespace qsim

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
