// Runs the reference's quantum-trajectory suite (tests/qtrajectory_testfixture.h)
// with QSimRunner + MultiQubitGateFuser (unchanged reference host code) on the
// B200 backend -- the drop-in check for lib/qtrajectory.h and lib/run_qsim.h.
#include "qtrajectory_testfixture.h"
#include "gtest/gtest.h"

#include "fuser_mqubit.h"
#include "gates_cirq.h"
#include "io.h"
#include "run_qsim.h"

#include "factory_b200.h"

namespace qsim {

#define B200_QT_TEST(Name)                                      \
  TEST(QTrajectoryB200Test, Name) {                             \
    using Factory = qsim::Factory<float>;                       \
    using Fuser = MultiQubitGateFuser<IO>;                      \
    using Runner = QSimRunner<IO, Fuser, Factory>;              \
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
