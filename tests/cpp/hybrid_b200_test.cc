// Runs the reference's qsimh (Schroedinger-Feynman hybrid) suite
// (tests/hybrid_testfixture.h) on the B200 backend: lib/hybrid.h calls the hot
// path through the same Simulator/StateSpace interface.
#include "hybrid_testfixture.h"
#include "gtest/gtest.h"

#include "factory_b200.h"

namespace qsim {

TEST(HybridB200Test, Hybrid2) {
  qsim::Factory<float> factory;
  TestHybrid2(factory);
}

TEST(HybridB200Test, Hybrid4) {
  qsim::Factory<float> factory;
  TestHybrid4(factory);
}

}  // namespace qsim

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
