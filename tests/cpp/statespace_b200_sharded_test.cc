// Runs the reference's own state-space suite (tests/statespace_testfixture.h,
// consumed in place) on a state SHARDED over several B200s (or several shards on one), float and double.
#include "statespace_testfixture.h"
#include "gtest/gtest.h"

#include "factory_b200_sharded.h"

namespace qsim {

template <class T>
class StateSpaceB200ShardedTest : public testing::Test {};

using fp_impl = ::testing::Types<float, double>;
TYPED_TEST_SUITE(StateSpaceB200ShardedTest, fp_impl);

#define B200_SS_TEST(Name)                             \
  TYPED_TEST(StateSpaceB200ShardedTest, Name) {               \
    qsim::Factory<TypeParam> factory;                  \
    Test##Name(factory);                               \
  }

B200_SS_TEST(Add)
B200_SS_TEST(NormSmall)
B200_SS_TEST(NormAndInnerProductSmall)
B200_SS_TEST(NormAndInnerProduct)
B200_SS_TEST(SamplingSmall)
B200_SS_TEST(SamplingCrossEntropyDifference)
B200_SS_TEST(Ordering)
B200_SS_TEST(MeasurementLarge)
B200_SS_TEST(Collapse)
B200_SS_TEST(BulkSetAmplitude)
B200_SS_TEST(BulkSetAmplitudeExclusion)
B200_SS_TEST(BulkSetAmplitudeDefault)

TEST(StateSpaceB200ShardedTest, MeasurementSmall) {
  qsim::Factory<float> factory;
  TestMeasurementSmall(factory, true);
}

TEST(StateSpaceB200ShardedTest, InvalidStateSize) {
  qsim::Factory<float> factory;
  TestInvalidStateSize(factory);
}

}  // namespace qsim

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
