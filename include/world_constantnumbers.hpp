// Source-compatibility header (worldb200): constants of /root/reference/include/world_constantnumbers.hpp.
#ifndef WORLD_CONSTANT_NUMBERS_HPP
#define WORLD_CONSTANT_NUMBERS_HPP

namespace world {
constexpr double kPi = 3.1415926535897932384;
constexpr double kMySafeGuardMinimum = 0.000000000001;
constexpr double kEps = 0.00000000000000022204460492503131;
constexpr double kFloorF0 = 71.0;
constexpr double kCeilF0 = 800.0;
constexpr double kDefaultF0 = 500.0;
constexpr double kLog2 = 0.69314718055994529;
constexpr double kMaximumValue = 100000.0;
constexpr int kHanning = 1;
constexpr int kBlackman = 2;
constexpr double kFrequencyInterval = 3000.0;
constexpr double kUpperLimit = 15000.0;
constexpr double kThreshold = 0.85;
constexpr double kFloorF0D4C = 47.0;
constexpr double kM0 = 1127.01048;
constexpr double kF0 = 700.0;
constexpr double kFloorFrequency = 40.0;
constexpr double kCeilFrequency = 20000.0;
}  // namespace world

#endif
