// Minimal stand-in for <ros/ros.h> (TEST INFRASTRUCTURE): just enough
// declarations for the reference's POMDP headers to parse when only its
// CUDA translation units are compiled for the oracle.  Nothing is linked.
#pragma once
#include <cmath>
#include <cstdio>
#include <iostream>
#include <string>
namespace ros {
class Publisher {};
class Subscriber {};
class ServiceServer {};
class Timer {};
class NodeHandle {
 public:
  template <typename T> bool getParam(const std::string&, T&) const { return false; }
  template <typename T> void param(const std::string&, T&, const T&) const {}
};
inline bool ok() { return true; }
struct Time {
  static Time now() { return Time(); }
  double toSec() const { return 0.0; }
  Time operator-(const Time&) const { return Time(); }
};
}  // namespace ros
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_INFO(...) ((void)0)
