// stand-in for <std_srvs/Trigger.h> (TEST INFRASTRUCTURE)
#pragma once
#include <string>
namespace std_srvs {
struct Trigger {
  struct Request {};
  struct Response { bool success; std::string message; };
};
}  // namespace std_srvs
