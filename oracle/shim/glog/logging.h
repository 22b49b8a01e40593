// TEST INFRASTRUCTURE: minimal stand-in for <glog/logging.h> so that the reference's okvis_matcher
// sources compile unmodified into oracle/_ref (glog is not installed in this image).
#pragma once
#include <iostream>
struct SvinNullLog {
  template <typename T>
  SvinNullLog& operator<<(const T&) { return *this; }
  SvinNullLog& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
#define LOG(x) SvinNullLog()
#define VLOG(x) SvinNullLog()
#define CHECK(x) SvinNullLog()
#define DLOG(x) SvinNullLog()
