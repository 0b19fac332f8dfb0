// Stand-in for <glog/logging.h> when compiling the reference's kernels outside its own build
// (glog is a vcpkg dependency that is not installed here): torch's c10 logging header provides
// the same CHECK / DCHECK / LOG macros the kernel sources use.
#pragma once
#include <c10/util/Logging.h>
