// Stand-in for the file the reference's CMake generates from
// src/version.cpp.in (configure_file, CMakeLists.txt:47); written by us so the
// reference can be compiled by oracle/Makefile without running its build system.
#include "version.h"
const std::string Version::GIT_SHA1 = "oracle-ref-build";
const std::string Version::GIT_DATE = "n/a";
const std::string Version::GIT_COMMIT_SUBJECT = "unmodified /root/reference/src compiled with oracle/shim (single-rank mpi.h, fixed-size Eigen)";
