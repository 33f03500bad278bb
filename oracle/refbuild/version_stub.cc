// Build stub (ours, not reference source): the reference's CMake generates version.cc into its source
// tree (riss/CMakeLists.txt:21, coprocessor/CMakeLists.txt:9), which is read-only here.  These four
// symbols per namespace are all that riss/utils/version.h:7-10 and coprocessor/version.h:7-10 declare.
namespace Riss {
const char* gitSHA1 = "oracle-ref";
const char* gitDate = "n/a";
const char* solverVersion = "7.1.0";
const char* signature = "riss oracle/_ref build";
}
namespace Coprocessor {
const char* gitSHA1 = "oracle-ref";
const char* gitDate = "n/a";
const char* coprocessorVersion = "7.1.0";
const char* signature = "coprocessor oracle/_ref build";
}
