// ref_host_platform.cpp -- TEST INFRASTRUCTURE ONLY.  The handful of FractalSharkPlatform entry points the reference's
// table builders link against (FractalSharkPlatform/Common/Environment.h:120-206).  The reference's own Linux
// implementation (FractalSharkPlatform/Linux/EnvironmentLinux.cpp) needs X11 headers this image does not have; none
// of these functions takes part in the arithmetic (thread names, a debugger trap, wall-clock counters, key state).
#include "Environment.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace Environment {
void *FileOpenDeleteOnClose(const wchar_t *) { return nullptr; }
void FileClose(void *) {}
void DebugBreakpoint() {
    std::fprintf(stderr, "oracle/_ref/libref_host.so: the reference hit Environment::DebugBreakpoint()\n");
    std::abort();
}
void SetCurrentThreadName(const wchar_t *) {}
uint64_t HighResCounter() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
uint64_t HighResFrequency() { return 1000000000ull; }
bool IsKeyDown(Key) { return false; }
std::pair<int, int> GetCursorPosition() { return {0, 0}; }
} // namespace Environment
