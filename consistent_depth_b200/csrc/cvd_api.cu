// Library-wide state of libcvd_sm100.so: version, last-error string, launch counter.
#include "cvd_common.cuh"

thread_local char g_cvd_err[512] = {0};
long long g_cvd_launches = 0;

extern "C" int cvd_version(void) { return CVD_VERSION; }
extern "C" const char* cvd_last_error(void) { return g_cvd_err; }
extern "C" long long cvd_launch_count(void) { return g_cvd_launches; }
