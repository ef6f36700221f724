// Error reporting and version for the C ABI (include/xdb200.h).
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static thread_local char g_err[512] = "";

void xd_set_error(const char* file, int line, const char* msg) {
    const char* base = strrchr(file, '/');
    snprintf(g_err, sizeof g_err, "%s:%d: %s", base ? base + 1 : file, line, msg);
}

extern "C" const char* xd_last_error(void) { return g_err; }
extern "C" int xd_abi_version(void) { return 1; }

bool xd_pdl_enabled() {
    // measured neutral under CUDA-graph replay (483 vs 482 img/s, DiT): off unless XDB200_PDL=1
    static const bool on = getenv("XDB200_PDL") && atoi(getenv("XDB200_PDL")) == 1;
    return on;
}
