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
extern "C" int xd_abi_version(void) { return 3; }

// Programmatic dependent launch: XDB200_PDL=0 off, 1 every kernel, 2 (default) only the tcgen05 GEMM / conv launches.
// Measured (profiles/README.md): GEMM-only gains ~2 % on the DiT step; on every kernel it costs the UNet 6 %.
static int pdl_mode() {
    static const int m = getenv("XDB200_PDL") ? atoi(getenv("XDB200_PDL")) : 2;
    return m;
}
bool xd_pdl_enabled() { return pdl_mode() == 1; }
bool xd_pdl_enabled_gemm() { return pdl_mode() >= 1; }
