// Library-level entry points of libsdt_b200: error string, version, math-mode switch.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
// process default of the convolution math mode: 3 (tcgen05 TF32, TMA operands, persistent kernel) unless SDT_CONV_MATH=0..4
int initial_conv_math() {
    const char* e = getenv("SDT_CONV_MATH");
    if (e != nullptr && e[0] >= '0' && e[0] <= '4' && e[1] == '\0') return e[0] - '0';
    return 3;
}
std::atomic<int> g_conv_math{initial_conv_math()};
std::atomic<long long> g_tc_launches{0};
}  // namespace

void sdt_note_tc_launch() { g_tc_launches.fetch_add(1); }

namespace sdt {
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("SDT_PDL");
        return !(e != nullptr && e[0] == '0');
    }();
    return on;
}
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace sdt

extern "C" const char* sdt_last_error(void) { return g_err; }
extern "C" int sdt_version(void) { return 100; }
extern "C" int sdt_set_conv_math(int mode) {
    SDT_REQUIRE(mode >= 0 && mode <= 4, "sdt_set_conv_math: mode must be 0 (fp32 FFMA), 1 (tcgen05 TF32), 2 (+ TMA operands), 3 (+ persistent kernel with operand reuse) or 4 (+ CTA pairs, experimental)");
    g_conv_math.store(mode);
    return SDT_OK;
}
extern "C" int sdt_get_conv_math(void) { return g_conv_math.load(); }
extern "C" int64_t sdt_tc_launches(void) { return g_tc_launches.load(); }
