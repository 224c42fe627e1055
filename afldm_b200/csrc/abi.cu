// Library identification, launch accounting and error strings of the afldm_b200 C ABI.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace afldm {
static std::atomic<unsigned long long> g_launches{0};
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("AFLDM_PDL");
        // on by default: +3.3 % on the step (298 vs 288 steps/s, three A/B runs on one B200) once the kernels were short
        // enough for the ~1.2 us inter-kernel gaps to matter; it was neutral at 5 ms per step.  AFLDM_PDL=0 disables.
        return e == nullptr || e[0] != '0';
    }();
    return on;
}
void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace afldm

extern "C" int afldm_abi_version(void) { return 1; }

extern "C" unsigned long long afldm_launch_count(void) {
    return afldm::g_launches.load(std::memory_order_relaxed);
}

extern "C" const char* afldm_error_string(int code) {
    switch (code) {
        case AFLDM_OK: return "ok";
        case AFLDM_E_SHAPE: return "afldm: shape / divisibility not supported by this kernel family";
        case AFLDM_E_ARG: return "afldm: bad argument (null or misaligned pointer, bad enum, bad pitch)";
        case AFLDM_E_NOKERNEL: return "afldm: no kernel specialisation for this size";
        case AFLDM_E_WORKSPACE: return "afldm: workspace too small";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "afldm: unknown error";
}
