// skrample_b200 - host-side helpers shared by the translation units of the library (defined in step_kernel.cu).
#pragma once

#include <atomic>
#include <cstdint>

namespace skr {

int fail(int code, const char* fmt, ...);   // records the thread-local error text, returns `code`
void count_launch(int kind);                // 0 block kernel, 1 interpreter, 2 noise kernels
int sm_count_or(int fallback);
int env_int(const char* name, int fallback);

struct DeviceInfo {
    int ordinal = 0;
    int sm_count = 0;
    int max_smem = 0;
    std::atomic<bool> attr_set[2] = {};  // interpreter instantiations (block kernels keep their own masks)
};
DeviceInfo* device_info(int* err);

// Pipeline shape for the block kernel: stages per CTA and CTAs per SM from the size of one staged tile.
struct PipeShape {
    int stages;
    int ctas_per_sm;
    bool ok;
};
PipeShape pick_shape(uint32_t stage_bytes, int max_smem, int max_ctas);

}  // namespace skr
