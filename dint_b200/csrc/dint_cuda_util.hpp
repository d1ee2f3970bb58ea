// Small host-side CUDA helpers shared by the translation units of libdint_b200.so (not installed).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <string>

#include "dint_internal.hpp"

#define CUDA_TRY(expr)                                                            \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            dint::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__)); \
            return DINT_E_CUDA;                                                   \
        }                                                                         \
    } while (0)

namespace dintk {

struct device_info {
    int sms = 0;
    int max_smem_optin = 0;
};
int get_device_info(int device, device_info* di);  // dint_kernels.cu

// kernels launched by the library in this process (dint_kernel_launches)
extern std::atomic<uint64_t> g_launches;

struct device_guard {
    int prev = -1;
    bool ok = false;
    explicit device_guard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~device_guard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace dintk
