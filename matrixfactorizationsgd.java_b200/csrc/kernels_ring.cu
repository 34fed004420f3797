// kernels_ring.cu -- subsystem (4), device side of the ring window (engine.cu "ring window"): the rotation of a one-process-
// per-GPU ring moves the Q slices with copy-engine writes into the neighbour's memory over NVLink and orders them with
// sequence-number flags. The waits are stream memory operations (cuStreamWaitValue32); this file holds the fallback for a
// driver that refuses them: one thread polling the flag (system scope), which needs a free SM slot -- hence a fallback.
#include "kernels.cuh"

namespace mfsgd {

namespace {

__global__ void __launch_bounds__(32) ring_wait_flag_kernel(const uint32_t* __restrict__ flag, uint32_t value) {
    if (threadIdx.x != 0) return;
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - value) >= 0) break;          // cyclic >=, like CU_STREAM_WAIT_VALUE_GEQ
        __nanosleep(200);
    }
}

}  // namespace

cudaError_t launch_ring_wait_flag(const uint32_t* flag, uint32_t value, cudaStream_t stream, int* launches) {
    ring_wait_flag_kernel<<<1, 32, 0, stream>>>(flag, value);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace mfsgd
