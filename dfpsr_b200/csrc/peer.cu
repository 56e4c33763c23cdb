// peer.cu — strip-sharded frames assembled over NVLink peer memory (SURVEY.md §8e "Screen strips").
//
// The reference splits one frame into row strips for its worker threads, all writing into the same target
// (ref: implementation/render/renderCore.cpp:449-480). Across GPUs the same split is one process per GPU with the presenting
// rank owning the frame: its memory is exported with a CUDA IPC handle, every other rank maps it and hands the mapped pointer to
// the renderer as its colour target, so the tile kernel's final 8-byte stores ARE the gather (peer stores through NVSwitch, no
// staging copy and no collective launch). Completion travels the same way: a flag per rank in the presenter's memory, written by
// a one-thread kernel at the end of the rank's stream after a system-wide fence, and a wait kernel on the presenter's stream.
#include "common.cuh"

namespace dfpsr {

struct FlagList { uint32_t *flag[DFPSR_PEER_MAX_RANKS]; };

__global__ void peer_signal_kernel(FlagList flags, int32_t count, uint32_t value) {
	// everything this stream wrote before (the tile kernel's peer stores included) is ordered before the flag
	__threadfence_system();
	if ((int32_t)threadIdx.x < count) {
		asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.flag[threadIdx.x]), "r"(value) : "memory");
	}
}

// status[0]: number of waits through this status block that gave up (a rank never signalled; the waiter must not hang the device),
// status[1]: the value the most recent of them was waiting for. Every wait site owns its block; dfpsr_peer_reset_status clears one.
__global__ void peer_wait_kernel(const uint32_t *flags, int32_t count, uint32_t value, uint64_t timeoutNs, uint32_t *status) {
	const int32_t i = (int32_t)threadIdx.x;
	bool timedOut = false;
	if (i < count) {
		uint64_t start;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
		for (;;) {
			uint32_t seen;
			asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + i) : "memory");
			if ((int32_t)(seen - value) >= 0) { break; } // frame numbers only grow; wrap-safe comparison
			uint64_t now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			if (now - start > timeoutNs) { timedOut = true; break; }
			__nanosleep(64);
		}
	}
	if (timedOut) { atomicAdd(status, 1u); atomicExch(status + 1, value); }
	__threadfence_system();
}

} // namespace dfpsr

using namespace dfpsr;

static_assert(sizeof(cudaIpcMemHandle_t) == DFPSR_PEER_HANDLE_BYTES, "CUDA IPC handle size");

extern "C" {

int dfpsr_peer_alloc(void **devicePtr, size_t bytes, uint8_t *handle) {
	DFPSR_REQUIRE(devicePtr != nullptr && handle != nullptr && bytes > 0, "peer_alloc: null argument or zero size");
	void *ptr = nullptr;
	// plain cudaMalloc: pooled / virtual-memory allocations cannot be exported through cudaIpcGetMemHandle
	DFPSR_CHECK_CUDA(cudaMalloc(&ptr, bytes));
	cudaError_t err = cudaMemset(ptr, 0, bytes);
	cudaIpcMemHandle_t h;
	if (err == cudaSuccess) { err = cudaIpcGetMemHandle(&h, ptr); }
	if (err != cudaSuccess) {
		cudaFree(ptr);
		set_error("peer_alloc: %s", cudaGetErrorString(err));
		return 1;
	}
	memcpy(handle, &h, sizeof(h));
	*devicePtr = ptr;
	return 0;
}

int dfpsr_peer_free(void *devicePtr) {
	if (devicePtr) { DFPSR_CHECK_CUDA(cudaFree(devicePtr)); }
	return 0;
}

int dfpsr_peer_open(void **devicePtr, const uint8_t *handle) {
	DFPSR_REQUIRE(devicePtr != nullptr && handle != nullptr, "peer_open: null argument");
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof(h));
	void *ptr = nullptr;
	// cudaIpcMemLazyEnablePeerAccess maps the exporting GPU's memory into this one over NVLink when they are different devices
	DFPSR_CHECK_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
	*devicePtr = ptr;
	return 0;
}

int dfpsr_peer_close(void *devicePtr) {
	if (devicePtr) { DFPSR_CHECK_CUDA(cudaIpcCloseMemHandle(devicePtr)); }
	return 0;
}

int dfpsr_peer_signal(uint32_t *const *flags, int32_t count, uint32_t value, void *stream) {
	DFPSR_REQUIRE(flags != nullptr && count >= 1 && count <= DFPSR_PEER_MAX_RANKS, "peer_signal: 1..%d flags, got %d", DFPSR_PEER_MAX_RANKS, count);
	FlagList list = {};
	for (int32_t i = 0; i < count; i++) {
		DFPSR_REQUIRE(flags[i] != nullptr, "peer_signal: flag %d is null", i);
		list.flag[i] = flags[i];
	}
	DFPSR_LAUNCH(peer_signal_kernel, 1, 32, 0, as_stream(stream), list, count, value);
	return 0;
}

int dfpsr_peer_reset_status(uint32_t *status, void *stream) {
	DFPSR_REQUIRE(status != nullptr, "peer_reset_status: null argument");
	DFPSR_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(uint32_t), as_stream(stream)));
	return 0;
}

int dfpsr_peer_wait(const uint32_t *flags, int32_t count, uint32_t value, uint32_t timeoutMs, uint32_t *status, void *stream) {
	DFPSR_REQUIRE(flags != nullptr && status != nullptr && count >= 1 && count <= DFPSR_PEER_MAX_RANKS, "peer_wait: 1..%d flags, got %d", DFPSR_PEER_MAX_RANKS, count);
	DFPSR_REQUIRE(timeoutMs >= 1 && timeoutMs <= 10000, "peer_wait: the time limit must be 1..10000 ms (a waiter must never hang the device)");
	DFPSR_LAUNCH(peer_wait_kernel, 1, 32, 0, as_stream(stream), flags, count, value, (uint64_t)timeoutMs * 1000000ull, status);
	return 0;
}

} // extern "C"
