// Does a kernel write to cudaMallocHost memory (through the host pointer, UVA) reach the host?  The reference's list
// staleness flag (CUDABaseBackend.cu:210, CUDA_mixed.cuh:68) depends on it.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(bool *f) { f[0] = true; }
__global__ void k2(bool *f, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if(i < n && i % 1000 == 7) f[0] = true; }
int main() {
	bool *f;
	cudaError_t e = cudaMallocHost(&f, sizeof(bool), cudaHostAllocDefault);
	printf("cudaMallocHost: %s\n", cudaGetErrorString(e));
	f[0] = false;
	k<<<1, 1>>>(f);
	e = cudaDeviceSynchronize();
	printf("kernel: %s, flag after kernel = %d\n", cudaGetErrorString(e), (int) f[0]);
	f[0] = false;
	k2<<<640, 128>>>(f, 81920);
	e = cudaDeviceSynchronize();
	printf("kernel2: %s, flag after kernel = %d\n", cudaGetErrorString(e), (int) f[0]);
	int v = 0;
	cudaDeviceGetAttribute(&v, cudaDevAttrCanUseHostPointerForRegisteredMem, 0); printf("CanUseHostPointerForRegisteredMem = %d\n", v);
	cudaDeviceGetAttribute(&v, cudaDevAttrUnifiedAddressing, 0); printf("UnifiedAddressing = %d\n", v);
	cudaDeviceGetAttribute(&v, cudaDevAttrCanMapHostMemory, 0); printf("CanMapHostMemory = %d\n", v);
	return 0;
}
