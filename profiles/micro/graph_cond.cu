// Micro-test: CUDA graph with fork/join + (nested) conditional IF nodes set from device code; launch throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o graph_cond graph_cond.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); return 1; } } while(0)

__global__ void k_work(float *p, int n, int iters) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	float x = p[i];
	for(int k = 0; k < iters; k++) x = x * 1.0001f + 0.5f;
	p[i] = x;
}
__global__ void k_decide(int *ctr, cudaGraphConditionalHandle h, cudaGraphConditionalHandle h2) {
	int c = atomicAdd(ctr, 1);
	cudaGraphSetConditional(h, (c % 10) == 0);
	cudaGraphSetConditional(h2, (c % 20) == 0);
}
__global__ void k_count(int *ctr) { atomicAdd(ctr, 1); }

int main() {
	const int n = 81920;
	float *a, *b, *c, *d;
	int *ctr;
	CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, n * 4)); CK(cudaMalloc(&d, n * 4));
	CK(cudaMalloc(&ctr, 16)); CK(cudaMemset(ctr, 0, 16));
	cudaStream_t s, s2, s3;
	CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
	cudaEvent_t ef, e2, e3;
	CK(cudaEventCreateWithFlags(&ef, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&e3, cudaEventDisableTiming));

	cudaGraph_t g;
	CK(cudaGraphCreate(&g, 0));
	cudaGraphConditionalHandle h, h2;
	CK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault));
	CK(cudaGraphConditionalHandleCreate(&h2, g, 0, cudaGraphCondAssignDefault));
	// capture the fork/join part into g
	CK(cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
	CK(cudaEventRecord(ef, s));
	CK(cudaStreamWaitEvent(s2, ef, 0)); CK(cudaStreamWaitEvent(s3, ef, 0));
	k_work<<<(n + 127) / 128, 128, 0, s>>>(a, n, 200);
	k_work<<<(n + 127) / 128, 128, 0, s2>>>(b, n, 200);
	k_work<<<(n + 127) / 128, 128, 0, s3>>>(c, n, 200);
	CK(cudaEventRecord(e2, s2)); CK(cudaEventRecord(e3, s3));
	CK(cudaStreamWaitEvent(s, e2, 0)); CK(cudaStreamWaitEvent(s, e3, 0));
	k_work<<<(n + 127) / 128, 128, 0, s>>>(d, n, 200);
	k_decide<<<1, 1, 0, s>>>(ctr, h, h2);
	// conditional node appended behind the capture's current dependencies
	cudaStreamCaptureStatus st; const cudaGraphNode_t *deps; size_t ndeps; cudaGraph_t cg;
	CK(cudaStreamGetCaptureInfo(s, &st, nullptr, &cg, &deps, &ndeps));
	cudaGraphNodeParams cp = {};
	cp.type = cudaGraphNodeTypeConditional;
	cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeIf; cp.conditional.size = 1;
	cudaGraphNode_t cond;
	CK(cudaGraphAddNode(&cond, cg, deps, ndeps, &cp));
	cudaGraph_t body = cp.conditional.phGraph_out[0];
	CK(cudaStreamUpdateCaptureDependencies(s, &cond, 1, cudaStreamSetCaptureDependencies));
	CK(cudaStreamEndCapture(s, &cg));
	// body: count + nested IF
	CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
	k_count<<<1, 1, 0, s>>>(ctr + 1);
	CK(cudaStreamGetCaptureInfo(s, &st, nullptr, &cg, &deps, &ndeps));
	cudaGraphNodeParams cp2 = {};
	cp2.type = cudaGraphNodeTypeConditional;
	cp2.conditional.handle = h2; cp2.conditional.type = cudaGraphCondTypeIf; cp2.conditional.size = 1;
	cudaGraphNode_t cond2;
	CK(cudaGraphAddNode(&cond2, cg, deps, ndeps, &cp2));
	cudaGraph_t body2 = cp2.conditional.phGraph_out[0];
	CK(cudaStreamUpdateCaptureDependencies(s, &cond2, 1, cudaStreamSetCaptureDependencies));
	k_work<<<(n + 127) / 128, 128, 0, s>>>(d, n, 10);
	CK(cudaStreamEndCapture(s, &cg));
	CK(cudaStreamBeginCaptureToGraph(s, body2, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
	k_count<<<1, 1, 0, s>>>(ctr + 2);
	CK(cudaStreamEndCapture(s, &cg));

	cudaGraphExec_t ge;
	CK(cudaGraphInstantiate(&ge, g, 0));
	cudaEvent_t t0, t1;
	CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
	for(int rep = 0; rep < 3; rep++) {
		CK(cudaMemsetAsync(ctr, 0, 16, s));
		CK(cudaEventRecord(t0, s));
		for(int i = 0; i < 1000; i++) CK(cudaGraphLaunch(ge, s));
		CK(cudaEventRecord(t1, s));
		CK(cudaEventSynchronize(t1));
		float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
		int h_ctr[4]; CK(cudaMemcpy(h_ctr, ctr, 16, cudaMemcpyDeviceToHost));
		printf("graph: 1000 launches %.3f ms (%.2f us/launch)  decide=%d if=%d nested=%d\n", ms, ms, h_ctr[0], h_ctr[1], h_ctr[2]);
	}
	// the same kernels as plain serial launches
	for(int rep = 0; rep < 2; rep++) {
		CK(cudaEventRecord(t0, s));
		for(int i = 0; i < 1000; i++) {
			k_work<<<(n + 127) / 128, 128, 0, s>>>(a, n, 200); k_work<<<(n + 127) / 128, 128, 0, s>>>(b, n, 200);
			k_work<<<(n + 127) / 128, 128, 0, s>>>(c, n, 200); k_work<<<(n + 127) / 128, 128, 0, s>>>(d, n, 200);
			k_count<<<1, 1, 0, s>>>(ctr + 3);
		}
		CK(cudaEventRecord(t1, s));
		CK(cudaEventSynchronize(t1));
		float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
		printf("serial launches: %.2f us per 5-kernel step\n", ms);
	}
	printf("OK\n");
	return 0;
}
