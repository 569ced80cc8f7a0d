// Micro-test: launch overhead of the step structures considered for the MD loop (sm_100a):
//  A serial stream launches; B graph of serial kernels; C graph with fork/join; D = C + device-set IF node;
//  E = C unrolled 8x per graph; F = WHILE node looping C on the device (one launch for all steps).
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
__global__ void k_decide(int *ctr, cudaGraphConditionalHandle h, int every) {
	int c = atomicAdd(ctr, 1);
	cudaGraphSetConditional(h, (c % every) == 0);
}
__global__ void k_loop(int *ctr, cudaGraphConditionalHandle h, int total) {
	int c = atomicAdd(ctr, 1);
	cudaGraphSetConditional(h, (c + 1) < total);
}
__global__ void k_count(int *ctr) { atomicAdd(ctr, 1); }

const int n = 81920, IT = 200;
float *a, *b, *c, *d;
int *ctr;
cudaStream_t s, s2, s3;
cudaEvent_t ef, e2, e3;

int step_forkjoin() {
	CK(cudaEventRecord(ef, s));
	CK(cudaStreamWaitEvent(s2, ef, 0)); CK(cudaStreamWaitEvent(s3, ef, 0));
	k_work<<<(n + 127) / 128, 128, 0, s>>>(a, n, IT);
	k_work<<<(n + 127) / 128, 128, 0, s2>>>(b, n, IT);
	k_work<<<(n + 127) / 128, 128, 0, s3>>>(c, n, IT);
	CK(cudaEventRecord(e2, s2)); CK(cudaEventRecord(e3, s3));
	CK(cudaStreamWaitEvent(s, e2, 0)); CK(cudaStreamWaitEvent(s, e3, 0));
	k_work<<<(n + 127) / 128, 128, 0, s>>>(d, n, IT);
	k_work<<<(n + 127) / 128, 128, 0, s>>>(d, n, IT);
	return 0;
}
int step_serial() {
	k_work<<<(n + 127) / 128, 128, 0, s>>>(a, n, IT); k_work<<<(n + 127) / 128, 128, 0, s>>>(b, n, IT);
	k_work<<<(n + 127) / 128, 128, 0, s>>>(c, n, IT); k_work<<<(n + 127) / 128, 128, 0, s>>>(d, n, IT);
	k_work<<<(n + 127) / 128, 128, 0, s>>>(d, n, IT);
	return 0;
}

int time_exec(const char *name, cudaGraphExec_t ge, int launches, int steps_per_launch) {
	cudaEvent_t t0, t1;
	CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
	for(int rep = 0; rep < 3; rep++) {
		CK(cudaMemsetAsync(ctr, 0, 16, s));
		CK(cudaEventRecord(t0, s));
		for(int i = 0; i < launches; i++) CK(cudaGraphLaunch(ge, s));
		CK(cudaEventRecord(t1, s));
		CK(cudaEventSynchronize(t1));
		float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
		int h[4]; CK(cudaMemcpy(h, ctr, 16, cudaMemcpyDeviceToHost));
		if(rep == 2) printf("%-44s %8.2f us/step   (ctr %d %d)\n", name, 1e3 * ms / (launches * steps_per_launch), h[0], h[1]);
	}
	return 0;
}

int add_if(cudaStream_t st, cudaGraphConditionalHandle h, enum cudaGraphConditionalNodeType type, cudaGraph_t *body) {
	cudaStreamCaptureStatus cs; const cudaGraphNode_t *deps; size_t ndeps; cudaGraph_t cg;
	CK(cudaStreamGetCaptureInfo(st, &cs, nullptr, &cg, &deps, &ndeps));
	cudaGraphNodeParams cp = {};
	cp.type = cudaGraphNodeTypeConditional;
	cp.conditional.handle = h; cp.conditional.type = type; cp.conditional.size = 1;
	cudaGraphNode_t cond;
	CK(cudaGraphAddNode(&cond, cg, deps, ndeps, &cp));
	*body = cp.conditional.phGraph_out[0];
	CK(cudaStreamUpdateCaptureDependencies(st, &cond, 1, cudaStreamSetCaptureDependencies));
	return 0;
}

int main() {
	CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, n * 4)); CK(cudaMalloc(&d, n * 4));
	CK(cudaMalloc(&ctr, 16)); CK(cudaMemset(ctr, 0, 16));
	CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
	CK(cudaEventCreateWithFlags(&ef, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&e3, cudaEventDisableTiming));
	cudaEvent_t t0, t1;
	CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
	cudaGraph_t g; cudaGraphExec_t ge;

	// A: plain stream launches
	for(int rep = 0; rep < 3; rep++) {
		CK(cudaEventRecord(t0, s));
		for(int i = 0; i < 1000; i++) step_serial();
		CK(cudaEventRecord(t1, s)); CK(cudaEventSynchronize(t1));
		float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
		if(rep == 2) printf("%-44s %8.2f us/step\n", "A  5 serial stream launches", ms);
	}
	// A2: fork/join with streams+events, no graph
	for(int rep = 0; rep < 3; rep++) {
		CK(cudaEventRecord(t0, s));
		for(int i = 0; i < 1000; i++) step_forkjoin();
		CK(cudaEventRecord(t1, s)); CK(cudaEventSynchronize(t1));
		float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
		if(rep == 2) printf("%-44s %8.2f us/step\n", "A2 fork/join on 3 streams, no graph", ms);
	}
	// B: graph of 5 serial kernels
	CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal)); step_serial(); CK(cudaStreamEndCapture(s, &g));
	CK(cudaGraphInstantiate(&ge, g, 0)); time_exec("B  graph, 5 serial kernels", ge, 1000, 1);
	// C: graph fork/join
	CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal)); step_forkjoin(); CK(cudaStreamEndCapture(s, &g));
	CK(cudaGraphInstantiate(&ge, g, 0)); time_exec("C  graph, fork/join (3 wide) + 2", ge, 1000, 1);
	// D: C + decide kernel + IF node (body = 1 kernel, taken every 10th)
	{
		CK(cudaGraphCreate(&g, 0));
		cudaGraphConditionalHandle h;
		CK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault));
		CK(cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		step_forkjoin();
		k_decide<<<1, 1, 0, s>>>(ctr, h, 10);
		cudaGraph_t body, cg;
		if(add_if(s, h, cudaGraphCondTypeIf, &body)) return 1;
		CK(cudaStreamEndCapture(s, &cg));
		CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		k_count<<<1, 1, 0, s>>>(ctr + 1);
		CK(cudaStreamEndCapture(s, &cg));
		CK(cudaGraphInstantiate(&ge, g, 0)); time_exec("D  C + decide + IF node", ge, 1000, 1);
	}
	// E: 8 steps of C per graph
	CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
	for(int k = 0; k < 8; k++) step_forkjoin();
	CK(cudaStreamEndCapture(s, &g));
	CK(cudaGraphInstantiate(&ge, g, 0)); time_exec("E  graph, 8 x fork/join step per launch", ge, 125, 8);
	// E2: 8 serial steps per graph
	CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
	for(int k = 0; k < 8; k++) step_serial();
	CK(cudaStreamEndCapture(s, &g));
	CK(cudaGraphInstantiate(&ge, g, 0)); time_exec("E2 graph, 8 x serial step per launch", ge, 125, 8);
	// F: WHILE node, body = serial step + loop kernel; one launch = 1000 steps
	{
		CK(cudaGraphCreate(&g, 0));
		cudaGraphConditionalHandle h;
		CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
		CK(cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		cudaGraph_t body, cg;
		if(add_if(s, h, cudaGraphCondTypeWhile, &body)) return 1;
		CK(cudaStreamEndCapture(s, &cg));
		CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		step_serial();
		k_loop<<<1, 1, 0, s>>>(ctr, h, 1000);
		CK(cudaStreamEndCapture(s, &cg));
		CK(cudaGraphInstantiate(&ge, g, 0)); time_exec("F  WHILE node, serial step body, 1 launch", ge, 1, 1000);
	}
	// G: WHILE node with fork/join body and nested IF
	{
		CK(cudaGraphCreate(&g, 0));
		cudaGraphConditionalHandle h, h2;
		CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
		CK(cudaGraphConditionalHandleCreate(&h2, g, 0, cudaGraphCondAssignDefault));
		CK(cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		cudaGraph_t body, body2, cg;
		if(add_if(s, h, cudaGraphCondTypeWhile, &body)) return 1;
		CK(cudaStreamEndCapture(s, &cg));
		CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		step_forkjoin();
		k_decide<<<1, 1, 0, s>>>(ctr + 1, h2, 10);
		if(add_if(s, h2, cudaGraphCondTypeIf, &body2)) return 1;
		k_loop<<<1, 1, 0, s>>>(ctr, h, 1000);
		CK(cudaStreamEndCapture(s, &cg));
		CK(cudaStreamBeginCaptureToGraph(s, body2, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		k_count<<<1, 1, 0, s>>>(ctr + 2);
		CK(cudaStreamEndCapture(s, &cg));
		CK(cudaGraphInstantiate(&ge, g, 0)); time_exec("G  WHILE{fork/join + decide + IF{..}}", ge, 1, 1000);
	}
	printf("OK\n");
	return 0;
}
