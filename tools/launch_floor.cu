// Measures the launch-to-launch floor of dependent kernel nodes in a CUDA graph on this GPU:
// an empty kernel, and a kernel that does one dependent DRAM round trip, both with the grid shape
// of the cfg2 step kernel (342 CTAs x 128 threads). Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o launch_floor tools/launch_floor.cu ; evidence for profiles/README.md.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void empty_kernel(float* p) {
    if (p == nullptr) return;
}
__global__ void touch_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + 1.0f;
}

static float run_graph(cudaStream_t s, int nodes, int reps, int which, float* a, float* b, size_t n, size_t ring) {
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < nodes; ++i) {
        size_t off = (size_t)(i % ring) * n;
        if (which == 0) empty_kernel<<<342, 128, 0, s>>>(a);
        else touch_kernel<<<342, 128, 0, s>>>(a + off, b + off, n);
    }
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
    return ms * 1000.0f / (nodes * reps);
}

int main() {
    cudaStream_t s;
    cudaStreamCreate(&s);
    const size_t n = 342 * 128, ring = 2048;   // ring * n * 4 B * 2 = 717 MB > L2
    float *a, *b;
    cudaMalloc(&a, n * ring * sizeof(float));
    cudaMalloc(&b, n * ring * sizeof(float));
    cudaMemset(a, 0, n * ring * sizeof(float));
    printf("empty kernel, graph of 1024 nodes:            %.3f us per node\n", run_graph(s, 1024, 20, 0, a, b, n, ring));
    printf("one DRAM round trip (cold ring), 1024 nodes:  %.3f us per node\n", run_graph(s, 1024, 20, 1, a, b, n, ring));
    printf("one round trip, L2 resident (ring 1):         %.3f us per node\n", run_graph(s, 1024, 20, 1, a, b, n, 1));
    return 0;
}
