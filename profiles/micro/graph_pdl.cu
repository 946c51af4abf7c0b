// Does programmatic dependent launch survive stream capture? A chain of tiny kernels launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, eager vs captured graph: edge types and time.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void tiny(int* p, int pdl) {
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
        asm volatile("griddepcontrol.wait;\n" ::: "memory");
    }
    if (threadIdx.x == 0) p[blockIdx.x] += 1;
}

static void launch(cudaStream_t st, int* p, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(8); cfg.blockDim = dim3(128); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, tiny, p, pdl ? 1 : 0);
}

int main() {
    int* p; cudaMalloc(&p, 1024); cudaMemset(p, 0, 1024);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int N = 40, REP = 50;
    for (int pdl = 0; pdl < 2; pdl++) {
        for (int w = 0; w < 3; w++) { for (int i = 0; i < N; i++) launch(st, p, pdl); }
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int r = 0; r < REP; r++) for (int i = 0; i < N; i++) launch(st, p, pdl);
        cudaEventRecord(e1, st); cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("eager pdl=%d: %.2f us per kernel\n", pdl, 1000 * ms / (REP * N));
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
        for (int i = 0; i < N; i++) launch(st, p, pdl);
        cudaError_t e = cudaStreamEndCapture(st, &g);
        printf("capture: %s\n", cudaGetErrorString(e));
        size_t ne = 0; cudaGraphGetEdges_v2(g, nullptr, nullptr, nullptr, &ne);
        std::vector<cudaGraphNode_t> from(ne), to(ne); std::vector<cudaGraphEdgeData> ed(ne);
        cudaGraphGetEdges_v2(g, from.data(), to.data(), ed.data(), &ne);
        int prog = 0; for (auto& d : ed) prog += d.type == cudaGraphDependencyTypeProgrammatic;
        printf("graph pdl=%d: %zu edges, %d programmatic\n", pdl, ne, prog);
        e = cudaGraphInstantiate(&ge, g, 0);
        printf("instantiate: %s\n", cudaGetErrorString(e));
        for (int w = 0; w < 3; w++) cudaGraphLaunch(ge, st);
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int r = 0; r < REP; r++) cudaGraphLaunch(ge, st);
        cudaEventRecord(e1, st); cudaStreamSynchronize(st);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("graph pdl=%d: %.2f us per kernel\n", pdl, 1000 * ms / (REP * N));
    }
    return 0;
}
