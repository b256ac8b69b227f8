// Experiment (round 2): can the kernel nodes inside a conditional WHILE body be re-parameterised in an instantiated
// graph (cudaGraphExecKernelNodeSetParams / cudaGraphExecUpdate), and what do capture / instantiate / update cost on the host?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cond_update_test cond_update_test.cu && ./cond_update_test
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Args { int* counter; int* out; int n; double pad[40]; };

__global__ void body_a(Args a) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < a.n) a.out[i] += 1; }
__global__ void body_b(Args a, cudaGraphConditionalHandle h) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { int c = --(*a.counter); cudaGraphSetConditional(h, c > 0 ? 1u : 0u); }
}

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int build(cudaStream_t st, const Args& a, int grid, cudaGraph_t* g_out, cudaGraph_t* body_out, cudaGraphConditionalHandle* h_out) {
    cudaGraph_t g;
    CK(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h;
    CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = h;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, g, nullptr, 0, &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    for (int k = 0; k < 4; k++) body_a<<<grid, 256, 0, st>>>(a);
    body_b<<<1, 32, 0, st>>>(a, h);
    cudaGraph_t cap;
    CK(cudaStreamEndCapture(st, &cap));
    *g_out = g; *body_out = body; *h_out = h;
    return 0;
}

int main() {
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const int N1 = 100000, N2 = 150000;
    int *c1, *c2, *o1, *o2;
    CK(cudaMalloc(&c1, 4)); CK(cudaMalloc(&c2, 4)); CK(cudaMalloc(&o1, N2 * 4)); CK(cudaMalloc(&o2, N2 * 4));
    CK(cudaMemset(o1, 0, N2 * 4)); CK(cudaMemset(o2, 0, N2 * 4));
    Args a1{c1, o1, N1}, a2{c2, o2, N2};
    cudaGraph_t g, body; cudaGraphConditionalHandle h;
    double t0 = now_us();
    if (build(st, a1, (N1 + 255) / 256, &g, &body, &h)) return 1;
    double t1 = now_us();
    cudaGraphExec_t ex;
    CK(cudaGraphInstantiate(&ex, g, 0));
    double t2 = now_us();
    printf("capture %.1f us, instantiate %.1f us\n", t1 - t0, t2 - t1);
    int five = 5;
    CK(cudaMemcpy(c1, &five, 4, cudaMemcpyHostToDevice));
    CK(cudaGraphLaunch(ex, st)); CK(cudaStreamSynchronize(st));
    int v; CK(cudaMemcpy(&v, o1 + N1 - 1, 4, cudaMemcpyDeviceToHost));
    printf("run 1: out[last] = %d (expect 20)\n", v);

    // (1) per-node parameter update inside the body
    size_t cnt = 0;
    CK(cudaGraphGetNodes(body, nullptr, &cnt));
    std::vector<cudaGraphNode_t> nodes(cnt);
    CK(cudaGraphGetNodes(body, nodes.data(), &cnt));
    printf("body has %zu nodes\n", cnt);
    double t3 = now_us();
    int ok = 1;
    for (size_t k = 0; k < cnt; k++) {
        cudaKernelNodeParams kp;
        CK(cudaGraphKernelNodeGetParams(nodes[k], &kp));
        void* pa[2] = {&a2, &h};
        kp.kernelParams = pa;
        if (kp.func == (void*)body_a) kp.gridDim = dim3((N2 + 255) / 256);
        cudaError_t e = cudaGraphExecKernelNodeSetParams(ex, nodes[k], &kp);
        if (e != cudaSuccess) { printf("ExecKernelNodeSetParams node %zu -> %s\n", k, cudaGetErrorString(e)); ok = 0; cudaGetLastError(); break; }
    }
    double t4 = now_us();
    printf("SetParams on %zu body nodes: %s, %.1f us\n", cnt, ok ? "OK" : "FAILED", t4 - t3);
    if (ok) {
        int three = 3;
        CK(cudaMemcpy(c2, &three, 4, cudaMemcpyHostToDevice));
        CK(cudaGraphLaunch(ex, st)); CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(&v, o2 + N2 - 1, 4, cudaMemcpyDeviceToHost));
        printf("run 2 (updated params): out2[last] = %d (expect 12)\n", v);
    }
    // (2) whole-graph update from a re-captured graph
    cudaGraph_t g2, body2; cudaGraphConditionalHandle h2;
    double t5 = now_us();
    if (build(st, a2, (N2 + 255) / 256, &g2, &body2, &h2)) return 1;
    double t6 = now_us();
    cudaGraphExecUpdateResultInfo info;
    cudaError_t e = cudaGraphExecUpdate(ex, g2, &info);
    double t7 = now_us();
    printf("re-capture %.1f us, ExecUpdate %.1f us -> %s (result %d)\n", t6 - t5, t7 - t6, cudaGetErrorString(e), (int)info.result);
    cudaGetLastError();
    if (e == cudaSuccess) {
        int two = 2;
        CK(cudaMemset(o2, 0, N2 * 4));
        CK(cudaMemcpy(c2, &two, 4, cudaMemcpyHostToDevice));
        CK(cudaGraphLaunch(ex, st)); CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(&v, o2 + N2 - 1, 4, cudaMemcpyDeviceToHost));
        printf("run 3 (ExecUpdate): out2[last] = %d (expect 8)\n", v);
    }
    // timing of a full re-create + instantiate cycle, warm
    double acc_c = 0, acc_i = 0;
    for (int r = 0; r < 20; r++) {
        cudaGraph_t g3, b3; cudaGraphConditionalHandle h3; cudaGraphExec_t e3;
        double a = now_us();
        if (build(st, a1, (N1 + 255) / 256, &g3, &b3, &h3)) return 1;
        double b = now_us();
        CK(cudaGraphInstantiate(&e3, g3, 0));
        double c = now_us();
        acc_c += b - a; acc_i += c - b;
        cudaGraphExecDestroy(e3); cudaGraphDestroy(g3);
    }
    printf("warm: capture %.1f us, instantiate %.1f us per graph (5 kernel nodes)\n", acc_c / 20, acc_i / 20);
    double acc_l = 0;
    for (int r = 0; r < 20; r++) {
        int one = 1;
        CK(cudaMemcpyAsync(c2, &one, 4, cudaMemcpyHostToDevice, st));
        double a = now_us();
        CK(cudaGraphLaunch(ex, st));
        acc_l += now_us() - a;
        CK(cudaStreamSynchronize(st));
    }
    printf("graph launch call: %.1f us\n", acc_l / 20);
    return 0;
}
