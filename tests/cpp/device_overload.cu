// Test program of include/connect6_device.hpp (built by tests/test_abi.py::build_device_overload with nvcc against the
// reference's own connect6.hpp in place): n games of self-play whose evaluator is a CUDA kernel reading the leaf batch in HBM.
// The evaluator is the constant one of the hand-derived trace (value 0.5, prior 0.25: oracle const_eval_callback), so the
// Python test can compare every stone with the oracle.  Prints "GAME <winner> <cell> <cell> ..." per game.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "connect6.hpp"
#include "connect6_device.hpp"

static int g_calls = 0, g_boards = 0;

__global__ void const_eval(float* values, float* policies, int len, int* bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len * 225) {
        const float b = policies[i];
        if (b != -1.f && b != 0.f && b != 1.f) atomicAdd(bad, 1);  // the batch must hold boards on entry
        policies[i] = 0.25f;
    }
    if (i < len) {
        if (values[i] != 0.f) atomicAdd(bad, 1);  // values arrive pre-zeroed (cppeval/mod.rs:34)
        values[i] = 0.5f;
    }
}

static void evaluate(int player, float* d_values, float* d_policies, int len, void* stream, void* user) {
    if (player != -1 && player != 1) std::abort();
    g_calls += 1, g_boards += len;
    const int n = len * 225;
    const_eval<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_values, d_policies, len, static_cast<int*>(user));
}

int main(int argc, char** argv) {
    const int games = argc > 1 ? std::atoi(argv[1]) : 3, sims = argc > 2 ? std::atoi(argv[2]) : 4;
    const unsigned seed = argc > 3 ? (unsigned)std::strtoul(argv[3], nullptr, 0) : 0xC6u;
    int* bad = nullptr;
    cudaMalloc(&bad, sizeof(int));
    cudaMemset(bad, 0, sizeof(int));
    Connect6::DeviceOptions opt;
    opt.seed = seed, opt.log2_records = 14;
    auto results = Connect6::self_play(evaluate, bad, Connect6::Param().NumSimulation(sims).NumGameThread(games), opt);
    int h_bad = -1;
    cudaMemcpy(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost);
    for (const auto& g : results) {
        std::printf("GAME %d", static_cast<int>(g.GetWinner()));
        for (const auto& p : g) std::printf(" %d", (int)(std::get<0>(p.GetPos()) * 15 + std::get<1>(p.GetPos())));
        std::printf("\n");
    }
    std::printf("CALLS %d BOARDS %d BAD %d\n", g_calls, g_boards, h_bad);
    return h_bad == 0 ? 0 : 1;
}
