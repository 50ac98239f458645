// Microbenchmark: does the packed FADD2/FMUL2 path (sm_100 f32x2) buy issue slots for the
// Jacobi row update?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cuda_runtime.h>
#include <cstdio>
template <int PACKED>
__global__ void __launch_bounds__(512, 2) k(const float4* __restrict__ in, float4* __restrict__ out, float c, int n, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float4 cc = in[i], up = in[i + n], dn = in[i + 2 * n], z = in[i + 3 * n];
    for (int s = 0; s < iters; ++s) {
        float lft = __shfl_up_sync(0xffffffffu, cc.w, 1), rgt = __shfl_down_sync(0xffffffffu, cc.x, 1);
        float t0 = __fadd_rn(lft, cc.y), t1 = __fadd_rn(cc.x, cc.z), t2 = __fadd_rn(cc.y, cc.w), t3 = __fadd_rn(cc.z, rgt);
        float4 nn;
        if (PACKED) {
            float2 s01 = __fadd2_rn(make_float2(t0, t1), make_float2(up.x, up.y));
            float2 s23 = __fadd2_rn(make_float2(t2, t3), make_float2(up.z, up.w));
            s01 = __fadd2_rn(s01, make_float2(dn.x, dn.y));
            s23 = __fadd2_rn(s23, make_float2(dn.z, dn.w));
            s01 = __fadd2_rn(make_float2(z.x, z.y), s01);
            s23 = __fadd2_rn(make_float2(z.z, z.w), s23);
            s01 = __fmul2_rn(s01, make_float2(c, c));
            s23 = __fmul2_rn(s23, make_float2(c, c));
            nn = make_float4(s01.x, s01.y, s23.x, s23.y);
        } else {
            nn.x = __fmul_rn(__fadd_rn(z.x, __fadd_rn(__fadd_rn(t0, up.x), dn.x)), c);
            nn.y = __fmul_rn(__fadd_rn(z.y, __fadd_rn(__fadd_rn(t1, up.y), dn.y)), c);
            nn.z = __fmul_rn(__fadd_rn(z.z, __fadd_rn(__fadd_rn(t2, up.z), dn.z)), c);
            nn.w = __fmul_rn(__fadd_rn(z.w, __fadd_rn(__fadd_rn(t3, up.w), dn.w)), c);
        }
        up = cc;
        cc = nn;
    }
    out[i] = cc;
}
int main() {
    const int n = 148 * 2 * 512 * 4, iters = 4096;
    float4 *in, *out;
    cudaMalloc(&in, sizeof(float4) * n * 4);
    cudaMalloc(&out, sizeof(float4) * n);
    cudaMemset(in, 0, sizeof(float4) * n * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int v = 0; v < 2; ++v) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (v) k<1><<<n / 512, 512>>>(in, out, 0.25f, n, iters); else k<0><<<n / 512, 512>>>(in, out, 0.25f, n, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double rows = (double)n * iters;  // float4-rows per thread-iteration
            printf("%s: %.3f ms  %.1f G float4-rows/s  (%.2f rows/clk/SM @1.9GHz)\n", v ? "packed f32x2" : "scalar", ms, rows / ms / 1e6,
                   rows / 32 / (ms * 1e-3) / 148 / 1.9e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
