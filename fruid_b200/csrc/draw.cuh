// draw.cuh -- device-side draw_density (reference src/main.rs:146-183), the first step past the
// solver on the caller's side: CMYK dye fields -> RGB vertex colours
//   total = sqrt(c^2 + m^2 + y^2);  col = log2((1 - f - k) / max(total, 1) + 2)   for f in c, m, y
// and the four vertices {pos[3], color[3]} the viewer pushes per cell, in the reference's order
// (i outer, j inner; tl, tr, bl, br), so that only the finished vertex buffer crosses PCIe instead
// of four full fields.  Positions are bit-exact; log2f is CUDA's (<= 1 ulp), glibc's on the CPU.
#pragma once
#include "common.cuh"

namespace fruid {

__device__ __forceinline__ float3 dye_color(float c, float m, float y, float k) {
    const float sum = fadd(fadd(fadd(0.0f, fmul(c, c)), fmul(m, m)), fmul(y, y));
    const float total = __fsqrt_rn(sum);
    const float denom = (total > 1.0f) ? total : 1.0f;  // f32::max(total, 1.): NaN -> 1
    float3 o;
    o.x = log2f(fadd(fdiv(fsub(fsub(1.0f, c), k), denom), 2.0f));
    o.y = log2f(fadd(fdiv(fsub(fsub(1.0f, m), k), denom), 2.0f));
    o.z = log2f(fadd(fdiv(fsub(fsub(1.0f, y), k), denom), 2.0f));
    return o;
}

// 32 x 32 cell tiles: fields are read with x fastest (coalesced), vertices are written with j
// fastest because the reference's push order is i outer / j inner (cell index i*h + j).
// VERTS: 24 floats per cell {pos, color} x 4;  otherwise rgb[x + y*w] (3 floats per cell).
template <bool VERTS>
__global__ void __launch_bounds__(256) k_draw_density(const float* __restrict__ c, const float* __restrict__ m,
                                                      const float* __restrict__ y, const float* __restrict__ k, Geom g,
                                                      float z, float cell_w, float cell_h, float* __restrict__ out) {
    __shared__ float3 tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + threadIdx.x, j = j0 + r;
        if (i < g.w && j < g.h) {
            const size_t idx = at(g, i, j);
            const float3 col = dye_color(c[idx], m[idx], y[idx], k[idx]);
            if (VERTS) tile[r][threadIdx.x] = col;
            else {
                float* o = out + ((size_t)i + (size_t)j * g.w) * 3;
                o[0] = col.x; o[1] = col.y; o[2] = col.z;
            }
        }
    }
    if (!VERTS) return;
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, j = j0 + threadIdx.x;  // j fastest across the warp
        if (i >= g.w || j >= g.h) continue;
        const float3 col = tile[threadIdx.x][r];
        const float x0 = fsub(fmul(fdiv((float)i, (float)g.w), 2.0f), 1.0f);   // main.rs:151
        const float y0 = fsub(fmul(fdiv((float)j, (float)g.h), 2.0f), 1.0f);   // main.rs:153
        const float x1 = fadd(x0, cell_w), y1 = fadd(y0, cell_h);
        float4* o = reinterpret_cast<float4*>(out + ((size_t)i * g.h + j) * 24);
        o[0] = make_float4(x0, y0, z, col.x);
        o[1] = make_float4(col.y, col.z, x1, y0);
        o[2] = make_float4(z, col.x, col.y, col.z);
        o[3] = make_float4(x0, y1, z, col.x);
        o[4] = make_float4(col.y, col.z, x1, y1);
        o[5] = make_float4(z, col.x, col.y, col.z);
    }
}

// draw_velocity_lines (src/main.rs:185-216): per cell the line {tail, tip} = 2 vertices {pos[3], color[3]},
// cell index i*h + j like draw_density.  Only +, *, /, sqrt: bit-exact against the CPU restatement.
__global__ void __launch_bounds__(256) k_draw_velocity_lines(const float* __restrict__ u, const float* __restrict__ v, Geom g,
                                                             float z, float cell_w, float cell_h, float* __restrict__ out) {
    __shared__ float2 tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + threadIdx.x, j = j0 + r;
        if (i < g.w && j < g.h) tile[r][threadIdx.x] = make_float2(u[at(g, i, j)], v[at(g, i, j)]);
    }
    __syncthreads();
    const float half_w = fdiv(cell_w, 2.0f), half_h = fdiv(cell_h, 2.0f);
    const float len_num = fmul(cell_h, 2.0f);
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, j = j0 + threadIdx.x;  // j fastest across the warp
        if (i >= g.w || j >= g.h) continue;
        const float2 uv = tile[threadIdx.x][r];
        const float speed = __fsqrt_rn(fadd(fmul(uv.x, uv.x), fmul(uv.y, uv.y)));                 // main.rs:197
        const float tail_x = fadd(fsub(fmul(fdiv((float)i, (float)g.w), 2.0f), 1.0f), half_w);   // main.rs:190,206
        const float tail_y = fadd(fsub(fmul(fdiv((float)j, (float)g.h), 2.0f), 1.0f), half_h);   // main.rs:192,207
        const float len = fdiv(len_num, speed);                                                   // main.rs:210
        const float tip_x = fadd(tail_x, fmul(uv.x, len)), tip_y = fadd(tail_y, fmul(uv.y, len));
        float4* o = reinterpret_cast<float4*>(out + ((size_t)i * g.h + j) * 12);
        o[0] = make_float4(tail_x, tail_y, z, speed);
        o[1] = make_float4(speed, speed, tip_x, tip_y);
        o[2] = make_float4(z, speed, speed, speed);
    }
}

}  // namespace fruid
