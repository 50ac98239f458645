// frame_e2e.cpp -- the reference viewer's frame loop (src/main.rs:36-52 setup, 89-124 frame), headless, through the
// C++ mirror of the crate API (include/fruid.hpp): the end-to-end leg of bench.py for the scene configurations.
// Every frame: one-cell writes through uv_mut(), density_mut().data_mut().fill(0) on the four dyes, sim.step,
// c/m/y/k.step(sim.uv(), ..), then what draw_density (main.rs:124,146-183) needs -- here the RGB of every cell,
// computed on the device and copied to the host.  Wall clock per frame, synchronous like the viewer's loop.
//
//   frame_e2e <width> <height> <frames> <warmup> [batched=1] [dump.bin]
// prints one JSON object: microseconds per frame and the host<->device traffic per frame the mirror caused.
// With a dump path the fields after the last frame are written as raw float32 (u, v, u_prev, v_prev, then dens and
// dens_prev of c, m, y, k -- the order of fruid_b200.scene.Scene) for the parity test; use warmup = 0 with it.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/fruid.hpp"

using namespace fruid;

int main(int argc, char** argv) {
    const std::size_t w = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 250;
    const std::size_t h = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 250;
    const int frames = argc > 3 ? std::atoi(argv[3]) : 100;
    const int warmup = argc > 4 ? std::atoi(argv[4]) : 3;
    const bool batched = argc > 5 ? std::atoi(argv[5]) != 0 : true;
    const char* dump = argc > 6 ? argv[6] : nullptr;
    try {
        FluidSim sim(w, h);                                        // main.rs:36
        DensitySim c(w, h), m(w, h), y(w, h), k(w, h);             // main.rs:38-41
        const float intensity = 80.0f * (float)(w * h);            // main.rs:42
        c.density_mut()[{w / 5, h / 2}] = intensity;               // main.rs:43-46
        k.density_mut()[{2 * w / 5, h / 2}] = intensity / 100.0f;
        m.density_mut()[{3 * w / 5, h / 2}] = intensity;
        y.density_mut()[{4 * w / 5, h / 2}] = intensity;
        auto step_dyes = [&](float dt, float diff) {
            if (batched) {
                DensitySim::step_all({&c, &m, &y, &k}, sim, dt, diff);
            } else {
                c.step(sim.uv(), dt, diff);                        // main.rs:115-118
                m.step(sim.uv(), dt, diff);
                y.step(sim.uv(), dt, diff);
                k.step(sim.uv(), dt, diff);
            }
        };
        sim.step(0.1f, 0.0f);                                      // main.rs:48
        step_dyes(0.1f, 0.0f);                                     // main.rs:49-52
        std::vector<float> rgb(w * h * 3);
        long frame_count = 0;
        auto frame = [&]() {
            frame_count += 1;                                      // main.rs:91
            const float time = (float)frame_count / 12.0f;
            const float x = (float)(w / 2) * (std::cos(time / 5.0f) + 1.0f);
            std::size_t px = (std::size_t)x;
            if (px >= w) px = w - 1;                               // the reference panics here (frame 377); keep running
            auto [u, v] = sim.uv_mut();                            // main.rs:98
            u[{px, h / 2}] = -4500.0f * std::cos(time * 3.0f);     // main.rs:101-102
            v[{px, h / 2}] = -4500.0f * std::sin(time * 3.0f);
            c.density_mut().data_mut().fill(0.0f);                 // main.rs:105-108
            m.density_mut().data_mut().fill(0.0f);
            y.density_mut().data_mut().fill(0.0f);
            k.density_mut().data_mut().fill(0.0f);
            sim.step(1e-2f, 0.0f);                                 // main.rs:114
            step_dyes(1e-2f, 0.0f);
            check(fruid_density_colors(c.handle(), m.handle(), y.handle(), k.handle(), rgb.data()));  // main.rs:124
            traffic().d2h_bytes += rgb.size() * sizeof(float);
            traffic().d2h_calls += 1;
        };
        for (int i = 0; i < warmup; ++i) frame();
        const Traffic t0 = traffic();
        const auto a = std::chrono::steady_clock::now();
        for (int i = 0; i < frames; ++i) frame();
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count() / frames;
        const Traffic t1 = traffic();
        if (dump) {
            std::FILE* f = std::fopen(dump, "wb");
            if (!f) { std::printf("{\"error\": \"cannot open %s\"}\n", dump); return 2; }
            auto put = [&](const Array2D& a) { std::fwrite(a.data().data(), sizeof(float), a.data().size(), f); };
            put(sim.uv().first); put(sim.uv().second);
            { auto [p, dv] = sim.uv_mut(); put(p); put(dv); }
            for (DensitySim* d : {&c, &m, &y, &k}) { put(d->density()); put(d->density_mut()); }
            std::fclose(f);
        }
        bool finite = true;
        for (std::size_t i = 0; i < rgb.size(); i += 97) finite = finite && std::isfinite(rgb[i]);
        std::printf("{\"us_per_frame\": %.3f, \"frames\": %d, \"width\": %zu, \"height\": %zu, \"batched\": %s, \"finite\": %s, "
                    "\"h2d_bytes_per_frame\": %.1f, \"d2h_bytes_per_frame\": %.1f, \"h2d_calls_per_frame\": %.2f, "
                    "\"d2h_calls_per_frame\": %.2f, \"injects_per_frame\": %.2f, \"device_fills_per_frame\": %.2f}\n",
                    us, frames, w, h, batched ? "true" : "false", finite ? "true" : "false",
                    (double)(t1.h2d_bytes - t0.h2d_bytes) / frames, (double)(t1.d2h_bytes - t0.d2h_bytes) / frames,
                    (double)(t1.h2d_calls - t0.h2d_calls) / frames, (double)(t1.d2h_calls - t0.d2h_calls) / frames,
                    (double)(t1.injects - t0.injects) / frames, (double)(t1.fills - t0.fills) / frames);
        return 0;
    } catch (const Error& e) {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 2;
    }
}
