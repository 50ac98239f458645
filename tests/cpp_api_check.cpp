// Compile-and-link check of include/fruid.hpp (the C++ mirror of the crate API) against
// libfruid_b200.so.  With a GPU it also runs one step; without one it expects the loud error.
#include <cstdio>
#include "../include/fruid.hpp"
int main() {
    try {
        fruid::FluidSim sim(64, 48);
        fruid::DensitySim c(sim.width(), sim.height());
        auto [u0, v0] = sim.uv_mut();
        u0[{32, 24}] = 100.0f;
        v0[{32, 24}] = -50.0f;
        c.density_mut()[{20, 24}] = 80.0f * 64 * 48;
        sim.step(0.1f, 0.0f);
        c.step(sim, 0.1f, 0.0f);
        const unsigned long long up0 = fruid_host_velocity_uploads();
        c.step(sim.uv(), 1e-2f, 0.0f);  // the reference's own call shape (main.rs:115): must not upload u, v
        if (fruid_host_velocity_uploads() != up0) { std::printf("c.step(sim.uv(), ..) uploaded the velocity\n"); return 1; }
        fruid::DensitySim m(sim.width(), sim.height());
        fruid::DensitySim::step_all({&c, &m}, sim, 1e-2f, 0.0f);  // main.rs:114-118, batched
        auto [u, v] = sim.uv();
        double s = 0;
        for (float x : u.data()) s += x;
        std::printf("ok sum_u=%g dens=%g\n", s, (double)c.density()[{20, 24}]);
        return 0;
    } catch (const fruid::Error& e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return e.code == FRUID_ERR_CUDA ? 3 : 1;
    }
}
