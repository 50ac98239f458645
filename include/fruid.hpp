// fruid.hpp -- C++ mirror of the reference crate API (src/lib.rs:1, 210-294) over the C ABI
// of fruid_b200.h.  The reference is compiled Rust and no Rust toolchain exists in the build
// image, so this header is the compiled-language host side of the drop-in boundary: same
// type names, method names, argument order and semantics as the Rust types
//
//   pub type Array2D = idek_basics::Array2D<f32>;                      lib.rs:1
//   DensitySim::{new, step((u, v), dt, diff), density, density_mut}    lib.rs:216-245
//   FluidSim::{new, step(dt, visc), uv, uv_mut, width, height}         lib.rs:255-294
//
// Host-mirror semantics (what a Rust shim implements the same way):
//   step()        uploads u_prev/v_prev if the host copy was handed out mutably, runs the
//                 device step, marks all host copies stale
//   uv()          downloads u, v if stale
//   uv_mut()      downloads u_prev, v_prev if stale (they hold pressure / divergence after a
//                 step, lib.rs:207) and marks them dirty
//   DensitySim::step((u, v), ..) uses the device-resident velocity when (u, v) are the
//                 mirrors of a live FluidSim (the c.step(sim.uv(), ..) pattern of main.rs:115)
// The reference API is infallible; like the Rust shim (panic!), this mirror throws on error.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <initializer_list>
#include <vector>

#include "fruid_b200.h"

namespace fruid {

struct Error : std::runtime_error {
    int code;
    Error(int c, const char* msg) : std::runtime_error(std::string("fruid_b200: ") + msg), code(c) {}
};
inline void check(int code) {
    if (code != FRUID_OK) throw Error(code, fruid_last_error());
}

// idek_basics::Array2D<f32>: width + flat Vec<f32>, index (x, y) -> x + y*width.
class Array2D {
public:
    Array2D(std::size_t width, std::size_t height) : width_(width), data_(width * height, 0.0f) {}
    std::size_t width() const { return width_; }
    std::size_t height() const { return width_ ? data_.size() / width_ : 0; }
    const std::vector<float>& data() const { return data_; }
    std::vector<float>& data_mut() { return data_; }
    float operator[](std::pair<std::size_t, std::size_t> p) const { return data_.at(p.first + p.second * width_); }
    float& operator[](std::pair<std::size_t, std::size_t> p) { return data_.at(p.first + p.second * width_); }

private:
    std::size_t width_;
    std::vector<float> data_;
};

class DensitySim;

class FluidSim {
public:
    FluidSim(std::size_t width, std::size_t height)  // FluidSim::new, lib.rs:256
        : u_(width, height), v_(width, height), u_prev_(width, height), v_prev_(width, height) {
        check(fruid_fluid_create(width, height, &h_));
        // uv() hands out these two arrays: c.step(sim.uv(), ..) then finds the velocity on the device
        check(fruid_fluid_bind_mirror(h_, FRUID_U, u_.data().data()));
        check(fruid_fluid_bind_mirror(h_, FRUID_V, v_.data().data()));
    }
    ~FluidSim() { fruid_fluid_destroy(h_); }
    FluidSim(const FluidSim&) = delete;
    FluidSim& operator=(const FluidSim&) = delete;

    void step(float dt, float visc) {  // lib.rs:267
        if (prev_dirty_) {
            check(fruid_fluid_upload(h_, FRUID_U_PREV, u_prev_.data().data()));
            check(fruid_fluid_upload(h_, FRUID_V_PREV, v_prev_.data().data()));
            prev_dirty_ = false;
        }
        check(fruid_fluid_step(h_, dt, visc));
        uv_stale_ = prev_stale_ = true;
    }
    std::pair<const Array2D&, const Array2D&> uv() {  // lib.rs:279
        if (uv_stale_) {
            check(fruid_fluid_download(h_, FRUID_U, u_.data_mut().data()));
            check(fruid_fluid_download(h_, FRUID_V, v_.data_mut().data()));
            uv_stale_ = false;
        }
        return {u_, v_};
    }
    std::pair<Array2D&, Array2D&> uv_mut() {  // lib.rs:283: (u_prev, v_prev)
        if (prev_stale_) {
            check(fruid_fluid_download(h_, FRUID_U_PREV, u_prev_.data_mut().data()));
            check(fruid_fluid_download(h_, FRUID_V_PREV, v_prev_.data_mut().data()));
            prev_stale_ = false;
        }
        prev_dirty_ = true;
        return {u_prev_, v_prev_};
    }
    std::size_t width() const { return u_.width(); }    // lib.rs:287
    std::size_t height() const { return u_.height(); }  // lib.rs:291

    fruid_fluid_t* handle() { return h_; }
    bool owns(const Array2D& u, const Array2D& v) const { return &u == &u_ && &v == &v_; }

private:
    Array2D u_, v_, u_prev_, v_prev_;
    fruid_fluid_t* h_ = nullptr;
    bool uv_stale_ = false, prev_stale_ = false, prev_dirty_ = false;
};

class DensitySim {
public:
    DensitySim(std::size_t width, std::size_t height) : dens_(width, height), dens_prev_(width, height) {  // lib.rs:217
        check(fruid_density_create(width, height, &h_));
    }
    ~DensitySim() { fruid_density_destroy(h_); }
    DensitySim(const DensitySim&) = delete;
    DensitySim& operator=(const DensitySim&) = delete;

    // DensitySim::step((u, v), dt, diff), lib.rs:226, with the velocity resident on the device.
    void step(FluidSim& vel, float dt, float diff) {
        push();
        check(fruid_density_step_dev(h_, vel.handle(), dt, diff));
        dens_stale_ = prev_stale_ = true;
    }
    // The literal reference signature, c.step(sim.uv(), dt, diff) (lib.rs:226, main.rs:115): when the pair is a
    // FluidSim's current uv() the library uses the device-resident velocity, otherwise it uploads the arrays.
    void step(std::pair<const Array2D&, const Array2D&> uv, float dt, float diff) {
        push();
        check(fruid_density_step_host(h_, uv.first.data().data(), uv.second.data().data(), dt, diff));
        dens_stale_ = prev_stale_ = true;
    }
    // c.step(sim.uv(), ..); m.step(sim.uv(), ..); ... (main.rs:114-118) as one batched call.
    static void step_all(std::initializer_list<DensitySim*> dyes, FluidSim& vel, float dt, float diff) {
        std::vector<fruid_density_t*> h;
        for (DensitySim* d : dyes) { d->push(); h.push_back(d->h_); }
        check(fruid_density_step_dev_batch(h.data(), (int)h.size(), vel.handle(), dt, diff));
        for (DensitySim* d : dyes) d->dens_stale_ = d->prev_stale_ = true;
    }
    const Array2D& density() {  // lib.rs:238
        if (dens_stale_) {
            check(fruid_density_download(h_, FRUID_DENS, dens_.data_mut().data()));
            dens_stale_ = false;
        }
        return dens_;
    }
    Array2D& density_mut() {  // lib.rs:242: dens_prev (the source field), not dens
        if (prev_stale_) {
            check(fruid_density_download(h_, FRUID_DENS_PREV, dens_prev_.data_mut().data()));
            prev_stale_ = false;
        }
        prev_dirty_ = true;
        return dens_prev_;
    }

private:
    void push() {
        if (prev_dirty_) {
            check(fruid_density_upload(h_, FRUID_DENS_PREV, dens_prev_.data().data()));
            prev_dirty_ = false;
        }
    }
    Array2D dens_, dens_prev_;
    fruid_density_t* h_ = nullptr;
    bool dens_stale_ = false, prev_stale_ = false, prev_dirty_ = false;
};

}  // namespace fruid
