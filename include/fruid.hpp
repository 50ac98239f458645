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
#include <algorithm>
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

// Host <-> device traffic of the mirrors in this process (bytes and calls), for the end-to-end bench.
struct Traffic { unsigned long long h2d_bytes = 0, d2h_bytes = 0, h2d_calls = 0, d2h_calls = 0, injects = 0, fills = 0; };
inline Traffic& traffic() { static Traffic t; return t; }

class FluidSim;
class DensitySim;

// idek_basics::Array2D<f32>: width + flat Vec<f32>, index (x, y) -> x + y*width.
//
// As a mirror of a device field it tracks HOW it was written so that the reference's frame loop (main.rs:94-108)
// does not move whole fields over PCIe:
//   a[{x, y}] = v            remembered as a single-cell write, forwarded as an injection (no download, no upload)
//   a.data_mut().fill(v)     remembered as a fill, performed on the device
//   any other mutable access (data_mut().data(), iteration, float x = a[{x, y}] on a non-const array)
//                            pulls the device copy if it is newer and marks the whole field dirty (uploaded before the
//                            next step) -- always correct, just not free
// Reads (const operator[], data()) download the field if the device copy is newer.
class Array2D {
public:
    Array2D(std::size_t width, std::size_t height) : width_(width), data_(width * height, 0.0f) {}
    std::size_t width() const { return width_; }
    std::size_t height() const { return width_ ? data_.size() / width_ : 0; }
    const std::vector<float>& data() const { pull(); return data_; }

    class CellRef {  // what a[{x, y}] yields on a mutable array
    public:
        CellRef& operator=(float v) { a_.write_cell(i_, v); return *this; }
        operator float() const { a_.pull(); return a_.data_[i_]; }
    private:
        friend class Array2D;
        CellRef(Array2D& a, std::size_t i) : a_(a), i_(i) {}
        Array2D& a_;
        std::size_t i_;
    };
    class DataMut {  // what data_mut() yields: the Vec<f32> of the reference, with fill() kept on the device
    public:
        void fill(float v) { a_.write_fill(v); }
        float* data() { a_.whole(); return a_.data_.data(); }
        float* begin() { return data(); }
        float* end() { return data() + a_.data_.size(); }
        float& operator[](std::size_t i) { a_.whole(); return a_.data_.at(i); }
        std::size_t size() const { return a_.data_.size(); }
    private:
        friend class Array2D;
        explicit DataMut(Array2D& a) : a_(a) {}
        Array2D& a_;
    };
    DataMut data_mut() { return DataMut(*this); }
    float operator[](std::pair<std::size_t, std::size_t> p) const { pull(); return data_.at(p.first + p.second * width_); }
    CellRef operator[](std::pair<std::size_t, std::size_t> p) {
        if (p.first >= width_ || p.second >= height()) throw std::out_of_range("Array2D index");  // Rust panics
        return CellRef(*this, p.first + p.second * width_);
    }

private:
    friend class FluidSim;
    friend class DensitySim;
    struct Link {  // how this mirror reaches its device field (set by the owning simulation)
        int (*download)(void* h, int field, float* host) = nullptr;
        int (*upload)(void* h, int field, const float* host) = nullptr;
        int (*inject)(void* h, int field, std::size_t x, std::size_t y, float v) = nullptr;
        int (*fill)(void* h, int field, float v) = nullptr;
        void* handle = nullptr;
        int field = 0;
        bool fluid = false;  // the handle is a fruid_fluid_t
    };
    void pull() const {  // device -> host if the device copy is newer
        if (!stale_ || !link_.download) return;
        check(link_.download(link_.handle, link_.field, const_cast<float*>(data_.data())));
        traffic().d2h_bytes += data_.size() * sizeof(float);
        traffic().d2h_calls += 1;
        stale_ = false;
        if (has_fill_) std::fill(const_cast<std::vector<float>&>(data_).begin(), const_cast<std::vector<float>&>(data_).end(), fill_);
        for (const auto& c : cells_) const_cast<std::vector<float>&>(data_)[c.first] = c.second;  // not-yet-pushed writes stay visible
    }
    void whole() { pull(); dirty_ = true; has_fill_ = false; cells_.clear(); }
    void write_cell(std::size_t i, float v) {
        if (!stale_) data_[i] = v;
        if (!dirty_) cells_.emplace_back(i, v);
    }
    void write_fill(float v) {
        std::fill(data_.begin(), data_.end(), v);
        stale_ = false; dirty_ = false; has_fill_ = true; fill_ = v; cells_.clear();
    }
    void push() {  // host -> device what was written since the last step
        if (dirty_) {
            check(link_.upload(link_.handle, link_.field, data_.data()));
            traffic().h2d_bytes += data_.size() * sizeof(float);
            traffic().h2d_calls += 1;
        } else {
            if (has_fill_) { check(link_.fill(link_.handle, link_.field, fill_)); traffic().fills += 1; }
            for (const auto& c : cells_) {
                check(link_.inject(link_.handle, link_.field, c.first % width_, c.first / width_, c.second));
                traffic().injects += 1;
                traffic().h2d_bytes += sizeof(float);
            }
        }
        dirty_ = false; has_fill_ = false; cells_.clear();
    }
    void mark_stale() { stale_ = true; dirty_ = false; has_fill_ = false; cells_.clear(); }

    std::size_t width_;
    std::vector<float> data_;
    Link link_;
    mutable bool stale_ = false;  // device newer than host
    bool dirty_ = false;          // host newer than device (whole field)
    bool has_fill_ = false;
    float fill_ = 0.0f;
    std::vector<std::pair<std::size_t, float>> cells_;
};

class FluidSim {
public:
    FluidSim(std::size_t width, std::size_t height)  // FluidSim::new, lib.rs:256
        : u_(width, height), v_(width, height), u_prev_(width, height), v_prev_(width, height) {
        check(fruid_fluid_create(width, height, &h_));
        int f = 0;
        for (Array2D* a : {&u_, &v_, &u_prev_, &v_prev_}) {
            a->link_.download = [](void* h, int fl, float* p) { return fruid_fluid_download((fruid_fluid_t*)h, fl, p); };
            a->link_.upload = [](void* h, int fl, const float* p) { return fruid_fluid_upload((fruid_fluid_t*)h, fl, p); };
            a->link_.inject = [](void* h, int fl, std::size_t x, std::size_t y, float v) { return fruid_fluid_inject((fruid_fluid_t*)h, fl, x, y, v); };
            a->link_.fill = [](void* h, int fl, float v) { return fruid_fluid_fill((fruid_fluid_t*)h, fl, v); };
            a->link_.handle = h_;
            a->link_.field = f++;
            a->link_.fluid = true;
        }
        // uv() hands out these two arrays: c.step(sim.uv(), ..) then finds the velocity on the device
        check(fruid_fluid_bind_mirror(h_, FRUID_U, u_.data_.data()));
        check(fruid_fluid_bind_mirror(h_, FRUID_V, v_.data_.data()));
    }
    ~FluidSim() { fruid_fluid_destroy(h_); }
    FluidSim(const FluidSim&) = delete;
    FluidSim& operator=(const FluidSim&) = delete;

    void step(float dt, float visc) {  // lib.rs:267
        for (Array2D* a : {&u_, &v_, &u_prev_, &v_prev_}) a->push();
        check(fruid_fluid_step(h_, dt, visc));
        for (Array2D* a : {&u_, &v_, &u_prev_, &v_prev_}) a->mark_stale();
    }
    // lib.rs:279.  The arrays are downloaded when (and if) they are read.
    std::pair<const Array2D&, const Array2D&> uv() const { return {u_, v_}; }
    // lib.rs:283: (u_prev, v_prev) -- the next step's sources; after a step they hold pressure and divergence.
    std::pair<Array2D&, Array2D&> uv_mut() { return {u_prev_, v_prev_}; }
    std::size_t width() const { return u_.width(); }    // lib.rs:287
    std::size_t height() const { return u_.height(); }  // lib.rs:291

    fruid_fluid_t* handle() { return h_; }
    bool owns(const Array2D& u, const Array2D& v) const { return &u == &u_ && &v == &v_; }

private:
    Array2D u_, v_, u_prev_, v_prev_;
    fruid_fluid_t* h_ = nullptr;
};

class DensitySim {
public:
    DensitySim(std::size_t width, std::size_t height) : dens_(width, height), dens_prev_(width, height) {  // lib.rs:217
        check(fruid_density_create(width, height, &h_));
        int f = 0;
        for (Array2D* a : {&dens_, &dens_prev_}) {
            a->link_.download = [](void* h, int fl, float* p) { return fruid_density_download((fruid_density_t*)h, fl, p); };
            a->link_.upload = [](void* h, int fl, const float* p) { return fruid_density_upload((fruid_density_t*)h, fl, p); };
            a->link_.inject = [](void* h, int fl, std::size_t x, std::size_t y, float v) { return fruid_density_inject((fruid_density_t*)h, fl, x, y, v); };
            a->link_.fill = [](void* h, int fl, float v) { return fruid_density_fill((fruid_density_t*)h, fl, v); };
            a->link_.handle = h_;
            a->link_.field = f++;
        }
    }
    ~DensitySim() { fruid_density_destroy(h_); }
    DensitySim(const DensitySim&) = delete;
    DensitySim& operator=(const DensitySim&) = delete;

    // DensitySim::step((u, v), dt, diff), lib.rs:226, with the velocity resident on the device.
    void step(FluidSim& vel, float dt, float diff) {
        push();
        check(fruid_density_step_dev(h_, vel.handle(), dt, diff));
        dens_.mark_stale(); dens_prev_.mark_stale();
    }
    // The literal reference signature, c.step(sim.uv(), dt, diff) (lib.rs:226, main.rs:115).  When the pair is the
    // (u, v) mirror pair of a FluidSim -- the arrays know which device field they shadow -- the step runs on the
    // device-resident velocity and nothing is downloaded or uploaded; any other pair of arrays is uploaded.
    void step(std::pair<const Array2D&, const Array2D&> uv, float dt, float diff) {
        const Array2D::Link &lu = uv.first.link_, &lv = uv.second.link_;
        push();
        if (lu.fluid && lv.fluid && lu.handle == lv.handle && lu.field == FRUID_U && lv.field == FRUID_V) {
            check(fruid_density_step_dev(h_, (fruid_fluid_t*)lu.handle, dt, diff));
        } else {
            check(fruid_density_step_host(h_, uv.first.data().data(), uv.second.data().data(), dt, diff));
            traffic().h2d_bytes += 2 * uv.first.data().size() * sizeof(float);
            traffic().h2d_calls += 2;
        }
        dens_.mark_stale(); dens_prev_.mark_stale();
    }
    // c.step(sim.uv(), ..); m.step(sim.uv(), ..); ... (main.rs:114-118) as one batched call.
    static void step_all(std::initializer_list<DensitySim*> dyes, FluidSim& vel, float dt, float diff) {
        std::vector<fruid_density_t*> h;
        for (DensitySim* d : dyes) { d->push(); h.push_back(d->h_); }
        check(fruid_density_step_dev_batch(h.data(), (int)h.size(), vel.handle(), dt, diff));
        for (DensitySim* d : dyes) { d->dens_.mark_stale(); d->dens_prev_.mark_stale(); }
    }
    const Array2D& density() const { return dens_; }   // lib.rs:238 (downloaded when read)
    Array2D& density_mut() { return dens_prev_; }      // lib.rs:242: dens_prev (the source field), not dens
    fruid_density_t* handle() { return h_; }

private:
    void push() { dens_.push(); dens_prev_.push(); }
    Array2D dens_, dens_prev_;
    fruid_density_t* h_ = nullptr;
};

}  // namespace fruid
