//! Drop-in shim: the public API of the reference crate (reference src/lib.rs:1, 210-294) over the
//! C ABI of include/fruid_b200.h.  UNCOMPILED in the build image (no Rust toolchain); kept small
//! and literal on purpose.  `src/main.rs` of the reference builds against it unchanged.
//!
//! Host-mirror rules: the `Array2D`s handed out are host copies.  `step` uploads what was handed
//! out mutably and marks everything stale; `uv()`/`density()` download lazily (interior
//! mutability, so the types are `!Sync` -- a documented auto-trait difference from the reference).
use std::cell::{Cell, UnsafeCell};
use std::os::raw::{c_char, c_int};

pub type Array2D = idek_basics::Array2D<f32>; // reference src/lib.rs:1

#[repr(C)] pub struct FruidFluid { _p: [u8; 0] }
#[repr(C)] pub struct FruidDensity { _p: [u8; 0] }

extern "C" {
    fn fruid_last_error() -> *const c_char;
    fn fruid_fluid_create(w: usize, h: usize, out: *mut *mut FruidFluid) -> c_int;
    fn fruid_fluid_destroy(f: *mut FruidFluid) -> c_int;
    fn fruid_fluid_upload(f: *mut FruidFluid, field: c_int, host: *const f32) -> c_int;
    fn fruid_fluid_download(f: *mut FruidFluid, field: c_int, host: *mut f32) -> c_int;
    fn fruid_fluid_step(f: *mut FruidFluid, dt: f32, visc: f32) -> c_int;
    fn fruid_density_create(w: usize, h: usize, out: *mut *mut FruidDensity) -> c_int;
    fn fruid_density_destroy(d: *mut FruidDensity) -> c_int;
    fn fruid_density_upload(d: *mut FruidDensity, field: c_int, host: *const f32) -> c_int;
    fn fruid_density_download(d: *mut FruidDensity, field: c_int, host: *mut f32) -> c_int;
    fn fruid_density_step_host(d: *mut FruidDensity, u: *const f32, v: *const f32, dt: f32, diff: f32) -> c_int;
    fn fruid_fluid_bind_mirror(f: *mut FruidFluid, field: c_int, host: *const f32) -> c_int;
    fn fruid_host_register(p: *mut std::ffi::c_void, bytes: usize) -> c_int;
    fn fruid_host_unregister(p: *mut std::ffi::c_void) -> c_int;
}

const U: c_int = 0; const V: c_int = 1; const U_PREV: c_int = 2; const V_PREV: c_int = 3;
const DENS: c_int = 0; const DENS_PREV: c_int = 1;

fn check(code: c_int) {
    if code != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(fruid_last_error()) }.to_string_lossy().into_owned();
        panic!("fruid_b200 error {code}: {msg}"); // the reference API is infallible: misuse panics
    }
}

fn pin(a: &mut Array2D) {
    let d = a.data_mut();
    unsafe { fruid_host_register(d.as_mut_ptr() as *mut _, d.len() * 4) }; // best effort
}
fn unpin(a: &mut Array2D) { unsafe { fruid_host_unregister(a.data_mut().as_mut_ptr() as *mut _) }; }

pub struct DensitySim { // reference src/lib.rs:210-214 (scratch lives on the device)
    dens: UnsafeCell<Array2D>,
    dens_prev: Array2D,
    handle: *mut FruidDensity,
    dens_stale: Cell<bool>,
    prev_stale: bool,
    prev_dirty: bool,
}

impl DensitySim {
    pub fn new(width: usize, height: usize) -> Self { // src/lib.rs:217
        let mut handle = std::ptr::null_mut();
        check(unsafe { fruid_density_create(width, height, &mut handle) });
        let mut s = Self {
            dens: UnsafeCell::new(Array2D::new(width, height)), dens_prev: Array2D::new(width, height),
            handle, dens_stale: Cell::new(false), prev_stale: false, prev_dirty: false,
        };
        pin(s.dens.get_mut()); pin(&mut s.dens_prev);
        s
    }

    pub fn step(&mut self, (u, v): (&Array2D, &Array2D), dt: f32, diff: f32) { // src/lib.rs:226
        if self.prev_dirty {
            check(unsafe { fruid_density_upload(self.handle, DENS_PREV, self.dens_prev.data().as_ptr()) });
            self.prev_dirty = false;
        }
        // (u, v) are host arrays here.  When they are a live FluidSim's uv() -- `c.step(sim.uv(), ..)`, reference
        // src/main.rs:115 -- the library recognises the bound mirror pointers (fruid_fluid_bind_mirror in
        // FluidSim::new) and uses the device-resident velocity; any other pair is uploaded.
        check(unsafe { fruid_density_step_host(self.handle, u.data().as_ptr(), v.data().as_ptr(), dt, diff) });
        self.dens_stale.set(true);
        self.prev_stale = true;
    }

    pub fn density(&self) -> &Array2D { // src/lib.rs:238
        if self.dens_stale.get() {
            let a = unsafe { &mut *self.dens.get() };
            check(unsafe { fruid_density_download(self.handle, DENS, a.data_mut().as_mut_ptr()) });
            self.dens_stale.set(false);
        }
        unsafe { &*self.dens.get() }
    }

    pub fn density_mut(&mut self) -> &mut Array2D { // src/lib.rs:242: dens_prev, not dens
        if self.prev_stale {
            check(unsafe { fruid_density_download(self.handle, DENS_PREV, self.dens_prev.data_mut().as_mut_ptr()) });
            self.prev_stale = false;
        }
        self.prev_dirty = true;
        &mut self.dens_prev
    }
}

impl Drop for DensitySim {
    fn drop(&mut self) { unpin(self.dens.get_mut()); unpin(&mut self.dens_prev); unsafe { fruid_density_destroy(self.handle) }; }
}

pub struct FluidSim { // reference src/lib.rs:247-253
    u: UnsafeCell<Array2D>,
    v: UnsafeCell<Array2D>,
    u_prev: Array2D,
    v_prev: Array2D,
    handle: *mut FruidFluid,
    uv_stale: Cell<bool>,
    prev_stale: bool,
    prev_dirty: bool,
}

impl FluidSim {
    pub fn new(width: usize, height: usize) -> Self { // src/lib.rs:256
        let mut handle = std::ptr::null_mut();
        check(unsafe { fruid_fluid_create(width, height, &mut handle) });
        let arr = || Array2D::new(width, height);
        let mut s = Self { u: UnsafeCell::new(arr()), v: UnsafeCell::new(arr()), u_prev: arr(), v_prev: arr(), handle,
                           uv_stale: Cell::new(false), prev_stale: false, prev_dirty: false };
        pin(s.u.get_mut()); pin(s.v.get_mut()); pin(&mut s.u_prev); pin(&mut s.v_prev);
        // the Vec buffers never move: uv() always hands out these two arrays (see DensitySim::step)
        check(unsafe { fruid_fluid_bind_mirror(s.handle, U, s.u.get_mut().data().as_ptr()) });
        check(unsafe { fruid_fluid_bind_mirror(s.handle, V, s.v.get_mut().data().as_ptr()) });
        s
    }

    pub fn step(&mut self, dt: f32, visc: f32) { // src/lib.rs:267
        if self.prev_dirty {
            check(unsafe { fruid_fluid_upload(self.handle, U_PREV, self.u_prev.data().as_ptr()) });
            check(unsafe { fruid_fluid_upload(self.handle, V_PREV, self.v_prev.data().as_ptr()) });
            self.prev_dirty = false;
        }
        check(unsafe { fruid_fluid_step(self.handle, dt, visc) });
        self.uv_stale.set(true);
        self.prev_stale = true;
    }

    pub fn uv(&self) -> (&Array2D, &Array2D) { // src/lib.rs:279
        if self.uv_stale.get() {
            let (u, v) = unsafe { (&mut *self.u.get(), &mut *self.v.get()) };
            check(unsafe { fruid_fluid_download(self.handle, U, u.data_mut().as_mut_ptr()) });
            check(unsafe { fruid_fluid_download(self.handle, V, v.data_mut().as_mut_ptr()) });
            self.uv_stale.set(false);
        }
        unsafe { (&*self.u.get(), &*self.v.get()) }
    }

    pub fn uv_mut(&mut self) -> (&mut Array2D, &mut Array2D) { // src/lib.rs:283: (u_prev, v_prev)
        if self.prev_stale { // after a step they hold the pressure and the divergence (src/lib.rs:207)
            check(unsafe { fruid_fluid_download(self.handle, U_PREV, self.u_prev.data_mut().as_mut_ptr()) });
            check(unsafe { fruid_fluid_download(self.handle, V_PREV, self.v_prev.data_mut().as_mut_ptr()) });
            self.prev_stale = false;
        }
        self.prev_dirty = true;
        (&mut self.u_prev, &mut self.v_prev)
    }

    pub fn width(&self) -> usize { unsafe { &*self.u.get() }.width() }   // src/lib.rs:287
    pub fn height(&self) -> usize { unsafe { &*self.u.get() }.height() } // src/lib.rs:291
}

impl Drop for FluidSim {
    fn drop(&mut self) {
        unpin(self.u.get_mut()); unpin(self.v.get_mut()); unpin(&mut self.u_prev); unpin(&mut self.v_prev);
        unsafe { fruid_fluid_destroy(self.handle) };
    }
}
