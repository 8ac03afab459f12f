//! `extern "C"` binding of libtod_b200.so (include/tod.h) plus the two safe wrappers the reference's
//! frame loop needs.  The wrappers keep the reference's signatures and its panic-on-error behaviour:
//!
//!   reference (icf3ver/tiny-object-detection)          this crate
//!   -------------------------------------------------  ------------------------------------------
//!   `Yolact::init() -> Yolact`        (yolact.rs:17)    `Yolact::init() -> Yolact`
//!   `Yolact::classify(&mut [u32])`    (yolact.rs:39)    `Yolact::classify(&mut [u32])`
//!   GPU half of `append_scene`        (scene.rs:147)    `SceneGpu::append(&[u16], &[u16]) -> Scene`
//!   `struct Scene { height, pos, balls, connections }`  same fields, same element types
//!
//! Not compiled in the build container of this repository (no Rust toolchain there); the same ABI is
//! exercised by tools/frame_loop.cpp and the pytest suite.
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_float, c_int, c_void, CStr, CString};

pub mod sys {
    use super::*;

    #[repr(C)]
    pub struct tod_scene { _private: [u8; 0] }
    #[repr(C)]
    pub struct tod_yolact { _private: [u8; 0] }
    #[repr(C)]
    pub struct tod_pool { _private: [u8; 0] }

    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct tod_scene_params {
        pub width: i32,
        pub height: i32,
        pub max_depth_in: c_float,
        pub x_fov: c_float,
        pub y_fov: c_float,
        pub bot_avoidance_const: c_float,
        pub bot_norm_const: i32,
        pub terrain_norm_const: i32,
        pub bump_err: c_float,
        pub sample_shift: i32,
        pub weights_mode: i32,
        pub max_batch: i32,
    }

    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct tod_yolact_options {
        pub max_tiles: i32,
        pub id_mode: i32,
        pub conf_thresh: c_float,
        pub nms_thresh: c_float,
        pub top_k: i32,
        pub max_dets: i32,
        pub use_cuda_graph: i32,
        pub conv_impl: i32,
        pub fusion: i32,
        pub use_pdl: i32,
        pub batches_in_flight: i32,
    }

    extern "C" {
        pub fn tod_last_error() -> *const c_char;
        pub fn tod_abi_version() -> c_int;
        pub fn tod_scene_default_params(p: *mut tod_scene_params);
        pub fn tod_scene_create(device: c_int, params: *const tod_scene_params, out: *mut *mut tod_scene) -> c_int;
        pub fn tod_scene_destroy(s: *mut tod_scene);
        pub fn tod_scene_append_batch(s: *mut tod_scene, depth: *const u16, target: *const u16, n: c_int, map: *mut u32,
                                      world4: *mut f32, conn0: *mut f32, conn1: *mut f32, balls4: *mut f32) -> c_int;
        pub fn tod_scene_materialize(s: *mut tod_scene, frame: c_int, height: *mut f32, pos3: *mut f32, balls2: *mut i32,
                                     connections8: *mut f32) -> c_int;
        pub fn tod_yolact_default_options(o: *mut tod_yolact_options);
        pub fn tod_yolact_create(path: *const c_char, device: c_int, opts: *const tod_yolact_options,
                                 out: *mut *mut tod_yolact) -> c_int;
        pub fn tod_yolact_destroy(y: *mut tod_yolact);
        pub fn tod_yolact_classify(y: *mut tod_yolact, frame: *mut u32, width: c_int, height: c_int) -> c_int;
        pub fn tod_yolact_classify_batch(y: *mut tod_yolact, frames: *mut u32, n: c_int, width: c_int, height: c_int) -> c_int;
        // frame sharder over every GPU of the box (include/tod.h "Pool")
        pub fn tod_pool_create(path: *const c_char, devices: *const i32, n_devices: c_int, depth: c_int,
                               opts: *const tod_yolact_options, out: *mut *mut tod_pool) -> c_int;
        pub fn tod_pool_destroy(p: *mut tod_pool);
        pub fn tod_pool_classify_batch(p: *mut tod_pool, frames: *mut u32, n: c_int, width: c_int, height: c_int) -> c_int;
        pub fn tod_pool_rgbd_batch(p: *mut tod_pool, scene_params: *const tod_scene_params, frames: *mut u32, depth: *const u16, n: c_int,
                                   map: *mut u32, world4: *mut f32, conn0: *mut f32, conn1: *mut f32, balls4: *mut f32) -> c_int;
        // the Scene's consumer (path.rs)
        pub fn tod_path_modify(device: c_int, width: c_int, height_px: c_int, height: *const f32, pos3: *const f32, balls2: *const i32,
                               connections8: *const f32, cost_out: *mut f32, pred_out: *mut i32, directions: *mut f32, cap: c_int,
                               n_directions: *mut i32) -> c_int;
        pub fn tod_path_serialize(created_secs: u64, directions: *const f32, n: c_int, out: *mut u8, cap: usize, bytes: *mut usize) -> c_int;
    }
}

fn expect(rc: c_int, what: &str) {
    // the reference `.expect()`s / `.unwrap()`s every fallible call (yolact.rs:20-35,163; scene.rs:152-282)
    if rc < 0 {
        let msg = unsafe { CStr::from_ptr(sys::tod_last_error()) }.to_string_lossy().into_owned();
        panic!("{}: {} ({})", what, msg, rc);
    }
}

/// Drop-in for `yolact::Yolact` (src/yolact.rs:13-41).
pub struct Yolact {
    handle: *mut sys::tod_yolact,
}

// One owner at a time, like `&mut self` on the reference type; the handle may move between tokio workers.
unsafe impl Send for Yolact {}

impl Yolact {
    /// `Yolact::init()` (yolact.rs:17).  The reference hard-codes the EdgeTPU model path (yolact.rs:19);
    /// this build runs the CPU model of the same network.
    pub fn init() -> Yolact {
        Self::init_with("data/FRC_model.tflite", 0)
    }

    pub fn init_with(model: &str, device: i32) -> Yolact {
        let path = CString::new(model).expect("model path");
        let mut opts = std::mem::MaybeUninit::<sys::tod_yolact_options>::uninit();
        let mut handle = std::ptr::null_mut();
        unsafe {
            sys::tod_yolact_default_options(opts.as_mut_ptr());
            expect(sys::tod_yolact_create(path.as_ptr(), device, opts.as_ptr(), &mut handle), "Yolact::init");
        }
        Yolact { handle }
    }

    /// `classify(&mut self, frame_buffer: &mut [u32])` (yolact.rs:39): 640x480 pixels `r<<24|g<<16|b<<8`,
    /// replaced in place by `[class,0,0,0]` big-endian (yolact.rs:231-233).
    pub fn classify(&mut self, frame_buffer: &mut [u32]) {
        assert_eq!(frame_buffer.len(), 640 * 480, "frame buffer must hold 640x480 pixels (yolact.rs:233 copy_from_slice)");
        expect(unsafe { sys::tod_yolact_classify(self.handle, frame_buffer.as_mut_ptr(), 640, 480) }, "Yolact::classify");
    }
}

impl Drop for Yolact {
    fn drop(&mut self) {
        unsafe { sys::tod_yolact_destroy(self.handle) }
    }
}

/// `scene::Scene` (src/scene.rs:122-132).
pub struct Scene {
    pub height: Vec<f32>,
    pub pos: Vec<(f32, f32, f32)>,
    pub balls: Vec<(i32, i32)>,
    pub connections: Vec<[f32; 8]>,
}

/// The GPU half of `append_scene` (src/scene.rs:152-327): what the two Vulkan dispatches, the uploads and the four
/// read-backs do, with the per-call Vulkan objects hoisted into the handle.
pub struct SceneGpu {
    handle: *mut sys::tod_scene,
}

unsafe impl Send for SceneGpu {}

impl SceneGpu {
    pub fn new(device: i32) -> SceneGpu {
        let mut p = std::mem::MaybeUninit::<sys::tod_scene_params>::uninit();
        let mut handle = std::ptr::null_mut();
        unsafe {
            sys::tod_scene_default_params(p.as_mut_ptr());
            expect(sys::tod_scene_create(device, p.as_ptr(), &mut handle), "SceneGpu::new");
        }
        SceneGpu { handle }
    }

    /// depth / target: the two `[u16; 640*480]` popped at scene.rs:186-187.  Blocks like `future.wait` (scene.rs:282).
    pub fn append(&mut self, depth: &[u16], target: &[u16]) -> Scene {
        const N: usize = 640 * 480;
        assert!(depth.len() == N && target.len() == N);
        let null = std::ptr::null_mut::<c_void>();
        let mut height = vec![0f32; N];
        // The C side writes flat scalars.  Arrays have a guaranteed layout; Rust tuples do not (`repr(Rust)` may reorder
        // or pad), so the tuple-typed `Scene` fields are filled by an explicit conversion, exactly as scene.rs:316-322 does.
        let mut pos3 = vec![[0f32; 3]; N];
        let mut balls2 = vec![[0i32; 2]; 100];
        let mut connections = vec![[0f32; 8]; N];
        unsafe {
            expect(sys::tod_scene_append_batch(self.handle, depth.as_ptr(), target.as_ptr(), 1, null as *mut u32, null as *mut f32,
                                               null as *mut f32, null as *mut f32, null as *mut f32), "append_scene");
            expect(sys::tod_scene_materialize(self.handle, 0, height.as_mut_ptr(), pos3.as_mut_ptr() as *mut f32,
                                              balls2.as_mut_ptr() as *mut i32, connections.as_mut_ptr() as *mut f32), "append_scene");
        }
        let pos = pos3.iter().map(|p| (p[0], p[1], p[2])).collect();
        let balls = balls2.iter().map(|b| (b[0], b[1])).collect();
        Scene { height, pos, balls, connections }
    }
}

impl Drop for SceneGpu {
    fn drop(&mut self) {
        unsafe { sys::tod_scene_destroy(self.handle) }
    }
}

/// `path::Path` (src/path.rs:11-22) and the planner step `modify_path` (path.rs:25-120, intent mode: the reference
/// function indexes 224*224-element arrays with 640x480 node numbers and panics on every input).
pub struct Path {
    pub created: std::time::SystemTime,
    pub directions: Vec<(f32, f32)>,
}

impl Path {
    /// path.rs:17-21: big-endian seconds since the epoch, then big-endian (magnitude, rotation) pairs.
    pub fn serialize(&self) -> Vec<u8> {
        let secs = self.created.duration_since(std::time::UNIX_EPOCH).expect("Incorrect System Time").as_secs();
        let flat: Vec<[f32; 2]> = self.directions.iter().map(|d| [d.0, d.1]).collect();
        let mut out = vec![0u8; 8 + 8 * flat.len()];
        let mut n = 0usize;
        expect(unsafe { sys::tod_path_serialize(secs, flat.as_ptr() as *const f32, flat.len() as c_int, out.as_mut_ptr(), out.len(), &mut n) },
               "Path::serialize");
        out.truncate(n);
        out
    }
}

/// `modify_path(path, scene)`: overwrites `path` with the route from the start node to the nearest ball.
pub fn modify_path(path: &mut Path, scene: &Scene, device: i32) {
    let pos3: Vec<[f32; 3]> = scene.pos.iter().map(|p| [p.0, p.1, p.2]).collect();
    let balls2: Vec<[i32; 2]> = scene.balls.iter().map(|b| [b.0, b.1]).collect();
    let cap = 640 * 480;
    let mut dirs = vec![[0f32; 2]; cap];
    let mut n: i32 = 0;
    expect(unsafe {
        sys::tod_path_modify(device, 640, 480, scene.height.as_ptr(), pos3.as_ptr() as *const f32, balls2.as_ptr() as *const i32,
                             scene.connections.as_ptr() as *const f32, std::ptr::null_mut(), std::ptr::null_mut(),
                             dirs.as_mut_ptr() as *mut f32, cap as c_int, &mut n)
    }, "modify_path");
    dirs.truncate(n.max(0) as usize);
    *path = Path { created: std::time::SystemTime::now(), directions: dirs.iter().map(|d| (d[0], d[1])).collect() };
}
