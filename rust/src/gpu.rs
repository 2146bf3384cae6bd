//! gpu.rs — copy to src/gpu.rs of pathtrace-rs (`mod gpu;` behind feature "gpu", rust/patches/06-src-main_rs.diff).  Thin FFI over
//! include/ptgpu.h (ABI v4): flattens a `Scene` whose world is a `Hitable::List` of spheres / moving spheres, uploads it to
//! one or several GPUs and serves `Scene::update` from the CUDA library.  NOT compiled in this repository's CI (no Rust
//! toolchain in the build image); the identical C ABI is exercised by the C++ mirror and the ctypes tests.
//!
//! The accessors it needs are added by rust/patches/*.diff (all `pub(crate)`, the fields are private):
//!   Camera::to_ffi(&self) -> PtCamera                                   03-src-camera_rs.diff        (src/camera.rs:8-19)
//!   Scene::world(&self) -> &Hitable, Scene::sky_colour(&self) -> Option<Vec3>, Scene::attach_gpu     09-src-scene_rs.diff
//!   HitableList::hitables(&self) -> &[Hitable]                          04-...hitable_list_rs.diff   (hitable_list.rs:9-11)
//!   MovingSphere::endpoints(&self) -> (centre0, centre1, time0, time1)  05-...moving_sphere_rs.diff  (moving_sphere.rs:7-26)
//!   Perlin::tables(&self) -> (&[Vec3], &[u32], &[u32], &[u32])          08-src-perlin_rs.diff        (src/perlin.rs:7-12)
//!   RgbImage::raw(&self) -> (u32, u32, &[u8])                           10-src-texture_rs.diff       (src/texture.rs:6-10)
#![cfg(feature = "gpu")]
use crate::{camera::Camera, collision::Hitable, material::Material, params::Params, perlin::Perlin, scene::Scene, texture::{RgbImage, Texture}};
use std::{collections::HashMap, ffi::CStr, os::raw::{c_char, c_int, c_void}, ptr};

#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct PtParams { pub width: u32, pub height: u32, pub samples: u32, pub max_depth: u32, pub random_seed: u8, pub use_bvh: u8, pub _pad: [u8; 6], pub seed_salt: u64 }
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct PtCamera { pub origin: [f32; 3], pub lower_left_corner: [f32; 3], pub horizontal: [f32; 3], pub vertical: [f32; 3], pub u: [f32; 3], pub v: [f32; 3], pub w: [f32; 3], pub time0: f32, pub time1: f32, pub lens_radius: f32 }
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct PtTexture { pub kind: i32, pub color: [f32; 3], pub odd: i32, pub even: i32, pub scale: f32, pub image: i32 }
/// src/texture.rs:6-10: `image.to_rgb8().into_raw()` (copied by pt_scene_create)
#[repr(C)] #[derive(Copy, Clone)]
pub struct PtImage { pub width: u32, pub height: u32, pub data: *const u8 }
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct PtMaterial { pub kind: i32, pub texture: i32, pub albedo: [f32; 3], pub fuzz: f32, pub ref_idx: f32, pub _pad: i32 }
#[repr(C)]
pub struct PtPerlin { pub randvec: [[f32; 3]; 256], pub perm_x: [u32; 256], pub perm_y: [u32; 256], pub perm_z: [u32; 256] }
#[repr(C)]
pub struct PtSceneDesc {
    pub struct_size: u32, pub n_spheres: u32,
    pub centre_x: *const f32, pub centre_y: *const f32, pub centre_z: *const f32, pub radius: *const f32, pub material_index: *const i32,
    pub n_materials: u32, pub n_textures: u32, pub materials: *const PtMaterial, pub textures: *const PtTexture, pub perlin: *const PtPerlin,
    pub has_sky: u32, pub sky: [f32; 3],
    pub motion: *const PtMotion, // per sphere, or null when nothing moves
    pub n_images: u32, pub _pad: u32, pub images: *const PtImage,
}
/// src/collision/moving_sphere.rs:16-26 inverted to its constructor arguments (centre0/radius travel in the sphere arrays)
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct PtMotion { pub centre1: [f32; 3], pub time0: f32, pub time1: f32, pub moving: u32 }
/// explicit launch options; the shim passes NULL (every field "library decides") except `tile_rows`
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct PtOptions { pub struct_size: u32, pub force_stream_tile_blocks: i32, pub stream_ctas: i32, pub chunk_samples: i32, pub spatial_order: i32, pub tile_rows: u32, pub resident_kernel: u32 }
#[repr(C)] pub struct PtScene { _private: [u8; 0] }

extern "C" {
    fn pt_abi_version() -> c_int;
    fn pt_abi_struct_size(which: c_int) -> u32;
    fn pt_last_error() -> *const c_char;
    fn pt_scene_create_multi(desc: *const PtSceneDesc, devices: *const c_int, n_devices: u32, options: *const PtOptions, out: *mut *mut PtScene) -> c_int;
    fn pt_scene_destroy(scene: *mut PtScene);
    fn pt_render(scene: *mut PtScene, params: *const PtParams, camera: *const PtCamera, frame_num: u32, rgb_inout: *mut f32, ray_count_out: *mut u64) -> c_int;
    fn pt_render_progressive(scene: *mut PtScene, params: *const PtParams, camera: *const PtCamera, frame_num: u32, rgb_out: *mut f32, rgb8_out: *mut u8, ray_count_out: *mut u64) -> c_int;
}

fn last_error() -> String { unsafe { CStr::from_ptr(pt_last_error()).to_string_lossy().into_owned() } }

/// Device copy of a flattened `Scene`.  Dropping it frees the device memory (ties `pt_scene_destroy` to a Rust `Drop`).
pub struct GpuScene { handle: *mut PtScene }
// `Scene` must stay `Sync` (its CPU `update` hands `&self` to rayon); the handle is only ever used by `update`, which the
// reference calls from one thread at a time (main thread offline, one worker thread in the windowed mode)
unsafe impl Send for GpuScene {}
unsafe impl Sync for GpuScene {}

impl GpuScene {
    /// The GPU arm of `Params::new_scene` (src/params.rs:29-46): walk `Hitable::List`, reject anything that is not a sphere
    /// (same message as the panic in src/collision/spheres_soa.rs:49-51), dedupe arena pointers into indices, upload.
    /// `devices`: CUDA device ordinals.  With more than one, every `update` is split over them by interleaved row tiles inside
    /// the library (one host thread per GPU, each copying its rows from / into `buffer`); the image does not depend on the list.
    pub fn new(scene: &Scene, perlin: &Perlin, devices: &[i32]) -> GpuScene {
        assert_eq!(unsafe { pt_abi_version() }, 4);
        assert_eq!(unsafe { pt_abi_struct_size(11) } as usize, std::mem::size_of::<PtOptions>());
        assert!(!devices.is_empty(), "empty GPU device list");
        assert_eq!(unsafe { pt_abi_struct_size(5) } as usize, std::mem::size_of::<PtSceneDesc>());
        let hitables = match scene.world() { Hitable::List(list) => list.hitables(), other => panic!("Expected Hitable::List, got {:?}", other) };
        let (mut cx, mut cy, mut cz, mut radius, mut mat_index) = (vec![], vec![], vec![], vec![], vec![]);
        let (mut materials, mut textures): (Vec<PtMaterial>, Vec<PtTexture>) = (vec![], vec![]);
        let mut motion: Vec<PtMotion> = vec![];
        let mut any_moving = false;
        let mut mat_ids: HashMap<*const Material, i32> = HashMap::new();
        let mut tex_ids: HashMap<*const Texture, i32> = HashMap::new();
        let mut uses_noise = false;
        let mut images: Vec<PtImage> = vec![];
        let mut image_ids: HashMap<*const RgbImage, i32> = HashMap::new();
        struct Tables<'t> { textures: &'t mut Vec<PtTexture>, ids: &'t mut HashMap<*const Texture, i32>, images: &'t mut Vec<PtImage>,
                            image_ids: &'t mut HashMap<*const RgbImage, i32>, uses_noise: &'t mut bool }
        fn texture_id(t: &Texture, tb: &mut Tables) -> i32 {
            if let Some(id) = tb.ids.get(&(t as *const Texture)) { return *id; }
            let mut f = PtTexture { odd: -1, even: -1, image: -1, ..Default::default() };
            match t {
                Texture::Constant { color } => { f.kind = 0; f.color = [color.x, color.y, color.z]; }
                Texture::Checker { odd, even } => { f.kind = 1; f.odd = texture_id(odd, tb); f.even = texture_id(even, tb); }
                Texture::Noise { scale, .. } => { f.kind = 2; f.scale = *scale; *tb.uses_noise = true; }
                Texture::Image { image } => {
                    f.kind = 3;
                    let images = &mut *tb.images;
                    f.image = *tb.image_ids.entry(*image as *const RgbImage).or_insert_with(|| {
                        let (width, height, data) = image.raw();
                        images.push(PtImage { width, height, data: data.as_ptr() });
                        images.len() as i32 - 1
                    });
                }
            }
            let (textures, ids) = (&mut *tb.textures, &mut *tb.ids);
            textures.push(f);
            let id = textures.len() as i32 - 1;
            ids.insert(t as *const Texture, id);
            id
        }
        for hitable in hitables {
            // Hitable::Sphere and Hitable::MovingSphere (src/collision/hitable.rs:16-17) are the two arms the GPU path takes.
            let material = match hitable {
                Hitable::Sphere(sphere, material) => {
                    cx.push(sphere.centre().x); cy.push(sphere.centre().y); cz.push(sphere.centre().z); radius.push(sphere.radius());
                    motion.push(PtMotion::default());
                    Some(material)
                }
                Hitable::MovingSphere(ms, material) => {
                    let (c0, c1, time0, time1) = ms.endpoints();
                    cx.push(c0.x); cy.push(c0.y); cz.push(c0.z); radius.push(ms.radius());
                    motion.push(PtMotion { centre1: [c1.x, c1.y, c1.z], time0, time1, moving: 1 });
                    any_moving = true;
                    Some(material)
                }
                _ => None,
            };
            if let Some(material) = material {
                let key = *material as *const Material;
                let id = *mat_ids.entry(key).or_insert_with(|| {
                    let mut f = PtMaterial { texture: -1, ..Default::default() };
                    match material {
                        Material::Lambertian { albedo } => { f.kind = 0; f.texture = texture_id(albedo, &mut Tables { textures: &mut textures, ids: &mut tex_ids, images: &mut images, image_ids: &mut image_ids, uses_noise: &mut uses_noise }); }
                        Material::Metal { albedo, fuzz } => { f.kind = 1; f.albedo = [albedo.x, albedo.y, albedo.z]; f.fuzz = *fuzz; }
                        Material::Dielectric { ref_idx } => { f.kind = 2; f.ref_idx = *ref_idx; }
                        Material::DiffuseLight { emit } => { f.kind = 3; f.texture = texture_id(emit, &mut Tables { textures: &mut textures, ids: &mut tex_ids, images: &mut images, image_ids: &mut image_ids, uses_noise: &mut uses_noise }); }
                        Material::Isotropic { .. } => panic!("Material::Isotropic is not supported on the GPU path"),
                    }
                    materials.push(f);
                    materials.len() as i32 - 1
                });
                mat_index.push(id);
            } else {
                panic!("Expected Hitable::Sphere, got {:?}", hitable);
            }
        }
        let tables = if uses_noise {
            let (rv, px, py, pz) = perlin.tables();
            let mut t = Box::new(PtPerlin { randvec: [[0.0; 3]; 256], perm_x: [0; 256], perm_y: [0; 256], perm_z: [0; 256] });
            for i in 0..256 { t.randvec[i] = [rv[i].x, rv[i].y, rv[i].z]; t.perm_x[i] = px[i]; t.perm_y[i] = py[i]; t.perm_z[i] = pz[i]; }
            Some(t)
        } else { None };
        let sky = scene.sky_colour();
        let desc = PtSceneDesc {
            struct_size: std::mem::size_of::<PtSceneDesc>() as u32, n_spheres: cx.len() as u32,
            centre_x: cx.as_ptr(), centre_y: cy.as_ptr(), centre_z: cz.as_ptr(), radius: radius.as_ptr(), material_index: mat_index.as_ptr(),
            n_materials: materials.len() as u32, n_textures: textures.len() as u32, materials: materials.as_ptr(), textures: textures.as_ptr(),
            perlin: tables.as_ref().map_or(ptr::null(), |t| &**t as *const PtPerlin),
            has_sky: sky.is_some() as u32, sky: sky.map_or([0.0; 3], |s| [s.x, s.y, s.z]),
            motion: if any_moving { motion.as_ptr() } else { ptr::null() },
            n_images: images.len() as u32, _pad: 0, images: if images.is_empty() { ptr::null() } else { images.as_ptr() },
        };
        let mut handle: *mut PtScene = ptr::null_mut();
        let rc = unsafe { pt_scene_create_multi(&desc, devices.as_ptr(), devices.len() as u32, ptr::null(), &mut handle) };
        if rc != 0 { panic!("pt_scene_create_multi failed: {}", last_error()); }
        GpuScene { handle }
    }

    /// Same signature and return as `Scene::update` (src/scene.rs:73-79).
    pub fn update(&self, params: &Params, camera: &Camera, frame_num: u32, buffer: &mut [(f32, f32, f32)]) -> usize {
        assert_eq!(std::mem::size_of::<(f32, f32, f32)>(), 12); // tuple layout is not guaranteed: checked, then reinterpreted
        assert_eq!(buffer.len(), (params.width * params.height) as usize);
        let p = PtParams { width: params.width, height: params.height, samples: params.samples, max_depth: params.max_depth,
                           random_seed: params.random_seed as u8, use_bvh: params.use_bvh as u8, _pad: [0; 6],
                           seed_salt: if params.random_seed { rand::random() } else { 0 } };
        let cam = camera.to_ffi();
        let mut rays = 0u64;
        let rc = unsafe { pt_render(self.handle, &p, &cam, frame_num, buffer.as_mut_ptr() as *mut f32, &mut rays) };
        if rc != 0 { panic!("pt_render failed: {}", last_error()); }
        rays as usize
    }
}

impl GpuScene {
    /// The worker loop of the windowed mode (src/glium_window.rs:96-131) with the accumulation buffer resident on the GPU:
    /// `frame_num == 0` restarts, `frame_num > 0` must follow the previous call.  `rgb8` receives what the window uploads
    /// (top-down sRGB bytes, src/glium_window.rs:108-121 / src/offline.rs:43-51); pass `None` for frames that are not shown.
    pub fn update_progressive(&self, params: &Params, camera: &Camera, frame_num: u32, rgb8: Option<&mut [u8]>) -> usize {
        let p = PtParams { width: params.width, height: params.height, samples: params.samples, max_depth: params.max_depth,
                           random_seed: params.random_seed as u8, use_bvh: params.use_bvh as u8, _pad: [0; 6],
                           seed_salt: if params.random_seed { rand::random() } else { 0 } };
        let cam = camera.to_ffi();
        let mut rays = 0u64;
        let out = match rgb8 {
            Some(buf) => { assert_eq!(buf.len(), (params.width * params.height * 3) as usize); buf.as_mut_ptr() }
            None => ptr::null_mut(),
        };
        let rc = unsafe { pt_render_progressive(self.handle, &p, &cam, frame_num, ptr::null_mut(), out, &mut rays) };
        if rc != 0 { panic!("pt_render_progressive failed: {}", last_error()); }
        rays as usize
    }
}

impl Drop for GpuScene {
    fn drop(&mut self) { unsafe { pt_scene_destroy(self.handle) } }
}
#[allow(dead_code)] fn _unused(_: *mut c_void) {}
